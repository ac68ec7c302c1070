#!/bin/bash
# (ncu cannot replay the cooperative launch of the apply kernel: the profiled runs use the plain cluster launch, RANGE_PC_COOP=0)
# One GPU-box visit (round 2): parity tests, bench, smoke, host-path / launch-mode / beta-sweep timings, ncu launch list
# + `--set full` capture of the retrieval kernels.   usage: tools/gpu_round2.sh <tag>   (outputs: gpurun_out/<tag>_*)
tag=${1:-r2}
mkdir -p gpurun_out
o=gpurun_out/${tag}
timeout 1500 python -m pytest tests -m gpu -q ${PYTEST_ARGS:-} > ${o}_pytest.log 2>&1; echo "pytest exit $?" | tee -a ${o}_pytest.log
tail -15 ${o}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > ${o}_bench.json 2> ${o}_bench.err; echo "bench exit $?"
cat ${o}_bench.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${o}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 ${o}_smoke.log
SETTINGS=copy:24576:6144:0.5,packed:24576:6144:0.5 timeout 400 python tools/time_e2e.py > ${o}_time_e2e.log 2>&1; echo "time_e2e exit $?"; grep -v "Using RANGE" ${o}_time_e2e.log | tail -14
timeout 300 python tools/time_apply.py > ${o}_time_apply.log 2>&1; tail -2 ${o}_time_apply.log
RANGE_PC_COOP=0 timeout 300 python tools/time_apply.py 2>&1 | tail -1 | sed 's/^/plain cluster launch (RANGE_PC_COOP=0): /' | tee -a ${o}_time_apply.log
RANGE_PC_PERSIST=1 timeout 300 python tools/time_apply.py 2>&1 | tail -1 | sed 's/^/persisting-L2 window over the scratch (RANGE_PC_PERSIST=1): /' | tee -a ${o}_time_apply.log
PROF=1 RANGE_APPLY_KERNEL=pc timeout 300 python tools/time_apply.py 2>&1 | tail -9 > ${o}_apply_roles.log; cat ${o}_apply_roles.log
CONFIGS=3 timeout 600 python tools/configs.py > ${o}_config3_1gpu.jsonl 2> ${o}_config3.err; echo "config3 exit $?"; cat ${o}_config3_1gpu.jsonl
[ -x build/probe_ex2 ] && build/probe_ex2 | tee ${o}_probe_ex2.log
if [ "${NCU:-1}" = "1" ]; then
RANGE_PC_COOP=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${o}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${o}_ncu_bench.log 2>&1; echo "ncu list exit $?"
RANGE_PC_COOP=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'range_(apply_pc|stats_pc)_kernel' -s 4 -c 2 \
  -o ${o}_k2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${o}_ncu_full.log 2>&1; echo "ncu full exit $?"
RANGE_PC_COOP=0 RANGE_PC_PERSIST=1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:'range_apply_pc_kernel' -s 2 -c 1 --csv --log-file ${o}_apply_persist.csv python tools/time_apply.py > /dev/null 2>&1; echo "ncu persist exit $?"
fi
