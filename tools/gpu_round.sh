#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, ncu --set full of the retrieval kernels.
# usage: tools/gpu_round.sh <tag>     (outputs under gpurun_out/<tag>_*)
tag=${1:-r}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
cat gpurun_out/${tag}_bench.json
timeout 300 python tools/time_apply.py > gpurun_out/${tag}_time_apply.log 2>&1; cat gpurun_out/${tag}_time_apply.log | tail -3
RANGE_PC_COOP=1 timeout 300 python tools/time_apply.py 2>&1 | tail -1
if [ "${NCU:-1}" = "1" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'range_(apply_pc|stats_pc)_kernel' -s 4 -c 2 \
  -o gpurun_out/${tag}_k2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full exit $?"
fi
if [ "${EXTRA:-0}" = "1" ]; then
# round-end extras: smoke(), the reference arm (short), one rank's share of the north-star case
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference arm exit $?"; cat gpurun_out/${tag}_bench_reference.json
CONFIGS=ns timeout 300 python tools/configs.py > gpurun_out/${tag}_north_star_share.jsonl 2> gpurun_out/${tag}_north_star_share.err; echo "ns exit $?"; cat gpurun_out/${tag}_north_star_share.jsonl
fi
