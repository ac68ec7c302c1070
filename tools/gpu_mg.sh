#!/bin/bash
# Multi-GPU visit (inside gpurun --gpus N): M-sharded check, bench with the m_sharded section, optional extras.
# usage: tools/gpu_mg.sh <n_gpus> <tag>
n=${1:-2}; tag=${2:-r2mg}
o=gpurun_out/${tag}
mkdir -p gpurun_out
run() { timeout ${T:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
M=${CHECK_M:-200000} N=${CHECK_N:-20000} T=600 run 29511 tools/multi_gpu_check.py > ${o}_check.log 2>&1; echo "check exit $?"; grep -v "Using RANGE\|\*\*\*\|OMP_NUM" ${o}_check.log | tail -8
NCCL_DEBUG=${NCCL_DEBUG:-WARN} T=900 run 29512 bench.py --gpus $n --steps ${STEPS:-10} --warmup 3 > ${o}_bench_${n}gpu.json 2> ${o}_bench_${n}gpu.err; echo "bench exit $?"; cat ${o}_bench_${n}gpu.json; tail -3 ${o}_bench_${n}gpu.err
if [ "${WALL:-0}" = "1" ]; then
T=600 run 29513 tools/d2h_wall.py > ${o}_d2h_wall.jsonl 2> ${o}_d2h_wall.err; echo "d2h wall exit $?"; cat ${o}_d2h_wall.jsonl
fi
if [ -n "${CONFIGS:-}" ]; then
CONFIGS=$CONFIGS T=900 run 29514 tools/configs.py > ${o}_configs_${n}gpu.jsonl 2> ${o}_configs.err; echo "configs exit $?"; cat ${o}_configs_${n}gpu.jsonl
fi
