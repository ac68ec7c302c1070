# usage: tools/run_mg.sh <n_gpus>   (inside gpurun --gpus N)
n=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py 2>&1 | grep -v "Using RANGE\|\*\*\*\|OMP_NUM" | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 5 --warmup 3 2>gpurun_out/mg_bench.err | tee gpurun_out/bench_${n}gpu.json | cut -c1-330
CONFIGS=${CONFIGS:-2,3,4,5} M_MAX=${M_MAX:-3000000} python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 tools/configs.py 2>gpurun_out/mg_configs.err | tee gpurun_out/configs_${n}gpu.jsonl
