n=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 5 --warmup 3 2>gpurun_out/mg_bench.err | tee gpurun_out/bench_${n}gpu.json | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py 2>&1 | grep "^world" | tail -1
CONFIGS=4 M_MAX=1000000 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 tools/configs.py 2>gpurun_out/mg_configs.err | tee gpurun_out/configs_${n}gpu.jsonl | cut -c1-260
