#!/bin/bash
# usage (inside gpurun --gpus 2): tools/e2e_diag.sh
f() { grep "^rank" ; }
echo "== one process"; python tools/e2e_diag.py 2>/dev/null | f
echo "== two independent processes"; (LOCAL_RANK=0 python tools/e2e_diag.py 2>/dev/null | f) & (LOCAL_RANK=1 python tools/e2e_diag.py 2>/dev/null | f); wait
echo "== torchrun x2, no process group"; NO_DIST=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/e2e_diag.py 2>/dev/null | f
echo "== torchrun x2, NCCL process group"; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 tools/e2e_diag.py 2>/dev/null | f
echo "== torchrun x2, NCCL, OMP_NUM_THREADS=8"; OMP_NUM_THREADS=8 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 tools/e2e_diag.py 2>/dev/null | f
nproc; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"; nvidia-smi topo -m | head -8
