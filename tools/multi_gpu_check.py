"""Run under torchrun on N >= 2 GPUs: (1) M-sharded database over NCCL == unsharded; (2) query sharding == one GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from argparse import Namespace
from oracle import range_oracle as O
from range_b200.range import LocationEncoder
from range_b200.distributed import shard_rows

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
M, N = int(os.environ.get("M", 200_000)), int(os.environ.get("N", 40_000))
rng = np.random.default_rng(0)
db = dict(locs=O.area_uniform(M, rng), satclip_embeddings=rng.standard_normal((M, 256), dtype=np.float32),
          image_embeddings=rng.standard_normal((M, 1024), dtype=np.float32) + 0.5)
enc = dict(L=40, dims=[1600, 512, 512, 256], weights=O.siren_init(40, 512, 2, 256, seed=0))
coords = torch.tensor(O.area_uniform(N, np.random.default_rng(1)), device=dev)

def rel(a, b):
    return ((a - b).norm(dim=1) / b.norm(dim=1)).max().item()

full = LocationEncoder(Namespace(location_model_name="RANGE+", pretrained_path=enc, device=dev, range_db=db, beta=0.5))
ref = full.embed(coords)                                           # every rank: unsharded database, all queries
shard = LocationEncoder(Namespace(location_model_name="RANGE+", pretrained_path=enc, device=dev, range_db=db, beta=0.5,
                                  db_shard=(rank, world), db_group=dist.group.WORLD))
out = shard.embed(coords)                                          # M-sharded: stats SUM/MAX + output SUM over NCCL
e1 = rel(out[:, :1024], ref[:, :1024])
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(3):
    out = shard.embed(coords)
torch.cuda.synchronize(); dist.barrier()
t_sh = (time.perf_counter() - t0) / 3
t0 = time.perf_counter()
for _ in range(3):
    ref = full.embed(coords)
torch.cuda.synchronize(); dist.barrier()
t_full = (time.perf_counter() - t0) / 3
lo, hi = shard_rows(N, rank, world)                                # query sharding: my slab only, then gather
mine = full.embed(coords[lo:hi].contiguous())
parts = [torch.empty(shard_rows(N, r, world)[1] - shard_rows(N, r, world)[0], 1280, device=dev) for r in range(world)]
dist.all_gather(parts, mine)
e2 = rel(torch.cat(parts)[:, :1024], ref[:, :1024])
if rank == 0:
    print(f"world {world}: M-sharded vs unsharded max rel-row err {e1:.2e} (time {t_sh*1e3:.1f} ms vs {t_full*1e3:.1f} ms unsharded, "
          f"N={N} M={M}); query-sharded + all_gather vs one GPU {e2:.2e}")
    assert e1 < 1e-3 and e2 < 1e-3
dist.destroy_process_group()
