"""Run under torchrun on N >= 2 GPUs: the M-sharded database path (range_b200/distributed.py) against the unsharded one.
Every rank owns its own queries (ragged counts, one rank's set clustered in a small region, several steps), the database
is sharded along M; both merges are checked: 'peer' (the apply kernel's epilogue stores partial rows into the owners'
receive buffers over NVLink) and 'reduce_scatter' (NCCL).  Also: query sharding + all_gather == one GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
"""
import contextlib, os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from argparse import Namespace
from range_b200 import synthetic as S          # seeded input generators
from range_b200.range import LocationEncoder
from range_b200.distributed import shard_rows

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
M, N = int(os.environ.get("M", 200_000)), int(os.environ.get("N", 20_000))
rng = np.random.default_rng(0)
db = dict(locs=S.area_uniform(M, rng), satclip_embeddings=rng.standard_normal((M, 256), dtype=np.float32),
          image_embeddings=rng.standard_normal((M, 1024), dtype=np.float32) + 0.5)
enc = dict(L=40, dims=[1600, 512, 512, 256], weights=S.siren_init(40, 512, 2, 256, seed=0))
n_mine = N + 1500 * rank                                            # ragged: the last step of the other ranks is padded
r2 = np.random.default_rng(100 + rank)
mine = S.area_uniform(n_mine, r2)
if rank == world - 1:                                               # a regional query set: thousands of queries per cell
    mine[: n_mine // 2] = np.stack([r2.uniform(10, 12, n_mine // 2), r2.uniform(45, 47, n_mine // 2)], 1)
coords = torch.tensor(mine, device=dev)


def rel(a, b):
    return ((a - b).norm(dim=1) / b.norm(dim=1)).max().item()


def model(**kw):
    with contextlib.redirect_stdout(sys.stderr):
        return LocationEncoder(Namespace(location_model_name="RANGE+", pretrained_path=enc, device=dev, range_db=db, beta=0.5, **kw))


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    return (time.perf_counter() - t0) / reps


full = model()
ref = full.embed(coords)                                            # unsharded database, my queries
t_full = timed(lambda: full.embed(coords))
report = {}
for merge in ("peer", "reduce_scatter"):
    sh = model(db_shard=(rank, world), db_group=None, db_merge=merge)
    sh.sharded.MAX_STEP_ROWS = int(os.environ.get("STEP_ROWS", 16_384 * world))     # several steps per call
    out = sh.embed(coords)
    err = torch.tensor([rel(out[:, :1024], ref[:, :1024]), (out[:, 1024:] - ref[:, 1024:]).abs().max().item()], device=dev)
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    again = sh.embed(coords)
    same = torch.tensor([float(torch.equal(out, again))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    t = timed(lambda: sh.embed(coords))
    host = sh(coords.cpu()) if merge == "peer" else None            # the public call, collective
    if host is not None:
        assert host.shape == (n_mine, 1280) and np.allclose(host[:, :1024], out[:, :1024].cpu().numpy(), rtol=1e-6, atol=1e-7)
    report[merge] = (sh.sharded.merge, float(err[0]), float(err[1]), bool(same[0] > 0), t)
    sh.sharded.close()
    del sh
lo, hi = shard_rows(N, rank, world)                                # query sharding: my slab of a common set, then gather
common = torch.tensor(S.area_uniform(N, np.random.default_rng(1)), device=dev)
part = full.embed(common[lo:hi].contiguous())
parts = [torch.empty(shard_rows(N, r, world)[1] - shard_rows(N, r, world)[0], 1280, device=dev) for r in range(world)]
dist.all_gather(parts, part)
e2 = rel(torch.cat(parts)[:, :1024], full.embed(common)[:, :1024])
if rank == 0:
    for merge, (used, e_o, e_q, same, t) in report.items():
        print(f"world {world}: M-sharded [{merge} -> ran {used}] vs unsharded: max rel-row err {e_o:.2e}, location columns {e_q:.1e}, "
              f"repeatable {same}; {t*1e3:.1f} ms per call vs {t_full*1e3:.1f} ms with the database replicated "
              f"(N={N}+1500 r per rank, M={M})")
        assert e_o < 1e-3 and e_q == 0.0 and same
    print(f"query-sharded + all_gather vs one GPU {e2:.2e}")
    assert e2 < 1e-3
dist.destroy_process_group()
