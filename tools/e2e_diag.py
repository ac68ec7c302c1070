"""model(locs) when several ranks share a host: aggregate queries/s per host path.  Run plain (1 process) or under torchrun.
SETTINGS="copy,packed" (host_path); NO_DIST=1 skips process-group creation."""
import contextlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from argparse import Namespace
from range_b200.range import LocationEncoder
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
use_dist = world > 1 and os.environ.get("NO_DIST", "0") != "1"
if use_dist:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
db, weights, coords = bench.synthetic_inputs(rank)
enc = dict(L=40, dims=[1600, 512, 512, 256], weights=weights)
h = torch.tensor(coords).pin_memory()
K = int(os.environ.get("CALLS", 8))


def barrier():
    torch.cuda.synchronize()
    if use_dist:
        dist.barrier()


ddb = None
for setting in os.environ.get("SETTINGS", "copy,packed").split(","):
    path, _, share = setting.partition(":")
    with contextlib.redirect_stdout(sys.stderr):
        ns = Namespace(location_model_name="RANGE+", pretrained_path=enc, device=dev, range_db=db if ddb is None else ddb, beta=0.5,
                       host_path=path)
        if share:
            ns.packed_share = float(share)
        model = LocationEncoder(ns)
        ddb = model.engine.db                       # the prepared device layout is reused by the next setting
    for _ in range(3):
        r = model(h)
    barrier()
    ts = []
    t0 = time.perf_counter()
    for _ in range(K):
        t = time.perf_counter(); r = model(h); ts.append(time.perf_counter() - t)
    barrier()
    total = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if use_dist:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{setting:12s} world {world}: {world * len(coords) * K / float(total[0]) / 1e6:.2f} M queries/s aggregate "
              f"(rank 0 calls: min {min(ts)*1e3:.1f} mean {sum(ts)/K*1e3:.1f} max {max(ts)*1e3:.1f} ms; host threads {model.host_threads})",
              flush=True)
    del model, r
if use_dist:
    dist.barrier(); dist.destroy_process_group()
