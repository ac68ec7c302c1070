"""Where does model(locs) lose time when several ranks share a host?  Run plain (1 process) or under torchrun.
Per call: wall time until the last launch is enqueued (host side), until the device has finished, until the copies have
landed; GPU time per piece from CUDA events.  NO_DIST=1 skips process-group creation (independent processes)."""
import os, sys, time
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
import contextlib
with contextlib.redirect_stdout(sys.stderr):
    model = LocationEncoder(Namespace(location_model_name="RANGE+", pretrained_path=enc, device=dev, range_db=db, beta=0.5))
h = torch.tensor(coords).pin_memory()
for _ in range(3): model(h)
torch.cuda.synchronize()
if use_dist: dist.barrier()
ts = []
for _ in range(8):
    t = time.perf_counter(); r = model(h); ts.append(time.perf_counter() - t)
d = torch.tensor(coords, device=dev)
torch.cuda.synchronize()
te = []
for _ in range(4):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); model.embed(d, out_dtype=torch.float32); b.record(); torch.cuda.synchronize(); te.append(a.elapsed_time(b))
# host-side enqueue time alone: same call with the copy stream idle (device work still queued)
big = torch.empty(100000, 1280, dtype=torch.float64, pin_memory=True)
dd = torch.empty(100000, 1280, dtype=torch.float64, device=dev)
tc = []
for _ in range(4):
    torch.cuda.synchronize(); t = time.perf_counter(); big.copy_(dd, non_blocking=True); torch.cuda.synchronize(); tc.append(time.perf_counter() - t)
print(f"rank {rank}/{world} dist={use_dist} OMP={os.environ.get('OMP_NUM_THREADS')} cpus={os.cpu_count()} affinity={len(os.sched_getaffinity(0))}: "
      f"model(h) {min(ts)*1e3:.1f} ms (runs {[round(x*1e3,1) for x in ts]}); embed device {min(te):.1f} ms; "
      f"D2H 1 GB {min(tc)*1e3:.1f} ms", flush=True)
if use_dist:
    dist.barrier(); dist.destroy_process_group()
