"""timeline of the 'stream' path of model(locs): when the device work ends vs when the copies end"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from argparse import Namespace
from range_b200.range import LocationEncoder
db, weights, coords = bench.synthetic_inputs()
enc = dict(L=40, dims=[1600, 512, 512, 256], weights=weights)
h = torch.tensor(coords).pin_memory()
m = LocationEncoder(Namespace(location_model_name="RANGE+", pretrained_path=enc, device="cuda:0", range_db=db, beta=0.5, host_path="stream"))
for _ in range(3): m(h)
eng = m.engine
dc = torch.tensor(coords, device="cuda:0")
# (1) the same device work without any copy: sort pieces + encode + one retrieve_concat, with / without progress counters
def device_only(progress):
    cuts = m._stream_pieces(100000, 6144, 16, 4)
    parts = [eng.sort_queries(dc[lo:hi]) for lo, hi in cuts]
    sub = torch.cat([p[0] for p in parts]); perm = torch.cat([p[1] + lo for (lo, _), p in zip(cuts, parts)])
    q64, q16, qxyz = eng.encode(sub)
    if progress: m._progress.zero_(); eng.set_progress(m._progress)
    m._retrieve_concat(q16, qxyz, q64, m._dev_result[: 100000 * 1280 * 8].view(torch.float64).view(100000, 1280), torch.float64, perm)
    eng.set_progress(None)
for progress in (False, True):
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); t = time.perf_counter(); device_only(progress); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    print(f"device work only (sorted in stream pieces), progress counters {progress}: min {min(ts)*1e3:.1f} ms")
# (2) the full call, wall clock split: enqueue / device drained / copies done
import types
orig = m._forward_stream
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = m(h); t1 = time.perf_counter()
    print(f"model(h) stream: {(t1-t0)*1e3:.1f} ms")
# (3) D2H of 1 GB while an apply kernel runs: does the concurrent copy slow the kernel?
big = torch.empty(100000, 1280, dtype=torch.float64, pin_memory=True)
dd = m._dev_result[: 100000 * 1280 * 8].view(torch.float64).view(100000, 1280)
cs = torch.cuda.Stream()
for with_copy in (False, True):
    ts = []
    for _ in range(4):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); device_only(False); b.record()
        if with_copy:
            with torch.cuda.stream(cs):
                big.copy_(dd, non_blocking=True)
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(f"device work (CUDA events) with a concurrent 1 GB D2H copy {with_copy}: min {min(ts):.1f} ms")
