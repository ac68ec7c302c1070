"""where does model(locs) spend its time? (host-pinned in -> numpy out)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from argparse import Namespace
from range_b200.range import LocationEncoder
db, weights, coords = bench.synthetic_inputs()
enc = dict(L=40, dims=[1600, 512, 512, 256], weights=weights)
model = LocationEncoder(Namespace(location_model_name="RANGE+", pretrained_path=enc, device="cuda:0", range_db=db, beta=0.5))
h = torch.tensor(coords).pin_memory()
for _ in range(2): model(h)
torch.cuda.synchronize()
for name, fn in [("model(h) full", lambda: model(h)),
                 ("pinned alloc 1GB", lambda: torch.empty((100000, 1280), dtype=torch.float64, pin_memory=True)),
                 ("embed device fp64", lambda: model.embed(torch.tensor(coords, device="cuda:0"), out_dtype=torch.float64))]:
    ts = []
    for _ in range(4):
        torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    print(f"{name}: {min(ts)*1e3:.1f} ms (runs {[round(x*1e3,1) for x in ts]})")
d = model.embed(torch.tensor(coords, device="cuda:0"), out_dtype=torch.float64)
hp = torch.empty((100000, 1280), dtype=torch.float64, pin_memory=True)
ts = []
for _ in range(4):
    torch.cuda.synchronize(); t = time.perf_counter(); hp.copy_(d, non_blocking=True); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
print(f"D2H 1.02 GB pinned: {min(ts)*1e3:.1f} ms -> {1.024/min(ts):.1f} GB/s")
