"""where does model(locs) spend its time? (host-pinned in -> numpy float64 out), per host path and piece plan of
range.py:_forward_host"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from argparse import Namespace
from range_b200.range import LocationEncoder
db, weights, coords = bench.synthetic_inputs()
enc = dict(L=40, dims=[1600, 512, 512, 256], weights=weights)
h = torch.tensor(coords).pin_memory()
dc = torch.tensor(coords, device="cuda:0")


def best(fn, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    return min(ts), ts


model = None
# host path x (chunk, tail, taper) of the pipelined pieces
settings = [("copy", 24576, 6144, 0.5), ("copy", 24576, 2048, 0.5), ("copy", 24576, 3072, 0.5), ("copy", 36864, 3072, 0.5),
            ("copy", 24576, 3072, 0.6), ("copy", 18432, 3072, 0.5), ("copy", 24576, 6144, 0.5), ("packed", 24576, 6144, 0.5)]
if os.environ.get("SETTINGS"):       # "path:chunk:tail:taper,..."
    settings = [(a, int(b), int(c), float(d)) for a, b, c, d in (x.split(":") for x in os.environ["SETTINGS"].split(","))]
for path, chunk, tail, taper in settings:
    del model
    model = LocationEncoder(Namespace(location_model_name="RANGE+", pretrained_path=enc, device="cuda:0", range_db=db, beta=0.5,
                                      host_path=path, chunk=chunk, tail=tail, taper=taper))
    for _ in range(3): model(h)
    t, ts = best(lambda: model(h), reps=10)
    pieces = [hi - lo for lo, hi in model._chunks(len(coords), chunk, tail, taper)]
    print(f"model(h) host_path={path} chunk={chunk} tail={tail} taper={taper}: {t*1e3:.1f} ms = {len(coords)/t/1e6:.2f} M q/s "
          f"(mean {sum(ts)/len(ts)*1e3:.1f}; runs {[round(x*1e3,1) for x in ts]}; pieces {pieces}; host threads {model.host_threads})")
if os.environ.get("TRACE", "1") == "1":       # timeline of one call with the last setting's model: when each piece is computed / copied
    model = LocationEncoder(Namespace(location_model_name="RANGE+", pretrained_path=enc, device="cuda:0", range_db=db, beta=0.5))
    for _ in range(3): model(h)
    for rep in range(2):
        model.trace = []
        torch.cuda.synchronize(); w = time.perf_counter(); model(h); wall = time.perf_counter() - w
        t0 = model.trace[0][1]
        print(f"timeline (wall {wall*1e3:.1f} ms): " + " | ".join(
            f"{rows}: computed {t0.elapsed_time(a):.1f} copied {t0.elapsed_time(b):.1f}" for rows, a, b in model.trace[1:]))
    model.trace = None
t, ts = best(lambda: model.embed(dc, out_dtype=torch.float64))
print(f"embed device fp64: {t*1e3:.1f} ms")
t, ts = best(lambda: model.embed(dc, out_dtype=torch.float32))
print(f"embed device fp32: {t*1e3:.1f} ms")
d = model.embed(dc, out_dtype=torch.float64)
hp = torch.empty((100000, 1280), dtype=torch.float64, pin_memory=True)
t, _ = best(lambda: hp.copy_(d, non_blocking=True))
print(f"D2H 1.02 GB pinned: {t*1e3:.1f} ms -> {1.024/t:.1f} GB/s")
pk = torch.empty((100000, 6144), dtype=torch.uint8, pin_memory=True)
res = np.empty((100000, 1280), np.float64)
for th in (1, 4, 8, 16):
    t, _ = best(lambda: model.engine.lib.range_host_unpack(pk.data_ptr(), 100000, res.ctypes.data, th), reps=3)
    print(f"host unpack 100k rows, {th} threads: {t*1e3:.1f} ms = {0.1/t:.2f} M rows/s")
