"""model(locs) latency at BASELINE config 1 (RANGE, M = 50 000, 10 000 page-locked queries per call) for a few piece
plans (chunk / tail of LocationEncoder): min / mean ms per call and the device-resident embed() time beside it.
    python tools/config1_latency.py
Measured (B200): 2.8-2.9 ms for every plan - the call is bound by the 102 MB device-to-host copy (1.9 ms) plus the first
piece's computation; the first configuration's mean includes the one-time page-locked allocations."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench, contextlib
from argparse import Namespace
from range_b200.range import LocationEncoder
db, weights, c = bench.synthetic_inputs(M=50000, n=10000)
enc = dict(L=40, dims=[1600, 512, 512, 256], weights=weights)
h = torch.tensor(c).pin_memory(); d = torch.tensor(c, device='cuda:0')
ddb=None
for kw in [dict(), dict(chunk=2048, tail=2048), dict(chunk=4096, tail=2048), dict(chunk=3072, tail=3072), dict(chunk=4096, tail=4096), dict(chunk=6144, tail=2048)]:
    with contextlib.redirect_stdout(sys.stderr):
        m = LocationEncoder(Namespace(location_model_name="RANGE", pretrained_path=enc, device='cuda:0', range_db=db if ddb is None else ddb, **kw)); ddb = m.engine.db
    for _ in range(3): m(h)
    ts=[]
    for _ in range(10):
        torch.cuda.synchronize(); t=time.perf_counter(); r=m(h); ts.append(time.perf_counter()-t)
    te=[]
    for _ in range(5):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record(); m.embed(d, out_dtype=torch.float64); b.record(); torch.cuda.synchronize(); te.append(a.elapsed_time(b))
    print(kw, [hi-lo for lo,hi in m._chunks(10000, min(m.chunk,10000), m.tail, m.taper)], f"model(h) min {min(ts)*1e3:.2f} mean {sum(ts)/10*1e3:.2f} ms; embed device {min(te):.2f} ms")
