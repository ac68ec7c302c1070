"""The location encoder alone (K1 harmonics + K1b SIREN layers + K3 normalisation) on N area-uniform queries.
    python tools/time_encoder.py [N]         -> ms per call (CUDA events, 20 calls after 3 warm-ups)
    ONCE=1 python tools/time_encoder.py      -> a single call after one warm-up (what an `ncu -k regex:sh_|siren_` capture wants)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from range_b200 import synthetic                                # noqa: E402
from range_b200.engine import RangeEngine                      # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
dev = torch.device("cuda", 0)
enc = dict(L=40, dims=[1600, 512, 512, 256], weights=synthetic.siren_init(seed=7))
eng = RangeEngine(dev, encoder=enc)
lonlat = torch.tensor(synthetic.area_uniform(N, np.random.default_rng(5)), device=dev)
out = eng.encode(lonlat)
torch.cuda.synchronize()
if os.environ.get("ONCE"):
    eng.encode(lonlat, *out)
    torch.cuda.synchronize()
    sys.exit(0)
ts = []
for _ in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.encode(lonlat, *out); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"encode N={N}: min {min(ts):.3f} ms, median {sorted(ts)[10]:.3f} ms")
