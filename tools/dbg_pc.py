import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import range_oracle as O
from range_b200.engine import RangeEngine
from range_b200.database import DeviceDatabase
from range_b200.sh_table import load_entries
DEV="cuda:0"
N, M = 12_300 + 37, 3000 + 5
db = O.synthetic_db(M, seed=6, kind="iid")
ws = O.siren_init(40, 64, 2, 256, seed=3)
eng = RangeEngine(DEV, encoder=dict(L=40, dims=[1600, 64, 64, 256], weights=ws), database=DeviceDatabase(db, DEV))
c = O.area_uniform(N, np.random.default_rng(21))
cs, perm = eng.sort_queries(torch.tensor(c))
q64, q16, qxyz = eng.encode(cs)
big = eng.retrieve("RANGE+", q16, qxyz, 12.0, 40.0, 0.5)
for n0, n1 in [(0, 1000), (0, 128), (5000, 6000), (12000, N)]:
    small = eng.retrieve("RANGE+", q16[n0:n1].contiguous(), qxyz[n0:n1].contiguous(), 12.0, 40.0, 0.5)
    d = ((big[n0:n1] - small).norm(dim=1) / small.norm(dim=1))
    print(n0, n1, "max rel", d.max().item(), "argmax", d.argmax().item(), "mean", d.mean().item())
sb, mb = eng.retrieve_stats("RANGE+", q16, qxyz, 12.0, 40.0)
ss, ms = eng.retrieve_stats("RANGE+", q16[:1000].contiguous(), qxyz[:1000].contiguous(), 12.0, 40.0)
print("stats diff", ((sb[:1000]-ss).abs()/ss).max().item(), (mb[:1000]-ms).abs().max().item())
