"""Does the P.V accumulation lose mass as the database grows?  Values = 1 everywhere, so every output element is
sum_j P_j = 1 in exact arithmetic; the deviation isolates the fp16 rounding of P' (unbiased) from the tensor core's
fp32 accumulation over the database axis."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from range_b200 import synthetic as O          # seeded input generators
from range_b200.utils import rad_to_cart
from range_b200.engine import RangeEngine
from range_b200.database import DeviceDatabase
dev = "cuda:0"
N = 12_288
g = torch.Generator(device="cpu").manual_seed(1)
q = torch.randn(N, 256, generator=g); q = (q / q.norm(dim=1, keepdim=True)).half().to(dev)
c = O.area_uniform(N, np.random.default_rng(1))
for M in [10_000, 100_000, 1_000_000, 3_000_000]:
    d = DeviceDatabase.synthetic(M, dev, seed=3)
    d.Vt.fill_(0); d.Vt[:, :M] = 1.0 * d.vscale                      # V = 1
    eng = RangeEngine(dev, L=40, database=d)
    cs = eng.sort_queries(torch.tensor(c))[0].cpu()
    xyz = torch.zeros(N, 4); xyz[:, :3] = torch.tensor(rad_to_cart(cs.numpy() * np.pi / 180)).float(); xyz = xyz.to(dev)
    for mode, beta, t in (("RANGE", None, 15.0), ("RANGE+", 0.5, 12.0)):
        out = eng.retrieve(mode, q, xyz, t, 40.0, beta).double()
        dev_ = out - 1.0
        print(f"M={M:>8d} {mode:6s}: mean(O - 1) {dev_.mean().item():+.3e}  rms {dev_.pow(2).mean().sqrt().item():.3e}  "
              f"min {dev_.min().item():+.3e} max {dev_.max().item():+.3e}")
    del eng, d
