"""Repeat the producer/consumer retrieval on one set of inputs and demand bit-identical results every time: a
hand-off race in the P' ring (stale slot, early release) would show up as sporadically different rows."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from range_b200 import synthetic as O          # seeded input generators
from range_b200.utils import rad_to_cart
from range_b200.engine import RangeEngine
from range_b200.database import DeviceDatabase
dev = "cuda:0"
reps = int(os.environ.get("REPS", 30))
for N, M in [(100_000, 100_000), (24_576, 30_011), (6_144, 777), (13_000, 200_000)]:
    eng = RangeEngine(dev, L=40, database=DeviceDatabase.synthetic(M, dev, seed=N))
    g = torch.Generator(device="cpu").manual_seed(1)
    q = torch.randn(N, 256, generator=g); q = (q / q.norm(dim=1, keepdim=True)).half().to(dev)
    c = eng.sort_queries(torch.tensor(O.area_uniform(N, np.random.default_rng(1))))[0].cpu()
    xyz = torch.zeros(N, 4); xyz[:, :3] = torch.tensor(rad_to_cart(c.numpy() * np.pi / 180)).float(); xyz = xyz.to(dev)
    first, bad = None, 0
    for r in range(reps):
        out = eng.retrieve("RANGE+", q, xyz, 12.0, 40.0, 0.5)
        if first is None:
            first = out.clone()
            assert torch.isfinite(first).all()
        elif not torch.equal(out, first):
            bad += 1
            d = (out - first).abs().amax(dim=1)
            print(f"  rep {r}: {int((d > 0).sum())} rows differ, max abs {d.max().item():.3e}")
    print(f"N={N} M={M}: {reps} repetitions, {bad} differing")
    assert bad == 0
print("ok")
