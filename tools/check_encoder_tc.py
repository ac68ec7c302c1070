import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import range_oracle as O
from range_b200.engine import RangeEngine
from range_b200.sh_table import load_entries
dev = "cuda:0"
ws = O.siren_init(40, 512, 2, 256, seed=0)
enc = dict(L=40, dims=[1600, 512, 512, 256], weights=ws)
e64 = RangeEngine(dev, encoder=enc, encoder_precision="fp64")
etc = RangeEngine(dev, encoder=enc, encoder_precision="f16x3")
print("precisions:", e64.precision, etc.precision)
c = O.area_uniform(5000, np.random.default_rng(7))
c[:4] = [[0, 90], [0, -90], [180, 0], [-180, 0]]
ct = torch.tensor(c)
qa, qa16, xa = e64.encode(ct)
qb, qb16, xb = etc.encode(ct)
d = (qa - qb).abs().max(1).values.cpu().numpy()
print("f16x3 vs fp64 (GPU): max abs", d.max(), " mean", d.mean(), " nan", int(torch.isnan(qb).sum()))
ref = O.RangeOracle.__new__(O.RangeOracle); ref.L, ref.entries, ref.weights = 40, load_entries(40), ws
qr = ref.encode(ct).numpy()
lat = np.abs(c[:, 1])
for nm, q in (("fp64", qa), ("f16x3", qb)):
    dd = np.abs(q.cpu().numpy() - qr).max(1)
    print(f"{nm} vs oracle: |lat|<60 {dd[lat < 60].max():.3e}  >=60 {dd[lat >= 60].max():.3e}")
# ragged N
for N in (1, 127, 129, 40000):
    cc = torch.tensor(O.area_uniform(N, np.random.default_rng(N)))
    a = e64.encode(cc)[0]; b = etc.encode(cc)[0]
    print(N, "max abs", (a - b).abs().max().item())
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
big = torch.tensor(O.area_uniform(100000, np.random.default_rng(1)), device=dev)
print("encode 100k: fp64 %.2f ms   f16x3 %.2f ms" % (timeit(lambda: e64.encode(big)), timeit(lambda: etc.encode(big))))
