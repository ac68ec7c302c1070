"""Error of the full path at the bench shape (100 000 x 100 000, H = 512) against the CPU oracle on sampled rows:
reference arithmetic (fp32 retrieval) and the fp64-exact restatement; iid and structured databases."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from argparse import Namespace
from oracle import range_oracle as O
from range_b200.range import LocationEncoder
from range_b200.sh_table import load_entries
dev = "cuda:0"
N = M = 100_000
entries = load_entries(40)
ws = O.siren_init(40, 512, 2, 256, seed=0)
enc = dict(L=40, dims=[1600, 512, 512, 256], weights=ws)
c = O.area_uniform(N, np.random.default_rng(1))
sub = np.linspace(0, N - 1, 256).astype(np.int64)
helper = O.RangeOracle.__new__(O.RangeOracle); helper.L, helper.entries, helper.weights = 40, entries, [(torch.as_tensor(W), torch.as_tensor(b)) for W, b in ws]
for kind in ("iid", "structured"):
    db = O.synthetic_db(M, seed=0, kind=kind, encoder=helper.encode if kind == "structured" else None)
    for name, beta in (("RANGE+", 0.5), ("RANGE", None)):
        m = LocationEncoder(Namespace(location_model_name=name, pretrained_path=enc, device=dev, range_db=db, beta=beta))
        got = m.embed(torch.tensor(c, device=dev), out_dtype=torch.float64)[torch.tensor(sub, device=dev)].cpu().numpy()
        for label, exact in (("fp32 reference arithmetic", False), ("fp64 exact", True)):
            ref = O.RangeOracle(name, ws, entries, db, beta=beta, exact=exact)(c[sub])
            rel = np.linalg.norm(got[:, :1024] - ref[:, :1024], axis=1) / np.linalg.norm(ref[:, :1024], axis=1)
            cos = (got[:, :1024] * ref[:, :1024]).sum(1) / np.linalg.norm(got[:, :1024], axis=1) / np.linalg.norm(ref[:, :1024], axis=1)
            dq = np.abs(got[:, 1024:] - ref[:, 1024:])
            lat = np.abs(c[sub, 1])
            print(f"{kind:10s} {name:6s} vs {label:26s}: O rel-row mean {rel.mean():.2e} max {rel.max():.2e}  min cos {cos.min():.8f}"
                  f"  | q max-abs |lat|<60 {dq[lat < 60].max():.2e}  >=60 {dq[lat >= 60].max():.2e}")
        del m
