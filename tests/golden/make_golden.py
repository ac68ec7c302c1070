"""Generate tests/golden/range_golden.npz by running the UNMODIFIED reference (mvrl/RANGE) on the CPU.

Runs only in the build container (needs /root/reference and sympy); the output is committed.

What it does (steps 1-3 live in oracle/ref_harness.py, shared with bench.py's reference arm):
  1. regenerates `spherical_harmonics_ylm.py` with the reference's own generator into oracle/_ref/
     (the file is stripped from the mount; .MISSING_LARGE_BLOBS) and pre-registers it under the module
     name the reference imports;
  2. installs permissive stub modules for third-party packages the reference imports at module scope but
     never uses on this path (lightning, timm, torchgeo, rasterio, matplotlib, albumentations, ...);
  3. fabricates a random-init SatCLIP-L40 checkpoint through the reference's own
     `SatCLIPLightningModule` and a synthetic `.npz` database;
  4. calls `range.load_model.load_model(...)` / `model(locs)` for RANGE and RANGE+ (beta grid) and stores
     inputs + outputs;
  5. runs oracle/range_oracle.py on the same inputs and asserts agreement.

    python tests/golden/make_golden.py            # writes tests/golden/range_golden.npz
    python tests/golden/make_golden.py --big      # additionally cross-checks N=10000, M=20000, H=512
"""
import hashlib
import os
import sys
import tempfile
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
BETAS = (0.0, 0.25, 0.5, 0.75, 1.0)


from oracle.ref_harness import fabricate_ckpt, import_reference      # noqa: E402  (the unmodified reference)


def special_points():
    pts = [(0.0, 0.0), (0.0, 90.0), (0.0, -90.0), (180.0, 0.0), (-180.0, 0.0), (179.999999, 45.0),
           (-179.999999, -45.0), (12.5, 89.9999), (-77.0, -89.9999), (90.0, 60.0), (-90.0, -60.0),
           (45.0, 85.0), (-135.0, -85.0), (0.0, 1e-9), (1e-9, 0.0), (123.456789012345, -33.3333333333333)]
    return np.asarray(pts, np.float64)


def checksum(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def run_reference(load_model, ckpt, dbfile, coords, name, beta):
    model = load_model(name, ckpt, device="cpu", db_path=dbfile, beta=beta)
    with torch.no_grad():
        out = model(torch.tensor(coords))
    assert isinstance(out, np.ndarray) and out.dtype == np.float64 and out.shape == (len(coords), 1280)
    return out, model


def main():
    from oracle import range_oracle as O
    from range_b200.sh_table import load_entries
    load_model, SatCLIPLightningModule = import_reference()
    entries = load_entries(40)
    tmp = tempfile.mkdtemp()

    # ---- committed golden: H = 64 (keeps the weights small), M = 4096, N = 64 ---------------------
    H, M, N = 64, 4096, 64
    ckpt = os.path.join(tmp, "satclip_h64.ckpt")
    weights = fabricate_ckpt(SatCLIPLightningModule, ckpt, H)
    db = O.synthetic_db(M, seed=0, kind="iid")
    db = {k: v.astype(np.float32).astype(np.float64) for k, v in db.items()}      # fp32-exact values
    dbfile = os.path.join(tmp, "db.npz")
    np.savez(dbfile, **db)
    rng = np.random.default_rng(1)
    coords = np.concatenate([special_points(), db["locs"][:4], O.area_uniform(N - 20, rng)])
    assert coords.shape == (N, 2)

    gold = dict(coords=coords, H=H, M=M, db_seed=0,
                db_checksum=checksum(db["locs"], db["satclip_embeddings"], db["image_embeddings"]),
                betas=np.asarray(BETAS))
    for i, (w, b) in enumerate(weights):
        gold[f"W{i}"], gold[f"b{i}"] = w.numpy(), b.numpy()
    out, model = run_reference(load_model, ckpt, dbfile, coords, "RANGE", None)
    gold["q"] = out[:, 1024:]
    gold["O_range"] = out[:, :1024].astype(np.float32)
    assert np.array_equal(gold["O_range"].astype(np.float64), out[:, :1024])
    with torch.no_grad():
        gold["Y"] = model.loc_model.posenc(torch.tensor(coords)).numpy()
    for beta in BETAS:
        outp, _ = run_reference(load_model, ckpt, dbfile, coords, "RANGE+", beta)
        assert np.array_equal(outp[:, 1024:], gold["q"])
        gold[f"O_plus_{beta}"] = outp[:, :1024].astype(np.float32)

    # ---- oracle vs reference on the same inputs ---------------------------------------------------
    Y = O.sh_analytic(coords, 40, entries).numpy()
    print("SH  oracle vs reference: max abs", np.abs(Y - gold["Y"]).max(),
          " max |Y|", np.abs(gold["Y"]).max())
    orc = O.RangeOracle("RANGE", weights, entries, db)
    o = orc(coords)
    print("RANGE  q max abs", np.abs(o[:, 1024:] - gold["q"]).max(),
          " O max abs", np.abs(o[:, :1024] - gold["O_range"]).max())
    for beta in BETAS:
        o = O.RangeOracle("RANGE+", weights, entries, db, beta=beta)(coords)
        print(f"RANGE+ beta={beta}  O max abs", np.abs(o[:, :1024] - gold[f'O_plus_{beta}']).max())
    path = os.path.join(HERE, "range_golden.npz")
    np.savez_compressed(path, **gold)
    print("wrote", path, os.path.getsize(path), "bytes")

    # ---- harmonics_calculation='closed-form' (spherical_harmonics_closed_form.py): separate small fixture ----------
    from range.location_models.satclip.positional_encoding.spherical_harmonics import SphericalHarmonics
    from range.location_models.satclip.location_encoder import get_neural_network
    with torch.no_grad():
        posenc = SphericalHarmonics(legendre_polys=40, harmonics_calculation="closed-form").double()
        Ycf = posenc(torch.tensor(coords)).numpy()
        nnet = get_neural_network("siren", input_dim=1600, num_classes=256, dim_hidden=H, num_layers=2).double()
        sd = {"layers.0.weight": weights[0][0], "layers.0.bias": weights[0][1], "layers.1.weight": weights[1][0],
              "layers.1.bias": weights[1][1], "last_layer.weight": weights[2][0], "last_layer.bias": weights[2][1]}
        nnet.load_state_dict(sd)
        nnet.eval()                                      # dropout in the hidden Siren layers is identity in eval mode
        ecf = nnet(torch.tensor(Ycf))
        qcf = (ecf / ecf.norm(p=2, dim=-1, keepdim=True)).numpy()
    Yo = O.sh_closed_form(coords, 40).numpy()
    qo = O.RangeOracle("RANGE", weights, entries, db, harmonics="closed-form").encode(torch.tensor(coords)).numpy()
    print("closed-form SH oracle vs reference: max abs", np.abs(Yo - Ycf).max(), " q max abs", np.abs(qo - qcf).max(),
          " closed-form vs analytic max abs", np.abs(Ycf - gold["Y"]).max())
    cf_path = os.path.join(HERE, "closed_form_golden.npz")
    np.savez_compressed(cf_path, coords=coords, Y=Ycf, q=qcf)
    print("wrote", cf_path, os.path.getsize(cf_path), "bytes")

    if "--big" in sys.argv:
        H, M, N = 512, 20000, 10000
        ckpt = os.path.join(tmp, "satclip_h512.ckpt")
        weights = fabricate_ckpt(SatCLIPLightningModule, ckpt, H)
        db = O.synthetic_db(M, seed=0, kind="iid")
        np.savez(dbfile, **db)
        coords = O.area_uniform(N, np.random.default_rng(1))
        for name in ("RANGE", "RANGE+"):
            t = time.time()
            ref, _ = run_reference(load_model, ckpt, dbfile, coords, name, 0.5)
            t_ref = time.time() - t
            t = time.time()
            o = O.RangeOracle(name, weights, entries, db, beta=0.5)(coords)
            t_or = time.time() - t
            rel = np.linalg.norm(o[:, :1024] - ref[:, :1024], axis=1) / np.linalg.norm(ref[:, :1024], axis=1)
            print(f"BIG {name}: q max abs {np.abs(o[:, 1024:] - ref[:, 1024:]).max():.3e}  "
                  f"O max abs {np.abs(o[:, :1024] - ref[:, :1024]).max():.3e}  O rel-row max {rel.max():.3e}  "
                  f"reference {t_ref:.2f}s  oracle {t_or:.2f}s")


if __name__ == "__main__":
    main()
