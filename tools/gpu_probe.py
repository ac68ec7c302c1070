"""One-shot GPU diagnostic: every kernel against the CPU oracle, then first timings.  Prints, never asserts,
so a single `gpurun` call shows everything.   python tools/gpu_probe.py [--skip-timing]"""
import os, sys, time, traceback
import numpy as np, torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import range_oracle as O
from range_b200.sh_table import load_entries
from range_b200.engine import RangeEngine
from range_b200.database import DeviceDatabase

dev = "cuda:0"
entries = load_entries(40)
g = np.load(os.path.join(ROOT, "tests/golden/range_golden.npz"))
weights = [(torch.tensor(g[f"W{i}"]), torch.tensor(g[f"b{i}"])) for i in range(3)]
enc = dict(L=40, dims=[1600, 64, 64, 256], weights=weights)


def section(name):
    print(f"\n===== {name} =====", flush=True)


def rel_rows(a, b):
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-30)


def run(fn, name):
    section(name)
    try:
        fn()
    except Exception:
        traceback.print_exc()
    torch.cuda.synchronize()


def t_sh():
    eng = RangeEngine(dev, L=40)
    coords = np.concatenate([g["coords"], O.area_uniform(4000, np.random.default_rng(5))])
    Y = eng.sh_features(torch.tensor(coords)).cpu().numpy()
    Yr = O.sh_analytic(coords, 40, entries).numpy()
    d = np.abs(Y - Yr)
    lat = np.abs(coords[:, 1])
    for lo, hi in [(0, 60), (60, 85), (85, 90.01)]:
        m = (lat >= lo) & (lat < hi)
        print(f"|lat| in [{lo},{hi}): n={m.sum()} max abs dev {d[m].max():.3e}  (l<20: {d[m][:, :400].max():.3e})")
    print("golden Y max abs dev", np.abs(Y[:64] - g["Y"]).max())


def t_encode():
    eng = RangeEngine(dev, encoder=enc)
    q64, q16, qxyz = eng.encode(torch.tensor(g["coords"]))
    q = q64.cpu().numpy()
    print("q vs golden: max abs", np.abs(q - g["q"]).max(), " min cos", (q * g["q"]).sum(1).min())
    print("q16 vs q64 max abs", (q16.double() - q64).abs().max().item())
    xyz = O.rad_to_cart(g["coords"] * np.pi / 180).astype(np.float32)
    print("qxyz max abs", np.abs(qxyz.cpu().numpy()[:, :3] - xyz).max())
    # bigger, H=512 random init
    ws = O.siren_init(40, 512, 2, 256, seed=0)
    eng2 = RangeEngine(dev, encoder=dict(L=40, dims=[1600, 512, 512, 256], weights=ws))
    c = O.area_uniform(3000, np.random.default_rng(7))
    q64b, _, _ = eng2.encode(torch.tensor(c))
    ref = O.RangeOracle.__new__(O.RangeOracle)
    ref.L, ref.entries, ref.weights = 40, entries, ws
    qr = ref.encode(torch.tensor(c)).numpy()
    d = np.abs(q64b.cpu().numpy() - qr).max(1)
    lat = np.abs(c[:, 1])
    print(f"H=512 N=3000: max abs |lat|<60: {d[lat < 60].max():.3e}   >=60: {d[lat >= 60].max():.3e}")


def make_db(M, seed=0):
    db = O.synthetic_db(M, seed=seed, kind="iid")
    return {k: v.astype(np.float32).astype(np.float64) for k, v in db.items()}


def t_retrieve_golden():
    db = make_db(int(g["M"]))
    eng = RangeEngine(dev, encoder=enc, database=DeviceDatabase(db, dev))
    q64, q16, qxyz = eng.encode(torch.tensor(g["coords"]))
    Ot = eng.retrieve("RANGE", q16, qxyz, 15.0, 0.0, None).cpu().numpy()
    r = rel_rows(Ot, g["O_range"])
    print(f"RANGE  vs golden: max abs {np.abs(Ot - g['O_range']).max():.3e} rel-row max {r.max():.3e} mean {r.mean():.3e}")
    for beta in g["betas"]:
        Ot = eng.retrieve("RANGE+", q16, qxyz, 12.0, 40.0, float(beta)).cpu().numpy()
        ref = g[f"O_plus_{beta}"]
        r = rel_rows(Ot, ref)
        print(f"RANGE+ beta={beta}: max abs {np.abs(Ot - ref).max():.3e} rel-row max {r.max():.3e} mean {r.mean():.3e}"
              f" nan={np.isnan(Ot).sum()}")


def t_retrieve_ragged():
    for (N, M) in [(300, 5000), (1000, 20001), (129, 128), (5, 77)]:
        db = make_db(M, seed=3)
        ws = O.siren_init(40, 64, 2, 256, seed=1)
        e = dict(L=40, dims=[1600, 64, 64, 256], weights=ws)
        eng = RangeEngine(dev, encoder=e, database=DeviceDatabase(db, dev))
        c = O.area_uniform(N, np.random.default_rng(11))
        c[: min(N, 3)] = db["locs"][: min(N, 3)]
        q64, q16, qxyz = eng.encode(torch.tensor(c))
        for name, beta in [("RANGE", None), ("RANGE+", 0.5)]:
            orc = O.RangeOracle(name, ws, entries, db, beta=beta, exact=True)
            ref = orc(c)
            Ot = eng.retrieve(name, q16, qxyz, orc.temp, 40.0, beta).cpu().numpy()
            r = rel_rows(Ot, ref[:, :1024])
            print(f"N={N} M={M} {name}: rel-row max {r.max():.3e} mean {r.mean():.3e} nan={np.isnan(Ot).sum()}"
                  f" q max abs {np.abs(q64.cpu().numpy() - ref[:, 1024:]).max():.2e}")


def t_timing():
    M = 100_000
    rng = np.random.default_rng(0)
    db = dict(locs=O.area_uniform(M, rng), satclip_embeddings=rng.standard_normal((M, 256), dtype=np.float32),
              image_embeddings=rng.standard_normal((M, 1024), dtype=np.float32))
    ws = O.siren_init(40, 512, 2, 256, seed=0)
    eng = RangeEngine(dev, encoder=dict(L=40, dims=[1600, 512, 512, 256], weights=ws), database=DeviceDatabase(db, dev))
    for N in (18944, 100_000):
        c = torch.tensor(O.area_uniform(N, np.random.default_rng(1)), device=dev)
        q64, q16, qxyz = eng.encode(c)

        def timeit(fn, reps=3):
            fn(); torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            return min(ts)
        t_enc = timeit(lambda: eng.encode(c))
        t_st = timeit(lambda: eng.retrieve_stats("RANGE+", q16, qxyz, 12.0, 40.0))
        sums, maxs = eng.retrieve_stats("RANGE+", q16, qxyz, 12.0, 40.0)
        t_ap = timeit(lambda: eng.retrieve_apply("RANGE+", q16, qxyz, 12.0, 40.0, 0.5, sums, maxs))
        t_r = timeit(lambda: eng.retrieve("RANGE", q16, qxyz, 15.0, 0.0, None))
        fl = 2566.0 * N * M
        print(f"N={N} M={M}: encode {t_enc:.2f} ms | RANGE+ stats {t_st:.2f} ms apply {t_ap:.2f} ms -> "
              f"{fl / ((t_st + t_ap) * 1e-3) / 1e12:.1f} TFLOP/s algorithmic ({fl / ((t_st + t_ap) * 1e-3) / 1364.5e12:.3f} of sustained peak)"
              f" | RANGE total {t_r:.2f} ms | q/s e2e-device {N / ((t_enc + t_st + t_ap) * 1e-3):.0f}")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count, "SMs")
    run(t_sh, "K1 spherical harmonics vs oracle")
    run(t_encode, "encoder vs golden / oracle")
    run(t_retrieve_golden, "retrieval vs golden (reference outputs)")
    run(t_retrieve_ragged, "retrieval ragged sizes vs fp64-exact oracle")
    if "--skip-timing" not in sys.argv:
        run(t_timing, "timing")
