"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the reference's golden vectors.

Tolerances (SURVEY.md section 8c recommends: retrieved columns relative row error <= 1e-3 and cosine >= 0.99999,
location columns max-abs <= 1e-5 for |lat| < 60 deg / <= 1e-3 elsewhere; measured noise floors in DESIGN.md section 2):
  * retrieved columns O (fp16 operands, fp32 accumulate), against the reference's CPU fp32 output and the fp64-exact
    restatement: relative row error <= 1e-3 (TOL_O) whenever the geographic softmax takes part (RANGE+ with beta < 1;
    measured 3.4e-4 max at the bench shape) and <= 2e-4 on the structured database (non-zero-mean values, like real
    SatMAE features; measured 1.5e-4) - also at full size; cosine >= 0.99999 everywhere.
    NAMED EXCEPTION (TOL_O_IID_SEM = 2e-3): the purely semantic softmax (RANGE, or RANGE+ at beta = 1) on the iid
    database - flat weights over zero-mean values, so the fp16 rounding of q and K shows undamped (logit error =
    temperature x 2.5e-5): measured 1.2e-3 max, mean <= 6e-4.  (The reference itself runs these matmuls in TF32 on
    CUDA: 5.9e-3, SURVEY.md Appendix B.)
  * location columns q: max-abs <= 3e-5 for |lat| < 60 deg, <= 1e-3 elsewhere.  Section 8c's 1e-5 is the reference's OWN
    fp64 rounding noise on its 15-digit polynomials there (1e-5 / 4e-4; measured here 1.2e-5 /
    2.9e-4): "equal up to the reference's irreproducible noise", not a looser implementation.  Features with l < 20
    (no cancellation) must agree to 1e-9.
"""
import numpy as np
import pytest
import torch

from oracle import range_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_O, TOL_O_IID_SEM, TOL_O_STRUCTURED = 1e-3, 2e-3, 2e-4
TOL_Q, TOL_Q_POLAR = 3e-5, 1e-3


def tol_o(name, beta):
    """iid database: the named exception applies to the purely semantic softmax only"""
    return TOL_O_IID_SEM if (name == "RANGE" or beta == 1.0) else TOL_O


def rel_rows(a, b):
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-30)


def cos_rows(a, b):
    return (a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))


@pytest.fixture(scope="module")
def gold_setup(golden, sh_entries):
    from range_b200.engine import RangeEngine
    from range_b200.database import DeviceDatabase
    g = golden
    weights = [(torch.tensor(g[f"W{i}"]), torch.tensor(g[f"b{i}"])) for i in range(3)]
    db = O.synthetic_db(int(g["M"]), seed=int(g["db_seed"]), kind="iid")
    db = {k: v.astype(np.float32).astype(np.float64) for k, v in db.items()}
    eng = RangeEngine(DEV, encoder=dict(L=40, dims=[1600, 64, 64, 256], weights=weights),
                      database=DeviceDatabase(db, DEV))
    return g, weights, db, eng


def test_extension_is_loaded_and_counts_launches(gold_setup):
    from range_b200 import _lib
    g, _, _, eng = gold_setup
    before = _lib.launch_count()
    eng.sh_features(torch.tensor(g["coords"]))
    assert _lib.launch_count() == before + 1


def test_sh_features_vs_reference(gold_setup, sh_entries):
    g, _, _, eng = gold_setup
    Y = eng.sh_features(torch.tensor(g["coords"])).cpu().numpy()
    dev = np.abs(Y - g["Y"])
    assert dev[:, :400].max() < 1e-9                   # l < 20: no cancellation, agrees to rounding
    lat = np.abs(g["coords"][:, 1])
    assert dev[lat < 60].max() < 5e-3                  # reference's own noise there: 7.5e-4
    assert dev.max() < 0.2                             # polar: reference's own noise 4e-2
    # stratified random points against the oracle
    pts = O.area_uniform(20000, np.random.default_rng(5))
    Y = eng.sh_features(torch.tensor(pts)).cpu().numpy()
    Yr = O.sh_analytic(pts, 40, sh_entries).numpy()
    lat = np.abs(pts[:, 1])
    d = np.abs(Y - Yr)
    assert d[:, :400].max() < 1e-9
    assert d[lat < 60].max() < 5e-3 and d.max() < 0.2


def test_sh_known_answers(gold_setup):
    _, _, _, eng = gold_setup
    pts = np.array([[10.0, 20.0], [-120.0, -45.0], [0.0, 0.0], [180.0, 90.0]])
    Y = eng.sh_features(torch.tensor(pts)).cpu().numpy()
    phi, theta = np.deg2rad(pts[:, 0] + 180), np.deg2rad(pts[:, 1] + 90)
    assert np.allclose(Y[:, 0], 0.886226925452758, rtol=0, atol=1e-15)
    assert np.allclose(Y[:, 2], 1.53499006191973 * np.cos(theta), rtol=0, atol=2e-15)
    assert np.allclose(Y[:, 3], 0.48860251190292 * np.sin(theta) * np.cos(phi), rtol=0, atol=2e-15)
    assert np.allclose(Y[:, 1], 0.48860251190292 * np.sin(theta) * np.sin(phi), rtol=0, atol=2e-15)


def test_encoder_vs_reference(gold_setup):
    g, _, _, eng = gold_setup
    q64, q16, qxyz = eng.encode(torch.tensor(g["coords"]))
    q = q64.cpu().numpy()
    lat = np.abs(g["coords"][:, 1])
    d = np.abs(q - g["q"]).max(1)
    assert d[lat < 60].max() <= TOL_Q
    assert d.max() <= TOL_Q_POLAR
    assert np.allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-14)
    assert (q16.double() - q64).abs().max().item() < 1e-3
    xyz = O.rad_to_cart(g["coords"] * np.pi / 180).astype(np.float32)
    assert np.abs(qxyz.cpu().numpy()[:, :3] - xyz).max() <= 1.2e-7


def test_encoder_split_vs_fp64(sh_entries):
    """tensor-core SIREN (split fp16 operands, fp32-class) against the fp64 DMMA path and the oracle, H = 512; ragged N"""
    from range_b200.engine import RangeEngine
    ws = O.siren_init(40, 512, 2, 256, seed=0)
    enc = dict(L=40, dims=[1600, 512, 512, 256], weights=ws)
    e64 = RangeEngine(DEV, encoder=enc, encoder_precision="fp64")
    etc = RangeEngine(DEV, encoder=enc, encoder_precision="auto")
    assert e64.precision == "fp64" and etc.precision == "f16x3"
    for N in (1, 3, 129, 255, 257, 3000):       # partial 4-query harmonics batches, partial 256-row CTA-pair tiles
        c = O.area_uniform(N, np.random.default_rng(N))
        c[0] = [0.0, 90.0]
        a, b = e64.encode(torch.tensor(c))[0], etc.encode(torch.tensor(c))[0]
        assert torch.isfinite(b).all()
        assert (a - b).abs().max().item() <= 1e-5            # measured 2.7e-6
    helper = O.RangeOracle.__new__(O.RangeOracle)
    helper.L, helper.entries, helper.weights = 40, sh_entries, ws
    qr = helper.encode(torch.tensor(c)).numpy()
    lat = np.abs(c[:, 1])
    d = np.abs(b.cpu().numpy() - qr).max(1)
    assert d[lat < 60].max() <= TOL_Q and d.max() <= TOL_Q_POLAR
    with pytest.raises(Exception):
        RangeEngine(DEV, encoder=dict(L=40, dims=[1600, 64, 64, 256], weights=O.siren_init(40, 64, 2, 256)),
                    encoder_precision="f16x3")


def test_retrieval_vs_reference_golden(gold_setup):
    g, _, _, eng = gold_setup
    q64, q16, qxyz = eng.encode(torch.tensor(g["coords"]))
    Ot = eng.retrieve("RANGE", q16, qxyz, 15.0, 0.0, None).cpu().numpy()
    assert rel_rows(Ot, g["O_range"]).max() <= TOL_O_IID_SEM and cos_rows(Ot, g["O_range"]).min() >= 0.99999
    for beta in g["betas"]:
        Ot = eng.retrieve("RANGE+", q16, qxyz, 12.0, 40.0, float(beta)).cpu().numpy()
        ref = g[f"O_plus_{beta}"]
        assert np.isfinite(Ot).all()
        assert rel_rows(Ot, ref).max() <= tol_o("RANGE+", float(beta)) and rel_rows(Ot, ref).mean() <= 6e-4, beta
        assert cos_rows(Ot, ref).min() >= 0.99999, beta


@pytest.mark.parametrize("N,M", [(300, 5000), (1000, 20001), (129, 128), (5, 77), (128, 4096)])
def test_retrieval_ragged_vs_exact_oracle(N, M, sh_entries):
    from range_b200.engine import RangeEngine
    from range_b200.database import DeviceDatabase
    db = O.synthetic_db(M, seed=3, kind="iid")
    ws = O.siren_init(40, 64, 2, 256, seed=1)
    eng = RangeEngine(DEV, encoder=dict(L=40, dims=[1600, 64, 64, 256], weights=ws), database=DeviceDatabase(db, DEV))
    c = O.area_uniform(N, np.random.default_rng(11))
    k = min(N, 3)
    c[:k] = db["locs"][:k]                   # queries sitting exactly on database entries
    q64, q16, qxyz = eng.encode(torch.tensor(c))
    for name, beta in [("RANGE", None), ("RANGE+", 0.5), ("RANGE+", 0.0), ("RANGE+", 1.0)]:
        orc = O.RangeOracle(name, ws, sh_entries, db, beta=beta, exact=True)
        ref = orc(c)[:, :1024]
        Ot = eng.retrieve(name, q16, qxyz, orc.temp, 40.0, beta).cpu().numpy()
        assert np.isfinite(Ot).all()
        r = rel_rows(Ot, ref)
        assert r.max() <= tol_o(name, beta) and r.mean() <= 6e-4, (name, beta, r.max(), r.mean())
        assert cos_rows(Ot, ref).min() >= 0.99999, (name, beta)


def test_structured_db_and_properties(sh_entries):
    """peaky softmax (structured DB); beta limits; DB row permutation invariance; shard merge == unsharded"""
    from range_b200.engine import RangeEngine
    from range_b200.database import DeviceDatabase
    ws = O.siren_init(40, 64, 2, 256, seed=2)
    helper = O.RangeOracle.__new__(O.RangeOracle)
    helper.L, helper.entries, helper.weights = 40, sh_entries, ws
    M, N = 6000, 257
    db = O.synthetic_db(M, seed=4, kind="structured", encoder=helper.encode)
    c = O.area_uniform(N, np.random.default_rng(12))
    enc = dict(L=40, dims=[1600, 64, 64, 256], weights=ws)
    eng = RangeEngine(DEV, encoder=enc, database=DeviceDatabase(db, DEV))
    q64, q16, qxyz = eng.encode(torch.tensor(c))
    ref = O.RangeOracle("RANGE+", ws, sh_entries, db, beta=0.5, exact=True)(c)[:, :1024]
    full = eng.retrieve("RANGE+", q16, qxyz, 12.0, 40.0, 0.5).cpu().numpy()
    assert rel_rows(full, ref).max() <= TOL_O_STRUCTURED
    # beta = 1 is the semantic softmax alone at temperature 12, beta = 0 the geographic one
    sem = eng.retrieve("RANGE+", q16, qxyz, 12.0, 40.0, 1.0).cpu().numpy()
    geo = eng.retrieve("RANGE+", q16, qxyz, 12.0, 40.0, 0.0).cpu().numpy()
    assert rel_rows(0.5 * sem + 0.5 * geo, full).max() <= 1e-3
    only_sem = eng.retrieve("RANGE", q16, qxyz, 12.0, 0.0, None).cpu().numpy()
    assert rel_rows(sem, only_sem).max() <= 2e-4
    # permutation of database rows
    perm = np.random.default_rng(0).permutation(M)
    dbp = {k: v[perm] for k, v in db.items()}
    engp = RangeEngine(DEV, encoder=enc, database=DeviceDatabase(dbp, DEV))
    fullp = engp.retrieve("RANGE+", q16, qxyz, 12.0, 40.0, 0.5).cpu().numpy()
    assert rel_rows(fullp, full).max() <= 1e-3
    # M-sharded (3 shards emulated on one GPU): SUM/MAX merge of stats, SUM of partial outputs
    shards = [RangeEngine(DEV, encoder=enc, database=DeviceDatabase(db, DEV, shard=(r, 3))) for r in range(3)]
    stats = [s.retrieve_stats("RANGE+", q16, qxyz, 12.0, 40.0) for s in shards]
    sums = torch.stack([s for s, _ in stats]).sum(0)
    maxs = torch.stack([m for _, m in stats]).max(0).values
    merged = sum(s.retrieve_apply("RANGE+", q16, qxyz, 12.0, 40.0, 0.5, sums, maxs) for s in shards).cpu().numpy()
    assert rel_rows(merged, full).max() <= 2e-4


def test_load_model_api(tmp_path, golden, sh_entries):
    """the drop-in boundary: load_model(...) / model(locs) -> numpy float64 (N, 1280)"""
    from range_b200.load_model import load_model
    g = golden
    sd = {}
    names = ["layers.0", "layers.1", "last_layer"]
    for i, nm in enumerate(names):
        sd[f"model.location.nnet.{nm}.weight"] = torch.tensor(g[f"W{i}"])
        sd[f"model.location.nnet.{nm}.bias"] = torch.tensor(g[f"b{i}"])
    hp = dict(embed_dim=256, legendre_polys=40, le_type="sphericalharmonics", pe_type="siren",
              harmonics_calculation="analytic", num_hidden_layers=2, capacity=64, eval_downstream=False,
              air_temp_data_path=None, election_data_path=None)
    ckpt = tmp_path / "satclip.ckpt"
    torch.save({"hyper_parameters": hp, "state_dict": sd}, ckpt)
    db = O.synthetic_db(int(g["M"]), seed=int(g["db_seed"]), kind="iid")
    db = {k: v.astype(np.float32).astype(np.float64) for k, v in db.items()}
    dbfile = tmp_path / "db.npz"
    np.savez(dbfile, **db)
    with pytest.raises(ValueError):
        load_model("RANGE+", None, device="cuda")
    with pytest.raises(AssertionError):
        load_model("RANGE+", str(ckpt), device="cuda")
    model = load_model("RANGE+", str(ckpt), device="cuda", db_path=str(dbfile), beta=0.5, chunk=24)
    assert model.location_feature_dim == 1280 and model.args.temp == 12.0 and model.args.geo_temp == 40.0
    out = model(torch.tensor(g["coords"]).to("cuda"))          # chunk=24 -> exercises the pipelined path
    assert isinstance(out, np.ndarray) and out.dtype == np.float64 and out.shape == (64, 1280)
    assert rel_rows(out[:, :1024], g["O_plus_0.5"]).max() <= TOL_O
    assert np.abs(out[:, 1024:] - g["q"]).max() <= TOL_Q_POLAR
    m2 = load_model("RANGE", str(ckpt), device="cuda", db_path=str(dbfile))
    out2 = m2(torch.tensor(g["coords"]))
    assert m2.args.temp == 15.0 and rel_rows(out2[:, :1024], g["O_range"]).max() <= TOL_O_IID_SEM
    with pytest.raises(ValueError):
        load_model("RANGE++", str(ckpt), device="cuda", db_path=str(dbfile))


def test_large_batch_producer_consumer_apply(sh_entries):
    """batches of >= 48 query-tile pairs run the role-specialised apply kernel (retrieval_pc.cu): ragged N and M,
    spatially batched queries (geo-term skipping active), against the exact oracle"""
    from range_b200.engine import RangeEngine
    from range_b200.database import DeviceDatabase
    N, M = 12_300 + 37, 3000 + 5
    db = O.synthetic_db(M, seed=6, kind="iid")
    ws = O.siren_init(40, 64, 2, 256, seed=3)
    eng = RangeEngine(DEV, encoder=dict(L=40, dims=[1600, 64, 64, 256], weights=ws), database=DeviceDatabase(db, DEV))
    c = O.area_uniform(N, np.random.default_rng(21))
    cs, perm = eng.sort_queries(torch.tensor(c))
    q64, q16, qxyz = eng.encode(cs)
    p = perm.cpu().numpy().astype(np.int64)
    sub = np.linspace(0, N - 1, 400).astype(np.int64)               # oracle on a spread of sorted rows
    for name, beta in [("RANGE+", 0.5), ("RANGE", None), ("RANGE+", 0.0)]:
        orc = O.RangeOracle(name, ws, sh_entries, db, beta=beta, exact=True)
        Ot = eng.retrieve(name, q16, qxyz, orc.temp, 40.0, beta).cpu().numpy()
        assert np.isfinite(Ot).all()
        ref = orc(c[p[sub]])[:, :1024]
        r = rel_rows(Ot[sub], ref)
        assert r.max() <= tol_o(name, beta) and r.mean() <= 6e-4, (name, beta, r.max(), r.mean())
    # and the same rows through the single-role kernel (small batch): same arithmetic, same P' rounding
    small = eng.retrieve("RANGE+", q16[:1000], qxyz[:1000], 12.0, 40.0, 0.5)
    big = eng.retrieve("RANGE+", q16, qxyz, 12.0, 40.0, 0.5)[:1000]
    assert rel_rows(big.cpu().numpy(), small.cpu().numpy()).max() <= 5e-4     # fp16 rounding of P' (row sums differ in the last bit)


def test_full_size_properties(sh_entries):
    """BASELINE.json config 2 at full size (100 000 queries x 100 000 entries, H = 512) through the public module:
    sampled rows against the exact oracle, linearity in beta over every row, unit-norm location columns."""
    from argparse import Namespace
    from range_b200.range import LocationEncoder
    N = M = 100_000
    rng = np.random.default_rng(0)
    db = dict(locs=O.area_uniform(M, rng), satclip_embeddings=rng.standard_normal((M, 256), dtype=np.float32),
              image_embeddings=rng.standard_normal((M, 1024), dtype=np.float32))
    ws = O.siren_init(40, 512, 2, 256, seed=0)
    enc = dict(L=40, dims=[1600, 512, 512, 256], weights=ws)
    c = O.area_uniform(N, np.random.default_rng(1))
    dc = torch.tensor(c, device=DEV)
    outs = {}
    for beta in (0.0, 0.5, 1.0):
        m = LocationEncoder(Namespace(location_model_name="RANGE+", pretrained_path=enc, device=DEV, range_db=db, beta=beta))
        outs[beta] = m.embed(dc, out_dtype=torch.float64)
        del m
    mid = outs[0.5][:, :1024]
    lin = 0.5 * outs[0.0][:, :1024] + 0.5 * outs[1.0][:, :1024]
    rel = ((mid - lin).norm(dim=1) / mid.norm(dim=1))
    assert rel.max().item() <= 2e-3 and rel.mean().item() <= 3e-4, (rel.max().item(), rel.mean().item())
    qn = outs[0.5][:, 1024:].norm(dim=1)
    assert (qn - 1).abs().max().item() < 1e-12
    assert torch.equal(outs[0.0][:, 1024:], outs[1.0][:, 1024:])
    sub = np.linspace(0, N - 1, 192).astype(np.int64)
    ref = O.RangeOracle("RANGE+", ws, sh_entries, db, beta=0.5, exact=True)(c[sub])
    got = outs[0.5][torch.tensor(sub, device=DEV)].cpu().numpy()
    r = rel_rows(got[:, :1024], ref[:, :1024])
    assert r.max() <= TOL_O and r.mean() <= 6e-4, (r.max(), r.mean())
    assert cos_rows(got[:, :1024], ref[:, :1024]).min() >= 0.99999
    lat = np.abs(c[sub, 1])
    dq = np.abs(got[:, 1024:] - ref[:, 1024:])
    assert dq[lat < 60].max() <= TOL_Q and dq.max() <= TOL_Q_POLAR
    # the structured database (peaky softmax over non-zero-mean values, like real SatMAE features) at the same size:
    # keys = the encoder's own embedding of the entry location + noise
    rng = np.random.default_rng(3)
    eng = LocationEncoder(Namespace(location_model_name="RANGE+", pretrained_path=enc, device=DEV, range_db=db, beta=0.5)).engine
    E = eng.encode(torch.tensor(db["locs"], device=DEV))[0].cpu().numpy()
    K = E + 0.1 * rng.standard_normal((M, 256))
    V = (K @ rng.standard_normal((256, 1024)) / 16 + 0.5 + 0.2 * rng.standard_normal((M, 1024))).astype(np.float32)
    sdb = dict(locs=db["locs"], satclip_embeddings=K.astype(np.float32), image_embeddings=V)
    del eng
    for name, beta in (("RANGE+", 0.5), ("RANGE", None)):
        m = LocationEncoder(Namespace(location_model_name=name, pretrained_path=enc, device=DEV, range_db=sdb, beta=beta))
        got = m.embed(dc, out_dtype=torch.float64)[torch.tensor(sub, device=DEV)].cpu().numpy()
        ref = O.RangeOracle(name, ws, sh_entries, sdb, beta=beta, exact=True)(c[sub])
        r = rel_rows(got[:, :1024], ref[:, :1024])
        assert r.max() <= TOL_O_STRUCTURED, (name, r.max())
        del m


@pytest.mark.parametrize("N,M", [(13_000, 77), (12_288, 128), (20_000, 129)])
def test_large_batch_tiny_database(N, M, sh_entries):
    """producer/consumer kernels with one or two database tiles (most softmax groups idle, ragged last tile)"""
    from range_b200.engine import RangeEngine
    from range_b200.database import DeviceDatabase
    db = O.synthetic_db(M, seed=9, kind="iid")
    d = DeviceDatabase(db, DEV)
    eng = RangeEngine(DEV, L=40, database=d)
    g = torch.Generator(device="cpu").manual_seed(5)
    q = torch.randn(N, 256, generator=g); q = (q / q.norm(dim=1, keepdim=True)).half().to(DEV)
    c = O.area_uniform(N, np.random.default_rng(2))
    xyz = torch.zeros(N, 4); xyz[:, :3] = torch.tensor(O.rad_to_cart(c * np.pi / 180)).float(); xyz = xyz.to(DEV)
    K = d.Kh[:M].float(); V = d.Vt[:, :M].float().t() / d.vscale; X = d.xyz[:M, :3]
    out = eng.retrieve("RANGE+", q, xyz, 12.0, 40.0, 0.25)
    P = 0.25 * torch.softmax((q.float() @ K.t()) * 12.0, dim=1) + 0.75 * torch.softmax((xyz[:, :3] @ X.t()) * 40.0, dim=1)
    ref = P @ V
    rel = ((out - ref).norm(dim=1) / ref.norm(dim=1))
    assert torch.isfinite(out).all() and rel.max().item() <= TOL_O, rel.max().item()


def test_beta_sweep_and_database_cache(tmp_path, sh_entries):
    """embed_sweep (two apply passes - the geographic and the semantic end - blended per beta, range.py:238 is linear
    in beta) against one model per beta and the exact oracle; the on-disk device layout round-trips"""
    from argparse import Namespace
    from range_b200.range import LocationEncoder
    db = O.synthetic_db(4000, seed=8, kind="iid")
    ws = O.siren_init(40, 64, 2, 256, seed=4)
    enc = dict(L=40, dims=[1600, 64, 64, 256], weights=ws)
    c = torch.tensor(O.area_uniform(700, np.random.default_rng(5)), device=DEV)
    cache = str(tmp_path / "db_layout")           # no '.npz': the cache file name is normalised
    mk = lambda beta, **kw: LocationEncoder(Namespace(location_model_name="RANGE+", pretrained_path=enc, device=DEV,
                                                      range_db=db, beta=beta, **kw))
    m = mk(0.5, db_cache=cache)                   # builds and writes the cache
    import os
    assert os.path.exists(cache + ".npz")
    betas = [0.0, 0.25, 0.5, 0.75, 1.0]
    sweep = m.embed_sweep(c, betas)
    for beta, got in zip(betas, sweep):
        one = mk(beta, db_cache=cache).embed(c)   # reads the cache
        assert torch.equal(got[:, 1024:], one[:, 1024:]), beta
        assert rel_rows(got[:, :1024].cpu().numpy(), one[:, :1024].cpu().numpy()).max() <= 5e-4, beta
        ref = O.RangeOracle("RANGE+", ws, sh_entries, db, beta=beta, exact=True)(c.cpu().numpy())
        # the iid database is the worst case of the purely semantic softmax (flat weights over zero-mean values: the
        # fp16 rounding of q and K shows undamped): 2e-3 there, 1e-3 as soon as the geographic term takes part
        assert rel_rows(got[:, :1024].cpu().numpy(), ref[:, :1024]).max() <= tol_o("RANGE+", beta), beta
    fresh = mk(0.25).embed(c)
    assert torch.equal(fresh, mk(0.25, db_cache=cache).embed(c))


def test_sort_is_a_pure_function_of_the_coordinates():
    """the spatial batching permutation is a stable sort by cell: deterministic for any cell occupancy (clustered
    query sets, raster chunks, duplicates), so every rank of an M-sharded run derives the same row order"""
    from range_b200.engine import RangeEngine
    eng = RangeEngine(DEV, L=40)
    rng = np.random.default_rng(3)
    same = torch.tensor(np.tile([[12.5, 47.25]], (5000, 1)), device=DEV)
    _, perm = eng.sort_queries(same)
    assert torch.equal(perm.long(), torch.arange(5000, device=DEV))            # one cell: the order is kept
    region = np.stack([rng.uniform(10, 11, 30000), rng.uniform(45, 46, 30000)], 1)          # ~30000 points in a few cells
    lat, lon = np.meshgrid(np.linspace(60, 58, 96), np.linspace(-180, 180, 256), indexing="ij")
    raster = np.stack([lon.ravel(), lat.ravel()], 1)                                        # a lat-major raster chunk
    for pts in (region, raster, O.area_uniform(100_000, rng)):
        c = torch.tensor(pts, device=DEV)
        s1, p1 = eng.sort_queries(c)
        s2, p2 = eng.sort_queries(c.clone())
        assert torch.equal(p1, p2) and torch.equal(s1, s2)
        assert torch.equal(torch.sort(p1.long()).values, torch.arange(len(pts), device=DEV))
        assert torch.equal(s1, c[p1.long()])
    twice = torch.tensor(np.concatenate([region[:4000], region[:4000]]), device=DEV)      # equal keys: index order kept
    _, p = eng.sort_queries(twice)
    pos = torch.empty(8000, dtype=torch.long, device=DEV)
    pos[p.long()] = torch.arange(8000, device=DEV)
    assert (pos[:4000] < pos[4000:]).all()


@pytest.mark.parametrize("slab", [256, 4224])
def test_m_sharded_routed_merge_emulated(slab, sh_entries):
    """The M-sharded pipeline of range_b200/distributed.py with its ranks emulated one after the other on one GPU
    (the kernels never wait for another rank): every rank's own slab of sorted queries, statistics per shard, SUM of the
    exp-sums with LOCAL maxima, the apply pass storing its partial rows through the route (peer pointers = the owners'
    receive buffers), per-owner merge in rank order.  slab 256: single-role kernels + row router; slab 4224: the
    producer/consumer kernel's epilogue routes the rows itself."""
    import ctypes
    from range_b200 import _lib
    from range_b200.engine import RangeEngine
    from range_b200.database import DeviceDatabase
    P, M = 3, 5000
    ws = O.siren_init(40, 64, 2, 256, seed=2)
    helper = O.RangeOracle.__new__(O.RangeOracle)
    helper.L, helper.entries, helper.weights = 40, sh_entries, ws
    db = O.synthetic_db(M, seed=4, kind="structured", encoder=helper.encode)
    enc = dict(L=40, dims=[1600, 64, 64, 256], weights=ws)
    whole = RangeEngine(DEV, encoder=enc, database=DeviceDatabase(db, DEV))
    shards = [RangeEngine(DEV, encoder=enc, database=DeviceDatabase(db, DEV, shard=(r, P))) for r in range(P)]
    coords = [torch.tensor(O.area_uniform(slab, np.random.default_rng(30 + r)), device=DEV) for r in range(P)]
    own = [shards[r].sort_queries(coords[r]) for r in range(P)]
    encd = [shards[r].encode(own[r][0]) for r in range(P)]
    q16_all = torch.cat([e[1] for e in encd]).contiguous()
    qxyz_all = torch.cat([e[2] for e in encd]).contiguous()
    recv = [torch.full((P, slab, 1024), float("nan"), device=DEV) for _ in range(P)]
    ptrs = (ctypes.c_void_p * _lib.RANGE_MAX_RANKS)(*[b.data_ptr() for b in recv])
    for name, beta, temp in [("RANGE+", 0.5, 12.0), ("RANGE", None, 15.0)]:
        stats = [s.retrieve_stats(name, q16_all, qxyz_all, temp, 40.0) for s in shards]
        sums = torch.stack([a for a, _ in stats]).sum(0)
        for r in range(P):
            shards[r].retrieve_apply_routed(name, q16_all, qxyz_all, temp, 40.0, beta, sums, stats[r][1],
                                            _lib.Route(P, r, slab, ptrs))
        for r in range(P):
            assert torch.isfinite(recv[r]).all()                     # every slot of every owner was written
            got = shards[r].combine_concat([recv[r][k] for k in range(P)], None, encd[r][0], perm=own[r][1],
                                           dtype=torch.float64)
            q64, q16, qxyz = whole.encode(coords[r])
            ref = whole.concat(whole.retrieve(name, q16, qxyz, temp, 40.0, beta), q64)
            assert torch.equal(got[:, 1024:], ref[:, 1024:])
            rel = rel_rows(got[:, :1024].cpu().numpy(), ref[:, :1024].cpu().numpy())
            assert rel.max() <= 3e-4, (name, r, rel.max())
        for b in recv:
            b.fill_(float("nan"))


def test_host_paths_agree(sh_entries):
    """model(locs) hands the rows to the host in two ways (range.py:_forward_host): float64 rows copied chunk by chunk
    into the page-locked result, or packed rows (fp32 features) widened on the host - also into the caller's own array
    (embed_into, what save.embed_to_npy uses on a memory-mapped .npy)"""
    from argparse import Namespace
    from range_b200.range import LocationEncoder
    db = O.synthetic_db(3000, seed=6, kind="iid")
    ws = O.siren_init(40, 64, 2, 256, seed=3)
    enc = dict(L=40, dims=[1600, 64, 64, 256], weights=ws)
    c = torch.tensor(O.area_uniform(30_000, np.random.default_rng(7)))
    outs = {}
    for path in ("copy", "packed"):
        m = LocationEncoder(Namespace(location_model_name="RANGE+", pretrained_path=enc, device=DEV, range_db=db, beta=0.5,
                                      host_path=path, chunk=12288, tail=6144, super_batch=24576))
        outs[path] = m(c)
        assert outs[path].dtype == np.float64 and outs[path].shape == (30_000, 1280)
    assert np.array_equal(outs["copy"], outs["packed"])              # same kernels; the widening is exact
    m32 = LocationEncoder(Namespace(location_model_name="RANGE+", pretrained_path=enc, device=DEV, range_db=db, beta=0.5,
                                    out_dtype=np.float32, chunk=12288, tail=6144, super_batch=24576))
    o32 = m32(c)                                                      # opt-in: float32 rows (not the reference's dtype)
    assert o32.dtype == np.float32 and np.array_equal(o32[:, :1024], outs["copy"][:, :1024].astype(np.float32))
    assert np.array_equal(o32[:, 1024:], outs["copy"][:, 1024:].astype(np.float32))
    sub = np.linspace(0, 29_999, 100).astype(np.int64)
    ref = O.RangeOracle("RANGE+", ws, sh_entries, db, beta=0.5, exact=True)(c.numpy()[sub])
    assert rel_rows(outs["copy"][sub, :1024], ref[:, :1024]).max() <= 1e-3
    # the streaming writer of save.py on the real model: rows land in a memory-mapped .npy
    import os
    import tempfile
    from range_b200.save import embed_to_npy
    with tempfile.TemporaryDirectory() as d:
        mm = embed_to_npy(m, c.numpy(), os.path.join(d, "emb.npy"), batch=20_000)
        got = np.asarray(mm)
        # other batch boundaries -> other query tiles: the fp16 rounding of the weights differs in the last bit
        assert mm.shape == (30_000, 1280) and np.array_equal(got[:, 1024:], outs["copy"][:, 1024:])
        assert rel_rows(got[:, :1024], outs["copy"][:, :1024]).max() <= 5e-4
        del mm
        assert np.array_equal(np.load(os.path.join(d, "emb.npy"), mmap_mode="r")[:100], got[:100])


def test_closed_form_harmonics_vs_reference(golden):
    """harmonics_calculation='closed-form' through both encoder paths against the reference's own output"""
    import os
    from range_b200.engine import RangeEngine
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "closed_form_golden.npz"))
    weights = [(torch.tensor(golden[f"W{i}"]), torch.tensor(golden[f"b{i}"])) for i in range(3)]
    enc = dict(L=40, dims=[1600, 64, 64, 256], weights=weights, harmonics_calculation="closed-form")
    eng = RangeEngine(DEV, encoder=enc, encoder_precision="fp64")
    assert eng.harmonics == "closed-form"
    Y = eng.sh_features(torch.tensor(g["coords"])).cpu().numpy()
    # the recurrence is evaluated in the reference's operation order: only libm-vs-CUDA cos / sin / sqrt differ
    assert np.abs(Y - g["Y"]).max() <= 1e-12 * max(1.0, np.abs(g["Y"]).max())
    q64, _, _ = eng.encode(torch.tensor(g["coords"]))
    assert np.abs(q64.cpu().numpy() - g["q"]).max() <= 1e-10
    # tensor-core encoder (needs widths % 256 == 0): H = 512 against the oracle
    ws = O.siren_init(40, 512, 2, 256, seed=0)
    etc = RangeEngine(DEV, encoder=dict(L=40, dims=[1600, 512, 512, 256], weights=ws, harmonics_calculation="closed-form"))
    assert etc.precision == "f16x3"
    c = O.area_uniform(777, np.random.default_rng(4))
    q, _, _ = etc.encode(torch.tensor(c))
    ref = O.RangeOracle.__new__(O.RangeOracle)
    ref.L, ref.entries, ref.harmonics = 40, None, "closed-form"
    ref.weights = [(torch.as_tensor(W, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)) for W, b in ws]
    assert np.abs(q.cpu().numpy() - ref.encode(torch.tensor(c)).numpy()).max() <= 5e-6


@pytest.mark.parametrize("L", [10, 20, 32])
def test_other_legendre_degrees(L, sh_entries):
    """checkpoints with legendre_polys != 40: L = 10 (100 features -> zero-padded first layer, fp64 encoder), L = 20
    (fp64 encoder), L = 32 (1024 features: tensor-core encoder) - against the oracle, both harmonics flavours"""
    from range_b200.engine import RangeEngine
    H = 256
    ws = O.siren_init(L, H, 2, 256, seed=L)
    c = O.area_uniform(300, np.random.default_rng(L))
    for flavour in ("analytic", "closed-form"):
        eng = RangeEngine(DEV, encoder=dict(L=L, dims=[L * L, H, H, 256], weights=ws, harmonics_calculation=flavour))
        assert eng.precision == ("f16x3" if (L * L) % 64 == 0 else "fp64")
        q64, q16, qxyz = eng.encode(torch.tensor(c))
        ref = O.RangeOracle.__new__(O.RangeOracle)
        ref.L, ref.entries, ref.harmonics = L, sh_entries, flavour
        ref.weights = [(torch.as_tensor(W, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)) for W, b in ws]
        d = np.abs(q64.cpu().numpy() - ref.encode(torch.tensor(c)).numpy()).max()
        assert d <= (5e-6 if eng.precision == "f16x3" else (1e-9 if flavour == "closed-form" or L <= 20 else 1e-5)), (L, flavour, d)
        Y = eng.sh_features(torch.tensor(c)).cpu().numpy()
        Yref = (O.sh_analytic(c, L, sh_entries) if flavour == "analytic" else O.sh_closed_form(c, L)).numpy()
        assert Y.shape == (300, L * L)
        lo = min(L, 20) ** 2                       # l < 20: no cancellation in the generated polynomials
        assert np.abs(Y - Yref)[:, :lo].max() <= 1e-9
        # l >= 20 ('analytic' only): the reference's own fp64 noise on its cancelling 15-digit polynomials
        assert np.abs(Y - Yref).max() <= (1e-3 if flavour == "analytic" else 1e-9)


@pytest.mark.parametrize("N,M", [(24_576, 30_011), (13_000, 200_000)])
def test_producer_consumer_retrieval_is_deterministic(N, M):
    """a hand-off race in the P' ring (stale slot, early release) would show up as sporadically different rows:
    repeated launches must be bit-identical (tools/stress_pc.py runs more shapes and repetitions)"""
    from range_b200.engine import RangeEngine
    from range_b200.database import DeviceDatabase
    eng = RangeEngine(DEV, L=40, database=DeviceDatabase.synthetic(M, DEV, seed=N))
    g = torch.Generator(device="cpu").manual_seed(1)
    q = torch.randn(N, 256, generator=g); q = (q / q.norm(dim=1, keepdim=True)).half().to(DEV)
    c = eng.sort_queries(torch.tensor(O.area_uniform(N, np.random.default_rng(1))))[0].cpu()
    xyz = torch.zeros(N, 4); xyz[:, :3] = torch.tensor(O.rad_to_cart(c.numpy() * np.pi / 180)).float(); xyz = xyz.to(DEV)
    first = eng.retrieve("RANGE+", q, xyz, 12.0, 40.0, 0.5).clone()
    assert torch.isfinite(first).all()
    for _ in range(8):
        assert torch.equal(eng.retrieve("RANGE+", q, xyz, 12.0, 40.0, 0.5), first)


@pytest.mark.parametrize("N", [1_000, 12_288])
def test_no_mass_loss_over_a_long_database_axis(N):
    """The tensor core adds each K = 16 block into its fp32 accumulator with truncation; summed naively over 300 000
    entries that loses 1.5e-3 of every row's mass (6e-3 at 1 M).  With values = 1 every output element is sum_j P_j = 1:
    the accumulation windows (large batches) / database splits (small batches) must keep the deficit at the 1e-5 level."""
    from range_b200.engine import RangeEngine
    from range_b200.database import DeviceDatabase
    M = 300_000
    d = DeviceDatabase.synthetic(M, DEV, seed=3)
    d.Vt.fill_(0)
    d.Vt[:, :M] = d.vscale                                              # V = 1
    eng = RangeEngine(DEV, L=40, database=d)
    g = torch.Generator(device="cpu").manual_seed(1)
    q = torch.randn(N, 256, generator=g); q = (q / q.norm(dim=1, keepdim=True)).half().to(DEV)
    c = eng.sort_queries(torch.tensor(O.area_uniform(N, np.random.default_rng(1))))[0].cpu()
    xyz = torch.zeros(N, 4); xyz[:, :3] = torch.tensor(O.rad_to_cart(c.numpy() * np.pi / 180)).float(); xyz = xyz.to(DEV)
    for mode, beta, t in (("RANGE", None, 15.0), ("RANGE+", 0.5, 12.0)):
        dev = eng.retrieve(mode, q, xyz, t, 40.0, beta).double() - 1.0
        assert abs(dev.mean().item()) < 1e-4 and dev.abs().max().item() < 5e-4, (mode, dev.mean().item(), dev.abs().max().item())


def test_m_sharded_on_two_gpus():
    """tools/multi_gpu_check.py under torchrun on 2 GPUs (skipped on a one-GPU box): the M-sharded pipeline with real
    peer memory and NCCL against the unsharded database, ragged / clustered queries, both merges"""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(root, "tools", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=900, cwd=root, env=dict(os.environ, M="60000", N="9000"))
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-3000:])
    assert "M-sharded [peer -> ran peer]" in p.stdout, p.stdout[-2000:]


@pytest.mark.parametrize("N,M", [(12_337, 3_005), (300, 5_000)])
def test_retrieve_concat_is_stats_plus_apply(N, M):
    """range_retrieve_concat (one call, what model.embed uses) == range_retrieve_stats + range_retrieve_apply_concat"""
    from range_b200.engine import RangeEngine
    from range_b200.database import DeviceDatabase
    db = O.synthetic_db(M, seed=6, kind="iid")
    ws = O.siren_init(40, 64, 2, 256, seed=3)
    eng = RangeEngine(DEV, encoder=dict(L=40, dims=[1600, 64, 64, 256], weights=ws), database=DeviceDatabase(db, DEV))
    cs, perm = eng.sort_queries(torch.tensor(O.area_uniform(N, np.random.default_rng(21))))
    q64, q16, qxyz = eng.encode(cs)
    for name, beta, temp in [("RANGE+", 0.5, 12.0), ("RANGE", None, 15.0)]:
        sums, maxs = eng.retrieve_stats(name, q16, qxyz, temp, 40.0)
        for dt in (torch.float32, torch.float64, torch.uint8):
            two = eng.retrieve_apply_concat(name, q16, qxyz, temp, 40.0, beta, sums, maxs, q64, dtype=dt, perm=perm)
            one = eng.retrieve_concat(name, q16, qxyz, temp, 40.0, beta, q64, dtype=dt, perm=perm)
            assert torch.equal(one, two), (name, dt)
