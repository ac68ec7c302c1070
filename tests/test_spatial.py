"""Spatial ordering of the database / queries and the geographic tile-skip bound.

The reference's result does not depend on row order (range/range.py:213-238 are row-wise sums), so these
are properties of OUR layout: the Hilbert order is a permutation, every tile's cap contains its entries, and a
skipped (query tile, database tile) pair only ever holds pairs whose geo weight is below 2^-24 of the row's
normaliser."""
import numpy as np
import pytest
import torch

from oracle import range_oracle as O


def test_hilbert_order_and_caps_cpu():
    from range_b200.database import hilbert_order, tile_caps, prepare_reference_arrays
    db = O.synthetic_db(5000, seed=7, kind="iid")
    _, _, xyz = prepare_reference_arrays(db)
    order = hilbert_order(xyz)
    assert sorted(order.tolist()) == list(range(5000))
    s = xyz[order]
    caps = tile_caps(s)
    assert caps.shape == (40, 4) and caps.dtype == np.float32
    for t in range(40):
        pts = s[t * 128:(t + 1) * 128].astype(np.float64)
        ang = np.arccos(np.clip(pts @ caps[t, :3].astype(np.float64), -1, 1))
        assert ang.max() <= caps[t, 3] + 1e-5
    # locality: Hilbert tiles are far smaller than tiles of the unsorted rows
    assert np.median(caps[:, 3]) < 0.25 * np.median(tile_caps(xyz)[:, 3])
    # identical points / tiny inputs
    one = np.repeat(xyz[:1], 3, axis=0)
    assert tile_caps(one)[0, 3] < 1e-3 and hilbert_order(one).tolist() == [0, 1, 2]


@pytest.mark.parametrize("regional", [False, True])
def test_geo_skip_rule_is_exact_cpu(regional):
    """the skip decision (numpy restatement of geo_mask_kernel, oracle/geo_skip.py) never drops an entry that carries more
    than 2^-24 / M of its row's geo mass - statistics-pass and apply-pass flavours, global and regional databases"""
    from oracle import geo_skip as GS
    from range_b200.database import hilbert_order, prepare_reference_arrays
    M, N = 6000, 3000
    db = O.synthetic_db(M, seed=11, kind="iid")
    if regional:
        db["locs"][:, 1] = 30.0 + 0.5 * db["locs"][:, 1]
        db["locs"][:, 0] = 0.25 * db["locs"][:, 0]
    _, _, xyz = prepare_reference_arrays(db)
    xyz = xyz[hilbert_order(xyz)].astype(np.float64)
    q = O.rad_to_cart(O.area_uniform(N, np.random.default_rng(12)) * np.pi / 180).reshape(-1, 3)
    q = q[hilbert_order(q)]
    m1 = GS.skip_mask(q, xyz, M)
    slack, lg = GS.exactness_slack(q, xyz, M, m1)
    assert slack >= -1e-9
    m2 = GS.skip_mask(q, xyz, M, lg=lg)
    assert (m2 | ~m1).all() and m2.sum() >= m1.sum()
    assert GS.exactness_slack(q, xyz, M, m2)[0] >= -1e-9
    if not regional:
        assert m2.mean() > 0.1                      # 24 x 47 tiles over the globe: a fair share is skippable
    # what is dropped really is below fp32 resolution of the geo softmax
    G = q @ xyz.T
    w = np.exp(40.0 * (G - 1)) / lg[:, None]
    pad = (-N) % 128, (-M) % 128
    wt = np.pad(w, ((0, pad[0]), (0, pad[1]))).reshape(len(m2), 128, m2.shape[1], 128)
    dropped = (wt * m2[:, None, :, None]).sum((2, 3))
    assert dropped.max() < 2.0 ** -24


@pytest.mark.gpu
@pytest.mark.parametrize("N", [1, 5, 129, 4096, 100_000])
def test_sort_queries_is_a_deterministic_permutation(N):
    from range_b200.engine import RangeEngine
    eng = RangeEngine("cuda:0", L=40)
    c = torch.tensor(O.area_uniform(N, np.random.default_rng(N)))
    s1, p1 = eng.sort_queries(c)
    s2, p2 = eng.sort_queries(c)
    p = p1.cpu().numpy()
    assert sorted(p.tolist()) == list(range(N))
    assert torch.equal(p1, p2) and torch.equal(s1, s2)
    assert np.array_equal(s1.cpu().numpy(), c.numpy()[p])
    if N >= 4096:
        # consecutive sorted queries are neighbours: 128-query tiles are compact
        x = O.rad_to_cart(s1.cpu().numpy() * np.pi / 180).reshape(-1, 3)
        T = N // 128
        x = x[: T * 128].reshape(T, 128, 3)
        cen = x.mean(1); cen /= np.linalg.norm(cen, axis=1, keepdims=True)
        rad = np.arccos(np.clip(np.einsum("tbi,ti->tb", x, cen), -1, 1)).max(1)
        assert np.median(rad) < 4.0 * np.sqrt(4 * np.pi * 128 / N / np.pi)     # a few times the ideal disc radius


@pytest.mark.gpu
def test_sort_queries_identical_points():
    from range_b200.engine import RangeEngine
    eng = RangeEngine("cuda:0", L=40)
    c = torch.tensor(np.tile(np.array([[12.5, 41.9]]), (3000, 1)))
    s, p = eng.sort_queries(c)
    assert sorted(p.cpu().tolist()) == list(range(3000)) and torch.equal(s.cpu(), c)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["global", "regional_db"])
def test_geo_skip_bound_and_equivalence(kind, sh_entries):
    """every skipped tile pair satisfies g <= g_max(row) - delta for all its pairs; skipping changes nothing
    beyond fp32 rounding; and it does skip when queries are batched spatially"""
    from range_b200.database import DeviceDatabase
    from range_b200.engine import RangeEngine
    M, N = 20_000, 16_384
    db = O.synthetic_db(M, seed=5, kind="iid")
    if kind == "regional_db":                       # database confined to a cap: far queries have g_max << 1
        db["locs"][:, 1] = 30.0 + 0.5 * db["locs"][:, 1]
        db["locs"][:, 0] = 0.25 * db["locs"][:, 0]
    ws = O.siren_init(40, 64, 2, 256, seed=1)
    enc = dict(L=40, dims=[1600, 64, 64, 256], weights=ws)
    dsort = DeviceDatabase(db, "cuda:0")
    dplain = DeviceDatabase(db, "cuda:0", spatial_sort=False)
    assert dsort.caps is not None and dplain.caps is None
    eng = RangeEngine("cuda:0", encoder=enc, database=dsort)
    ref = RangeEngine("cuda:0", encoder=enc, database=dplain)
    c, perm = eng.sort_queries(torch.tensor(O.area_uniform(N, np.random.default_rng(3))))
    q64, q16, qxyz = eng.encode(c)
    mask = eng.geo_mask(qxyz, 40.0)                                   # (128 query tiles, 157 database tiles)
    assert mask.shape == (N // 128, dsort.Mpad // 128)
    frac = mask.float().mean().item()
    if kind == "global":
        assert frac > 0.1, frac                # 128 query tiles over the globe; the bench shape (782 tiles) skips ~45 %
    G = qxyz[:, :3] @ dsort.xyz[: dsort.M, :3].t()                    # (N, M) cosines
    pad = dsort.Mpad - dsort.M
    Gt = torch.nn.functional.pad(G, (0, pad), value=-2.0).reshape(N // 128, 128, -1, 128)
    tile_max = Gt.amax(dim=3)                                          # (qtile, row, dbtile)
    # an entry is negligible for row i iff exp(T (g - 1)) <= 2^-24 l_g,i / M, i.e. g <= thr_i (exact l_g in fp64)
    lg = torch.exp(40.0 * (G.double() - 1)).sum(1)
    thr_row = (1 + (torch.log(lg) - np.log(M) - 24 * np.log(2)) / 40.0).float().reshape(N // 128, 128, 1)
    slack = thr_row - tile_max                                         # must be >= 0 wherever a tile is skipped
    assert slack[mask.unsqueeze(1).expand_as(slack)].min().item() >= -1e-5
    # apply pass: the normalisers are known (statistics-pass mask: lower bounds from the tile caps only)
    sums, _ = eng.retrieve_stats("RANGE+", q16, qxyz, 12.0, 40.0)
    mask2 = eng.geo_mask(qxyz, 40.0, sums=sums)
    assert bool((mask2 | ~mask).all())                                # never skips less than the statistics-pass mask
    assert mask2.float().mean().item() >= frac
    assert slack[mask2.unsqueeze(1).expand_as(slack)].min().item() >= -1e-5
    # the kernel's decisions against the numpy restatement (fp32 vs fp64 trigonometry: a handful of borderline tiles)
    from oracle import geo_skip as GS
    m_ref = GS.skip_mask(qxyz[:, :3].double().cpu().numpy(), dsort.xyz[: dsort.M, :3].double().cpu().numpy(), M)
    assert (m_ref != mask.cpu().numpy()).mean() < 2e-3
    a = eng.retrieve("RANGE+", q16, qxyz, 12.0, 40.0, 0.5)
    b = ref.retrieve("RANGE+", q16, qxyz, 12.0, 40.0, 0.5)
    rel = ((a - b).norm(dim=1) / b.norm(dim=1)).max().item()
    assert rel < 1e-3, rel                    # database row order differs -> fp16/fp32 summation order only
    a0 = eng.retrieve("RANGE+", q16, qxyz, 12.0, 40.0, 0.0)           # geo only: the skipped mass is all there is to lose
    b0 = ref.retrieve("RANGE+", q16, qxyz, 12.0, 40.0, 0.0)
    assert ((a0 - b0).norm(dim=1) / b0.norm(dim=1)).max().item() < 1e-3
    sa, ma = eng.retrieve_stats("RANGE+", q16, qxyz, 12.0, 40.0)
    sb, mb = ref.retrieve_stats("RANGE+", q16, qxyz, 12.0, 40.0)
    assert torch.allclose(sa, sb, rtol=2e-5, atol=0) and torch.allclose(ma, mb, atol=1e-6)
    # scatter-concat puts row i back at the caller's position perm[i]
    out = eng.concat(a, q64, perm=perm, dtype=torch.float32)
    assert torch.equal(out[perm.long(), :1024], a)
