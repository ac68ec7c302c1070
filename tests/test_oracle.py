"""The CPU oracle against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py) - this is what pins the oracle."""
import numpy as np
import pytest
import torch

from oracle import range_oracle as O


def _weights(g):
    return [(torch.tensor(g[f"W{i}"]), torch.tensor(g[f"b{i}"])) for i in range(3)]


def _db(g):
    from tests.golden.make_golden import checksum
    db = O.synthetic_db(int(g["M"]), seed=int(g["db_seed"]), kind="iid")
    db = {k: v.astype(np.float32).astype(np.float64) for k, v in db.items()}
    if checksum(db["locs"], db["satclip_embeddings"], db["image_embeddings"]) != str(g["db_checksum"]):
        pytest.skip("numpy RNG stream differs from the one the golden DB was drawn with")
    return db


def test_sh_matches_reference(golden, sh_entries):
    Y = O.sh_analytic(golden["coords"], 40, sh_entries).numpy()
    assert np.abs(Y - golden["Y"]).max() <= 4e-16 * np.abs(golden["Y"]).max()


def test_sh_closed_forms(sh_entries):
    pts = np.array([[10.0, 20.0], [-120.0, -45.0], [0.0, 0.0]])
    Y = O.sh_analytic(pts, 3, sh_entries).numpy()
    phi, theta = np.deg2rad(pts[:, 0] + 180), np.deg2rad(pts[:, 1] + 90)
    assert np.allclose(Y[:, 0], 0.886226925452758, rtol=0, atol=1e-15)
    assert np.allclose(Y[:, 2], 1.53499006191973 * np.cos(theta), rtol=0, atol=1e-15)
    assert np.allclose(Y[:, 3], 0.48860251190292 * np.sin(theta) * np.cos(phi), rtol=0, atol=1e-15)
    assert np.allclose(Y[:, 1], 0.48860251190292 * np.sin(theta) * np.sin(phi), rtol=0, atol=1e-15)


def test_reference_deviates_from_exact_harmonics_only_at_high_degree(sh_entries):
    """documents SURVEY.md Appendix B: the 15-digit polynomials are the reference's function"""
    pts = O.area_uniform(512, np.random.default_rng(3))
    Y = O.sh_analytic(pts, 40, sh_entries).numpy()
    Ye = O.sh_exact(pts, 40)
    dev = np.abs(Y - Ye).max(0)
    assert dev[: 20 * 20].max() < 1e-7
    assert dev.max() > 1e-6          # l >= 26 features are visibly off the exact values


def test_forward_matches_reference(golden, sh_entries):
    g = golden
    db, w = _db(g), _weights(g)
    out = O.RangeOracle("RANGE", w, sh_entries, db)(g["coords"])
    assert out.dtype == np.float64 and out.shape == (len(g["coords"]), 1280)
    assert np.abs(out[:, 1024:] - g["q"]).max() < 1e-15
    assert np.abs(out[:, :1024] - g["O_range"]).max() < 1e-6
    for beta in g["betas"]:
        out = O.RangeOracle("RANGE+", w, sh_entries, db, beta=float(beta))(g["coords"])
        assert np.abs(out[:, :1024] - g[f"O_plus_{beta}"]).max() < 1e-6
    assert np.allclose(np.linalg.norm(out[:, 1024:], axis=1), 1.0, atol=1e-14)


def test_beta_limits_and_exact_mode(golden, sh_entries):
    g = golden
    db, w = _db(g), _weights(g)
    c = g["coords"][:16]
    p1 = O.RangeOracle("RANGE+", w, sh_entries, db, beta=1.0)(c)
    ex = O.RangeOracle("RANGE+", w, sh_entries, db, beta=1.0, exact=True)(c)
    rel = np.linalg.norm(p1[:, :1024] - ex[:, :1024], axis=1) / np.linalg.norm(ex[:, :1024], axis=1)
    assert rel.max() < 1e-5
    with pytest.raises(ValueError):
        O.RangeOracle("RANGE++", w, sh_entries, db)


def test_closed_form_harmonics_match_reference(golden):
    """harmonics_calculation='closed-form' (spherical_harmonics_closed_form.py): fixture written by the unmodified
    reference (tests/golden/make_golden.py)"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "closed_form_golden.npz"))
    assert np.array_equal(g["coords"], golden["coords"])
    Y = O.sh_closed_form(g["coords"], 40).numpy()
    assert np.abs(Y - g["Y"]).max() <= 4e-16 * np.abs(g["Y"]).max()
    q = O.RangeOracle("RANGE", _weights(golden), None, _db(golden), harmonics="closed-form").encode(
        torch.tensor(g["coords"])).numpy()
    assert np.abs(q - g["q"]).max() <= 1e-14
    # known answers: orthonormal Y_00, and the Condon-Shortley sign of m = 1
    pts = np.array([[10.0, 20.0], [-120.0, -45.0]])
    y = O.sh_closed_form(pts, 2).numpy()
    theta, phi = np.deg2rad(pts[:, 1] + 90), np.deg2rad(pts[:, 0] + 180)
    assert np.allclose(y[:, 0], 0.5 / np.sqrt(np.pi))
    assert np.allclose(y[:, 3], -np.sqrt(3 / (4 * np.pi)) * np.sin(theta) * np.cos(phi))
    from range_b200.sh_table import closed_form_norms
    assert np.array_equal(closed_form_norms(40), O.closed_form_norms(40))


def test_product_side_generators_match_the_oracles():
    """bench.py / smoke() draw their synthetic inputs from range_b200.synthetic (the product never imports oracle/)"""
    import torch
    from range_b200 import synthetic as S
    a, b = S.area_uniform(1000, np.random.default_rng(4)), O.area_uniform(1000, np.random.default_rng(4))
    assert np.array_equal(a, b)
    for (w1, b1), (w2, b2) in zip(S.siren_init(40, 64, 2, 256, seed=3), O.siren_init(40, 64, 2, 256, seed=3)):
        assert torch.equal(w1, w2) and torch.equal(b1, b2)
