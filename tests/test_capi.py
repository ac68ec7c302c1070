"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/range_b200.h declares; the host-side API mirrors the reference's argument handling.
No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "range_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(range_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from range_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.PROTOTYPES) == names       # the ctypes table covers the header exactly
    assert lib.range_version() == 100
    assert lib.range_launch_count() == 0


def test_errors_without_gpu_are_loud():
    import torch
    from range_b200 import _lib
    from range_b200.engine import RangeEngine
    with pytest.raises(_lib.RangeError):
        RangeEngine("cpu", L=40)
    if not torch.cuda.is_available():
        with pytest.raises(_lib.RangeError):
            RangeEngine("cuda", L=40)
    lib = _lib.load()
    assert lib.range_ctx_set_db(None, 10, 128, None, None, None, ctypes.c_float(1.0)) == -1
    assert b"bad database" in lib.range_last_error()
    assert lib.range_retrieve_workspace_bytes(None, 100) == 0


def test_load_model_argument_contract():
    from range_b200.load_model import load_model
    with pytest.raises(ValueError, match="pretrained model path"):
        load_model("RANGE+", None)
    with pytest.raises(AssertionError, match="db_path is required"):
        load_model("RANGE", "x.ckpt")
    with pytest.raises(NotImplementedError):
        load_model("SatCLIP", "x.ckpt")


def test_checkpoint_reader(tmp_path, golden):
    import torch
    from range_b200.checkpoint import load_satclip_location_encoder
    g = golden
    sd = {"model.visual.junk": torch.zeros(3), "model.logit_scale": torch.ones(())}
    for i, nm in enumerate(["layers.0", "layers.1", "last_layer"]):
        sd[f"model.location.nnet.{nm}.weight"] = torch.tensor(g[f"W{i}"])
        sd[f"model.location.nnet.{nm}.bias"] = torch.tensor(g[f"b{i}"])
    hp = dict(embed_dim=256, legendre_polys=40, le_type="sphericalharmonics", pe_type="siren",
              harmonics_calculation="analytic", num_hidden_layers=2, capacity=64, eval_downstream=False,
              air_temp_data_path=None, election_data_path=None)
    p = tmp_path / "c.ckpt"
    torch.save({"hyper_parameters": hp, "state_dict": sd}, p)
    enc = load_satclip_location_encoder(str(p))
    assert enc["L"] == 40 and enc["dims"] == [1600, 64, 64, 256]
    assert all(w.dtype == torch.float64 for w, _ in enc["weights"])
    hp2 = dict(hp)
    hp2.pop("eval_downstream")                 # the reference pops these keys unconditionally (load.py:5-7)
    torch.save({"hyper_parameters": hp2, "state_dict": sd}, p)
    with pytest.raises(KeyError):
        load_satclip_location_encoder(str(p))
    hp3 = dict(hp, harmonics_calculation="closed-form")
    torch.save({"hyper_parameters": hp3, "state_dict": sd}, p)
    assert load_satclip_location_encoder(str(p))["harmonics_calculation"] == "closed-form"
    hp4 = dict(hp, harmonics_calculation="discretized")
    torch.save({"hyper_parameters": hp4, "state_dict": sd}, p)
    with pytest.raises(NotImplementedError):
        load_satclip_location_encoder(str(p))


def test_database_preparation_matches_reference_rules():
    import numpy as np
    from oracle import range_oracle as O
    from range_b200.database import prepare_reference_arrays
    db = O.synthetic_db(300, seed=9)
    K, V, xyz = prepare_reference_arrays(db)
    Kr, Vr, xr = O.prepare_db(db["locs"], db["satclip_embeddings"], db["image_embeddings"])
    assert K.dtype == np.float32 and xyz.dtype == np.float32
    assert np.array_equal(K, Kr) and np.array_equal(V, Vr) and np.array_equal(xyz, xr)
    assert np.allclose(np.linalg.norm(K, axis=1), 1, atol=1e-6)


def test_apply_work_plan_covers_every_query_tile_pair_once():
    """host-side decomposition of the producer/consumer apply kernel (csrc/retrieval_pc.cu: pc_plan / pc_work):
    every query-tile pair is processed for every database tile exactly once, by one unit per round"""
    import ctypes
    from range_b200 import _lib
    lib = _lib.load()
    lib.range_debug_apply_plan.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_int32)]
    lib.range_debug_apply_plan.restype = None
    for sm, N, M in [(148, 100_000, 100_000), (148, 6_144, 50_000), (148, 24_576, 3_005), (148, 12_337, 77),
                     (148, 7_900, 1_000_000), (132, 100_000, 100_000), (148, 300, 5_000), (148, 1 << 20, 10_000_000)]:
        out = (ctypes.c_int32 * 7)()
        lib.range_debug_apply_plan(sm, N, M, out)
        units, full, tail_pairs, split, tail_tiles, row0, windows = list(out)
        T = -(-M // 128)
        qp = (-(-N // 128) + 1) // 2
        assert units == (sm // 2) // 3 and full * units + tail_pairs == qp and 0 <= tail_pairs < units
        assert row0 == full * units * 256
        assert 1 <= split <= 4 and (tail_pairs == 0 or tail_pairs * split <= units)
        assert split == 1 or T // split >= 32                      # a split range keeps at least 32 tiles
        covered = {}
        for unit in range(units):                                   # the device-side pc_work(), restated
            for r in range(full + (1 if unit < tail_pairs * split else 0)):
                if r < full:
                    pair, t0, t1 = r * units + unit, 0, T
                else:
                    s = unit % split
                    pair, t0, t1 = full * units + unit // split, s * tail_tiles, min(T, (s + 1) * tail_tiles)
                assert t0 < t1
                covered.setdefault(pair, []).append((t0, t1))
        assert sorted(covered) == list(range(qp))
        for pair, ranges in covered.items():
            ranges.sort()
            assert ranges[0][0] == 0 and ranges[-1][1] == T and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        longest = full * T + (0 if tail_pairs == 0 else (tail_tiles if split > 1 else T))
        assert windows >= -(-longest // 64)


def test_host_unpack_widens_packed_rows_exactly():
    """range_host_unpack is a HOST entry point (no GPU needed): packed rows (1024 fp32 + 256 fp64 = 6144 B) -> the
    float64 (N,1280) rows of range/range.py:222,240; exact, any thread count, aligned or not, ragged row counts"""
    import numpy as np
    from range_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for N in (0, 1, 255, 256, 1000, 5000):
        f = rng.standard_normal((N, 1024)).astype(np.float32)
        q = rng.standard_normal((N, 256))
        if N:
            f[0, :4] = [0.0, -0.0, np.float32(1e-45), np.float32(3.4e38)]          # zero, signed zero, subnormal, near max
        packed = np.empty((N, 6144), np.uint8)
        packed[:, :4096] = f.view(np.uint8).reshape(N, 4096)
        packed[:, 4096:] = q.view(np.uint8).reshape(N, 2048)
        want = np.concatenate([f.astype(np.float64), q], axis=1)
        for threads in (1, 3, 16):
            for off in (0, 1):                                   # off = 1: result not 32-byte aligned (scalar path)
                buf = np.full(N * 1280 + 1, np.nan)
                out = buf[off:off + N * 1280].reshape(N, 1280)
                assert lib.range_host_unpack(packed.ctypes.data if N else None, N, out.ctypes.data if N else None, threads) == 0
                assert np.array_equal(out, want) and np.array_equal(np.signbit(out), np.signbit(want))
    assert lib.range_host_unpack(None, 5, None, 1) == -1


def test_encoder_feature_layout_is_a_permutation_of_the_reference_features():
    """host-side plan of the tensor-core encoder's first-layer input (csrc/capi.cu: plan_sh_layout): every reference feature
    l*l + l +- |m| appears in exactly one column, the other columns are zero columns, the column -> entry map agrees with
    it, and the rounds are ordered by Horner-chain length (what sh_rounds_kernel's table relies on)"""
    import ctypes
    import numpy as np
    from range_b200 import _lib
    from range_b200.sh_table import build_table
    lib = _lib.load()
    I32P = ctypes.POINTER(ctypes.c_int32)
    lib.range_debug_sh_layout.argtypes = [ctypes.c_int, I32P, ctypes.c_int, I32P, I32P, I32P, I32P, ctypes.c_int]
    lib.range_debug_sh_layout.restype = ctypes.c_int
    ZERO = 1 << 25
    for L in (8, 32, 40):
        t = build_table(L)
        off = np.ascontiguousarray(t["off"], np.int32)
        E = L * (L + 1) // 2
        assert off.shape == (E + 1,)
        ent = [(l, am) for am in range(L) for l in range(am, L)]            # the table's |m|-major entry order
        for want_rounds in (1, 0):
            cap = L * (L + 1) + 64
            perm, fmap = np.full(cap, -7, np.int32), np.full(cap, -7, np.int32)
            rounds, tab = ctypes.c_int32(-1), ctypes.c_int32(-1)
            K0 = lib.range_debug_sh_layout(L, off.ctypes.data_as(I32P), want_rounds, ctypes.byref(rounds), ctypes.byref(tab),
                                           perm.ctypes.data_as(I32P), fmap.ctypes.data_as(I32P), cap)
            assert K0 == (64 * -(-E // 32) if want_rounds else L * L) and K0 % 2 == 0
            assert rounds.value == (-(-E // 32) if want_rounds else 0)
            perm, fmap = perm[:K0], fmap[:K0]
            real = perm[perm >= 0]
            assert sorted(real.tolist()) == list(range(L * L))                # each reference feature exactly once
            assert ((perm < 0) == (fmap == ZERO)).all() and (perm >= -1).all()
            for f in np.nonzero(perm >= 0)[0]:
                e, am, is_sin = fmap[f] & 0xffff, (fmap[f] >> 16) & 0xff, (fmap[f] >> 24) & 1
                l, m = ent[e]
                assert m == am and perm[f] == l * l + l + (-am if is_sin else am) and (am > 0 or not is_sin)
            if want_rounds:
                # column 64 r + 2 i (cos) and + 1 (sin) belong to the same chain; chain lengths never grow from round to round
                cos, sin = fmap[0::2], fmap[1::2]
                same = (cos != ZERO) & (sin != ZERO)
                assert ((cos[same] & 0xffffff) == (sin[same] & 0xffffff)).all() and ((sin[same] >> 24) & 1).all()
                assert ((cos == ZERO) <= (sin == ZERO)).all()                  # an unused slot has neither column
                length = np.diff(off)
                per_round = [[length[c & 0xffff] for c in cos[32 * r:32 * r + 32] if c != ZERO] for r in range(rounds.value)]
                assert all(min(a) >= max(b) for a, b in zip(per_round, per_round[1:]) if a and b)
                assert tab.value == 32 * sum(1 + max(p) for p in per_round)
