"""Host-side database layout (range_b200/database.py) on CPU tensors: the reference's preparation rules
(range/range.py:79-95), padding, transposition, power-of-two value scale, spatial row order, shards, disk cache."""
import math

import numpy as np
import torch

from oracle import range_oracle as O


def _db(M=1000, seed=3):
    return O.synthetic_db(M, seed=seed, kind="iid")


def test_layout_matches_reference_preparation():
    from range_b200.database import DeviceDatabase, prepare_reference_arrays
    db = _db(1000)
    K, V, xyz = prepare_reference_arrays(db)
    # same arithmetic as the oracle's restatement of range.py:79-95
    Ko, Vo, xo = O.prepare_db(db["locs"], db["satclip_embeddings"], db["image_embeddings"])
    assert np.array_equal(K, np.asarray(Ko)) and np.array_equal(V, np.asarray(Vo)) and np.array_equal(xyz, np.asarray(xo))
    d = DeviceDatabase(db, "cpu")
    assert (d.M, d.Mpad) == (1000, 1024) and d.Kh.shape == (1024, 256) and d.Vt.shape == (1024, 1024)
    assert d.xyz.shape == (1024, 4) and d.caps.shape == (8, 4)
    assert d.Kh.dtype == torch.float16 and d.Vt.dtype == torch.float16 and d.xyz.dtype == torch.float32
    # rows are the reference's rows in Hilbert order; padding is zero
    order = d.order
    assert sorted(order.tolist()) == list(range(1000))
    assert torch.equal(d.Kh[:1000], torch.from_numpy(K[order]).half())
    assert torch.equal(d.xyz[:1000, :3], torch.from_numpy(np.ascontiguousarray(xyz[order])))
    assert float(d.Kh[1000:].abs().max()) == 0.0 and float(d.Vt[:, 1000:].abs().max()) == 0.0
    assert float(d.xyz[1000:].abs().max()) == 0.0 and float(d.xyz[:, 3].abs().max()) == 0.0
    # values: transposed, scaled by a power of two that keeps fp16 in range
    assert math.log2(d.vscale) == int(math.log2(d.vscale))
    assert torch.equal(d.Vt[:, :1000].t().contiguous(), (torch.from_numpy(V[order]) * d.vscale).half())
    assert 128.0 < float(d.Vt.abs().max()) <= 256.0


def test_unsorted_layout_and_shards_tile_the_database():
    from range_b200.database import DeviceDatabase
    db = _db(777)
    plain = DeviceDatabase(db, "cpu", spatial_sort=False)
    assert plain.order is None and plain.caps is None and plain.Mpad == 896
    full = DeviceDatabase(db, "cpu")
    parts = [DeviceDatabase(db, "cpu", shard=(r, 3)) for r in range(3)]
    assert [p.row_range for p in parts] == [(0, 259), (259, 518), (518, 777)]
    assert all(p.M_total == 777 for p in parts)
    assert torch.equal(torch.cat([p.Kh[:p.M] for p in parts]), full.Kh[:777])
    # every shard scales its values by its own power of two; undo it before comparing
    assert torch.allclose(torch.cat([p.Vt[:, :p.M].float() / p.vscale for p in parts], dim=1),
                          full.Vt[:, :777].float() / full.vscale, rtol=2e-3, atol=0)


def test_cache_round_trip(tmp_path):
    from range_b200.database import DeviceDatabase
    d = DeviceDatabase(_db(300), "cpu")
    path = str(tmp_path / "layout.npz")
    d.save_cache(path)
    e = DeviceDatabase.from_cache(path, "cpu")
    for name in ("Kh", "Vt", "xyz", "caps"):
        assert torch.equal(getattr(d, name), getattr(e, name)), name
    assert (d.M, d.Mpad, d.M_total, d.row_range, d.vscale) == (e.M, e.Mpad, e.M_total, e.row_range, e.vscale)
    assert np.array_equal(d.order, e.order)


def test_forward_chunking_covers_every_row():
    from range_b200.range import LocationEncoder
    for N in (1, 5, 6144, 6145, 12288, 13000, 24576, 24577, 100000, 1 << 20):
        cuts = LocationEncoder._chunks(N, 24576, 6144)
        assert cuts[0][0] == 0 and cuts[-1][1] == N and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        assert all(0 < hi - lo <= 24576 for lo, hi in cuts)
        assert cuts[-1][1] - cuts[-1][0] <= 12288            # the unoverlapped last copy stays short
    for N, chunk, tail in [(64, 24, 24), (64, 24, 5), (7, 3, 3), (100, 1, 1)]:          # degenerate settings still tile [0, N)
        cuts = LocationEncoder._chunks(N, chunk, tail)
        assert cuts[0][0] == 0 and cuts[-1][1] == N and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
