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


def test_cache_is_validated(tmp_path):
    """a cache is used only for the source, shard and row order it was built for (one file per shard, '.npz' or not)"""
    import os
    from range_b200.database import DeviceDatabase
    db, other = _db(300), _db(301)
    fp = DeviceDatabase.source_fingerprint(db)
    assert fp == DeviceDatabase.source_fingerprint(db) != DeviceDatabase.source_fingerprint(other)
    base = str(tmp_path / "layout")                    # no extension: save and lookup must still agree
    assert DeviceDatabase.from_cache(base, "cpu") is None            # nothing there yet
    whole = DeviceDatabase(db, "cpu")
    written = whole.save_cache(base, fingerprint=fp)
    assert written == base + ".npz" and os.path.exists(written)
    assert [f for f in os.listdir(tmp_path) if "tmp" in f] == []    # the temporary file is gone
    assert DeviceDatabase.from_cache(base, "cpu", fingerprint=fp) is not None
    assert DeviceDatabase.from_cache(base + ".npz", "cpu", fingerprint=fp) is not None
    assert DeviceDatabase.from_cache(base, "cpu", fingerprint=DeviceDatabase.source_fingerprint(other)) is None   # stale
    assert DeviceDatabase.from_cache(base, "cpu", fingerprint=fp, spatial_sort=False) is None
    # shards: their own files, never the whole database's or another rank's
    assert DeviceDatabase.from_cache(base, "cpu", shard=(1, 2), fingerprint=fp) is None
    parts = [DeviceDatabase(db, "cpu", shard=(r, 2)) for r in range(2)]
    files = [p.save_cache(base, fingerprint=fp) for p in parts]
    assert len(set(files + [written])) == 3
    for r in range(2):
        e = DeviceDatabase.from_cache(base, "cpu", shard=(r, 2), fingerprint=fp)
        assert e.row_range == parts[r].row_range and torch.equal(e.Kh, parts[r].Kh) and e.shard == (r, 2)
    os.replace(files[0], files[1])                      # rank 1 finds rank 0's rows under its name: refused
    assert DeviceDatabase.from_cache(base, "cpu", shard=(1, 2), fingerprint=fp) is None
    # file sources: path + size + mtime
    f = str(tmp_path / "db.npz")
    np.savez(f, **db)
    fp1 = DeviceDatabase.source_fingerprint(f)
    np.savez(f, **other)
    assert DeviceDatabase.source_fingerprint(f) != fp1


def test_synthetic_shards_tile_the_database():
    from range_b200.database import DeviceDatabase
    whole = DeviceDatabase.synthetic(1000, "cpu", seed=3)
    parts = [DeviceDatabase.synthetic(1000, "cpu", seed=3, shard=(r, 4)) for r in range(4)]
    assert [p.row_range for p in parts] == [(0, 250), (250, 500), (500, 750), (750, 1000)]
    assert all(p.M_total == 1000 and p.vscale == whole.vscale for p in parts)
    assert torch.equal(torch.cat([p.xyz[:p.M] for p in parts]), whole.xyz[:1000])       # same locations, same order


def test_forward_chunking_covers_every_row():
    from range_b200.range import LocationEncoder
    for N in (1, 5, 6144, 6145, 12288, 13000, 24576, 24577, 100000, 1 << 20):
        for taper in (0.5, 0.6):
            cuts = LocationEncoder._chunks(N, 24576, 6144, taper)
            assert cuts[0][0] == 0 and cuts[-1][1] == N and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert all(0 < hi - lo <= 24576 + 6143 for lo, hi in cuts) and all(hi - lo <= 24576 for lo, hi in cuts[1:])
            assert cuts[-1][1] - cuts[-1][0] <= 6144 + 3072          # the unoverlapped last copy stays short
            if N >= 24576 + 6144:                                    # ragged rows travel with the first piece: whole rounds after it
                assert all((hi - lo) % 6144 == 0 for lo, hi in cuts[1:])
            sizes = [hi - lo for lo, hi in cuts]
            assert all(a >= b or b <= 6144 + 3072 for a, b in zip(sizes, sizes[1:]))     # pieces shrink towards the end
    cuts = LocationEncoder._chunks(100_000, 49152, 2048)
    assert [hi - lo for lo, hi in cuts] == [49152, 24576, 12288, 6144, 3840, 2048, 1952]
    for N, chunk, tail in [(64, 24, 24), (64, 24, 5), (7, 3, 3), (100, 1, 1)]:          # degenerate settings still tile [0, N)
        cuts = LocationEncoder._chunks(N, chunk, tail)
        assert cuts[0][0] == 0 and cuts[-1][1] == N and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        assert all(0 < hi - lo <= max(chunk, tail + tail // 2) for lo, hi in cuts)


def test_open_npz_memory_maps_the_reference_file_format(tmp_path):
    """generate_db.py:209-214 writes the database with np.savez (float64, uncompressed): open_npz maps the members in
    place and DeviceDatabase builds bit for bit the layout it builds from np.load's arrays"""
    from range_b200.database import DeviceDatabase, open_npz
    db = {k: np.asarray(v, dtype=np.float64) for k, v in _db(700).items()}
    path = str(tmp_path / "range_db.npz")
    np.savez(path, locs=db["locs"], image_embeddings=db["image_embeddings"], satclip_embeddings=db["satclip_embeddings"])
    z = open_npz(path)
    assert set(z) == {"locs", "image_embeddings", "satclip_embeddings"}
    assert all(isinstance(v, np.memmap) for v in z.values())               # nothing was read
    ref = np.load(path)
    for k in z:
        assert z[k].dtype == ref[k].dtype and np.array_equal(np.asarray(z[k]), ref[k])
    for kw in (dict(), dict(shard=(1, 3)), dict(spatial_sort=False), dict(spatial_sort=False, shard=(2, 3))):
        a, b = DeviceDatabase(z, "cpu", **kw), DeviceDatabase(ref, "cpu", **kw)
        assert a.vscale == b.vscale and a.row_range == b.row_range
        for name in ("Kh", "Vt", "xyz"):
            assert torch.equal(getattr(a, name), getattr(b, name)), (name, kw)
    # compressed files and odd members take np.load's path
    cpath = str(tmp_path / "compressed.npz")
    np.savez_compressed(cpath, locs=db["locs"], image_embeddings=db["image_embeddings"],
                        satclip_embeddings=db["satclip_embeddings"], note=np.array("x"), empty=np.zeros((0, 3)))
    c = open_npz(cpath)
    assert not any(isinstance(v, np.memmap) for v in c.values())
    assert np.array_equal(c["image_embeddings"], db["image_embeddings"]) and c["empty"].shape == (0, 3)
    upath = str(tmp_path / "mixed.npz")
    np.savez(upath, locs=db["locs"], empty=np.zeros((0, 3)), scalar=np.float64(2.5))
    u = open_npz(upath)
    assert isinstance(u["locs"], np.memmap) and u["empty"].shape == (0, 3) and float(u["scalar"]) == 2.5


def test_block_wise_value_conversion_matches_the_reference_rule():
    """range.py:90 converts the whole value matrix to fp32 at once; the block-wise conversion gives the same scale and
    the same fp16 layout, including a NaN value (scale falls back to 1 like before)"""
    from range_b200.database import DeviceDatabase, _abs_max_fp32, prepare_reference_arrays
    db = _db(500)
    _, V, _ = prepare_reference_arrays(db)
    assert _abs_max_fp32(db["image_embeddings"], step=64) == float(np.abs(V).max())
    d = DeviceDatabase(db, "cpu")
    assert torch.equal(d.Vt[:, :500].t().contiguous(), (torch.from_numpy(V[d.order]) * d.vscale).half())
    bad = dict(db)
    bad["image_embeddings"] = np.array(db["image_embeddings"], dtype=np.float64)
    bad["image_embeddings"][17, 5] = np.nan
    assert math.isnan(_abs_max_fp32(bad["image_embeddings"], step=64))
    assert DeviceDatabase(bad, "cpu").vscale == 1.0


def test_layout_does_not_depend_on_the_staging_block(monkeypatch):
    """keys and values reach the device in blocks of STAGING_ROWS rows; row-wise arithmetic, so any block size gives
    the same bits (here: 1000 rows in blocks of 96 against one block)"""
    from range_b200.database import DeviceDatabase
    db = _db(1000)
    one = DeviceDatabase(db, "cpu")
    monkeypatch.setattr(DeviceDatabase, "STAGING_ROWS", 96)
    for kw in (dict(), dict(shard=(1, 2)), dict(spatial_sort=False)):
        monkeypatch.setattr(DeviceDatabase, "STAGING_ROWS", 1 << 18)
        ref = DeviceDatabase(db, "cpu", **kw)
        monkeypatch.setattr(DeviceDatabase, "STAGING_ROWS", 96)
        blk = DeviceDatabase(db, "cpu", **kw)
        for name in ("Kh", "Vt", "xyz"):
            assert torch.equal(getattr(ref, name), getattr(blk, name)), (name, kw)
    assert one.M == 1000


def test_inconsistent_database_is_rejected():
    from range_b200.database import DeviceDatabase
    import pytest
    db = dict(_db(50))
    db["image_embeddings"] = db["image_embeddings"][:40]
    with pytest.raises(ValueError):
        DeviceDatabase(db, "cpu")
    db = dict(_db(50))
    db["satclip_embeddings"] = db["satclip_embeddings"][:, :128]
    with pytest.raises(ValueError):
        DeviceDatabase(db, "cpu")


def test_forward_buffers_fit_every_piece():
    """the staging buffers of model(locs) are sized for the largest piece of any super-batch (tail > chunk / 2: a short
    last super-batch ends in a longer piece than the first one does)"""
    from range_b200.range import LocationEncoder
    for N, chunk, tail, sb in [(16384 + 12288, 8192, 6144, 16384), (100000, 24576, 6144, 1 << 20), (5, 24576, 6144, 1 << 20),
                               (3 * 49152 + 11111, 24576, 20000, 49152), (1000, 64, 60, 128)]:
        batches, plan, rows = LocationEncoder._pieces(N, min(chunk, N), tail, sb)
        assert batches[0][0] == 0 and batches[-1][1] == N and all(a[1] == b[0] for a, b in zip(batches, batches[1:]))
        for (s0, s1), cuts in zip(batches, plan):
            assert cuts[0][0] == 0 and cuts[-1][1] == s1 - s0 and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert all(0 < hi - lo <= rows for lo, hi in cuts)
    # buffers must cover the largest piece of ANY super-batch, not of the first one: search the small settings for a
    # case where a later (shorter) super-batch ends in a longer piece than the first, and check it is covered
    found = 0
    for N in range(200, 1200, 7):
        for chunk, tail, sb in [(64, 60, 128), (96, 90, 192), (50, 45, 100)]:
            _, plan, rows = LocationEncoder._pieces(N, chunk, tail, sb)
            assert rows == max(hi - lo for cuts in plan for lo, hi in cuts)
            found += max(hi - lo for lo, hi in plan[0]) < rows
    assert found > 0
