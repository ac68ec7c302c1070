"""Dense lat/lon rasters (BASELINE config 5; the reference's grids: evaluation/visualize_embeddings.py:29-39).

CPU: the raster axes reproduce coord_grid's coordinates exactly.  GPU: the separable raster encoder
(csrc/encoder_raster.cu) is bit-identical to the per-point encoder, and embed_raster / forward_raster return exactly
what embed / forward return on the materialised coordinate list."""
from argparse import Namespace

import numpy as np
import pytest
import torch

from oracle import range_oracle as O

DEV = "cuda:0"


def coord_grid(grid_size):
    """restatement of visualize_embeddings.py:29-39 (float32 array of meshgrid(linspace(-180, 180, W), linspace(90, -90, H)))"""
    feats = np.zeros((grid_size[0], grid_size[1], 2), dtype=np.float32)
    mg = np.meshgrid(np.linspace(-180, 180, feats.shape[1]), np.linspace(90, -90, feats.shape[0]))
    feats[:, :, 0] = mg[0]
    feats[:, :, 1] = mg[1]
    return feats.reshape(feats.shape[0] * feats.shape[1], 2)


@pytest.mark.parametrize("grid", [(1, 1), (2, 3), (45, 90), (180, 361)])
def test_axes_reproduce_coord_grid(grid):
    from range_b200.range import LocationEncoder
    lon, lat = LocationEncoder.coord_grid_axes(grid)
    H, W = grid
    assert lon.dtype == torch.float64 and lon.shape == (W,) and lat.shape == (H,)
    p = np.arange(H * W)
    pts = np.stack((lon.numpy()[p % W], lat.numpy()[p // W]), axis=1)          # point p = i * W + j
    assert np.array_equal(pts, coord_grid(grid).astype(np.float64))


def _model(width, name="RANGE+", **kw):
    from range_b200.range import LocationEncoder
    enc = dict(L=40, dims=[1600, width, width, 256], weights=O.siren_init(40, width, 2, 256, seed=0))
    db = O.synthetic_db(3000, seed=0)
    return LocationEncoder(Namespace(location_model_name=name, pretrained_path=enc, device=DEV, range_db=db, beta=0.5,
                                     chunk=1024, tail=256, **kw))


@pytest.mark.gpu
def test_raster_encoder_is_bit_identical():
    from range_b200.range import LocationEncoder
    model = _model(256)
    eng = model.engine
    assert eng.precision == "f16x3" and eng.raster_supported()
    H, W = 45, 90                                     # includes both poles and the antimeridian twice
    lon, lat = LocationEncoder.coord_grid_axes((H, W))
    coords = torch.from_numpy(coord_grid((H, W))).double()
    tables = eng.raster_tables(lon, lat)
    ij = model._raster_ij(W, 0, H * W)
    ll, q64, q16, qxyz = eng.encode_raster(tables, ij)
    assert torch.equal(ll.cpu(), coords)
    r64, r16, rxyz = eng.encode(coords)
    assert torch.isfinite(q64).all()
    assert torch.equal(q64, r64) and torch.equal(q16, r16) and torch.equal(qxyz, rxyz)
    # index / coordinate rows built on the device (range_raster_points), whole raster and a permuted window of it
    ij2 = torch.empty(H * W, 2, dtype=torch.int32, device=DEV)
    ll2 = torch.empty(H * W, 2, dtype=torch.float64, device=DEV)
    eng.raster_points(tables, 0, H * W, ij=ij2, lonlat=ll2)
    assert torch.equal(ij2, ij) and torch.equal(ll2.cpu(), coords)
    pw = torch.randperm(1500, generator=torch.Generator().manual_seed(1)).to(torch.int32).to(DEV)
    ijw = torch.empty(1500, 2, dtype=torch.int32, device=DEV)
    eng.raster_points(tables, 1000, 1500, perm=pw, ij=ijw)
    assert torch.equal(ijw, ij[1000:2500][pw.long()])
    sel = torch.randperm(H * W, generator=torch.Generator().manual_seed(0))[:1000].to(DEV)      # any subset, any order
    s64 = eng.encode_raster(tables, ij[sel])[1]
    assert torch.equal(s64, r64[sel])
    # the whole path: same rows in, same bits out
    for m in (model, _model(256, name="RANGE")):
        a = m.embed(coords.to(DEV))
        assert torch.equal(m.embed_raster(lon, lat), a)
        part = m.embed_raster(lon, lat, rows=(1000, 2500), tables=m.raster_tables(lon, lat))
        assert torch.equal(part, m.embed(coords[1000:2500].to(DEV)))
        out = m.forward_raster(lon, lat)
        assert isinstance(out, np.ndarray) and out.dtype == np.float64 and out.shape == (H * W, 1280)
        assert np.array_equal(out, m(coords))
        # a slab of the raster into the caller's array (packed rows widened on the host), e.g. a memory-mapped file
        mine = np.full((1500, 1280), np.nan)
        assert m.forward_raster(lon, lat, rows=(1000, 2500), out=mine) is mine
        assert np.array_equal(mine, m(coords[1000:2500]))
    bad = ij.clone()                                  # an index outside the raster gives a NaN row, nothing else changes
    bad[5, 0] = H
    b64 = eng.encode_raster(tables, bad)[1]
    assert torch.isnan(b64[5]).all() and torch.equal(b64[:5], r64[:5]) and torch.equal(b64[6:], r64[6:])


@pytest.mark.gpu
def test_raster_api_without_the_separable_encoder():
    """fp64 encoder (SIREN width 64): embed_raster / forward_raster run the per-point encoder on the coordinates"""
    from range_b200.range import LocationEncoder
    model = _model(64)
    assert not model.engine.raster_supported()
    lon, lat = LocationEncoder.coord_grid_axes((20, 41))
    coords = torch.from_numpy(coord_grid((20, 41))).double()
    assert torch.equal(model.embed_raster(lon, lat), model.embed(coords.to(DEV)))
    assert np.array_equal(model.forward_raster(lon, lat), model(coords))
