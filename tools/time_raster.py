"""Raster encoder (csrc/encoder_raster.cu) against the per-point encoder on a coord_grid raster: encoder alone and the
whole device-resident path (RANGE+ beta = 0.5, 100 000-entry database), CUDA events, min of 3."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from argparse import Namespace
from range_b200 import synthetic as O          # seeded input generators
from range_b200.utils import rad_to_cart
from range_b200.range import LocationEncoder
dev = "cuda:0"
H = int(os.environ.get("RASTER_H", 384)); W = 2 * H          # 294 912 points = 3 chunks of 98 304
M = int(os.environ.get("M", 100_000))
rng = np.random.default_rng(0)
db = dict(locs=O.area_uniform(M, rng), satclip_embeddings=rng.standard_normal((M, 256), dtype=np.float32),
          image_embeddings=rng.standard_normal((M, 1024), dtype=np.float32))
enc = dict(L=40, dims=[1600, 512, 512, 256], weights=O.siren_init(40, 512, 2, 256, seed=0))
m = LocationEncoder(Namespace(location_model_name="RANGE+", pretrained_path=enc, device=dev, range_db=db, beta=0.5))
eng = m.engine
lon, lat = LocationEncoder.coord_grid_axes((H, W))
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
t_tab = timeit(lambda: eng.raster_tables(lon, lat))
tables = eng.raster_tables(lon, lat)
N = H * W
ij = m._raster_ij(W, 0, N)
coords = m._raster_coords(tables, ij)
t_pt = timeit(lambda: eng.encode(coords))
t_ra = timeit(lambda: eng.encode_raster(tables, ij))
same = torch.equal(eng.encode(coords)[0], eng.encode_raster(tables, ij)[1])
CH = 98_304
out = torch.empty(CH, 1280, dtype=torch.float32, device=dev)
def per_point():
    for lo in range(0, N, CH): m.embed(coords[lo:lo + CH], out=out[: min(CH, N - lo)])
def raster():
    for lo in range(0, N, CH): m.embed_raster(lon, lat, rows=(lo, min(N, lo + CH)), out=out[: min(CH, N - lo)], tables=tables)
t_e, t_r = timeit(per_point), timeit(raster)
print(f"raster {H} x {W} = {N} points: tables {t_tab:.3f} ms | encoder per-point {t_pt:.2f} ms, raster {t_ra:.2f} ms "
      f"(bit-identical: {same}) | whole path per-point {t_e:.1f} ms, raster {t_r:.1f} ms ({N / t_r * 1e3 / 1e6:.2f} M points/s)")
