"""BASELINE.json configs 2-5 on however many GPUs torchrun gives (1 GPU: plain python).  Prints one JSON line per
config; device-resident timing (CUDA events, max over ranks), synthetic data, random-init SatCLIP-L40 (H = 512).
  C2  RANGE+ beta=0.5, 100k queries x 100k entries                       (bench.py's workload, here for reference)
  C3  RANGE+ beta in {0, .25, .5, .75, 1}, 1M queries, query-sharded
  C4  database scaling M = 100k .. M_MAX, 100k queries; with > 1 rank the database is sharded along M (NCCL merge)
  ns  one rank's share of the north-star case (1.25 M queries x 1 M entries), CONFIGS=ns
  C5  dense lat/lon raster (visualize_embeddings.py:29-39 coord_grid), RASTER_POINTS points, query-sharded
"""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from argparse import Namespace
from range_b200 import synthetic as O          # seeded input generators (area_uniform, siren_init)
from range_b200.range import LocationEncoder
from range_b200.distributed import shard_rows

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
M_MAX = int(os.environ.get("M_MAX", 1_000_000))
RASTER = int(os.environ.get("RASTER_POINTS", 10_000_000))
CHUNK = 98_304                                    # 16 rounds of the producer/consumer apply kernel
enc = dict(L=40, dims=[1600, 512, 512, 256], weights=O.siren_init(40, 512, 2, 256, seed=0))


def make_db(M, seed=0):
    rng = np.random.default_rng(seed)
    return dict(locs=O.area_uniform(M, rng), satclip_embeddings=rng.standard_normal((M, 256), dtype=np.float32),
                image_embeddings=rng.standard_normal((M, 1024), dtype=np.float32))


def model_for(db, beta, **kw):
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        return LocationEncoder(Namespace(location_model_name="RANGE+", pretrained_path=enc, device=dev, range_db=db, beta=beta, **kw))


def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) * 1e-3


def run_chunks(model, coords, out):
    for lo in range(0, coords.shape[0], CHUNK):
        hi = min(coords.shape[0], lo + CHUNK)
        model.embed(coords[lo:hi], out=out[: hi - lo])


def emit(**kw):
    if rank == 0: print(json.dumps(kw), flush=True)


out = torch.empty(CHUNK, 1280, dtype=torch.float32, device=dev)
db = make_db(100_000)
which = os.environ.get("CONFIGS", "2,3,4,5").split(",")

if "2" in which or "3" in which:
    N = 1_000_000
    lo, hi = shard_rows(N, rank, world)
    coords = torch.tensor(O.area_uniform(N, np.random.default_rng(1))[lo:hi], device=dev)
    for beta in ([0.5] if "3" not in which else [0.0, 0.25, 0.5, 0.75, 1.0]):
        m = model_for(db, beta)
        t = timed(lambda: run_chunks(m, coords, out), reps=1)
        emit(config="C3 beta sweep, one pass per beta", beta=beta, queries=N, M=100_000, n_gpus=world, seconds=t,
             queries_per_s=N / t, parallelism=f"query-sharded x{world}")
    if "3" in which:                  # all five beta in one call: encoder + statistics shared, five apply passes
        betas = [0.0, 0.25, 0.5, 0.75, 1.0]
        def sweep():
            for lo in range(0, coords.shape[0], CHUNK):
                m.embed_sweep(coords[lo:lo + CHUNK], betas)
        t = timed(sweep, reps=1)
        emit(config="C3 beta sweep, embed_sweep (shared statistics)", betas=betas, queries=N, M=100_000, n_gpus=world,
             seconds=t, embeddings_per_s=N * len(betas) / t, parallelism=f"query-sharded x{world}")
    del m

if "4" in which:
    # database scaling: with > 1 rank the database is sharded along M (bench.py: run_m_sharded - every rank owns
    # 100 000 / world of the queries, exp-sums merged by one all-reduce, partial rows stored into the owners' receive
    # buffers over NVLink), each size next to the replicated-database run; one GPU: the unsharded path
    import bench
    sizes = [int(x) for x in os.environ.get("M_SIZES", "").split(",") if x]
    if not sizes:
        sizes = [M for M in (100_000, 300_000, 1_000_000, 3_000_000, 10_000_000) if M <= M_MAX]
    if world > 1:
        def bar():
            dist.barrier(); torch.cuda.synchronize()
        for row in bench.run_m_sharded(world, rank, dev, enc["weights"], 5, 2, bar, sizes):
            emit(config="C4 database scaling, sharded along M", n_gpus=world, **row)
    else:
        from range_b200.database import DeviceDatabase
        N = 100_000
        coords = torch.tensor(O.area_uniform(N, np.random.default_rng(1)), device=dev)
        for M in sizes:
            m = model_for(DeviceDatabase.synthetic(M, dev), 0.5)
            t = timed(lambda: run_chunks(m, coords, out), reps=1)
            emit(config="C4 database scaling", queries=N, M=M, n_gpus=1, seconds=t, queries_per_s=N / t,
                 pair_rate_per_s=N * M / t, parallelism="one GPU")
            del m

if "4big" in which:
    # M = 10 M entries on ONE GPU (25.7 GB of fp16 keys / values resident): database generated on the device
    from range_b200.database import DeviceDatabase
    from range_b200.engine import RangeEngine
    Mbig = int(os.environ.get("M_BIG", 10_000_000))
    N = 100_000
    t0 = time.time()
    eng = RangeEngine(dev, encoder=enc, database=DeviceDatabase.synthetic(Mbig, dev))
    t_build = time.time() - t0
    coords = torch.tensor(O.area_uniform(N, np.random.default_rng(1)), device=dev)
    def big():
        c, perm = eng.sort_queries(coords)
        q64, q16, qxyz = eng.encode(c)
        sums, maxs = eng.retrieve_stats("RANGE+", q16, qxyz, 12.0, 40.0)
        return eng.retrieve_apply_concat("RANGE+", q16, qxyz, 12.0, 40.0, 0.5, sums, maxs, q64, dtype=torch.float32, perm=perm)
    t = timed(big, reps=1)
    res = big()
    emit(config="C4 database scaling, 10 M entries on one GPU", queries=N, M=Mbig, n_gpus=1, seconds=t, queries_per_s=N / t,
         pair_rate_per_s=N * Mbig / t, database_build_s=t_build, database_bytes=eng.db.nbytes(),
         finite=bool(torch.isfinite(res).all()), unit_norm=float((res[:, 1024:].double().norm(dim=1) - 1).abs().max()))
    del eng, res

if "ns" in which:
    # One rank's share of the north-star case (10 M queries x 1 M entries, query-sharded over 8 GPUs: 1.25 M queries per
    # rank against the whole database, no collective).  NS_QUERIES / NS_M override the sizes.
    from range_b200.database import DeviceDatabase
    from range_b200.engine import RangeEngine
    Mns, Nns = int(os.environ.get("NS_M", 1_000_000)), int(os.environ.get("NS_QUERIES", 1_250_000))
    eng = RangeEngine(dev, encoder=enc, database=DeviceDatabase.synthetic(Mns, dev))
    coords = torch.tensor(O.area_uniform(Nns, np.random.default_rng(1 + rank)), device=dev)
    def share(n=Nns):
        for lo in range(0, n, CHUNK):
            c, perm = eng.sort_queries(coords[lo:lo + CHUNK])
            q64, q16, qxyz = eng.encode(c)
            sums, maxs = eng.retrieve_stats("RANGE+", q16, qxyz, 12.0, 40.0)
            eng.retrieve_apply_concat("RANGE+", q16, qxyz, 12.0, 40.0, 0.5, sums, maxs, q64, dtype=torch.float32, perm=perm,
                                      out=out[: c.shape[0]])
    share(CHUNK); torch.cuda.synchronize()          # warm-up on one chunk (timed() would repeat the whole pass)
    if world > 1: dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); share(); b.record(); torch.cuda.synchronize()
    tt = torch.tensor([a.elapsed_time(b)], device=dev)
    if world > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t = float(tt[0]) * 1e-3
    emit(config="north-star share: 10 M queries x 1 M entries over 8 GPUs = 1.25 M queries per rank", queries_per_rank=Nns,
         M=Mns, n_gpus=world, seconds=t, queries_per_s=Nns * world / t, pair_rate_per_s=Nns * world * Mns / t,
         tensor_roofline_frac=2566.0 * Nns * Mns / t / 1364.5e12, total_queries=Nns * world, parallelism=f"query-sharded x{world}, database replicated")
    del eng

if "5" in which:
    # coord_grid (visualize_embeddings.py:29-39): lon = linspace(-180, 180, W), lat = linspace(90, -90, H) stored as float32;
    # raster points sharded over the ranks, the harmonics evaluated separably (embed_raster); RASTER_API=0: the
    # materialised coordinate list through embed()
    H = int(round((RASTER / 2) ** 0.5)); W = 2 * H
    lon_axis, lat_axis = LocationEncoder.coord_grid_axes((H, W))
    lo, hi = shard_rows(H * W, rank, world)
    m = model_for(db, 0.5)
    use_raster = os.environ.get("RASTER_API", "1") == "1"
    tables = m.raster_tables(lon_axis, lat_axis)
    if use_raster:
        def raster():
            for p0 in range(lo, hi, CHUNK):
                p1 = min(hi, p0 + CHUNK)
                m.embed_raster(lon_axis, lat_axis, rows=(p0, p1), out=out[: p1 - p0], tables=tables)
        t = timed(raster, reps=1)
    else:
        coords = m._raster_coords(tables, m._raster_ij(W, lo, hi))
        t = timed(lambda: run_chunks(m, coords, out), reps=1)
    emit(config="C5 dense raster", H=H, W=W, queries=H * W, M=100_000, n_gpus=world, seconds=t,
         queries_per_s=H * W / t, parallelism=f"query-sharded x{world}", encoder="raster (separable harmonics)" if use_raster else "per point")
if "5file" in which:
    # config 5 end to end TO A FILE: every rank streams its slab of the raster into its own memory-mapped .npy
    # (forward_raster(rows=, out=): packed rows over PCIe, widened by host threads into the mapped pages).
    # RASTER_FILE_POINTS points in total (default 2 M = 20 GB of float64 rows), files under RASTER_DIR (default /dev/shm).
    import tempfile
    total = int(os.environ.get("RASTER_FILE_POINTS", 2_000_000))
    H = int(round((total / 2) ** 0.5)); W = 2 * H
    lon_axis, lat_axis = LocationEncoder.coord_grid_axes((H, W))
    lo, hi = shard_rows(H * W, rank, world)
    m = model_for(db, 0.5)
    d = tempfile.mkdtemp(prefix="range_raster_", dir=os.environ.get("RASTER_DIR", "/dev/shm"))
    path = os.path.join(d, f"emb_rank{rank}.npy")
    m.forward_raster(lon_axis, lat_axis, rows=(lo, min(hi, lo + 30_000)), out=np.empty((min(hi, lo + 30_000) - lo, 1280)))   # warm-up
    mm = np.lib.format.open_memmap(path, mode="w+", dtype=np.float64, shape=(hi - lo, 1280))
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    m.forward_raster(lon_axis, lat_axis, rows=(lo, hi), out=mm)
    mm.flush()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1: dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    ok = bool(np.isfinite(mm[:: max(1, (hi - lo) // 1000)]).all())
    del mm
    os.remove(path); os.rmdir(d)
    emit(config="C5 dense raster, end to end into memory-mapped .npy files (one per rank)", H=H, W=W, queries=H * W, M=100_000,
         n_gpus=world, seconds=float(dt[0]), queries_per_s=H * W / float(dt[0]), bytes_written=H * W * 10240, finite=ok,
         host_threads_per_rank=m.host_threads, parallelism=f"query-sharded x{world}")
if world > 1:
    dist.destroy_process_group()
