"""time the retrieval kernels at the bench workload (optionally under RANGE_DBG experiments)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from range_b200 import synthetic as O
from range_b200.utils import rad_to_cart
from range_b200.engine import RangeEngine
from range_b200.database import DeviceDatabase
dev = "cuda:0"
M = int(os.environ.get("M", 100_000)); N = int(os.environ.get("N", 100_000))
rng = np.random.default_rng(0)
db = dict(locs=O.area_uniform(M, rng), satclip_embeddings=rng.standard_normal((M, 256), dtype=np.float32),
          image_embeddings=rng.standard_normal((M, 1024), dtype=np.float32))
eng = RangeEngine(dev, L=40, database=DeviceDatabase(db, dev))
q = torch.randn(N, 256, device=dev); q = (q / q.norm(dim=1, keepdim=True)).half()
c = torch.tensor(O.area_uniform(N, np.random.default_rng(1)))
if os.environ.get("SORT", "1") == "1":      # spatially batched queries (what range.py does for RANGE+)
    c = eng.sort_queries(c)[0].cpu()
xyz = torch.zeros(N, 4); xyz[:, :3] = torch.tensor(rad_to_cart(c.numpy() * np.pi / 180)).float(); xyz = xyz.to(dev)
def timeit(fn, reps=int(os.environ.get("REPS", 3))):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
if eng.db.caps is not None:
    print(f"geo-skip fraction: {eng.geo_mask(xyz, 40.0).float().mean().item():.3f}")
sums, maxs = eng.retrieve_stats("RANGE+", q, xyz, 12.0, 40.0)
if eng.db.caps is not None:
    print(f"geo-skip fraction, apply pass (known normalisers): {eng.geo_mask(xyz, 40.0, sums=sums).float().mean().item():.3f}")
t_st = timeit(lambda: eng.retrieve_stats("RANGE+", q, xyz, 12.0, 40.0))
t_ap = timeit(lambda: eng.retrieve_apply("RANGE+", q, xyz, 12.0, 40.0, 0.5, sums, maxs))
s2, m2 = eng.retrieve_stats("RANGE", q, xyz, 15.0, 0.0)
t_st_r = timeit(lambda: eng.retrieve_stats("RANGE", q, xyz, 15.0, 0.0))
t_ap_r = timeit(lambda: eng.retrieve_apply("RANGE", q, xyz, 15.0, 0.0, None, s2, m2))
fl = 2566.0 * N * M
print(f"DBG={os.environ.get('RANGE_DBG','0')} N={N} M={M}: RANGE+ stats {t_st:.2f} apply {t_ap:.2f} ms ({fl/((t_st+t_ap)*1e-3)/1364.5e12:.3f} of peak) | RANGE stats {t_st_r:.2f} apply {t_ap_r:.2f} ms")
if os.environ.get("PROF"):
    import ctypes
    buf = torch.zeros(64, dtype=torch.int64, device=dev)
    eng.lib.range_debug_set_profile_buffer(ctypes.c_void_p(buf.data_ptr()))
    eng.retrieve_apply("RANGE+", q, xyz, 12.0, 40.0, 0.5, sums, maxs); torch.cuda.synchronize()
    eng.lib.range_debug_set_profile_buffer(None)
    b = buf.cpu().numpy().astype(float)
    if os.environ.get("RANGE_APPLY_KERNEL") == "pc":
        T = 17.0 * ((M + 127) // 128)          # rounds of unit 1 x tiles
        f = lambda lo, n: " ".join(f"{x / T:7.0f}" for x in b[lo:lo + n])
        print(f"per-tile cycles, unit 1 leader CTAs (T = {T:.0f} tiles):")
        print(f" producer tma      wait_stage_empty | issue K | wait_xyz_empty : {f(0, 3)}")
        print(f" producer mma      wait_s_empty | wait_stage_full | issue+commit : {f(8, 3)}")
        print(f" (publish+gate merged: not instrumented)")
        print(f" producer softmax  wait_s_full | tmem_ld | compute | wait_slot_free | store+arrive : {f(32, 5)}")
        print(f" consumer Vt load  wait_v_empty (per tile = 2 stages)            : {f(40, 1)}")
        print(f" consumer P load   wait_full flag | proxy fence | wait_p_empty   : {f(48, 3)}")
        print(f" consumer mma      wait_p_full | issue | wait_v_full (per tile)   : {f(56, 3)}   (second consumer pair: {f(24, 3)})")
        sys.exit(0)
    T = max(1, b[22])
    print(f"tiles {T}; per-tile cycles:")
    print(f" producer: wait_empty(K) {b[0]/T:.0f} wait_empty(V) {b[1]/T:.0f} total {b[2]/T:.0f}")
    print(f" mma: wait_stage_qk {b[8]/T:.0f} issue_qk {b[9]/T:.0f} wait_p {b[10]/T:.0f} wait_stage_pv {b[11]/T:.0f} issue_pv {b[12]/T:.0f} total {b[13]/T:.0f}")
    print(f" (128-entry tiles)\n softmax w0: wait_s {b[16]/T:.0f} tmem_ld {b[17]/T:.0f} wait_xyz {b[18]/T:.0f} compute {b[19]/T:.0f} st+arrive {b[20]/T:.0f} total {b[21]/T:.0f}")
