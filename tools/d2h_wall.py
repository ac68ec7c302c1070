"""How fast can the ranks of one box hand (N,1280) float64 results to the host?  The ceiling of model(locs)'s e2e rate.
Run under torchrun with one rank per GPU; every phase is measured with k = 1, 2, 4, .. world active ranks at once
(the others idle at the barrier), wall clock between barriers, aggregate over the active ranks:

  copy_f64     cudaMemcpyAsync device -> page-locked host, 1.02 GB (100 000 float64 rows)       -> rows/s, GB/s
  copy_packed  the same rows packed (6 144 B per row: RANGE_OUT_PACKED)                          -> rows/s, GB/s
  unpack       range_host_unpack of 100 000 packed rows into a pageable float64 array (host only, T threads per rank)
  pipeline     copy_packed of chunk i+1 overlapped with unpack of chunk i (what host_path='packed' does)
  direct       kernel stores straight into the page-locked float64 result (range_combine_concat with a host pointer:
               what host_path='direct' does inside the apply kernel's epilogue)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/d2h_wall.py
"""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from range_b200 import _lib                                    # noqa: E402
from range_b200.engine import RangeEngine                      # noqa: E402

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ROWS, REPS = 100_000, int(os.environ.get("REPS", 6))
threads = max(1, (os.cpu_count() or 1) // world)
eng = RangeEngine(dev, L=40)
lib = eng.lib


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def measure(name, fn, bytes_per_rep, active):
    """fn() = one repetition on this rank (must be complete when it returns or after torch.cuda.synchronize())"""
    if rank < active:
        fn(); torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    if rank < active:
        for _ in range(REPS):
            fn()
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt if rank < active else 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    barrier()
    if rank == 0:
        dt = float(t[0])
        print(json.dumps({"phase": name, "active_ranks": active, "rows_per_s": active * ROWS * REPS / dt,
                          "GB_per_s": active * bytes_per_rep * REPS / dt / 1e9, "ms_per_rep": dt / REPS * 1e3,
                          "host_threads_per_rank": threads}), flush=True)


d64 = torch.randn(ROWS, 1280, dtype=torch.float64, device=dev)
h64 = torch.empty(ROWS, 1280, dtype=torch.float64, pin_memory=True)
dpk = torch.randint(0, 255, (ROWS, 6144), dtype=torch.uint8, device=dev)
hpk = [torch.empty(ROWS // 4, 6144, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
hpk_all = torch.empty(ROWS, 6144, dtype=torch.uint8, pin_memory=True)
res = np.empty((ROWS, 1280), np.float64)
O32 = torch.randn(ROWS, 1024, dtype=torch.float32, device=dev)
q64 = torch.randn(ROWS, 256, dtype=torch.float64, device=dev)
copy_stream = torch.cuda.Stream(device=dev)


def pipeline():
    """4 chunks of 25 000 rows: copy chunk i+1 while the host widens chunk i"""
    n = ROWS // 4
    evs = []
    for i in range(4):
        with torch.cuda.stream(copy_stream):
            if i >= 2:
                pass                                   # slot i % 2 was unpacked below before we get here
            hpk[i % 2].copy_(dpk[i * n:(i + 1) * n], non_blocking=True)
            ev = torch.cuda.Event(); ev.record(copy_stream); evs.append(ev)
        if i >= 1:
            evs[i - 1].synchronize()
            lib.range_host_unpack(hpk[(i - 1) % 2].data_ptr(), n, res[(i - 1) * n:i * n].ctypes.data, threads)
    evs[3].synchronize()
    lib.range_host_unpack(hpk[1].data_ptr(), n, res[3 * n:].ctypes.data, threads)


def direct():
    P = (ctypes.c_void_p * 1)(O32.data_ptr())
    _lib.check(lib.range_combine_concat(eng.ctx, ROWS, 1, P, None, ctypes.c_void_p(q64.data_ptr()), None,
                                        ctypes.c_void_p(h64.data_ptr()), _lib.RANGE_OUT_F64,
                                        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))


ks = [k for k in (1, 2, 4, 8) if k <= world]
for k in ks:
    measure("copy_f64", lambda: h64.copy_(d64, non_blocking=True), ROWS * 10240, k)
for k in ks:
    measure("copy_packed", lambda: hpk_all.copy_(dpk, non_blocking=True), ROWS * 6144, k)
for k in ks:
    measure("unpack", lambda: lib.range_host_unpack(hpk_all.data_ptr(), ROWS, res.ctypes.data, threads), ROWS * 16384, k)
for k in ks:
    measure("pipeline", pipeline, ROWS * 6144, k)
for k in ks:
    measure("direct", direct, ROWS * 10240, k)
if world > 1:
    dist.destroy_process_group()
