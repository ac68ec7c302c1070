"""Host-side logic of the N > 1 paths on CPU (gloo, world_size 2): row sharding, and the M-sharded database pipeline
of range_b200/distributed.py (own-slab encoding, all-gather of the compact queries, SUM merge of the exp-sums with
local maxima, per-owner merge of the partial outputs, ragged query counts, several steps) driven through a torch
stand-in for the engine and checked against the unsharded softmax of range/range.py:213-238."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

TS, TG, BETA, D = 12.0, 40.0, 0.3, 1024


def _database(M):
    g = torch.Generator().manual_seed(5)
    K = torch.nn.functional.normalize(torch.randn(M, 256, generator=g, dtype=torch.float64), dim=1)
    V = torch.randn(M, D, generator=g, dtype=torch.float64)
    X = torch.nn.functional.normalize(torch.randn(M, 3, generator=g, dtype=torch.float64), dim=1)
    return K, V, X


def _encode(coords):
    """any deterministic row-wise map (the real encoder is tested on the GPU): (n,2) -> unit (n,256), unit (n,3)"""
    f = torch.arange(1, 257, dtype=torch.float64)
    e = torch.sin(coords[:, :1] * f * 0.013) + torch.cos(coords[:, 1:] * f * 0.007)
    lon, lat = torch.deg2rad(coords[:, 0]), torch.deg2rad(coords[:, 1])
    xyz = torch.stack([torch.cos(lat) * torch.cos(lon), torch.cos(lat) * torch.sin(lon), torch.sin(lat)], 1)
    return torch.nn.functional.normalize(e, dim=1), xyz


class TorchEngine:
    """what range_b200.engine.RangeEngine offers to ShardedRetriever, in torch on the CPU, over database rows [lo, hi)"""

    def __init__(self, K, V, X):
        self.K, self.V, self.X = K, V, X
        self.device = torch.device("cpu")
        self.lib, self.index = None, 0

    def sort_queries(self, c):
        perm = torch.argsort(c[:, 0], stable=True).to(torch.int32)
        return c[perm.long()].contiguous(), perm

    def encode(self, c):
        q, xyz = _encode(c)
        return q, q.half(), torch.cat([xyz, torch.zeros(len(c), 1, dtype=torch.float64)], 1).float()

    def _logits(self, q16, qxyz):
        return q16.double() @ self.K.t(), qxyz[:, :3].double() @ self.X.t()

    def retrieve_stats(self, mode, q16, qxyz, temp, geo_temp):      # fixed offset -1, no running max
        s, g = self._logits(q16, qxyz)
        sums = torch.stack([torch.exp(temp * (s - 1)).sum(1), torch.exp(geo_temp * (g - 1)).sum(1)], 1)
        maxs = torch.stack([s.max(1).values, g.max(1).values], 1)
        return sums, maxs

    def retrieve_apply(self, mode, q16, qxyz, temp, geo_temp, beta, sums, maxs):   # normalised by the GLOBAL sums
        s, g = self._logits(q16, qxyz)
        P = beta * torch.exp(temp * (s - 1)) / sums[:, :1] + (1 - beta) * torch.exp(geo_temp * (g - 1)) / sums[:, 1:]
        return (P @ self.V).contiguous()

    def combine_concat(self, parts, weights, q64, out=None, dtype=torch.float64, perm=None):
        O = sum(p for p in parts)
        rows = torch.cat([O, q64], 1)
        if perm is None:
            out.copy_(rows)
        else:
            out[perm.long()] = rows.to(out.dtype)
        return out


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from range_b200.distributed import ShardedRetriever, shard_rows
    M = 1001
    K, V, X = _database(M)
    lo, hi = shard_rows(M, rank, world)
    sr = ShardedRetriever(TorchEngine(K[lo:hi], V[lo:hi], X[lo:hi]), merge="reduce_scatter")
    sr.MAX_STEP_ROWS = 256 * world                      # slab 256: rank 0 takes 2 steps, rank 1 takes 1 and pads
    n = [300, 77][rank]
    rng = np.random.default_rng(10 + rank)
    coords = torch.tensor(np.stack([rng.uniform(-180, 180, n), rng.uniform(-90, 90, n)], 1))
    ok = True
    for sort in (False, True):
        got = sr.embed("RANGE+", coords, TS, TG, BETA, sort, out_dtype=torch.float64)
        q, xyz = _encode(coords)
        s, g = q.half().double() @ K.t(), xyz.float().double() @ X.t()
        ref = (BETA * torch.softmax(TS * s, 1) + (1 - BETA) * torch.softmax(TG * g, 1)) @ V
        ok = ok and got.shape == (n, 1280) and torch.allclose(got[:, :D], ref, rtol=1e-9, atol=1e-11)
        ok = ok and torch.equal(got[:, D:], q)
    ok = ok and sr.collectives > 0
    empty = sr.embed("RANGE+", coords[:0], TS, TG, BETA, True, out_dtype=torch.float64)     # every rank: no rows
    ok = ok and empty.shape == (0, 1280)
    # query sharding covers every row exactly once
    slabs = [shard_rows(37, r, world) for r in range(world)]
    ok = ok and slabs[0][0] == 0 and slabs[-1][1] == 37 and all(a[1] == b[0] for a, b in zip(slabs, slabs[1:]))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_m_sharded_pipeline_and_query_sharding_gloo():
    world = 2
    port = 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))
