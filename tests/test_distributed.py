"""Host-side logic of the N > 1 paths on CPU (gloo, world_size 2): row sharding, and the SUM / MAX merge
rule of an M-sharded database, checked against the unsharded softmax on the oracle's arithmetic."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from range_b200.distributed import merge_outputs, merge_stats, shard_rows
    rng = np.random.default_rng(0)
    N, M, D = 37, 1001, 16
    s = torch.tensor(rng.uniform(-1, 1, (N, M)))
    g = torch.tensor(rng.uniform(-1, 1, (N, M)))
    V = torch.tensor(rng.standard_normal((M, D)))
    beta, ts, tg = 0.3, 12.0, 40.0
    lo, hi = shard_rows(M, rank, world)
    # what range_retrieve_stats returns for this shard: fixed offset -1, no running max
    sums = torch.stack([torch.exp(ts * (s[:, lo:hi] - 1)).sum(1), torch.exp(tg * (g[:, lo:hi] - 1)).sum(1)], 1)
    maxs = torch.stack([s[:, lo:hi].max(1).values, g[:, lo:hi].max(1).values], 1)
    merge_stats(sums, maxs)
    # what range_retrieve_apply returns: this shard's contribution normalised by the GLOBAL sums
    P = beta * torch.exp(ts * (s[:, lo:hi] - 1)) / sums[:, :1] + (1 - beta) * torch.exp(tg * (g[:, lo:hi] - 1)) / sums[:, 1:]
    O = merge_outputs(P @ V[lo:hi])
    ref = (beta * torch.softmax(ts * s, 1) + (1 - beta) * torch.softmax(tg * g, 1)) @ V
    ok = torch.allclose(O, ref, rtol=1e-10, atol=1e-12) and torch.equal(maxs[:, 0], s.max(1).values)
    # query sharding covers every row exactly once
    slabs = [shard_rows(N, r, world) for r in range(world)]
    ok = ok and slabs[0][0] == 0 and slabs[-1][1] == N and all(a[1] == b[0] for a, b in zip(slabs, slabs[1:]))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_m_sharded_merge_and_query_sharding_gloo():
    world = 2
    port = 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))
