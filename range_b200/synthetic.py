"""Seeded synthetic inputs at the reference's shapes (no network: random-init SatCLIP-L40, synthetic database;
BASELINE.json / SURVEY.md 8d).  Used by bench.py, __graft_entry__.smoke() and tools/; the oracle keeps its own copy
(tests/test_oracle.py checks that the two generate the same numbers)."""
import math

import numpy as np
import torch


def area_uniform(n, rng):
    """(n,2) fp64 (lon, lat) degrees, uniform on the sphere"""
    lon = rng.uniform(-180, 180, n)
    lat = np.degrees(np.arcsin(rng.uniform(-1, 1, n)))
    return np.stack([lon, lat], 1)


def siren_init(L=40, H=512, n_hidden=2, out_dim=256, seed=0):
    """random-init SIREN with the reference's own init rule (location_encoder.py:137-144): first layer U(+-1/in),
    the others U(+-sqrt(6/in)/w0) with w0 = 1"""
    g = torch.Generator().manual_seed(seed)
    dims = [L * L] + [H] * n_hidden + [out_dim]
    ws = []
    for i in range(len(dims) - 1):
        din, dout = dims[i], dims[i + 1]
        std = (1.0 / din) if i == 0 else math.sqrt(6.0 / din) / 1.0
        W = (torch.rand(dout, din, generator=g, dtype=torch.float64) * 2 - 1) * std
        b = (torch.rand(dout, generator=g, dtype=torch.float64) * 2 - 1) * std
        ws.append((W, b))
    return ws


def iid_database(M, seed=0):
    """area-uniform locations, N(0,1) keys and values in fp32 (flat softmax: the worst case for fp16 operands)"""
    rng = np.random.default_rng(seed)
    return dict(locs=area_uniform(M, rng), satclip_embeddings=rng.standard_normal((M, 256), dtype=np.float32),
                image_embeddings=rng.standard_normal((M, 1024), dtype=np.float32))
