"""RANGE database: the reference's preparation (range/range.py:78-100) and the device-resident layout the
retrieval kernels stream.

Device layout (what include/range_b200.h:range_ctx_set_db takes):
  Kh  (Mpad, 256)  fp16  row-normalised SatCLIP keys, row-major  -> TMA boxes [128 entries x 64 dims]
  Vt  (1024, Mpad) fp16  values TRANSPOSED (entries contiguous) x vscale -> TMA boxes [256 dims x 64 entries];
                         K-major B operand of the P.V tensor-core product, same descriptor as the keys
  xyz (Mpad, 4)    fp32  unit vectors of the entry locations (x, y, z, 0)
Mpad = M rounded up to 128; padding is zero and masked in-kernel.
"""
import math

import numpy as np
import torch

from .utils import rad_to_cart

BLOCK = 128


def prepare_reference_arrays(db):
    """Exactly range/range.py:79-95, on the host in numpy fp32 (bit-identical to the reference's tensors)."""
    locs = np.asarray(db["locs"]).astype(np.float32)                                  # :79
    K = np.asarray(db["satclip_embeddings"]).astype(np.float32)                       # :85
    K = K / np.linalg.norm(K, ord=2, axis=1, keepdims=True)                           # :89
    V = np.asarray(db["image_embeddings"]).astype(np.float32)                         # :90
    xyz = rad_to_cart(locs * math.pi / 180)                                           # :93-95 (fp32)
    return K, V, xyz


class DeviceDatabase:
    def __init__(self, db, device, shard=None):
        """db: mapping with locs / satclip_embeddings / image_embeddings (an opened .npz works).
        shard=(rank, world): keep rows [rank*M/world, (rank+1)*M/world) only (M-sharding)."""
        K, V, xyz = prepare_reference_arrays(db)
        self.M_total = K.shape[0]
        if shard is not None:
            r, w = shard
            lo, hi = (self.M_total * r) // w, (self.M_total * (r + 1)) // w
            K, V, xyz = K[lo:hi], V[lo:hi], xyz[lo:hi]
            self.row_range = (lo, hi)
        else:
            self.row_range = (0, self.M_total)
        if K.shape[1] != 256 or V.shape[1] != 1024:
            raise ValueError(f"RANGE database must have 256-d keys and 1024-d values, got {K.shape[1]}, {V.shape[1]}")
        M = K.shape[0]
        if M == 0:
            raise ValueError("empty database (shard)")
        Mpad = (M + BLOCK - 1) // BLOCK * BLOCK
        dev = torch.device(device)
        self.M, self.Mpad = M, Mpad
        self.Kh = torch.zeros(Mpad, 256, dtype=torch.float16, device=dev)
        self.Kh[:M] = torch.from_numpy(K).to(dev).half()
        vmax = float(np.abs(V).max())
        self.vscale = 1.0 if vmax == 0.0 or not math.isfinite(vmax) else 2.0 ** math.floor(math.log2(256.0 / vmax))
        self.Vt = torch.zeros(1024, Mpad, dtype=torch.float16, device=dev)
        step = 1 << 18
        for lo in range(0, M, step):                       # bounded staging memory for 10M-entry databases
            hi = min(M, lo + step)
            self.Vt[:, lo:hi] = (torch.from_numpy(V[lo:hi]).to(dev) * self.vscale).half().t()
        self.xyz = torch.zeros(Mpad, 4, dtype=torch.float32, device=dev)
        self.xyz[:M, :3] = torch.from_numpy(np.ascontiguousarray(xyz)).to(dev)

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.Kh, self.Vt, self.xyz))
