"""Writers for the embeddings - the B200-side counterpart of the reference's caller range/utils/save.py:7-58
(`save_embeddings`: loop a loader, call `location_model(coords)`, `np.savez(coords=, embeddings=, y=)`).

`save_embeddings` keeps the reference's file format (an .npz with `coords`, `embeddings`, `y`) but embeds all
coordinates of a split in ONE model call, so the model's own pipelining (chunked retrieval overlapped with the
device->host copy) applies.  `embed_to_npy` is for sets too large for host memory (a 10 M-point raster is 102 GB of
float64): the result streams batch by batch into a memory-mapped .npy.
"""
import os

import numpy as np
import torch


def _collect(loader):
    coords, ys = [], []
    for data in loader:
        c, y = data
        coords.append(torch.as_tensor(c).double().cpu())
        ys.append(torch.as_tensor(y).cpu())
    return torch.cat(coords), torch.cat(ys)


def save_embeddings(location_model, train_loader, val_loader, embeddings_dir, task_name, model_name=None):
    """utils/save.py:7-58: writes <dir>/<model>/<task>_train.npz and _val.npz; returns the two paths"""
    name = model_name or getattr(location_model, "location_model_name", "RANGE")
    out_dir = os.path.join(embeddings_dir, name)
    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for split, loader in (("train", train_loader), ("val", val_loader)):
        coords, y = _collect(loader)
        emb = location_model(coords)                    # numpy float64 (N, 1280), like the reference's RANGE branch
        if hasattr(emb, "cpu"):
            emb = emb.cpu().numpy()
        path = os.path.join(out_dir, f"{task_name}_{split}.npz")
        np.savez(path, coords=coords.numpy(), embeddings=emb, y=y.numpy())
        paths.append(path)
    return tuple(paths)


def embed_to_npy(location_model, coords, path, batch=1 << 20):
    """coords (N,2) (lon, lat) degrees (tensor / array / memmap) -> float64 (N, D) .npy at `path`, written through a
    memory map `batch` rows at a time; returns the open memmap.  A range_b200 model fills the map itself
    (LocationEncoder.embed_into): its chunks cross PCIe packed into page-locked staging buffers and host threads widen
    them into the mapped pages while the GPU computes the next chunks - no (N, D) array is ever resident."""
    N = len(coords)
    out = None
    into = getattr(location_model, "embed_into", None)
    dim = getattr(location_model, "location_feature_dim", None)
    if into is not None and dim is not None and N > 0:
        out = np.lib.format.open_memmap(path, mode="w+", dtype=np.float64, shape=(N, dim))
    for lo in range(0, max(N, 1), batch):
        hi = min(N, lo + batch)
        c = torch.as_tensor(np.asarray(coords[lo:hi]), dtype=torch.float64)
        if into is not None and out is not None:
            into(c, out[lo:hi])
            continue
        emb = location_model(c)
        if hasattr(emb, "cpu"):
            emb = emb.cpu().numpy()
        if out is None:
            out = np.lib.format.open_memmap(path, mode="w+", dtype=np.float64, shape=(N, emb.shape[1]))
        out[lo:hi] = emb
    if out is not None:
        out.flush()
    return out
