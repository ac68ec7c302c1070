"""range_b200/save.py: file formats of the reference's caller (range/utils/save.py:7-58) with a stand-in model"""
import numpy as np
import torch


class _FakeModel:
    location_model_name = "RANGE+"

    def __call__(self, coords):
        c = torch.as_tensor(coords).double().numpy()
        out = np.zeros((len(c), 1280))
        out[:, 0], out[:, 1], out[:, 1279] = c[:, 0], c[:, 1], c.sum(1)
        return out


def _loader(n, bs, seed):
    rng = np.random.default_rng(seed)
    c, y = torch.tensor(rng.uniform(-90, 90, (n, 2))), torch.tensor(rng.integers(0, 5, n))
    return [(c[i:i + bs], y[i:i + bs]) for i in range(0, n, bs)], c, y


def test_save_embeddings_writes_the_reference_format(tmp_path):
    from range_b200.save import save_embeddings
    tr, ctr, ytr = _loader(23, 5, 0)
    va, cva, yva = _loader(7, 4, 1)
    ptr, pva = save_embeddings(_FakeModel(), tr, va, str(tmp_path), "toy")
    assert ptr.endswith("RANGE+/toy_train.npz") and pva.endswith("RANGE+/toy_val.npz")
    z = np.load(ptr)
    assert sorted(z.files) == ["coords", "embeddings", "y"]
    assert np.array_equal(z["coords"], ctr.numpy()) and np.array_equal(z["y"], ytr.numpy())
    assert z["embeddings"].shape == (23, 1280) and np.array_equal(z["embeddings"][:, 1279], ctr.numpy().sum(1))
    assert np.load(pva)["embeddings"].shape == (7, 1280)


def test_embed_to_npy_streams_in_batches(tmp_path):
    from range_b200.save import embed_to_npy
    rng = np.random.default_rng(2)
    c = rng.uniform(-90, 90, (1001, 2))
    out = embed_to_npy(_FakeModel(), c, str(tmp_path / "e.npy"), batch=128)
    back = np.load(tmp_path / "e.npy", mmap_mode="r")
    assert back.shape == (1001, 1280) and back.dtype == np.float64
    assert np.array_equal(back[:, 0], c[:, 0]) and np.array_equal(back[:, 1279], c.sum(1))
    del out
