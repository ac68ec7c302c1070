"""Runs the UNMODIFIED reference (mvrl/RANGE, pure Python) on the CPU.  TEST / BENCHMARK INFRASTRUCTURE ONLY:
tests/golden/make_golden.py (golden vectors) and bench.py's `--impl reference` / `cpu_baseline` legs use it; the
product (range_b200/) never imports it.

Where the reference comes from: /root/reference in the build container; on the GPU box that mount does not exist, so
`stage()` (called by __graft_entry__.build() here) copies the reference's `range/` package - as it is, no edits - into
the git-ignored oracle/_ref/reference/, next to the regenerated spherical_harmonics_ylm.py (the file is stripped from
the mount, .MISSING_LARGE_BLOBS; regenerated with the reference's own sympy generator, tools/make_sh_table.py).
oracle/_ref/ travels to the GPU box with the snapshot like the built .so files do.

What has to be faked around it: permissive stub modules for third-party packages the reference
imports at module scope but never uses on this path (lightning, timm, torchgeo, rasterio, matplotlib, ...), and a
random-init SatCLIP-L40 checkpoint fabricated through the reference's own SatCLIPLightningModule.
"""
import importlib
import importlib.util
import os
import shutil
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
MOUNT = "/root/reference"
STAGED = os.path.join(HERE, "_ref", "reference")
YLM = os.path.join(HERE, "_ref", "spherical_harmonics_ylm.py")
YLM_MODULE = "range.location_models.satclip.positional_encoding.spherical_harmonics_ylm"


def reference_root():
    """directory holding the reference's `range/` package, or None"""
    for root in (MOUNT, STAGED):
        if os.path.isfile(os.path.join(root, "range", "load_model.py")):
            return root
    return None


def available():
    return reference_root() is not None and os.path.exists(YLM)


def stage():
    """build container only: copy the reference's python package next to the oracle (git-ignored) so it travels"""
    if not os.path.isdir(os.path.join(MOUNT, "range")):
        return False
    if not os.path.exists(YLM):
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from make_sh_table import regenerate
        regenerate(YLM)
    dst = os.path.join(STAGED, "range")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(os.path.join(MOUNT, "range"), dst,
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.ipynb", "*.png", "*.jpg"))
    for root, dirs, files in os.walk(STAGED):            # the mount is read-only: make the copy removable
        for name in dirs + files:
            os.chmod(os.path.join(root, name), 0o755 if name in dirs else 0o644)
    return True


def install_stubs():
    class _Meta(type):
        def __getattr__(cls, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Meta(name, (_Any,), {})

    class _Any(metaclass=_Meta):
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Any()

        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Any()

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

    def make(name):
        mod = types.ModuleType(name)
        mod.__path__ = []

        def _getattr(attr):
            if attr.startswith("__"):
                raise AttributeError(attr)
            return _Meta(attr, (_Any,), {})

        mod.__getattr__ = _getattr
        sys.modules[name] = mod
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, mod)
        return mod

    for name in ["lightning", "lightning.pytorch", "lightning.pytorch.callbacks", "lightning.pytorch.cli",
                 "pytorch_lightning", "timm", "torchgeo", "torchgeo.models", "torchgeo.datasets",
                 "torchgeo.datasets.geo", "rasterio", "matplotlib", "matplotlib.pyplot", "albumentations",
                 "albumentations.core", "albumentations.core.transforms_interface", "albumentations.pytorch",
                 "huggingface_hub", "wandb", "geoclip", "rshf", "rshf.satmae", "cartopy", "skimage", "h5py"]:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:
            make(name)
    lp = sys.modules["lightning.pytorch"]
    if not (isinstance(lp.__dict__.get("LightningModule"), type)
            and issubclass(lp.__dict__["LightningModule"], torch.nn.Module)):
        lp.LightningModule = LightningModule
        sys.modules["pytorch_lightning"].LightningModule = LightningModule


def import_reference():
    """(load_model, SatCLIPLightningModule) of the unmodified reference"""
    root = reference_root()
    if root is None:
        raise RuntimeError("the reference is neither mounted at /root/reference nor staged under oracle/_ref/reference")
    if not os.path.exists(YLM):
        if root != MOUNT:
            raise RuntimeError(f"{YLM} is missing (regenerated by __graft_entry__.build() in the build container)")
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from make_sh_table import regenerate
        regenerate(YLM)
    spec = importlib.util.spec_from_file_location(YLM_MODULE, YLM)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[YLM_MODULE] = mod
    spec.loader.exec_module(mod)
    install_stubs()
    sys.path.insert(0, root)
    from range.load_model import load_model                     # the reference, unmodified
    from range.location_models.satclip.main_old import SatCLIPLightningModule
    return load_model, SatCLIPLightningModule


def fabricate_ckpt(SatCLIPLightningModule, path, capacity, seed=0):
    """random-init SatCLIP-L40 checkpoint in the Lightning format load.py:3-18 reads; returns the SIREN (W, b) list"""
    torch.manual_seed(seed)
    hp = dict(embed_dim=256, image_resolution=64, vision_layers=1, vision_width=64, vision_patch_size=32,
              in_channels=3, le_type="sphericalharmonics", pe_type="siren", frequency_num=16, max_radius=260,
              min_radius=1, legendre_polys=40, harmonics_calculation="analytic", sh_embedding_dims=32,
              learning_rate=1e-4, weight_decay=0.01, num_hidden_layers=2, capacity=capacity)
    module = SatCLIPLightningModule(**hp)
    hp.update(eval_downstream=False, air_temp_data_path=None, election_data_path=None)
    torch.save({"hyper_parameters": hp, "state_dict": module.state_dict()}, path)
    sd = module.state_dict()
    pre = "model.location.nnet."
    weights = [(sd[pre + "layers.0.weight"], sd[pre + "layers.0.bias"]),
               (sd[pre + "layers.1.weight"], sd[pre + "layers.1.bias"]),
               (sd[pre + "last_layer.weight"], sd[pre + "last_layer.bias"])]
    return [(w.double().clone(), b.double().clone()) for w, b in weights]
