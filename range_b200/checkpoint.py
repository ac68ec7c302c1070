"""SatCLIP checkpoint -> SIREN weights.

Reads the Lightning checkpoint format the reference loads at
range/location_models/satclip/load.py:3-18 (`hyper_parameters` + `state_dict`) without needing
lightning: only the location tower (model.location.nnet.*) is used on this path
(model_old.py:326-330); the vision tower in the file is ignored.
"""
import torch

_PREFIX = "model.location.nnet."


def load_satclip_location_encoder(ckpt_path):
    """Returns dict(L, dims, weights=[(W, b), ...] fp64 CPU tensors, harmonics_calculation)."""
    ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    hp = dict(ckpt["hyper_parameters"])
    for k in ("eval_downstream", "air_temp_data_path", "election_data_path"):    # load.py:5-7
        hp.pop(k)
    if hp.get("le_type", "sphericalharmonics") != "sphericalharmonics" or hp.get("pe_type", "siren") != "siren":
        raise NotImplementedError("range_b200 implements the SatCLIP spherical-harmonics + SIREN location encoder "
                                  f"only (got le_type={hp.get('le_type')}, pe_type={hp.get('pe_type')})")
    calc = hp.get("harmonics_calculation", "analytic")
    if calc not in ("analytic", "closed-form"):                                   # spherical_harmonics.py:22-25
        raise NotImplementedError(f"harmonics_calculation={calc!r}: expected 'analytic' or 'closed-form'")
    L = int(hp["legendre_polys"])
    sd = ckpt["state_dict"]
    n_hidden = int(hp.get("num_hidden_layers", 2))
    weights = []
    for i in range(n_hidden):
        weights.append((sd[f"{_PREFIX}layers.{i}.weight"], sd[f"{_PREFIX}layers.{i}.bias"]))
    weights.append((sd[_PREFIX + "last_layer.weight"], sd[_PREFIX + "last_layer.bias"]))
    weights = [(w.detach().double().contiguous(), b.detach().double().contiguous()) for w, b in weights]  # range.py:83
    dims = [L * L] + [w.shape[0] for w, _ in weights]
    assert weights[0][0].shape[1] == L * L, "first SIREN layer does not match legendre_polys"
    return dict(L=L, dims=dims, weights=weights, harmonics_calculation=calc, hyper_parameters=hp)
