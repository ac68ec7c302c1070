"""`load_model` - drop-in for range/load_model.py:16-51 (RANGE / RANGE+ only).

    from range_b200.load_model import load_model
    model = load_model(model_name='RANGE+', pretrained_path=ckpt, device='cuda', db_path=npz, beta=0.5)
    emb = model(locs)            # locs (N,2) float64 (lon, lat) degrees -> np.ndarray float64 (N, 1280)
"""
from argparse import Namespace

from .range import LocationEncoder


def load_model(model_name='RANGE+', pretrained_path=None, device='cuda', **kwargs):
    """Same arguments and errors as the reference:
      * `pretrained_path is None` -> ValueError                         (load_model.py:31-32)
      * names containing 'RANGE' need `db_path`                         (load_model.py:33-35)
      * `beta` defaults to 0.5                                          (load_model.py:37-40)
    Extra, optional keywords (not in the reference): `chunk` (queries per pipelined chunk),
    `db_shard=(rank, world)` + `db_group` (torch.distributed group) for an M-sharded database (model(locs) is then a
    collective call: every rank passes its own queries; `db_merge='peer'|'reduce_scatter'`), `db_cache` (path of the
    prepared device layout: written on first use, validated against source and shard, read afterwards), `host_path`
    ('auto' | 'copy' | 'packed': how the float64 result reaches the host, range.py:_forward_host),
    `out_dtype` (np.float64 like the reference, or np.float32: half the device->host bytes), `pinned_limit` (bytes),
    `host_threads`.
    """
    if pretrained_path is None:
        raise ValueError("Please provide the pretrained model path.")
    if 'RANGE' in model_name:
        assert 'db_path' in kwargs, "db_path is required for RANGE model."
        db_path = kwargs.get('db_path')
        beta = kwargs.get('beta') if 'beta' in kwargs else 0.5
    else:
        raise NotImplementedError(f"{model_name}: range_b200 implements the RANGE and RANGE+ encoders only")
    args = Namespace(location_model_name=model_name, pretrained_path=pretrained_path, device=device,
                     range_db=db_path, beta=beta)
    for k in ('chunk', 'db_shard', 'db_group', 'db_cache', 'db_merge', 'tail', 'taper', 'super_batch', 'host_path', 'out_dtype', 'pinned_limit',
              'host_threads'):
        if k in kwargs:
            setattr(args, k, kwargs[k])
    model = LocationEncoder(args)
    model.eval()
    model.to(device)
    return model
