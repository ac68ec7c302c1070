"""Host-side helpers mirroring range/utils/utils.py."""
import numpy as np


def rad_to_cart(locations):
    """(lon, lat) radians -> unit xyz; dtype-preserving numpy, as range/utils/utils.py:11-16."""
    x = np.cos(locations[:, 1]) * np.cos(locations[:, 0])
    y = np.cos(locations[:, 1]) * np.sin(locations[:, 0])
    z = np.sin(locations[:, 1])
    return np.stack([x, y, z], axis=1)
