"""range_b200: B200-native implementation of mvrl/RANGE's `load_model(...)` / `model(locs)` hot path."""
from .load_model import load_model  # noqa: F401

__all__ = ["load_model"]
