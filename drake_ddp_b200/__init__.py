"""B200-native batched iLQR/DDP (drop-in for vincekurtz/drake_ddp's ilqr.py hot path)."""
from .utils_derivs_interpolation import derivs_interpolation, index_tuple  # noqa: F401
from . import systems, problems  # noqa: F401
