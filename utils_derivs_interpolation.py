"""Drop-in module name of the reference's utils_derivs_interpolation.py."""
from drake_ddp_b200.utils_derivs_interpolation import derivs_interpolation, index_tuple  # noqa: F401
