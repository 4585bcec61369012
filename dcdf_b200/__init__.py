"""dcdf_b200 -- B200-native (sm_100a) Heuristic T-k^2-raster codec behind a C-ABI.

Only what the hot path needs lives here: csrc/ (CUDA kernels + the C-ABI, built into libdcdf_cuda.so),
_ffi.py (ctypes declarations of include/dcdf_cuda.h) and api.py (host-side mirror of the reference's
Chunk / Superchunk / MMArray3 interface).  There is no CPU implementation in this package.
"""
from ._ffi import build_library as build  # noqa: F401
from .api import Chunk, Context, DcdfError, MMArray3, Superchunk  # noqa: F401

__all__ = ["build", "Chunk", "Context", "DcdfError", "MMArray3", "Superchunk"]
