"""dcdf_b200 -- B200-native (sm_100a) Heuristic T-k^2-raster codec behind a C-ABI.

Only what the hot path needs lives here: csrc/ (CUDA kernels + the C-ABI, built into libdcdf_cuda.so),
_ffi.py (ctypes declarations of include/dcdf_cuda.h), api.py (host-side mirror of the reference's Chunk / Superchunk
interface), variable.py (how the product drives it: Variable.append, a device-resident cache of stored superchunks,
the MMArray3 / __getitem__ surface of py-dcdf) and the stored metadata above the superchunks: span.py (span tree) and
dataset.py (Dataset node, coordinates).  There is no CPU implementation of the codec in this package.
"""
from ._ffi import build_library as build  # noqa: F401
from .api import Chunk, Context, DcdfError, Superchunk  # noqa: F401
from .variable import ChunkCache, MMArray3, Variable  # noqa: F401
from .dataset import Coordinate, Dataset  # noqa: F401
from .store import DirStore  # noqa: F401

__all__ = ["build", "Chunk", "ChunkCache", "Context", "Coordinate", "Dataset", "DcdfError", "DirStore", "MMArray3", "Superchunk", "Variable"]
