"""faspsolver_b200 — B200-native solve-phase hot path of FASP behind FASP's own C API.

The product is faspsolver_b200/lib/libfasp_cuda.so (hand-written CUDA for sm_100a, C ABI in
include/fasp_cuda.h). This package is the Python host-side mirror used by tests and
benchmarks: ctypes bindings (api), struct mirrors (fasp_types), synthetic inputs (problems),
and the in-tree build (build).
"""
from . import fasp_types  # noqa: F401

__all__ = ["api", "build", "fasp_types", "problems"]
__version__ = "0.1.0"
