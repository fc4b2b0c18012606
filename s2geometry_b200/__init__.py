"""s2geometry_b200 — S2's data-parallel hot path (cell-id encoding, shape-index point containment) on B200.

Everything that computes goes through libs2b.so (hand-written sm_100a CUDA kernels behind the C ABI in
include/s2b_capi.h). Importing the package does not require a GPU; calling a batch entry point does.
"""
from . import _lib  # noqa: F401
from . import cellid  # noqa: F401
from . import index  # noqa: F401
from . import shapes  # noqa: F401
from ._lib import S2BError, launch_count  # noqa: F401

__all__ = ["cellid", "index", "shapes", "S2BError", "launch_count"]
