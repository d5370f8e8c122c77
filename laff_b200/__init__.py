"""laff_b200 — B200-native (sm_100a) implementation of the LAFF retrieval hot path.

Python host side mirroring the reference's model / loss / evaluation surface (ruc-aimc-lab/LAFF) over hand-written
CUDA kernels reached through the C ABI in ``include/laff_b200.h``.  See DESIGN.md and INTEGRATION.md.
"""
from ._capi import LaffError, LIB_PATH  # noqa: F401

__all__ = ["LaffError", "LIB_PATH"]
__version__ = "0.1.0"
