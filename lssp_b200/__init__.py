"""lssp_b200 -- B200-native implementation of the LSSP solve-loop hot path.

CUDA kernels (sm_100a) + C ABI live in ``csrc/`` and are built into
``liblsspg.so``; this package is the Python host mirror of the reference
interface on top of that C ABI.  No CPU fallback exists.
"""
from . import generators  # noqa: F401
from ._lib import LIB_PATH, LsspgError, lib  # noqa: F401
