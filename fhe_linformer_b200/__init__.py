"""fhe_linformer_b200 -- B200-native CKKS evaluation engine behind FHE-Linformer's FHEController.

The product is the C-ABI shared library `lib/libflckks.so` (include/fl_ckks.h), hand-written CUDA for
sm_100a.  This Python package is only a ctypes front-end used by tests and bench.py; there is no CPU
fallback -- importing `capi` raises if the library has not been built.
"""
from .capi import Engine, DevBuf, load_library, LIB_PATH, ReferenceParams  # noqa: F401
from .ckks import CKKS, Elem  # noqa: F401
