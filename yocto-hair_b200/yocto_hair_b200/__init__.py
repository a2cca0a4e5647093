"""yocto-hair on B200: Python harness over the C ABI of csrc/libyocto_hair_b200.so.

The product is the CUDA library; this package only binds it (ctypes) for the
tests, bench.py and the multi-GPU launcher.
"""
from . import _capi  # noqa: F401
