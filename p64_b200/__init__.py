"""p64_b200 -- B200-native (sm_100a) hot path of the PVRG-P64 H.261 encoder behind a C ABI.

The product is the shared library ``p64_b200/libp64b200.so`` (include/p64_b200.h).  This package is the
thin Python mirror used by tests and bench.py: it binds the ABI with ctypes and adds no compute of its own.
"""
from .y4m import IT_CIF, IT_NTSC, IT_QCIF, DIMS  # noqa: F401

ME_TSS, ME_FULL = 0, 1
__all__ = ["IT_CIF", "IT_NTSC", "IT_QCIF", "DIMS", "ME_TSS", "ME_FULL"]
