"""cabanamd_b200 — B200-native (sm_100a CUDA + NCCL) implementation of CabanaMD's
short-range LJ MD step behind a C ABI (include/cbmd_c_api.h).

The Python layer is only the test / bench harness over that ABI; the product is
lib/libcbmd_cuda.so (kernels + C ABI) and the C++ host classes under host/.
There is no CPU fallback: loading fails loudly when the library is missing and
context creation fails when no sm_100 device is present.
"""
from .capi import Context, Hub, CbmdError, load_library, library_path, declared_symbols  # noqa: F401

__version__ = "0.1"
