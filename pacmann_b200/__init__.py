"""pacmann_b200: B200 (sm_100a) implementation of Pacmann's data-parallel hot path.

Layers:
  csrc/                hand-written CUDA kernels + the C-ABI (libpacmann_cuda.so, include/pacmann_cuda.h)
  cabi.py              ctypes binding of that C-ABI (what a cgo bridge binds)
  pianopir.py          host-side mirror of the reference's `pianopir` Go package over the C-ABI
  graphann.py          host-side mirror of the reference's `graphann` Go package over the C-ABI

No CPU fallback: importing works without the shared library, every call raises without it.
"""
from . import cabi  # noqa: F401

__all__ = ["cabi"]
