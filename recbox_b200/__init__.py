"""recbox_b200 -- B200 (sm_100a) implementation of RecBox's embedding + feature-interaction hot path.

Layout:
  csrc/        hand-written CUDA kernels + the C ABI of include/recbox_b200.h (-> librecbox_b200.so)
  _lib.py      in-tree build + ctypes loader (no fallback: missing library = exception)
  ops.py       one-call-per-op tensor wrappers over the C ABI
  functional.py, layers.py   the reference's nn.Module operator API on top (same names/ctor args)
"""
__version__ = "0.1.0"

from ._lib import RbxError, build, load  # noqa: F401
