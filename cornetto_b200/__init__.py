"""cornetto_b200 -- B200-native implementation of cornetto's sequence-scan hot path.

The product is the C ABI in ``include/corn_gpu.h`` (CUDA kernels in ``csrc/``) and the drop-in
``cornetto`` host binary in ``host/``.  This Python package is plumbing for tests and bench.py:
``capi`` binds the shared library with ctypes, ``build`` compiles it in-tree.

There is no CPU fallback anywhere in this package: every scan call goes through
``libcorn_gpu.so`` and fails loudly when the library or a GPU is missing.
"""
from .build import build, lib_path, bin_path  # noqa: F401
