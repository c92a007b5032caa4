"""phylo_b200: B200-native (sm_100a) implementation of the VCSMC hot path of amoretti86/phylo.

Layout
  csrc/        hand-written CUDA kernels + the C ABI (include/vcsmc_b200.h) -> libvcsmc_b200.so
  _lib.py      ctypes binding of the C ABI (no fallback: raises if the library is not built)
  ops.py       torch-facing operators / custom ops with autograd over the C ABI
  loader.py    alignment loaders of runner.py:83-184
  vcsmc.py     `VCSMC(datadict, K, args).train(...)`, the reference's class interface (vcsmc.py:103-645)
  runner.py    the reference's CLI (runner.py:12-58, 197-212)
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (does not load the shared library until first use)

__all__ = ["_lib", "__version__"]
