"""B200-native Monte-Carlo LDPC decode path (drop-in for LDPCsimulation's C_implementations hot loop).

The product is the C-ABI shared library built from csrc/ (include/ldpc_gpu.h); this package is its
ctypes binding plus the build recipe.  Nothing here falls back to a CPU implementation.
"""
from . import abi  # noqa: F401

__all__ = ["abi"]
