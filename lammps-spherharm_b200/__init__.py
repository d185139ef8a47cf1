"""lammps-spherharm_b200 — B200-native SPHERHARM contact hot path.

Host-side Python mirror of the C-ABI in include/shgpu.h (ctypes, no torch types cross the
boundary).  The compute lives in csrc/ (CUDA sm_100a) behind libshgpu.so; there is no CPU
fallback: constructing ShGpu without a CUDA device raises.
"""
from .capi import ShGpu, ShGpuError, load_library, exported_symbols  # noqa: F401
from . import workloads  # noqa: F401


def load_decomp():
    """Domain-decomposition driver (imports torch.distributed lazily)."""
    import importlib
    _d = importlib.import_module(__name__ + ".decomp")
    return _d
