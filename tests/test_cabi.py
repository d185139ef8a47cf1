"""CPU tests of the C-ABI library: it loads, exports every symbol include/shgpu.h declares, and
refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os

import pytest
import shpkg

pkg = shpkg.load()


def test_library_exports_every_declared_symbol():
    lib = pkg.load_library()
    syms = pkg.exported_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.sh_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.ShGpuError):
        pkg.ShGpu()


def test_product_does_not_reference_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "lammps-spherharm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_py" not in txt and "sh_oracle" not in txt and "libshoracle" not in txt, f
