"""The C-ABI library must load on a CPU-only box and export every symbol include/mvedit_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'mvedit_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(mve_[a-zA-Z0-9_]+)\s*\(', text)))


def test_library_loads_and_exports_all_declared_symbols():
    from mvedit_b200 import build
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    syms = declared_symbols()
    assert len(syms) >= 10
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    lib.mve_version.restype = ctypes.c_int
    assert lib.mve_version() == 1


def test_no_undeclared_exports():
    """Every exported mve_* symbol is declared in the header (the header is the contract)."""
    import subprocess
    from mvedit_b200 import build
    out = subprocess.run(['nm', '-D', '--defined-only', build.build()], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r' T (mve_[a-zA-Z0-9_]+)', out)))
    assert set(exported) == set(declared_symbols()), (set(exported) ^ set(declared_symbols()))


def test_ops_fail_loudly_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('has CUDA')
    from mvedit_b200 import raymarching
    with pytest.raises((RuntimeError, AssertionError)):
        raymarching.morton3D(torch.zeros(4, 3, dtype=torch.int32))
