"""The C-ABI library loads on a CPU-only box and exports every symbol include/bricklib_b200.h declares."""
import ctypes
import os
import re

import pytest

from bricklib_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "bricklib_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bk_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SIGNATURES)


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_metadata_calls_work_without_gpu():
    L = _lib.load()
    assert b"sm_100a" in L.bk_version()
    assert [L.bk_stencil_radius(s) for s in range(5)] == [1, 1, 2, 4, 2]
    assert [L.bk_stencil_st_iter(s) for s in range(5)] == [8, 8, 4, 2, 4]
    assert [L.bk_stencil_points(s) for s in range(5)] == [7, 7, 13, 25, 125]
    assert L.bk_stencil_radius(9) < 0


def test_no_cpu_fallback():
    """without a device the compute path must fail loudly, not fall back"""
    import bricklib_b200 as bk
    if bk.have_gpu():
        pytest.skip("GPU present")
    with pytest.raises(bk.BrickError):
        bk.DeviceBuffer(1024)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bricklib_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
