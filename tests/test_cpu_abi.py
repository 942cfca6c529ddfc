"""CPU: the C-ABI library builds, loads and exports every symbol include/sphy_b200.h declares; no compute."""
import ctypes
import os
import re

import pytest

from sphy_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "sphy_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sphy_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    declared = header_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes prototype"


def test_struct_layouts_match(lib):
    for which, st in enumerate(_lib.ABI_STRUCTS):
        assert lib.sphy_abi_sizeof(which) == ctypes.sizeof(st), st.__name__
    assert lib.sphy_scan_tile_cells() == _lib.SCAN_TILE      # the host planner aligns partitions to the library's tile


def test_no_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = ctypes.c_void_p()
    rc = lib.sphy_ctx_create(ctypes.byref(ctx), 0, 4, 4, ctypes.c_float(1.0))
    assert rc < 0 and not ctx.value
    from sphy_b200.model import MapSource, SphyModel
    with pytest.raises(RuntimeError):
        SphyModel("does-not-matter.cfg", MapSource())


def test_bad_arguments_are_rejected(lib):
    assert lib.sphy_ctx_create(None, 0, 4, 4, ctypes.c_float(1.0)) == -1
    ctx = ctypes.c_void_p()
    assert lib.sphy_ctx_create(ctypes.byref(ctx), 0, 0, 4, ctypes.c_float(1.0)) == -1
    assert lib.sphy_ctx_create(ctypes.byref(ctx), 0, 4, 4, ctypes.c_float(0.0)) == -1
    assert lib.sphy_wb_step(None, None, None, None, None, 1, None) == -1
    assert lib.sphy_route_step(None, 1, None, None, _lib.SphyParam(None, 0.9), ctypes.c_float(0.1), 0, None) == -1


def test_product_does_not_import_the_oracle():
    """The product path must never route through oracle/ (checked statically over the package sources)."""
    pkg = os.path.join(ROOT, "sphy_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "oracle/_build" not in text and "libldd_oracle" not in text, fn


def test_built_for_sm100a_only():
    flags = " ".join(_build.NVCC_FLAGS)
    assert "arch=compute_100a,code=sm_100a" in flags and "-fmad=false" in flags


def test_every_entry_point_is_documented_for_integrators():
    """INTEGRATION.md names, for every exported entry point, the reference interface it replaces."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "sphy_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    syms = sorted(set(re.findall(r"\b(sphy_[a-z0-9_]+)\s*\(", hdr)))
    assert len(syms) > 30
    missing = [s for s in syms if s not in doc]
    assert not missing, missing
