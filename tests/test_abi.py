"""The C-ABI library loads and exports every symbol include/femder_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "femder_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fdb_[A-Za-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_groups():
    syms = declared_symbols()
    for s in ("fdb_set_volume_mesh", "fdb_pattern_get", "fdb_elem2nnz_get", "fdb_assemble_volume", "fdb_heron_areas",
              "fdb_assemble_surface", "fdb_sweep", "fdb_spmv", "fdb_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from femder_b200 import _lib
    lib = _lib.load()
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/femder_b200.h but not exported"
    assert set(declared_symbols()) == set(_lib.SIGNATURES), "ctypes signatures and header disagree"
    assert lib.fdb_abi_version() == 2
    assert lib.fdb_sizeof_sweep_params() == ctypes.sizeof(_lib.SweepParams)
    assert lib.fdb_sizeof_sweep_result() == ctypes.sizeof(_lib.SweepResult)


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run instead of computing on the CPU."""
    from femder_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    import femder_b200
    with pytest.raises(_lib.FemderB200Error, match="no CUDA device"):
        femder_b200.System()
    h = ctypes.c_void_p()
    assert _lib.load().fdb_system_create(0, None, ctypes.byref(h)) != 0
    assert b"no CPU fallback" in _lib.load().fdb_last_error()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "femder_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), (dirpath, f)
                assert "femder_oracle" not in src and "reference_shim" not in src, (dirpath, f)
