"""The C-ABI library loads and exports every entry point include/eqv2_b200.h declares (no compute)."""
import ctypes
import os
import re

from conftest import PKG, REPO


def declared_symbols():
    text = open(os.path.join(REPO, "include", "eqv2_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eqv2_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "eqv2_gemm_f32" in syms and "eqv2_radius_graph_pbc" in syms and len(syms) >= 25


def test_library_exports_every_declared_symbol():
    import importlib
    build = importlib.import_module(PKG + ".build")
    path = build.build()
    lib = ctypes.CDLL(path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    lib.eqv2_abi_version.restype = ctypes.c_int
    assert lib.eqv2_abi_version() >= 1


def test_binding_covers_header():
    import importlib
    _lib = importlib.import_module(PKG + "._lib")
    declared = set(declared_symbols()) - {"eqv2_last_error"}
    assert declared == set(_lib._PROTOS), declared ^ set(_lib._PROTOS)


def test_product_refuses_cpu_tensors():
    import importlib
    import pytest
    import torch
    _lib = importlib.import_module(PKG + "._lib")
    ops = importlib.import_module(PKG + ".ops")
    with pytest.raises(_lib.Eqv2Error):
        ops.linear(torch.zeros(4, 4), torch.zeros(4, 4), None)
