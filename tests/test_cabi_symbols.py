"""CPU-only: the C-ABI library loads and exports every symbol include/ultranest_b200.h declares,
the ctypes signature table covers all of them, and the product refuses to run without a GPU
(no CPU fallback).  No compute calls are made here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ultranest_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(unb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ("unb_find_nearby", "unb_compute_maxradiussq", "unb_inside_ellipsoid",
                 "unb_region_inside", "unb_region_bootstrap", "unb_loglike_gauss",
                 "unb_region_inside_loglike"):
        assert must in syms
    assert len(syms) >= 30


def test_library_exports_every_declared_symbol():
    from ultranest_b200 import build
    path = build.build()
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), "missing export: %s" % name
    lib.unb_abi_version.restype = ctypes.c_int
    assert lib.unb_abi_version() == 1


def test_ctypes_table_matches_header():
    from ultranest_b200 import _native
    bound = set(_native.SIGNATURES) | set(_native.FREE_SIGNATURES)
    assert bound == set(declared_symbols())
    _native.load_library()


def test_no_cpu_fallback():
    """Without a CUDA device the engine must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ultranest_b200 import _native, mlfriends
    import numpy as np
    with pytest.raises(RuntimeError):
        _native.Engine(0)
    with pytest.raises(RuntimeError):
        mlfriends.find_nearby(np.zeros((2, 2)), np.zeros((2, 2)), 1.0, np.zeros(2, dtype=np.int64))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under ultranest_b200/ may reference it."""
    pkg = os.path.join(ROOT, "ultranest_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "liboracle" not in text, fn
