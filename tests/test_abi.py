"""The C-ABI library builds, loads and exports exactly what include/chrysalis_b200.h declares
(no compute calls here - there is no GPU on the CPU box)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from chrysalis_b200 import _lib

HEADER = os.path.join(ROOT, "include", "chrysalis_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(chr_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def test_header_and_binding_agree():
    assert declared_functions() == _lib.exported_symbols()


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert isinstance(getattr(lib, name), ctypes._CFuncPtr), name


def test_version_and_error_string(lib):
    assert lib.chr_version() >= 100
    assert isinstance(lib.chr_last_error(), bytes)


def test_workspace_queries_need_no_gpu(lib):
    assert lib.chr_moran_workspace_bytes(4035, 13000, 6, 0) > 4035
    assert lib.chr_moran_workspace_bytes(700000, 18000, 8, 0) < (1 << 30)
    assert lib.chr_knn_workspace_bytes(700000, 8) > 700000 * 8
    assert 700000 * 8 < lib.chr_moran_plan_bytes(700000, 8) < (1 << 30)


def test_argument_validation_precedes_any_cuda_call(lib):
    # null / inconsistent arguments are rejected with a message before touching the device
    rc = lib.chr_moran_dense_f32(None, 100, 10, 10, None, 6, None, 0, None, 0, None, None, None, None, 0, None)
    assert rc != 0 and b"chr_moran_dense_f32" in lib.chr_last_error()
    rc = lib.chr_knn_build(None, 10, 3, None, None, None, 0, None)
    assert rc != 0 and b"chr_knn_build" in lib.chr_last_error()
    rc = lib.chr_moran_plan_build(None, 10, 3, None, 0, None, None, 0, None)
    assert rc != 0 and b"chr_moran_plan_build" in lib.chr_last_error()


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import scipy.sparse as sp
    import chrysalis_b200 as ch
    ad = ch.AnnDataLite(sp.random(50, 20, 0.5, format="csr", dtype=np.float32), obsm={"spatial": np.random.rand(50, 2)})
    with pytest.raises(_lib.ChrysalisNativeError):
        ch.detect_svgs(ad)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "chrysalis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
