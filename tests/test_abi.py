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


def test_binding_argument_counts_match_the_header():
    """Every ctypes signature in _lib.py has as many arguments as the prototype in the header."""
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    protos = dict(re.findall(r"\b(chr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S))
    assert sorted(protos) == _lib.exported_symbols()
    for name, args in protos.items():
        args = args.strip()
        n_args = 0 if args in ("", "void") else len(args.split(","))
        assert n_args == len(_lib._SIGNATURES[name][1]), name


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


def test_aa_and_pca_entry_points_validate_their_shapes(lib):
    # limits of the archetype kernels (<= 64 archetypes, <= 64 dimensions) and of the eigensolver are reported, not hit
    one = ctypes.c_double(0.0)
    p = ctypes.addressof(one)
    rc = lib.chr_aa_iterate(p, 100, 65, 8, 200.0, 0, p, p, p, None, p, 1 << 20, None)
    assert rc != 0 and b"chr_aa_iterate" in lib.chr_last_error()
    rc = lib.chr_aa_iterate(p, 100, 8, 1, 200.0, 0, p, p, p, None, p, 1 << 20, None)
    assert rc != 0 and b"n_archetypes" in lib.chr_last_error()
    rc = lib.chr_aa_iterate(p, 100, 8, 4, 200.0, 0, p, p, p, None, p, 16, None)
    assert rc != 0 and b"workspace too small" in lib.chr_last_error()
    rc = lib.chr_eigh_topk_f64(p, 50, 60, 60, p, p, p, p, 1 << 20, None)
    assert rc != 0 and b"chr_eigh_topk_f64" in lib.chr_last_error()
    assert lib.chr_eigh_default_steps(2000, 30) == 122 and lib.chr_eigh_default_steps(40, 30) == 40
    assert lib.chr_aa_workspace_bytes(700000, 30, 12) < (1 << 28)


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
