"""ctypes binding of ``libchrysalis_b200.so`` (the C ABI declared in ``include/chrysalis_b200.h``).

There is no CPU fallback: if the shared library is missing, or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# the product is the in-tree library and nothing else (A/B builds of scripts/build_variant.sh are copied over it by
# scripts/ab_kernels.sh on the GPU box's scratch copy)
LIB_PATH = os.path.join(_HERE, "libchrysalis_b200.so")

FLAG_GEARY = 0x1
FLAG_TIMING = 0x2
FLAG_PRESHIFTED = 0x4
VARIANT_SHIFT = 8
FAMILY_MORAN, FAMILY_GRAM, FAMILY_AA, FAMILY_KNN, FAMILY_PREP = range(5)


class ChrysalisNativeError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); pointers are passed as integers (tensor.data_ptr())
_P = c_void_p
_SIGNATURES = {
    "chr_version": (c_int, []),
    "chr_last_error": (c_char_p, []),
    "chr_last_kernel_ms": (c_float, [c_int]),
    "chr_host_threads": (c_int, []),
    "chr_host_upload": (c_int, [_P, _P, c_size_t, _P]),
    "chr_host_download": (c_int, [_P, _P, c_size_t, _P]),
    "chr_host_csr_select_count": (c_int, [_P, _P, c_int64, c_int64, _P, _P, _P]),
    "chr_host_csr_select_upload": (c_int, [_P, _P, _P, c_int, c_int64, _P, _P, _P, _P, _P]),
    "chr_knn_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "chr_spatial_bbox": (c_int, [_P, c_int64, _P, _P]),
    "chr_knn_build": (c_int, [_P, c_int64, c_int, _P, _P, _P, c_size_t, _P]),
    "chr_hilbert_keys": (c_int, [_P, c_int64, _P, _P, _P]),
    "chr_csr_gene_nnz": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P]),
    "chr_csr_decode_compact": (c_int, [_P, _P, c_int64, _P, _P, c_int64, _P, _P, _P]),
    "chr_csr_decode_compact_delta": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P, c_int64, _P, _P, c_int64, _P, _P, _P]),
    "chr_csr_spot_totals": (c_int, [_P, _P, _P, c_int64, _P, _P, _P]),
    "chr_csr_densify_log1p": (c_int, [_P, _P, _P, c_int64, _P, _P, _P, _P, c_int64, c_int, _P]),
    "chr_moran_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_uint]),
    "chr_colsum_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "chr_colsum_f64": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, c_size_t, _P]),
    "chr_gram_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "chr_gram_centered_bf16x3": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, c_uint, _P, c_size_t, _P]),
    "chr_eigh_default_steps": (c_int, [c_int64, c_int]),
    "chr_eigh_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "chr_eigh_topk_f64": (c_int, [_P, c_int64, c_int, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "chr_pca_scores_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "chr_pca_scores_f32": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, c_int, _P, c_int64, _P, c_size_t, _P]),
    "chr_aa_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "chr_aa_init_archetypes": (c_int, [_P, c_int64, c_int, _P, c_int, _P, _P]),
    "chr_aa_furthest_sum_workspace_bytes": (c_size_t, [c_int64]),
    "chr_aa_furthest_sum": (c_int, [_P, c_int64, c_int, c_int, c_int64, _P, _P, c_size_t, _P]),
    "chr_aa_iterate": (c_int, [_P, c_int64, c_int, c_int, c_double, c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
    "chr_local_moran_f32": (c_int, [_P, c_int64, _P, c_int, c_double, c_int, ctypes.c_uint64, _P, _P, _P, _P, _P, _P, _P]),
    "chr_moran_plan_bytes": (c_size_t, [c_int64, c_int]),
    "chr_moran_plan_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "chr_moran_plan_build": (c_int, [_P, c_int64, c_int, _P, c_size_t, _P, _P, c_size_t, _P]),
    "chr_moran_blockdiag_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "chr_moran_blockdiag_f32": (c_int, [_P, c_int64, c_int64, c_int64, c_int, _P, c_int, _P, _P, c_int, _P, c_int64, c_uint, _P, _P, _P,
                                        c_size_t, _P]),
    "chr_p2p_allgather_f64": (c_int, [_P, c_int64, _P, _P, c_int, c_int, c_int64, c_int, c_uint, _P]),
    "chr_perm_row_map": (c_int, [c_int64, ctypes.c_uint64, c_int, _P, _P]),
    "chr_moran_perm_lattice_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int64, c_int]),
    "chr_moran_perm_lattice_f32": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_int, _P, c_int, _P, c_int, _P, c_int64, _P, c_int64, _P,
                                           c_int, _P, _P, c_int, ctypes.c_uint64, c_uint, _P, _P, _P, _P, _P, c_size_t, _P]),
    "chr_moran_lattice_csr_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int, c_int64]),
    "chr_moran_lattice_csr_f32": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, _P, c_int64, _P, c_int, c_int64, c_int, _P, c_int, _P, c_int,
                                          _P, c_int64, _P, _P, c_int, c_uint, _P, _P, _P, _P, _P, c_size_t, _P]),
    "chr_moran_lattice_workspace_bytes": (c_size_t, [c_int64, c_int, c_int64]),
    "chr_moran_lattice_f32": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_int, _P, c_int, c_int, _P, c_int, _P, c_int64, _P, c_int64, _P,
                                      c_int, c_uint, _P, _P, _P, _P, c_size_t, _P]),
    "chr_moran_dense_f32": (c_int, [_P, c_int64, c_int64, c_int64, _P, c_int, _P, c_int, _P, c_uint, _P, _P, _P, _P, c_size_t, _P]),
}


def exported_symbols():
    """Every symbol ``include/chrysalis_b200.h`` declares (kept in sync by tests/test_abi.py)."""
    return sorted(_SIGNATURES)


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ChrysalisNativeError(
            f"{LIB_PATH} is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C chrysalis_b200/csrc`).  chrysalis_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_NVTX = os.environ.get("CHR_NVTX", "0") == "1"     # NVTX range around every C-ABI call (timeline tools)


def call(name: str, *args):
    """Call a status-returning entry point; raise with the library's message on failure."""
    lib = load()
    if _NVTX:
        import torch
        torch.cuda.nvtx.range_push(name)
        try:
            rc = getattr(lib, name)(*args)
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.chr_last_error().decode("utf-8", "replace")
        raise ChrysalisNativeError(f"{name} failed (status {rc}): {msg}")


def query(name: str, *args):
    return getattr(load(), name)(*args)
