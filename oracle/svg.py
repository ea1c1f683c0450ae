"""``detect_svgs`` restated on plain arrays (test infrastructure, see oracle/__init__.py).

Follows ``/root/reference/chrysalis/core.py:50-92`` step by step; the scanpy
calls at ``core.py:53-57`` are restated from scanpy's documented semantics
(scanpy is unpinned, ``pyproject.toml:29``, and not installed here):

* ``filter_genes(min_cells=m)``: keep gene g iff ``(X[:, g] > 0).sum() >= m``
* ``normalize_total``: T_i = row sum over the KEPT genes, target = median of
  T_i[T_i > 0], rows with T_i == 0 untouched, integer X cast to float32
* ``log1p``: natural log1p of the stored values (zeros stay zeros)
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import moran as _moran
from . import weights as _weights


def filter_genes_mask(X: sp.spmatrix, min_cells: int) -> np.ndarray:
    """scanpy ``filter_genes(min_cells=...)`` gene subset (core.py:53)."""
    n_cells = np.asarray((sp.csr_matrix(X) > 0).sum(axis=0)).ravel()
    return n_cells >= min_cells


def normalize_total_log1p(X: sp.csr_matrix) -> sp.csr_matrix:
    """scanpy ``normalize_total(inplace=True)`` + ``log1p`` on a copy (core.py:56-57)."""
    X = sp.csr_matrix(X, copy=True)
    counts = np.asarray(X.sum(axis=1)).ravel()
    if np.issubdtype(X.dtype, np.integer):
        X = X.astype(np.float32)
    after = np.median(counts[counts > 0], axis=0)
    counts = counts + (counts == 0)
    counts = counts / after
    scale = 1 / counts
    X.data = (X.data * np.repeat(scale, np.diff(X.indptr))).astype(np.float32)
    X.data = np.log1p(X.data)
    return X


def select_svgs(morans: np.ndarray, scored: np.ndarray, top_svg: int, min_morans: float) -> np.ndarray:
    """The selection rule of core.py:78-92 on arrays.

    ``morans``: (G,) Moran's I for every var (NaN where not scored);
    ``scored``: (G,) bool, genes that passed the ``min_spots`` filter.
    Returns the (G,) bool ``spatially_variable`` column.

    Ties in the descending sort: pandas ``sort_values`` (quicksort) leaves
    tie order unspecified; here ties keep original gene order (stable).
    """
    idx = np.flatnonzero(scored)
    vals = morans[idx]
    order = idx[np.argsort(-vals, kind="stable")]
    n_top = len(order[:top_svg])
    n_thr = int((vals > min_morans).sum())
    out = np.zeros(len(morans), dtype=bool)
    if n_top < n_thr:
        out[order[:top_svg]] = True
    else:
        out[idx[vals > min_morans]] = True
    return out


def detect_svgs(X: sp.spmatrix, spatial: np.ndarray, min_spots: float = 0.05, top_svg: int = 1000,
                min_morans: float = 0.20, neighbors: int = 6, geary: bool = False,
                already_log1p: bool = False) -> dict:
    """Array-level restatement of ``detect_svgs`` (core.py:12-92).

    Returns ``{"morans": (G,) float64 with NaN for filtered genes,
    "spatially_variable": (G,) bool, "gene_mask": (G,) bool, "nbr": (n,k),
    ["geary": (G,)]}``.  ``already_log1p`` mirrors the ``"log1p" in
    adata.uns_keys()`` guard at core.py:55.
    """
    assert 0 < min_spots < 1  # core.py:50
    X = sp.csr_matrix(X)
    n, G = X.shape
    mask = filter_genes_mask(X, int(n * min_spots))  # core.py:53
    ad = X[:, mask]
    if not already_log1p:  # core.py:55-57
        ad = normalize_total_log1p(ad)
    dense = np.asarray(ad.todense(), dtype=np.float32 if ad.dtype != np.float64 else np.float64)  # core.py:59
    nbr = _weights.knn_neighbors(_weights.reference_points(spatial), neighbors)  # core.py:61-65
    res = _moran.morans_I_loop(dense, nbr, geary=geary)  # core.py:71-76
    I, C = res if geary else (res, None)
    morans = np.full(G, np.nan)
    morans[mask] = I
    out = {"morans": morans, "gene_mask": mask, "nbr": nbr,
           "spatially_variable": select_svgs(morans, mask, top_svg, min_morans)}
    if geary:
        gc = np.full(G, np.nan)
        gc[mask] = C
        out["geary"] = gc
    return out
