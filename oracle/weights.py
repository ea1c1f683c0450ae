"""Spatial weights of the reference path (test infrastructure, see oracle/__init__.py).

Restates ``weights.KNN.from_array(points, k=neighbors); w.transform = 'R'``
(``/root/reference/chrysalis/core.py:61-65``).  libpysal (>=4.x, not vendored in
the reference; unpinned via ``pysal`` in ``pyproject.toml:28``) implements
``KNN.from_array`` as ``scipy.spatial.cKDTree(points).query(points, k=k+1)``
followed by dropping the query point itself, binary weights, and 'R' =
row-standardisation (every weight 1/k, s0 = n).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.spatial import cKDTree


def reference_points(spatial: np.ndarray) -> np.ndarray:
    """``points = adata.obsm['spatial'].copy(); points[:, 1] *= -1`` (core.py:61-62)."""
    pts = np.array(spatial, copy=True)
    pts[:, 1] = pts[:, 1] * -1
    return pts


def knn_neighbors(points: np.ndarray, k: int) -> np.ndarray:
    """(n, k) neighbour indices in cKDTree order, self removed (libpysal KNN).

    libpysal drops the column equal to the row id; where the self index does
    not appear among the k+1 hits (duplicate coordinates) the last column is
    dropped instead.
    """
    pts = np.asarray(points)
    n = pts.shape[0]
    if not 0 < k < n:
        raise ValueError("k must satisfy 0 < k < n")
    _, idx = cKDTree(pts).query(pts, k=k + 1, p=2)
    idx = np.asarray(idx, dtype=np.int64)
    out = np.empty((n, k), dtype=np.int64)
    rows = np.arange(n)
    not_self = idx != rows[:, None]
    has_self = ~not_self.all(axis=1)
    # rows containing self: keep the k non-self entries in order
    sel = not_self.copy()
    sel[~has_self, k] = False  # no self found -> drop last column
    out[:] = idx[sel].reshape(n, k)
    return out


def row_standardised_csr(nbr: np.ndarray, n: int | None = None) -> sp.csr_matrix:
    """``w.sparse`` after ``w.transform = 'R'``: float64 CSR, weights 1/k."""
    nbr = np.asarray(nbr)
    n = nbr.shape[0] if n is None else n
    k = nbr.shape[1]
    indptr = np.arange(0, n * k + 1, k, dtype=np.int64)
    data = np.full(n * k, 1.0 / k, dtype=np.float64)
    return sp.csr_matrix((data, nbr.reshape(-1).astype(np.int64), indptr), shape=(n, n))


def w_moments(w: sp.csr_matrix) -> tuple[float, float, float]:
    """libpysal ``W.s0, W.s1, W.s2`` (needed by esda's analytic moments)."""
    s0 = float(w.sum())
    t = (w + w.T).tocsr()
    s1 = float((t.multiply(t)).sum() / 2.0)
    s2 = float((np.asarray(w.sum(1)).ravel() + np.asarray(w.sum(0)).ravel()) ** 2 @ np.ones(w.shape[0]))
    return s0, s1, s2
