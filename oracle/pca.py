"""``pca`` of the reference path (test infrastructure, see oracle/__init__.py).

``/root/reference/chrysalis/core.py:126-137``: densify the SVG columns and
call ``sklearn.decomposition.PCA(n_components=n_pcs, svd_solver='arpack',
random_state=42).fit_transform``.  scikit-learn IS installed here, so this
is the reference's own call, not a restatement (sklearn is unpinned in
``pyproject.toml:30``; the version in this image is recorded in the golden
fixtures).
"""
from __future__ import annotations

import numpy as np
from sklearn.decomposition import PCA


def pca_arpack(pcs: np.ndarray, n_pcs: int = 50) -> dict:
    """core.py:127-137 on a dense (N, S) SVG matrix."""
    model = PCA(n_components=n_pcs, svd_solver="arpack", random_state=42)
    scores = model.fit_transform(pcs)
    return {"X_pca": scores, "variance_ratio": model.explained_variance_ratio_,
            "loadings": model.components_, "mean": model.mean_,
            "explained_variance": model.explained_variance_}


def pca_covariance_eigh(pcs: np.ndarray, n_pcs: int = 50) -> dict:
    """Second oracle: sklearn's own Gram + eigh formulation (``svd_solver='covariance_eigh'``,
    sklearn >= 1.5) - the same factorisation the GPU path uses."""
    model = PCA(n_components=n_pcs, svd_solver="covariance_eigh")
    scores = model.fit_transform(pcs)
    return {"X_pca": scores, "variance_ratio": model.explained_variance_ratio_,
            "loadings": model.components_, "mean": model.mean_}


def subspace_cosines(V_ref: np.ndarray, V_test: np.ndarray) -> np.ndarray:
    """Cosines of the principal angles between the row spaces of two (r, S) loading matrices."""
    qa, _ = np.linalg.qr(np.asarray(V_ref, dtype=np.float64).T)
    qb, _ = np.linalg.qr(np.asarray(V_test, dtype=np.float64).T)
    return np.clip(np.linalg.svd(qa.T @ qb, compute_uv=False), 0.0, 1.0)


def component_cosines(V_ref: np.ndarray, V_test: np.ndarray) -> np.ndarray:
    """|cos| between matching components (sign-free)."""
    a = np.asarray(V_ref, dtype=np.float64)
    b = np.asarray(V_test, dtype=np.float64)
    return np.abs(np.einsum("ij,ij->i", a, b)) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))
