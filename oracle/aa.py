"""Archetypal analysis of the reference path (test infrastructure, see oracle/__init__.py).

``/root/reference/chrysalis/core.py:186-191`` calls
``archetypes.AA(n_archetypes, n_init=3, max_iter=max_iter, tol=0.001,
random_state=42).fit(X)``.  ``archetypes`` is pinned to 0.4.2
(``/root/reference/pyproject.toml:24``) but its source is not in the
reference tree and the package is not installed here -> **PARITY UNPINNED**.

This file restates the published 0.4.x algorithm (Cutler & Breiman
alternating non-negative least squares with a penalty row for the
sum-to-one constraint, solved with ``scipy.optimize.nnls``):

    for each of n_init restarts (consecutive draws of one RandomState):
        B0 (A x N one-hot rows), A0 (N x A one-hot rows);  Z = B0 @ X
        repeat <= max_iter:
            A = optimize(X, Z)            # N NNLS problems, (d+1) x A each
            Z = pinv(A) @ X
            B = optimize(Z, X)            # A NNLS problems, (d+1) x N each
            Z = B @ X
            rss = ||X - A @ Z||_F^2
            stop when |rss_prev - rss| < tol
    keep the restart with the smallest rss

Items that could not be verified offline are marked UNVERIFIED; the GPU path
is compared with this file *from matched initialisation*, which removes the
dependence on the UNVERIFIED initialisation details.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import nnls

PENALTY_M = 200.0  # UNVERIFIED constant of archetypes 0.4.x (`constant_values=200`)


def optimize_coefs(B: np.ndarray, A: np.ndarray, M: float = PENALTY_M) -> np.ndarray:
    """Rows of ``B`` as (approximately) convex combinations of rows of ``A``.

    Each row solves  min_{c >= 0} || [A, M]^T c - [b, M] ||_2  (Lawson-Hanson NNLS),
    then rows are renormalised to sum to one (UNVERIFIED, believed present).
    """
    Bp = np.pad(np.asarray(B, dtype=np.float64), ((0, 0), (0, 1)), "constant", constant_values=M)
    Ap = np.pad(np.asarray(A, dtype=np.float64), ((0, 0), (0, 1)), "constant", constant_values=M)
    coef = np.empty((Bp.shape[0], Ap.shape[0]))
    At = np.ascontiguousarray(Ap.T)
    for j in range(coef.shape[0]):
        coef[j], _ = nnls(At, Bp[j])
    coef = np.maximum(coef, 0)
    return coef / coef.sum(1)[:, None]


def furthest_sum(X: np.ndarray, n_archetypes: int, rs: np.random.RandomState) -> np.ndarray:
    """FurthestSum initialisation (Morup & Hansen 2012) - UNVERIFIED as the 0.4.2 default.

    Starts from a random point, greedily adds the point with the largest summed
    distance to the chosen set, then re-picks the first choices (10 extra rounds).
    """
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    x2 = np.einsum("ij,ij->i", X, X)
    chosen = [int(rs.randint(n))]
    avail = np.ones(n, dtype=bool)
    avail[chosen[0]] = False
    sum_dist = np.zeros(n)

    def dist_to(i):
        return np.sqrt(np.maximum(x2 - 2.0 * (X @ X[i]) + x2[i], 0.0))

    ind_t = chosen[0]
    for k in range(1, n_archetypes + 11):
        if k > n_archetypes - 1:
            first = chosen.pop(0)
            sum_dist -= dist_to(first)
            avail[first] = True
        sum_dist += dist_to(ind_t)
        cand = np.flatnonzero(avail)
        ind_t = int(cand[np.argmax(sum_dist[cand])])
        chosen.append(ind_t)
        avail[ind_t] = False
    return np.array(chosen[:n_archetypes], dtype=np.int64)


def init_coefs(X: np.ndarray, n_archetypes: int, rs: np.random.RandomState, algorithm_init: str = "auto"):
    """(alphas0 N x A, betas0 A x N) one-hot initial coefficients."""
    n = X.shape[0]
    if algorithm_init == "auto":
        algorithm_init = "furthest_sum"
    if algorithm_init == "random":
        ind = rs.choice(n, n_archetypes, replace=False)
    elif algorithm_init == "furthest_sum":
        ind = furthest_sum(X, n_archetypes, rs)
    else:
        raise ValueError(algorithm_init)
    B = np.zeros((n_archetypes, n))
    B[np.arange(n_archetypes), ind] = 1.0
    A = np.zeros((n, n_archetypes))
    A[np.arange(n), rs.choice(n_archetypes, n, replace=True)] = 1.0
    return A, B


def fit_single(X: np.ndarray, alphas0: np.ndarray, betas0: np.ndarray, max_iter: int = 200,
               tol: float = 1e-3, M: float = PENALTY_M) -> dict:
    """One restart from a given initialisation (the unit the GPU kernel is compared with)."""
    X = np.asarray(X, dtype=np.float64)
    A = np.asarray(alphas0, dtype=np.float64)
    B = np.asarray(betas0, dtype=np.float64)
    Z = B @ X
    rss_0 = np.inf
    rss = np.inf
    n_iter = 0
    for n_iter in range(max_iter):
        A = optimize_coefs(X, Z, M)
        Z = np.linalg.pinv(A) @ X
        B = optimize_coefs(Z, X, M)
        Z = B @ X
        rss = float(np.linalg.norm(X - A @ Z) ** 2)
        if abs(rss_0 - rss) < tol:
            break
        rss_0 = rss
    return {"alphas": A, "betas": B, "archetypes": Z, "rss": rss, "n_iter": n_iter + 1 if max_iter else 0}


def aa_fit(X: np.ndarray, n_archetypes: int, n_init: int = 3, max_iter: int = 200, tol: float = 1e-3,
           random_state: int = 42, algorithm_init: str = "auto", inits=None) -> dict:
    """``AA(n_archetypes, n_init, max_iter, tol, random_state).fit(X)``.

    ``inits``: optional list of (alphas0, betas0) to use instead of drawing them
    (matched-initialisation parity).  The drawn inits are returned under ``"inits"``.
    """
    rs = np.random.RandomState(random_state)
    best = None
    used = []
    for r in range(n_init):
        a0, b0 = inits[r] if inits is not None else init_coefs(X, n_archetypes, rs, algorithm_init)
        used.append((a0, b0))
        res = fit_single(X, a0, b0, max_iter, tol)
        if best is None or res["rss"] < best["rss"]:
            best = res
    best = dict(best)
    best["inits"] = used
    return best
