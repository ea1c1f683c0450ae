"""GPU parity of the PCA stage (K6-K8) against the reference's own call, sklearn PCA(arpack)
(core.py:127).  Bound: subspace cosine >= 0.9999 (north_star), signs per sklearn's svd_flip."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import chrysalis_b200 as ch
from chrysalis_b200 import device as dv
from oracle import pca as opca

pytestmark = pytest.mark.gpu


def to_dense_dev(P):
    n, s = P.shape
    ld = dv.dense_ld(s)
    X = torch.zeros((n, ld), dtype=torch.float32, device="cuda")
    X[:, :s] = torch.from_numpy(np.ascontiguousarray(P, dtype=np.float32)).cuda()
    return X


def synthetic_svg_matrix(n, s, rank=8, seed=0, noise=0.6):
    rng = np.random.default_rng(seed)
    F = rng.normal(size=(n, rank)) * np.linspace(3.0, 0.8, rank)
    L = rng.normal(size=(rank, s))
    X = np.log1p(np.maximum(F @ L * 0.3 + 1.2 + rng.normal(0, noise, (n, s)), 0))
    return X.astype(np.float32)


@pytest.mark.parametrize("n,s", [(900, 64), (2000, 130), (4035, 1000), (5000, 257)])
def test_gram_matches_float64(n, s):
    P = synthetic_svg_matrix(n, s, seed=n)
    X = to_dense_dev(P)
    sums = dv.column_sums(X, s)
    assert np.allclose(sums.cpu().numpy(), P.astype(np.float64).sum(0), rtol=1e-12)
    mean = sums / n
    C = dv.gram_centered(X, s, mean).cpu().numpy()
    Z = P.astype(np.float64) - P.astype(np.float64).mean(0)
    ref = Z.T @ Z
    assert np.array_equal(C, C.T)
    scale = np.sqrt(np.outer(np.diag(ref), np.diag(ref)))
    # split-bf16 keeps 16 mantissa bits per operand and drops the lo*lo product: a bias of about
    # 2^-16 / 3 = 5e-6 of the diagonal (measured 0.9e-5), random elsewhere - far inside what the
    # 0.9999 subspace bound needs (the PCA tests below)
    assert np.abs(C - ref).max() / scale.max() < 2e-5
    assert np.abs(C - ref).max() / scale.max() > 0             # and it is not the fp64 reference itself


@pytest.mark.parametrize("n,s", [(900, 64), (2000, 130), (9000, 300), (20000, 640), (4035, 1000)])
def test_super_tile_gram_kernel_equals_the_tile_kernel_bit_for_bit(n, s):
    """The 256 x 256 super-tile kernel (default) issues, per 128 x 128 tile, the same 16-spot MMA slices in the same order as
    the tile kernel (variant 2): identical fp32 partials, identical fp64 sums.  300 columns = an odd number of 128-blocks
    (half-empty last super-tile), 640 = several diagonal and off-diagonal super-tiles."""
    X = to_dense_dev(synthetic_svg_matrix(n, s, seed=s))
    mean = dv.column_sums(X, s) / n
    a = dv.gram_centered(X, s, mean, variant=0)
    b = dv.gram_centered(X, s, mean, variant=2)
    assert torch.equal(a, b)


@pytest.mark.parametrize("s,r", [(64, 12), (300, 20), (1000, 50)])
def test_eigensolver_matches_lapack(s, r):
    rng = np.random.default_rng(s)
    A = rng.normal(size=(3 * s, s)) * np.linspace(4, 0.5, s)
    C = A.T @ A / (3 * s)
    w, v = np.linalg.eigh(C)
    w, v = w[::-1][:r], v[:, ::-1][:, :r].T
    evals, evecs, worst = dv.eigh_topk(torch.from_numpy(C).cuda(), r)
    assert worst < 1e-8
    assert np.allclose(evals.cpu().numpy(), w, rtol=1e-9)
    cos = np.abs(np.einsum("ij,ij->i", evecs.cpu().numpy(), v))
    assert cos.min() > 1 - 1e-8
    G = evecs.cpu().numpy() @ evecs.cpu().numpy().T
    assert np.abs(G - np.eye(r)).max() < 1e-12


def test_pca_matches_sklearn_arpack_golden(golden_pca_aa):
    z = golden_pca_aa
    pcs = z["pcs"]
    n, s = pcs.shape
    full = np.zeros((n, s + 7), dtype=np.float32)
    sel = np.zeros(s + 7, dtype=bool)
    cols = np.sort(np.random.default_rng(0).choice(s + 7, s, replace=False))
    full[:, cols] = pcs
    sel[cols] = True
    ad = ch.AnnDataLite(sp.csr_matrix(full), obsm={"spatial": np.zeros((n, 2))})
    ad.var["spatially_variable"] = sel
    ch.pca(ad, n_pcs=12)
    L = ad.uns["chr_pca"]["loadings"]
    assert L.shape == (12, s) and L.dtype == np.float32 and ad.obsm["chr_X_pca"].shape == (n, 12)
    assert opca.subspace_cosines(z["loadings"], L).min() >= 0.9999
    assert opca.component_cosines(z["loadings"][:6], L[:6]).min() >= 0.9999
    # sign convention of sklearn's svd_flip: compare signed on well separated components
    signed = np.einsum("ij,ij->i", z["loadings"][:6], L[:6])
    assert (signed > 0.9999).all()
    assert np.allclose(ad.uns["chr_pca"]["variance_ratio"], z["variance_ratio"], rtol=2e-4)
    sc = ad.obsm["chr_X_pca"][:, :6]
    assert np.abs(sc - z["X_pca"][:, :6]).max() < 2e-3 * np.abs(z["X_pca"][:, :6]).max()
    assert ad.uns["chr_pca"]["features"] == [f"gene{j}" for j in cols]


def test_pca_cfg1_shape_vs_sklearn():
    """cfg1-sized SVG matrix (4,035 x 1,000, 20 PCs): subspace agreement with the reference call."""
    P = synthetic_svg_matrix(4035, 1000, rank=10, seed=3)
    ref = opca.pca_arpack(P, 20)
    ad = ch.AnnDataLite(sp.csr_matrix(P), obsm={"spatial": np.zeros((4035, 2))})
    ad.var["spatially_variable"] = np.ones(1000, dtype=bool)
    ch.pca(ad, n_pcs=20)
    L = ad.uns["chr_pca"]["loadings"]
    assert opca.subspace_cosines(ref["loadings"], L).min() >= 0.9999
    assert opca.component_cosines(ref["loadings"][:10], L[:10]).min() >= 0.9999
    assert np.allclose(ad.uns["chr_pca"]["variance_ratio"], ref["variance_ratio"], rtol=1e-3)


def test_pca_errors_like_sklearn():
    P = synthetic_svg_matrix(50, 10)
    ad = ch.AnnDataLite(sp.csr_matrix(P), obsm={"spatial": np.zeros((50, 2))})
    ad.var["spatially_variable"] = np.ones(10, dtype=bool)
    with pytest.raises(ValueError):
        ch.pca(ad, n_pcs=10)                                    # arpack needs n_pcs < min(n, S)
    ad2 = ch.AnnDataLite(sp.csr_matrix(P), obsm={"spatial": np.zeros((50, 2))})
    with pytest.raises(KeyError):
        ch.pca(ad2, n_pcs=3)                                    # spatially_variable missing


def test_pca_properties_at_full_spot_count():
    """BASELINE cfg4 spot count (700k) with a 256-gene SVG slab: size-independent properties of the
    Gram -> eigensolver -> scores chain (the oracle would need minutes and 0.7 GB per call at this size)."""
    n, s, r = 700_000, 256, 12
    gen = torch.Generator(device="cuda").manual_seed(11)
    F = torch.randn((n, 6), device="cuda", generator=gen) * torch.linspace(3.0, 0.7, 6, device="cuda")
    L = torch.randn((6, s), device="cuda", generator=gen)
    P = torch.log1p((F @ L * 0.3 + 1.2 + 0.5 * torch.randn((n, s), device="cuda", generator=gen)).clamp_min(0)).contiguous()
    X = torch.zeros((n, dv.dense_ld(s)), device="cuda")
    X[:, :s] = P
    mean = dv.column_sums(X, s) / n
    assert torch.allclose(mean, P.double().mean(0), rtol=1e-12, atol=0)
    C = dv.gram_centered(X, s, mean)
    assert torch.equal(C, C.T)
    # linearity: the Gram of two halves adds up (split-bf16 rounding is per element, chains differ: tolerance 1e-5 of the diagonal)
    h = n // 2
    Ca = dv.gram_centered(X[:h], s, mean)
    Cb = dv.gram_centered(X[h:], s, mean)
    scale = float(torch.diagonal(C).max())
    assert float((Ca + Cb - C).abs().max()) < 5e-5 * scale
    # against a float64 Gram of a column block (exact reference for 32 genes)
    Zb = P[:, :32].double() - mean[:32]
    # at this depth the tensor core's fp32 accumulation (truncating adds, 512 MMAs per 8192-spot chain) shrinks
    # every entry by a few 1e-5 of its size - a nearly proportional error, harmless to eigenvectors and ratios
    assert float((C[:32, :32] - Zb.T @ Zb).abs().max()) < 1e-4 * scale
    evals, V, worst = dv.eigh_topk(C / (n - 1), r)
    assert worst < 1e-9
    G = V @ V.T
    assert float((G - torch.eye(r, device="cuda", dtype=torch.float64)).abs().max()) < 1e-12
    assert bool((evals[:-1] >= evals[1:]).all())
    Y = dv.pca_scores(X, s, mean, V)
    # scores are centred, uncorrelated, with variance = eigenvalue (float32 accumulation: 1e-3)
    Yd = Y.double()
    assert float(Yd.mean(0).abs().max()) < 1e-3 * float(evals[0].sqrt())
    cov = (Yd.T @ Yd) / (n - 1)
    assert torch.allclose(torch.diagonal(cov), evals, rtol=2e-3)
    off = cov - torch.diag(torch.diagonal(cov))
    assert float(off.abs().max()) < 2e-3 * float(evals[0])
