"""GPU parity of the archetypal-analysis stage (K9-K11) against the CPU restatement of
archetypes 0.4.x (oracle/aa.py), from matched initialisation.  Bound (north_star): archetype
proportions within 1e-3 absolute."""
import numpy as np
import pytest
import torch

import chrysalis_b200 as ch
from chrysalis_b200 import core
from chrysalis_b200 import device as dv
from oracle import aa as oaa

pytestmark = pytest.mark.gpu


def simplex_data(n, d, a, seed, noise=0.05):
    """Points near the convex hull of `a` random vertices in d dimensions (what PCA scores of zoned tissue look like)."""
    rng = np.random.default_rng(seed)
    V = rng.normal(size=(a, d)) * 3.0
    W = rng.dirichlet(np.full(a, 0.4), size=n)
    return (W @ V + rng.normal(0, noise, (n, d))).astype(np.float32)


def one_hot_betas(rows, n):
    B = np.zeros((len(rows), n))
    B[np.arange(len(rows)), rows] = 1.0
    return B


@pytest.mark.parametrize("n,d,a", [(600, 5, 4), (1500, 8, 6), (4035, 20, 8), (900, 24, 20), (700, 40, 33)])
def test_single_iteration_matches_oracle(n, d, a):
    X = simplex_data(n, d, a, seed=n)
    rows = np.random.default_rng(1).choice(n, a, replace=False)
    ref = oaa.fit_single(X, np.zeros((n, a)), one_hot_betas(rows, n), max_iter=1)
    st = dv.AAState(dv.to_device(X.astype(np.float64), torch.float64), a)
    st.init_from_rows(rows)
    rss = st.iterate()
    assert np.abs(st.alphas.cpu().numpy() - ref["alphas"]).max() < 1e-6
    assert np.abs(st.Z.cpu().numpy() - ref["archetypes"]).max() < 1e-6 * np.abs(ref["archetypes"]).max()
    assert abs(rss - ref["rss"]) < 1e-6 * ref["rss"]
    al = st.alphas.cpu().numpy()
    assert (al >= 0).all() and np.allclose(al.sum(1), 1.0, atol=1e-12)


def blob_data(n, d, seed):
    """Anisotropic Gaussian scores (no simplex structure): most spots need many archetypes, the supports keep changing."""
    rng = np.random.default_rng(seed)
    return rng.normal(size=(n, d)) * (3.0 / np.sqrt(1.0 + np.arange(d)))


@pytest.mark.parametrize("n,d,a", [(777, 30, 12), (1300, 11, 3), (2049, 16, 16), (1000, 9, 10), (130, 40, 13), (3000, 20, 8),
                                   (300, 64, 16), (64, 3, 2), (129, 2, 5)])
def test_alpha_step_variants_agree_with_each_other_and_the_oracle(n, d, a, monkeypatch):
    """The branch-free masked-Cholesky alpha kernel (n_archetypes <= 16) against the general active-set kernel
    (CHR_AA_ALPHA_GENERAL=1) and against the oracle: the NNLS minimisers are unique, so all agree to rounding."""
    X = blob_data(n, d, seed=n + a)
    rows = np.random.default_rng(3).choice(n, a, replace=False)
    Xd = dv.to_device(X, torch.float64)
    runs = {}
    for variant, env in (("default", None), ("general", "CHR_AA_ALPHA_GENERAL")):
        monkeypatch.delenv("CHR_AA_ALPHA_GENERAL", raising=False)
        if env:
            monkeypatch.setenv(env, "1")
        st = dv.AAState(Xd, a)
        st.init_from_rows(rows)
        rss = [st.iterate() for _ in range(4)]
        runs[variant] = (st.alphas.cpu().numpy(), st.Z.cpu().numpy(), rss)
    monkeypatch.delenv("CHR_AA_ALPHA_GENERAL", raising=False)
    al, Z, rss = runs["default"]
    assert (al >= 0).all() and np.allclose(al.sum(1), 1.0, atol=1e-12)
    for other in ("general",):
        assert np.abs(al - runs[other][0]).max() < 1e-8, other
        assert np.abs(Z - runs[other][1]).max() < 1e-8 * np.abs(Z).max(), other
        assert np.allclose(rss, runs[other][2], rtol=1e-9), other
    ref = oaa.fit_single(X, np.zeros((n, a)), one_hot_betas(rows, n), max_iter=4, tol=0.0)
    assert np.abs(al - ref["alphas"]).max() < 1e-6
    assert abs(rss[-1] - ref["rss"]) < 1e-6 * ref["rss"]


@pytest.mark.parametrize("n,d,a,iters", [(900, 8, 5, 30), (2000, 12, 7, 25)])
def test_fit_from_matched_init(n, d, a, iters):
    X = simplex_data(n, d, a, seed=7 * n)
    rows = [np.random.default_rng(s).choice(n, a, replace=False) for s in range(3)]
    inits = [(np.zeros((n, a)), one_hot_betas(r, n)) for r in rows]
    ref = oaa.aa_fit(X, a, n_init=3, max_iter=iters, inits=inits)
    got = core.aa_stage(X, a, max_iter=iters, init_rows=rows)
    assert got["n_iter"] == ref["n_iter"]
    assert abs(got["rss"] - ref["rss"]) < 1e-6 * ref["rss"]
    assert np.abs(got["alphas"] - ref["alphas"]).max() < 1e-3          # north_star bound
    assert np.abs(got["archetypes"] - ref["archetypes"]).max() < 1e-3 * np.abs(ref["archetypes"]).max()


def test_golden_fixture(golden_pca_aa):
    z = golden_pca_aa
    X = z["X_pca"][:, :8]
    rows = [np.argmax(b, axis=1) for b in z["betas0"]]
    got = core.aa_stage(X, 5, max_iter=30, init_rows=rows)
    assert np.abs(got["alphas"] - z["alphas"]).max() < 1e-3
    assert abs(got["rss"] - float(z["rss"])) < 1e-6 * float(z["rss"])
    assert got["n_iter"] == int(z["n_iter"])


def test_default_initialisation_follows_the_oracle_rng():
    """No init given: FurthestSum draws from RandomState(42) exactly like the restatement."""
    X = simplex_data(800, 6, 4, seed=3)
    ref = oaa.aa_fit(X, 4, n_init=3, max_iter=15)
    got = core.aa_stage(X, 4, max_iter=15)
    for (a0, b0), rows in zip(ref["inits"], got["init_rows"]):
        assert np.array_equal(np.argmax(b0, axis=1), rows)
    assert abs(got["rss"] - ref["rss"]) < 1e-6 * ref["rss"]
    # restarts that reach the same optimum differ by a relabelling of the archetypes and tie in rss to
    # rounding, so the best-of-3 pick may be another labelling: compare after matching columns
    perm = [int(np.argmin(np.abs(got["archetypes"] - z).sum(1))) for z in ref["archetypes"]]
    assert sorted(perm) == list(range(4))
    assert np.abs(got["alphas"][:, perm] - ref["alphas"]).max() < 1e-3


def test_aa_drop_in_fields():
    n, s, d, a = 700, 40, 10, 4
    rng = np.random.default_rng(0)
    scores = simplex_data(n, d, a, seed=11)
    ad = ch.AnnDataLite(np.zeros((n, s), dtype=np.float32), obsm={"spatial": np.zeros((n, 2)), "chr_X_pca": scores})
    ad.uns["chr_pca"] = {"loadings": rng.normal(size=(d, s)).astype(np.float32)}
    ch.aa(ad, n_archetypes=a, n_pcs=6, max_iter=10)
    assert ad.obsm["chr_aa"].shape == (n, a)
    u = ad.uns["chr_aa"]
    assert u["archetypes"].shape == (a, 6) and u["alphas"].shape == (n, a) and u["loadings"].shape == (a, s)
    assert np.allclose(u["loadings"], u["archetypes"] @ ad.uns["chr_pca"]["loadings"][:6, :])
    assert np.isfinite(u["RSS"]) and u["RSS"] > 0
    # default n_pcs = n_archetypes - 1, alternative embedding via pca_key
    ad.obsm["other"] = scores * 2.0
    ch.aa(ad, n_archetypes=a, pca_key="other", max_iter=5)
    assert ad.uns["chr_aa"]["archetypes"].shape == (a, a - 1)


@pytest.mark.parametrize("n,d,a", [(500, 7, 3), (5000, 20, 8), (20000, 12, 15)])
def test_device_furthest_sum_matches_host_statement(n, d, a):
    X = np.random.default_rng(n).normal(size=(n, d))
    rs = np.random.RandomState(42)
    ref = oaa.furthest_sum(X, a, rs)
    first = int(np.random.RandomState(42).randint(n))
    got = dv.aa_furthest_sum(dv.to_device(X, torch.float64), a, first)
    assert np.array_equal(got, ref)
    assert np.array_equal(got, core._furthest_sum(X, a, np.random.RandomState(42)))


def test_aa_properties_at_large_n():
    """200k spots x 20 PCs, 8 archetypes: simplex constraints, monotone RSS of the alternating minimisation,
    archetypes inside the data's bounding box, equivariance to relabelling the spots."""
    n, d, a = 200_000, 20, 8
    X = simplex_data(n, d, a, seed=1).astype(np.float64)
    rows = np.random.default_rng(0).choice(n, a, replace=False)
    Xd = dv.to_device(X, torch.float64)
    st = dv.AAState(Xd, a)
    st.init_from_rows(rows)
    rss = [st.iterate() for _ in range(8)]
    assert all(r2 <= r1 * (1 + 1e-9) for r1, r2 in zip(rss, rss[1:]))
    al = st.alphas.cpu().numpy()
    assert (al >= 0).all() and np.allclose(al.sum(1), 1.0, atol=1e-12)
    Z = st.Z.cpu().numpy()
    assert (Z >= X.min(0) - 1e-9).all() and (Z <= X.max(0) + 1e-9).all()      # convex combinations of data points
    # the residual the kernel reports is the residual of its own factors
    assert abs(rss[-1] - np.linalg.norm(X - al @ Z) ** 2) < 1e-8 * rss[-1]
    # relabelling the spots permutes the rows of alphas and leaves the archetypes alone
    perm = np.random.default_rng(1).permutation(n)
    inv = np.empty(n, dtype=np.int64); inv[perm] = np.arange(n)
    st2 = dv.AAState(dv.to_device(X[perm], torch.float64), a)
    st2.init_from_rows(inv[rows])
    rss2 = [st2.iterate() for _ in range(8)]
    assert abs(rss2[-1] - rss[-1]) < 1e-6 * rss[-1]
    assert np.abs(st2.Z.cpu().numpy() - Z).max() < 1e-6 * np.abs(Z).max()
    assert np.abs(st2.alphas.cpu().numpy() - al[perm]).max() < 1e-6


@pytest.mark.parametrize("n,d,a", [(777, 30, 12), (40_000, 20, 8), (1300, 11, 3), (300, 64, 16), (2049, 16, 20)])
def test_rss_from_the_gram_sums_equals_the_residual_pass(n, d, a, monkeypatch):
    """||X - A Z||^2 as ||X||^2 - 2 <A^T X, Z> + <A^T A, Z Z^T> (default: no pass over X) against the per-spot residual sum
    (CHR_AA_RSS_DIRECT=1): the same state, RSS equal to 1e-14 ||X||^2 - five orders below the fit's 1e-3 stop rule."""
    X = blob_data(n, d, seed=n + a) if n < 20_000 else simplex_data(n, d, a, seed=n).astype(np.float64)
    rows = np.random.default_rng(2).choice(n, a, replace=False)
    Xd = dv.to_device(X, torch.float64)
    runs = {}
    for variant in ("gram", "direct"):
        monkeypatch.delenv("CHR_AA_RSS_DIRECT", raising=False)
        if variant == "direct":
            monkeypatch.setenv("CHR_AA_RSS_DIRECT", "1")
        st = dv.AAState(Xd, a)
        st.init_from_rows(rows)
        rss = [st.iterate() for _ in range(10)]
        runs[variant] = (st.alphas.clone(), st.Z.clone(), np.array(rss))
    monkeypatch.delenv("CHR_AA_RSS_DIRECT", raising=False)
    assert torch.equal(runs["gram"][0], runs["direct"][0]) and torch.equal(runs["gram"][1], runs["direct"][1])
    assert np.abs(runs["gram"][2] - runs["direct"][2]).max() <= 1e-14 * float((X ** 2).sum())


def test_concurrent_restarts_equal_sequential_restarts():
    """The restarts of a fit run on their own streams; every one of them must be the launch chain a sequential loop makes."""
    X = simplex_data(30_000, 12, 6, seed=5).astype(np.float64)
    Xd = dv.to_device(X, torch.float64)
    seq = core._aa_fit_device(Xd, 6, 25, 3, 0.001, 42, restart_workers=1)
    par = core._aa_fit_device(Xd, 6, 25, 3, 0.001, 42)
    assert all(np.array_equal(a, b) for a, b in zip(seq["init_rows"], par["init_rows"]))
    assert seq["rss"] == par["rss"] and seq["n_iter"] == par["n_iter"]
    assert torch.equal(seq["alphas"], par["alphas"]) and torch.equal(seq["archetypes"], par["archetypes"])
