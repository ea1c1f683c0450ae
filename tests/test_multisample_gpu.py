"""cfg5-shaped parity: several hex samples, per-sample Moran's I with per-sample W (block-diagonal W overall),
union of the SVG sets, joint PCA + AA - against the oracle run sample by sample (utils.py:161-250 flow)."""
import numpy as np
import pytest
import scipy.sparse as sp

import chrysalis_b200 as ch
from chrysalis_b200 import synth
from oracle import pca as opca
from oracle import svg as osvg

pytestmark = pytest.mark.gpu


def test_integrate_then_joint_pca_aa():
    samples, refs = [], []
    g = 400
    for s_i in range(3):
        xy = synth.make_coords("hex", 700 + 50 * s_i, seed=20 + s_i)
        X, _ = synth.make_counts_csr(xy, g, n_zones=5, seed=20 + s_i)
        samples.append(ch.AnnDataLite(X, obsm={"spatial": xy}))
        refs.append(osvg.detect_svgs(X, xy, neighbors=6, top_svg=50))
    ad = ch.integrate_adatas(samples, sample_names=["s0", "s1", "s2"], calculate_svgs=True, neighbors=6, top_svg=50)
    n_total = sum(len(s) for s in samples)
    assert ad.shape == (n_total, g)
    mi = ad.varm["Moran's I"]
    union = np.zeros(g, dtype=bool)
    for s_i, ref in enumerate(refs):
        m = ref["gene_mask"]
        got = mi[f"Moran's I_s{s_i}"].to_numpy()
        assert np.array_equal(np.isnan(got), ~m)                         # per-sample filter, per-sample centring
        assert np.all(np.abs(got[m] - ref["morans"][m]) <= 1e-5 * np.abs(ref["morans"][m]) + 2e-7)
        assert np.array_equal(ad.varm["spatially_variable"][f"spatially_variable_s{s_i}"].to_numpy(bool), ref["spatially_variable"])
        union |= ref["spatially_variable"]
    assert np.array_equal(ad.var["spatially_variable"].to_numpy(), union)

    # joint PCA + AA on the concatenated, normalised object (test.py:114-122 flow)
    ad.X = osvg.normalize_total_log1p(sp.csr_matrix(ad.X))
    ad.uns["log1p"] = {"base": None}
    ch.pca(ad, n_pcs=10)
    ref_p = opca.pca_arpack(np.asarray(ad.X[:, union].todense(), dtype=np.float32), 10)
    assert opca.subspace_cosines(ref_p["loadings"], ad.uns["chr_pca"]["loadings"]).min() >= 0.9999
    ch.aa(ad, n_archetypes=5, n_pcs=8, max_iter=15)
    a = ad.obsm["chr_aa"]
    assert a.shape == (n_total, 5) and (a >= 0).all() and np.allclose(a.sum(1), 1.0, atol=1e-9)
    ch.estimate_compartments(ad, n_pcs=8, range_archetypes=(3, 6), max_iter=3)
    assert ad.obsm["entropy"].shape == (n_total, 3) and set(ad.uns["chr_aa"]["RSSs"]) == {3, 4, 5}
    rss = ad.uns["chr_aa"]["RSSs"]
    assert rss[3] > rss[4] > rss[5] > 0                                   # more archetypes fit better


def test_blockdiag_kernel_matches_the_per_sample_oracle():
    """chr_moran_blockdiag_f32 through the C ABI: blocks of different sizes (shorter and longer than one CTA row range,
    an exact multiple of it), W from the oracle's cKDTree kNN per sample, one launch - against the oracle run per sample."""
    import torch
    from chrysalis_b200 import device as dv
    from oracle import moran as om
    from oracle import weights as ow

    rng = np.random.default_rng(0)
    sizes, k, g = [300, 448, 1500, 2 * 448, 77], 6, 150
    block_rows, block_n, pad_rows = dv.block_layout(sizes)
    n_rows = int(block_rows[-1])
    assert n_rows % dv.BLOCK_ALIGN_ROWS == 0 and len(pad_rows) == n_rows - sum(sizes)
    dense = torch.full((n_rows, dv.dense_ld(g)), 123.0, dtype=torch.float32, device="cuda")       # garbage in the padding rows
    nbr_all = torch.arange(n_rows, dtype=torch.int32, device="cuda")[:, None].repeat(1, k)
    refs = []
    for b, n_b in enumerate(sizes):
        xy = rng.uniform(0, 1000, (n_b, 2))
        X = (np.log1p(rng.poisson(0.7, (n_b, g))) + np.sin(xy[:, :1] / 150.0) * rng.uniform(0, 1, g)[None] + 3.0 * b).astype(np.float32)
        nbr = ow.knn_neighbors(xy, k)
        r0 = int(block_rows[b])
        dense[r0:r0 + n_b, :g] = torch.from_numpy(X).cuda()
        nbr_all[r0:r0 + n_b] = torch.from_numpy(nbr.astype(np.int32)).cuda() + r0
        refs.append((om.morans_I_batched(X, nbr), om.gearys_C_batched(X, nbr)))
    plan = dv.moran_plan(nbr_all.contiguous())
    I, C = dv.moran_blockdiag(dense, plan, block_rows, block_n, torch.from_numpy(pad_rows).cuda(), g, geary=True)
    I, C = I.cpu().numpy(), C.cpu().numpy()
    for b, (rI, rC) in enumerate(refs):
        assert np.all(np.abs(I[b] - rI) <= 1e-5 * np.abs(rI) + 2e-7), f"block {b}: {np.abs(I[b] - rI).max():.3e}"
        assert np.all(np.abs(C[b] - rC) <= 1e-5 * np.abs(rC) + 2e-7), f"block {b} (C)"
    # a second call on the same buffer (its padding rows now hold the shift value) reproduces the first bit for bit
    I2, _ = dv.moran_blockdiag(dense, plan, block_rows, block_n, torch.from_numpy(pad_rows).cuda(), g, geary=True)
    assert np.array_equal(I2.cpu().numpy(), I)


def test_detect_svgs_multi_equals_single_calls():
    """One block-diagonal launch for all samples == ``detect_svgs`` on each sample (same filter, W, centring, selection)."""
    samples = []
    for s_i in range(3):
        xy = synth.make_coords("hex", 600 + 130 * s_i, seed=40 + s_i)
        X, _ = synth.make_counts_csr(xy, 300, n_zones=4, seed=40 + s_i)
        samples.append((X, xy))
    multi = [ch.AnnDataLite(X, obsm={"spatial": xy}) for X, xy in samples]
    from chrysalis_b200 import core
    core.detect_svgs_multi(multi, neighbors=6, top_svg=40, geary=True)
    for (X, xy), ad_m in zip(samples, multi):
        ad_s = ch.AnnDataLite(X, obsm={"spatial": xy})
        ch.detect_svgs(ad_s, neighbors=6, top_svg=40, geary=True)
        a, b = ad_s.var["Moran's I"].to_numpy(), ad_m.var["Moran's I"].to_numpy()
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.allclose(a, b, rtol=1e-6, atol=2e-8, equal_nan=True)            # another summation tree, same statistic
        assert np.allclose(ad_s.var["Geary's C"].to_numpy(), ad_m.var["Geary's C"].to_numpy(), rtol=1e-6, atol=2e-8, equal_nan=True)
        assert np.array_equal(ad_s.var["spatially_variable"].to_numpy(), ad_m.var["spatially_variable"].to_numpy())


def test_estimate_compartments_matches_the_oracle_fits():
    """The concurrent device sweep against ``archetypes``-style fits of the oracle (same seeding: RandomState(42) +
    FurthestSum per archetype count, 3 restarts, best RSS) - RSS and per-spot entropies."""
    from scipy.stats import entropy
    from oracle import aa as oaa
    rng = np.random.default_rng(3)
    n, d = 900, 8
    Z = rng.normal(size=(5, d)) * 3
    A = rng.dirichlet(np.ones(5) * 0.4, n)
    emb = (A @ Z + 0.05 * rng.normal(size=(n, d))).astype(np.float32)
    ad = ch.AnnDataLite(np.zeros((n, 3), np.float32), obsm={"spatial": np.zeros((n, 2)), "chr_X_pca": emb})
    ch.estimate_compartments(ad, n_pcs=d, range_archetypes=(3, 8), max_iter=4)
    for j, a in enumerate(range(3, 8)):
        ref = oaa.aa_fit(emb.astype(np.float64), a, n_init=3, max_iter=4)
        assert ad.uns["chr_aa"]["RSSs"][a] == pytest.approx(ref["rss"], rel=1e-6)
        assert np.abs(ad.obsm["entropy"][:, j] - entropy(ref["alphas"].T)).max() < 1e-5
