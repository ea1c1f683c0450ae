"""Parity at the FULL shapes of BASELINE.json's configs, through the public API:

* cfg2  Slide-seq-shaped   40,000 random beads x 20,000 genes, k = 8: detect_svgs -> pca(20) -> aa(8)
* cfg5  8-sample cohort    8 x 5,000 hex spots x 30,000 genes, k = 6: integrate_adatas (block-diagonal W)
* exact (un-jittered) square and hex lattices with holes: the device kNN breaks distance ties by spot id, the oracle's
  cKDTree by traversal order - how far apart the two selections are is measured and bounded here.

(cfg1 is covered at its spot count in tests/test_moran_gpu.py, cfg3 with its permutations in tests/test_moran_perm_gpu.py,
cfg4 by bench.py's parity block and the size-independent properties in tests/test_moran_gpu.py.)
The oracle scores every gene (vectorised float64 statement of the same statistic, oracle.moran.morans_I_batched, in
column chunks); the counts are generated on the device because 10^9 Poisson draws take minutes on the host."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import chrysalis_b200 as ch
from chrysalis_b200 import core, synth
from oracle import aa as oaa
from oracle import moran as om
from oracle import pca as opca
from oracle import svg as osvg
from oracle import weights as ow

pytestmark = pytest.mark.gpu


def oracle_svgs(X, xy, neighbors, top_svg=1000, min_morans=0.2, min_spots=0.05, chunk=1024):
    """``oracle.svg.detect_svgs`` with the gene loop replaced by the chunked vectorised statement (float64 arithmetic on
    the float32 normalised values - the 'truth' both the reference's float32 path and the kernel are judged by)."""
    X = sp.csr_matrix(X)
    n, G = X.shape
    mask = osvg.filter_genes_mask(X, int(n * min_spots))
    Xn = osvg.normalize_total_log1p(X[:, mask]).tocsc()
    nbr = ow.knn_neighbors(ow.reference_points(xy), neighbors)
    I = np.empty(Xn.shape[1])
    for c0 in range(0, Xn.shape[1], chunk):
        I[c0:c0 + chunk] = om.morans_I_batched(np.asarray(Xn[:, c0:c0 + chunk].todense(), dtype=np.float32), nbr)
    morans = np.full(G, np.nan)
    morans[mask] = I
    return {"morans": morans, "gene_mask": mask, "nbr": nbr, "spatially_variable": osvg.select_svgs(morans, mask, top_svg, min_morans)}


def assert_morans(got, ref, what=""):
    m = ref["gene_mask"]
    assert np.array_equal(np.isnan(got), ~m), f"{what}: filtered-gene mask differs"
    err = np.abs(got[m] - ref["morans"][m])
    bad = err > 1e-5 * np.abs(ref["morans"][m]) + 2e-7
    assert not bad.any(), f"{what}: {bad.sum()} of {m.sum()} genes off, worst {err.max():.3e}"


def test_cfg2_slideseq_shape_svg_pca_aa():
    n, g = 40_000, 20_000
    xy = synth.make_coords("random", n, seed=2)
    X, _ = synth.make_counts_csr(xy, g, n_zones=8, seed=2, device="cuda", log_rate_mean=-2.6, log_rate_sd=1.3)
    ad = ch.AnnDataLite(X, obsm={"spatial": xy})
    ch.detect_svgs(ad, neighbors=8)                                           # reference defaults: min_spots 0.05, top_svg 1000, min_morans 0.2
    ref = oracle_svgs(X, xy, 8)
    assert ref["gene_mask"].sum() > 5_000
    assert_morans(ad.var["Moran's I"].to_numpy(), ref, "cfg2")
    assert np.array_equal(ad.var["spatially_variable"].to_numpy(), ref["spatially_variable"])
    sel = ref["spatially_variable"]
    assert 50 < sel.sum() <= 1000
    # README flow: normalise the full matrix, then PCA of the SVG columns and AA on the scores
    ad.X = osvg.normalize_total_log1p(X)
    ad.uns["log1p"] = {"base": None}
    ch.pca(ad, n_pcs=20)
    ref_p = opca.pca_arpack(np.asarray(ad.X[:, sel].todense(), dtype=np.float32), 20)      # the reference's own sklearn call
    assert opca.subspace_cosines(ref_p["loadings"], ad.uns["chr_pca"]["loadings"]).min() >= 0.9999
    assert opca.component_cosines(ref_p["loadings"][:5], ad.uns["chr_pca"]["loadings"][:5]).min() >= 0.9999
    assert np.allclose(ad.uns["chr_pca"]["variance_ratio"], ref_p["variance_ratio"], rtol=1e-3)
    # AA from matched initialisation, a few iterations (the oracle's scipy.optimize.nnls loop takes ~1 s per iteration here)
    emb = np.asarray(ad.obsm["chr_X_pca"][:, :20])
    rows = np.random.default_rng(0).choice(n, 8, replace=False)
    fit = core.aa_stage(emb, 8, max_iter=3, n_init=1, init_rows=[rows], tol=0.0)
    B0 = np.zeros((8, n))
    B0[np.arange(8), rows] = 1.0
    oref = oaa.fit_single(emb.astype(np.float64), np.zeros((n, 8)), B0, max_iter=3, tol=0.0)
    assert np.abs(fit["alphas"] - oref["alphas"]).max() < 1e-3
    assert fit["rss"] == pytest.approx(oref["rss"], rel=1e-6)
    ch.aa(ad, n_archetypes=8, n_pcs=20)                                        # the full default fit runs and is a valid decomposition
    a = ad.obsm["chr_aa"]
    assert a.shape == (n, 8) and (a >= 0).all() and np.allclose(a.sum(1), 1.0, atol=1e-9)
    assert ad.uns["chr_aa"]["loadings"].shape == (8, int(sel.sum()))


def test_cfg5_cohort_shape_block_diagonal():
    n_s, g, B = 5_000, 30_000, 8
    samples, refs = [], []
    for b in range(B):
        xy = synth.make_coords("hex", n_s, seed=50 + b)
        X, _ = synth.make_counts_csr(xy, g, n_zones=8, seed=50 + b, device="cuda", log_rate_mean=-2.4, log_rate_sd=1.3)
        samples.append(ch.AnnDataLite(X, obsm={"spatial": xy}))
        refs.append(oracle_svgs(X, xy, 6))
    ad = ch.integrate_adatas(samples, sample_names=[f"s{b}" for b in range(B)], calculate_svgs=True, neighbors=6)
    assert ad.shape == (B * n_s, g)
    union = np.zeros(g, dtype=bool)
    for b, ref in enumerate(refs):
        assert_morans(ad.varm["Moran's I"][f"Moran's I_s{b}"].to_numpy(), ref, f"cfg5 sample {b}")
        assert np.array_equal(ad.varm["spatially_variable"][f"spatially_variable_s{b}"].to_numpy(bool), ref["spatially_variable"])
        union |= ref["spatially_variable"]
    assert np.array_equal(ad.var["spatially_variable"].to_numpy(), union)
    # joint PCA on the concatenated, normalised object (test.py:114-122 flow)
    ad.X = osvg.normalize_total_log1p(sp.csr_matrix(ad.X))
    ad.uns["log1p"] = {"base": None}
    ch.pca(ad, n_pcs=20)
    ref_p = opca.pca_arpack(np.asarray(ad.X[:, union].todense(), dtype=np.float32), 20)
    assert opca.subspace_cosines(ref_p["loadings"], ad.uns["chr_pca"]["loadings"]).min() >= 0.9999


def _blob(xy, pitch_units):
    c = xy.mean(0)
    d = xy - c
    ang = np.arctan2(d[:, 1], d[:, 0])
    rad = np.hypot(d[:, 0], d[:, 1]) / (1.0 + 0.2 * np.sin(3 * ang + 0.5))
    keep = rad < 0.42 * pitch_units
    keep &= ~((np.abs(d[:, 0] - 0.09 * pitch_units) < 0.045 * pitch_units) & (np.abs(d[:, 1] + 0.03 * pitch_units) < 0.025 * pitch_units))
    return xy[keep]


@pytest.mark.parametrize("kind,k", [("square", 8), ("hex", 6)])
def test_exact_lattice_ties_selection_flips_vs_ckdtree(kind, k, capsys):
    """On an exact lattice every tissue-edge bin has tied candidates for its last neighbours.  knn.cu takes the lower spot
    id, cKDTree whatever its traversal meets first, so W differs in a few edge rows.  Measured here: how many edge spots
    get another neighbour set, how far Moran's I moves, how many SVG selections flip (printed; bounded)."""
    side = 160
    xy = _blob(synth.make_coords(kind, side * side, seed=3, jitter=0.0), side * 100.0)
    n = len(xy)
    X, _ = synth.make_counts_csr(xy, 3000, n_zones=6, seed=9, device="cuda", log_rate_mean=-2.0)
    ad = ch.AnnDataLite(X, obsm={"spatial": xy})
    ch.detect_svgs(ad, neighbors=k, top_svg=300, min_morans=0.05)
    ref = oracle_svgs(X, xy, k, top_svg=300, min_morans=0.05)
    # the device's own W
    from chrysalis_b200 import device as dv
    pts = ow.reference_points(xy)
    nbr_dev = dv.knn_build(dv.to_device(pts, torch.float64), k).cpu().numpy()
    rows_differ = int((np.sort(nbr_dev, 1) != np.sort(ref["nbr"], 1)).any(1).sum())
    d_dev = np.sort(np.linalg.norm(pts[nbr_dev] - pts[:, None], axis=2), 1)
    d_ref = np.sort(np.linalg.norm(pts[ref["nbr"]] - pts[:, None], axis=2), 1)
    assert np.allclose(d_dev, d_ref, rtol=0, atol=1e-9), "not a tie: the neighbour DISTANCES must agree everywhere"
    got, m = ad.var["Moran's I"].to_numpy(), ref["gene_mask"]
    assert np.array_equal(np.isnan(got), ~m)
    dI = np.abs(got[m] - ref["morans"][m])
    flips = int((ad.var["spatially_variable"].to_numpy() != ref["spatially_variable"]).sum())
    n_sel = int(ref["spatially_variable"].sum())
    with capsys.disabled():
        print(f"\n[{kind} lattice, k={k}, {n} spots] rows with another tied neighbour set: {rows_differ} ({100.0 * rows_differ / n:.2f} %), "
              f"max |dI| {dI.max():.2e}, median |dI| {np.median(dI):.2e}, SVG selection flips: {flips} of {n_sel}")
    assert rows_differ < 0.08 * n                     # only edge rows
    assert dI.max() < 5e-3                            # W differs in O(perimeter) rows: I moves by O(perimeter / n)
    assert flips <= max(2, 0.02 * n_sel)
    # with the oracle's W handed to the kernel the statistic itself is the oracle's (ties are the whole difference)
    dense = np.asarray(osvg.normalize_total_log1p(X[:, m]).todense(), dtype=np.float32)[:, :256]
    Xd = torch.zeros((n, dv.dense_ld(256)), dtype=torch.float32, device="cuda")
    Xd[:, :256] = torch.from_numpy(dense).cuda()
    I_w, _, _ = dv.moran_dense(Xd, torch.from_numpy(ref["nbr"].astype(np.int32)).cuda(), 256)
    r256 = ref["morans"][m][:256]
    assert np.all(np.abs(I_w.cpu().numpy() - r256) <= 1e-5 * np.abs(r256) + 2e-7)
