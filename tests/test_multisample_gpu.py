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
