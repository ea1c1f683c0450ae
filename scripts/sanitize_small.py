"""Small end-to-end invocation of every kernel family for compute-sanitizer (memcheck): general W, lattice (dense and
CSR-fed), block-diagonal, permutations (both paths), LISA, PCA, AA, host upload paths."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chrysalis_b200 as ch  # noqa: E402
from chrysalis_b200 import device as dv  # noqa: E402
from chrysalis_b200 import synth  # noqa: E402
from oracle.svg import normalize_total_log1p  # noqa: E402

torch.cuda.set_device(0)
for kind, n, k in (("hex", 777, 6), ("random", 1201, 8), ("square", 450, 18)):
    xy = synth.make_coords(kind, n, seed=1)
    X, _ = synth.make_counts_csr(xy, 300, n_zones=5, seed=1)
    ad = ch.AnnDataLite(X, obsm={"spatial": xy})
    ch.detect_svgs(ad, neighbors=k, top_svg=40, geary=True)
    ch.morans_i_permutation(ad, neighbors=k, permutations=3, seed=0)
# square lattice with a hole, k = 8: CSR-fed and dense stencil kernels, lattice permutations, compact counts
xy = synth.make_coords("square", 70 * 70, seed=2)
c = xy.mean(0)
xy = xy[(np.hypot(*(xy - c).T) < 3100) & ~((np.abs(xy[:, 0] - c[0]) < 400) & (np.abs(xy[:, 1] - c[1]) < 300))]
X, _ = synth.make_counts_csr(xy, 300, n_zones=5, seed=2)
for Xi in (X, dv.CompactCSR.from_scipy(X)):
    ad = ch.AnnDataLite(Xi, obsm={"spatial": xy})
    ch.detect_svgs(ad, neighbors=8, top_svg=40, geary=True)
os.environ["CHR_MORAN_CSR"] = "0"
ad = ch.AnnDataLite(X, obsm={"spatial": xy})
ch.detect_svgs(ad, neighbors=8, top_svg=40, geary=True)
del os.environ["CHR_MORAN_CSR"]
ch.morans_i_permutation(ad, neighbors=8, permutations=5, seed=0)
ch.local_morans_i(ad, list(ad.var_names[:2]), neighbors=8, permutations=9, seed=1)
# block-diagonal multi-sample launch
samples = []
for s_i in range(3):
    xs = synth.make_coords("hex", 500 + 60 * s_i, seed=30 + s_i)
    Xs, _ = synth.make_counts_csr(xs, 200, n_zones=4, seed=30 + s_i)
    samples.append(ch.AnnDataLite(Xs, obsm={"spatial": xs}))
merged = ch.integrate_adatas(samples, sample_names=["a", "b", "c"], calculate_svgs=True, neighbors=6, top_svg=30)
ad.X = normalize_total_log1p(sp.csr_matrix(X))
ad.uns["log1p"] = {"base": None}
ch.pca(ad, n_pcs=8)
ch.aa(ad, n_archetypes=4, n_pcs=6, max_iter=5)
ch.estimate_compartments(ad, n_pcs=6, range_archetypes=(3, 5), max_iter=2)
torch.cuda.synchronize()
print("sanitize_small ok", ad.obsm["chr_aa"].shape, merged.shape)
