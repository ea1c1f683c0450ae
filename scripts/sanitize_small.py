"""Small end-to-end invocation of every kernel family for compute-sanitizer (memcheck)."""
import os, sys
import numpy as np, scipy.sparse as sp, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chrysalis_b200 as ch
from chrysalis_b200 import synth, device as dv
torch.cuda.set_device(0)
for kind, n, k in (("hex", 777, 6), ("random", 1201, 8), ("square", 450, 18)):
    xy = synth.make_coords(kind, n, seed=1)
    X, _ = synth.make_counts_csr(xy, 300, n_zones=5, seed=1)
    ad = ch.AnnDataLite(X, obsm={"spatial": xy})
    ch.detect_svgs(ad, neighbors=k, top_svg=40, geary=True)
    ch.morans_i_permutation(ad, neighbors=k, permutations=3, seed=0)
from oracle.svg import normalize_total_log1p
ad.X = normalize_total_log1p(sp.csr_matrix(ad.X)); ad.uns["log1p"] = {"base": None}
ch.pca(ad, n_pcs=8)
ch.aa(ad, n_archetypes=4, n_pcs=6, max_iter=5)
torch.cuda.synchronize()
print("sanitize_small ok", ad.obsm["chr_aa"].shape)
