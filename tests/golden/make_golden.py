"""Regenerates tests/golden/*.npz from the CPU oracle.

The reference itself cannot be imported in the build image (pysal/esda/scanpy/archetypes/
anndata are absent, no network - SURVEY.md section 8c) and its test-suite holds no assertions
or golden vectors, so these fixtures pin the ORACLE (a drift guard) and give the GPU tests
fixed inputs; they are not outputs of the reference.  sklearn PCA outputs ARE the reference's
own call (core.py:127).

    python tests/golden/make_golden.py            # everything
    python tests/golden/make_golden.py square     # only svg_square_k8.npz
"""
import os
import sys

import numpy as np
import scipy
import sklearn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from chrysalis_b200 import synth  # noqa: E402
from oracle import aa as oaa  # noqa: E402
from oracle import moran as omoran  # noqa: E402
from oracle import pca as opca  # noqa: E402
from oracle import svg as osvg  # noqa: E402
from oracle import weights as ow  # noqa: E402


def svg_case(name, kind, n, g, k, seed, **kw):
    xy = synth.make_coords(kind, n, seed=seed)
    X, _ = synth.make_counts_csr(xy, g, n_zones=6, seed=seed)
    r = osvg.detect_svgs(X, xy, neighbors=k, geary=True, **kw)
    np.savez_compressed(
        os.path.join(HERE, f"{name}.npz"), xy=xy, indptr=X.indptr.astype(np.int64), indices=X.indices.astype(np.int32),
        data=X.data.astype(np.int32), shape=np.array(X.shape), k=k, morans=r["morans"], geary=r["geary"],
        gene_mask=r["gene_mask"], spatially_variable=r["spatially_variable"], nbr=r["nbr"].astype(np.int32),
        top_svg=kw.get("top_svg", 1000), min_morans=kw.get("min_morans", 0.2),
        versions=np.array([np.__version__, scipy.__version__, sklearn.__version__]))
    print(name, X.shape, "scored", int(r["gene_mask"].sum()), "svgs", int(r["spatially_variable"].sum()))
    return X, xy, r


def main():
    if sys.argv[1:] == ["square"]:     # add the lattice case without rewriting the other fixtures
        svg_case("svg_square_k8", "square", 1600, 300, 8, 11, top_svg=40, min_morans=0.05)
        return
    # cfg1-shaped (hex, k=6), cfg2-shaped (random beads, k=8) and cfg3/cfg4-shaped (square lattice of bins, k=8), down-scaled
    X, xy, r = svg_case("svg_hex_k6", "hex", 900, 400, 6, 42, top_svg=60)
    svg_case("svg_random_k8", "random", 1200, 300, 8, 7, top_svg=40, min_morans=0.05)
    svg_case("svg_square_k8", "square", 1600, 300, 8, 11, top_svg=40, min_morans=0.05)

    # PCA + AA on the hex case's SVG matrix (reference flow: normalise full adata, then pca/aa)
    Xn = osvg.normalize_total_log1p(X)
    sel = osvg.select_svgs(r["morans"], r["gene_mask"], 64, 0.0)   # 64 top-ranked genes as the SVG set
    pcs = np.asarray(Xn[:, sel].todense(), dtype=np.float32)
    p = opca.pca_arpack(pcs, 12)
    fit = oaa.aa_fit(p["X_pca"][:, :8], 5, n_init=3, max_iter=30)
    np.savez_compressed(
        os.path.join(HERE, "pca_aa_hex.npz"), pcs=pcs, sel=sel, X_pca=p["X_pca"], variance_ratio=p["variance_ratio"],
        loadings=p["loadings"], alphas=fit["alphas"], archetypes=fit["archetypes"], rss=fit["rss"],
        betas0=np.stack([b for _, b in fit["inits"]]), n_iter=fit["n_iter"])
    print("pca_aa_hex", pcs.shape, "rss", fit["rss"], "iters", fit["n_iter"])

    # permutation inference known answers (analytic moments are exact, p_sim statistical)
    w = ow.row_standardised_csr(r["nbr"])
    dense = np.asarray(osvg.normalize_total_log1p(X[:, r["gene_mask"]]).todense(), dtype=np.float32)
    an = [omoran.moran_analytic(dense[:, j], w) for j in range(8)]
    np.savez_compressed(os.path.join(HERE, "moran_moments_hex.npz"),
                        EI=np.array([a["EI"] for a in an]), VI_norm=np.array([a["VI_norm"] for a in an]),
                        VI_rand=np.array([a["VI_rand"] for a in an]))


if __name__ == "__main__":
    main()
