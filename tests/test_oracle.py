"""CPU tests of the oracle itself: pinned against outputs of the reference's own ``get_moransI``
(chrysalis/functions.py:15-35, EXECUTED - tests/golden/make_reference_vectors.py), hand-computed
lattices and the committed golden vectors."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import golden_csr
from oracle import moran as om
from oracle import svg as osvg
from oracle import weights as ow


def lattice_nbr(rows, cols, offsets):
    """ELL neighbour list of a torus lattice for the given (dr, dc) stencil."""
    r, c = np.divmod(np.arange(rows * cols), cols)
    return np.stack([((r + dr) % rows) * cols + (c + dc) % cols for dr, dc in offsets], 1)


ROOK = [(-1, 0), (1, 0), (0, -1), (0, 1)]
QUEEN = ROOK + [(-1, -1), (-1, 1), (1, -1), (1, 1)]


def test_known_answers_on_torus():
    rows, cols = 12, 10
    r, c = np.divmod(np.arange(rows * cols), cols)
    w4 = ow.row_standardised_csr(lattice_nbr(rows, cols, ROOK))
    w8 = ow.row_standardised_csr(lattice_nbr(rows, cols, QUEEN))
    checker = ((r + c) % 2).astype(np.float32)
    stripes = (r % 2).astype(np.float32)
    assert om.morans_I(checker, w4) == pytest.approx(-1.0, abs=1e-6)   # all 4 neighbours opposite
    assert om.morans_I(checker, w8) == pytest.approx(0.0, abs=1e-6)    # 4 opposite + 4 equal
    assert om.morans_I(stripes, w8) == pytest.approx(-0.5, abs=1e-6)   # 2 equal + 6 opposite
    assert om.morans_I(stripes, w4) == pytest.approx(0.0, abs=1e-6)
    assert om.gearys_C(checker, w4) == pytest.approx(2.0 * (rows * cols - 1) / (rows * cols), rel=1e-9)


def test_matches_in_repo_dense_formula(golden_hex):
    """chrysalis/functions.py:15-35 get_moransI (dense W, rounded to 8 dp)."""
    z = golden_hex
    X = osvg.normalize_total_log1p(golden_csr(z)[:, z["gene_mask"]])
    w = ow.row_standardised_csr(z["nbr"])
    wd = np.asarray(w.todense())
    for j in range(0, X.shape[1], 37):
        y = np.asarray(X[:, j].todense()).ravel()
        assert om.morans_I_dense_formula(wd, y) == pytest.approx(om.morans_I(y, w), abs=2e-7)
        assert om.morans_I_dense_formula(wd, y) == pytest.approx(om.morans_I_f64(y, w), abs=1e-8)


REFERENCE_FIXTURES = ["svg_hex_k6", "svg_random_k8", "svg_square_k8"]


def _reference_case(name):
    """Inputs of one reference-run fixture: (w CSR, dense float32 scored-gene matrix, scored ids, reference vectors)."""
    from conftest import load_golden
    ref = load_golden("reference_get_moransI")
    z = load_golden(name)
    mask = z["gene_mask"]
    dense = np.asarray(osvg.normalize_total_log1p(golden_csr(z)[:, mask]).todense(), dtype=np.float32)
    w = ow.row_standardised_csr(ow.knn_neighbors(ow.reference_points(z["xy"]), int(z["k"])))
    cols = np.searchsorted(np.flatnonzero(mask), ref[f"{name}_genes"])
    return w, dense, cols, ref[f"{name}_I_f32in"], ref[f"{name}_I_f64in"]


@pytest.mark.parametrize("name", REFERENCE_FIXTURES)
def test_oracle_pinned_to_reference_run_get_moransI(name):
    """``oracle.moran.morans_I`` against numbers the REFERENCE produced: ``get_moransI`` of
    /root/reference/chrysalis/functions.py:15-35 was cut out with ``ast`` and executed on these inputs
    (tests/golden/make_reference_vectors.py).  The reference rounds to 8 decimals (functions.py:35)."""
    w, dense, cols, ref32, ref64 = _reference_case(name)
    got64 = np.array([om.morans_I_f64(dense[:, j], w) for j in cols])
    assert np.abs(got64 - ref64).max() <= 5.1e-9                      # float64 in, float64 out: the rounding only
    got32 = np.array([om.morans_I(dense[:, j], w) for j in cols])      # the dtype path of core.py:59,72
    assert np.abs(got32 - ref64).max() <= 2e-7                        # float32 centring (core.py:59 hands esda float32): ~1e-7
    # fed float32, get_moransI itself forms D_i D_j in float32 and adds D^2 with Python's sequential float32 sum():
    # its own float32 noise reaches ~7e-6 (n = 1600), well above esda's (float64 lag); the oracle sits inside that band
    assert np.abs(got32 - ref32).max() <= 2e-5


@pytest.mark.skipif(not os.path.exists("/root/reference/chrysalis/functions.py"), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("name", REFERENCE_FIXTURES)
def test_committed_reference_vectors_reproduce_live(name):
    """Where the reference tree is mounted, run its ``get_moransI`` again: the committed vectors are its output."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_reference_vectors", os.path.join(os.path.dirname(__file__), "golden", "make_reference_vectors.py"))
    mrv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mrv)
    fn, lines = mrv.load_reference_get_moransI()
    w, dense, cols, ref32, ref64 = _reference_case(name)
    wd = np.asarray(w.todense())
    for t in range(0, len(cols), 9):
        assert fn(wd, dense[:, cols[t]].astype(np.float64)) == ref64[t]
        assert float(fn(wd, dense[:, cols[t]])) == pytest.approx(ref32[t], abs=1e-12)


def test_knn_is_libpysal_shaped():
    rng = np.random.default_rng(0)
    pts = rng.normal(size=(500, 2))
    nbr = ow.knn_neighbors(pts, 6)
    assert nbr.shape == (500, 6) and not (nbr == np.arange(500)[:, None]).any()
    d = np.linalg.norm(pts[:, None] - pts[None], axis=2)
    np.fill_diagonal(d, np.inf)
    assert np.array_equal(np.sort(nbr, 1), np.sort(np.argsort(d, 1)[:, :6], 1))
    w = ow.row_standardised_csr(nbr)
    assert np.allclose(np.asarray(w.sum(1)).ravel(), 1.0) and w.sum() == pytest.approx(500.0)
    # duplicate coordinates: self may be absent from the k+1 hits -> last column dropped
    pts[:8] = 0.0
    assert ow.knn_neighbors(pts, 3).shape == (500, 3)


def test_preprocessing_semantics():
    X = sp.csr_matrix(np.array([[0, 2, 0, 1], [3, 0, 0, 1], [0, 0, 0, 0], [1, 1, 0, 6]], dtype=np.int32))
    assert osvg.filter_genes_mask(X, 2).tolist() == [True, True, False, True]
    Y = osvg.normalize_total_log1p(X)
    tot = np.array([3, 4, 0, 8.0])
    med = np.median(tot[tot > 0])
    exp = np.log1p(X.toarray() * (med / np.where(tot == 0, 1, tot))[:, None])
    assert Y.dtype == np.float32 and np.allclose(Y.toarray(), exp, rtol=1e-6)
    assert Y.nnz == X.nnz


def test_selection_rule():
    I = np.array([0.9, 0.5, np.nan, 0.3, 0.25, 0.1, -0.2])
    scored = ~np.isnan(I)
    # fewer above threshold than top_svg -> threshold rule (strict >)
    assert osvg.select_svgs(I, scored, 5, 0.25).tolist() == [True, True, False, True, False, False, False]
    # more above threshold than top_svg -> top-k rule
    assert osvg.select_svgs(I, scored, 2, 0.05).tolist() == [True, True, False, False, False, False, False]
    # equal counts -> threshold rule
    assert osvg.select_svgs(I, scored, 3, 0.25).tolist() == [True, True, False, True, False, False, False]


@pytest.mark.parametrize("name", ["golden_hex", "golden_random", "golden_square"])
def test_oracle_reproduces_golden(name, request):
    z = request.getfixturevalue(name)
    r = osvg.detect_svgs(golden_csr(z), z["xy"], neighbors=int(z["k"]), top_svg=int(z["top_svg"]),
                         min_morans=float(z["min_morans"]), geary=True)
    assert np.array_equal(r["gene_mask"], z["gene_mask"])
    assert np.array_equal(np.sort(r["nbr"], 1), np.sort(z["nbr"], 1))
    m = z["gene_mask"]
    assert np.allclose(r["morans"][m], z["morans"][m], rtol=1e-6, atol=1e-9)
    assert np.allclose(r["geary"][m], z["geary"][m], rtol=1e-6)
    assert np.array_equal(r["spatially_variable"], z["spatially_variable"])


def test_batched_equals_loop(golden_hex):
    z = golden_hex
    X = np.asarray(osvg.normalize_total_log1p(golden_csr(z)[:, z["gene_mask"]]).todense(), dtype=np.float32)
    I_loop, C_loop = om.morans_I_loop(X, z["nbr"], geary=True)
    assert np.allclose(om.morans_I_batched(X, z["nbr"]), I_loop, rtol=2e-5, atol=2e-8)
    assert np.allclose(om.gearys_C_batched(X, z["nbr"]), C_loop, rtol=1e-9)


def test_analytic_moments_golden(golden_hex):
    from conftest import load_golden
    zm = load_golden("moran_moments_hex")
    z = golden_hex
    w = ow.row_standardised_csr(z["nbr"])
    X = np.asarray(osvg.normalize_total_log1p(golden_csr(z)[:, z["gene_mask"]]).todense(), dtype=np.float32)
    for j in range(8):
        a = om.moran_analytic(X[:, j], w)
        assert a["EI"] == pytest.approx(zm["EI"][j]) and a["VI_norm"] == pytest.approx(zm["VI_norm"][j])
        assert a["VI_rand"] == pytest.approx(zm["VI_rand"][j])
    p = om.moran_permutation(X[:, 0], w, 199, np.random.default_rng(1))
    assert 0 < p["p_sim"] <= 0.5 and p["sim"].shape == (199,)


def test_pca_oracle_matches_golden(golden_pca_aa):
    from oracle import pca as opca
    z = golden_pca_aa
    p = opca.pca_arpack(z["pcs"], 12)
    assert opca.subspace_cosines(z["loadings"], p["loadings"]).min() > 0.999999
    assert np.allclose(p["variance_ratio"], z["variance_ratio"], rtol=1e-4)
    q = opca.pca_covariance_eigh(z["pcs"], 12)
    assert opca.component_cosines(p["loadings"], q["loadings"]).min() > 0.9999


def test_aa_oracle_matches_golden(golden_pca_aa):
    from oracle import aa as oaa
    z = golden_pca_aa
    X = z["X_pca"][:, :8]
    inits = [(None, b) for b in z["betas0"]]
    fit = oaa.aa_fit(X, 5, n_init=3, max_iter=30, inits=[(np.zeros((X.shape[0], 5)), b) for _, b in inits])
    assert fit["rss"] == pytest.approx(float(z["rss"]), rel=1e-6)
    assert np.abs(fit["alphas"] - z["alphas"]).max() < 1e-6
    assert np.allclose(fit["alphas"].sum(1), 1.0) and (fit["alphas"] >= 0).all()


def test_local_moran_sums_to_global_and_conditional_null():
    """Moran_Local restatement: sum_i Is_i = (n - 1) I, quadrants partition the spots, p_sim in (0, 0.5]."""
    rs = np.random.RandomState(5)
    n = 200
    xy = rs.rand(n, 2)
    w = ow.row_standardised_csr(ow.knn_neighbors(xy, 6))
    y = np.sin(6 * xy[:, 0]) + 0.5 * rs.randn(n)
    r = om.local_moran(y, w, 39)
    assert abs(r["Is"].sum() / (n - 1) - om.morans_I_f64(y, w)) < 1e-12
    assert set(np.unique(r["q"])) <= {1, 2, 3, 4}
    assert r["p_sim"].min() >= 1 / 40 and r["p_sim"].max() <= 0.5 + 1 / 40
    # conditional null mean of spot i: -(n - 1) z_i^2 ... / den averaged over the other spots = -z_i * (sum z - z_i)/(n-1) * (n-1)/den
    z = r["z"]
    expect = (n - 1) * z * ((z.sum() - z) / (n - 1)) / (z * z).sum()
    assert np.abs(r["EI_sim"] - expect).max() < 6 * np.sqrt(r["VI_sim"].max() / 39)
