"""GPU parity of the SVG stage (K1-K4) against the CPU oracle - all calls go through the C ABI.

Tolerance: Moran's I within 1e-5 relative (north_star) plus an absolute floor of 2e-7: the
reference's input is float32 (core.py:59 ad.to_df()), and float32 rounding of z alone moves I
by ~1e-7 - the oracle's float32 and float64 statements differ by that much on the golden data
(tests/test_oracle.py::test_matches_in_repo_dense_formula), so nothing tighter is meaningful
for |I| < 0.02.  Exact known answers on lattices are held to 1e-12.
"""
import numpy as np
import scipy.sparse as sp
import pytest
import torch

import chrysalis_b200 as ch
from chrysalis_b200 import device as dv
from chrysalis_b200 import synth
from conftest import golden_csr
from oracle import moran as om
from oracle import svg as osvg
from oracle import weights as ow

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 2e-7


def assert_close(got, ref, rtol=RTOL, atol=ATOL, what="I"):
    got, ref = np.asarray(got), np.asarray(ref)
    err = np.abs(got - ref)
    bad = err > rtol * np.abs(ref) + atol
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} off; worst abs {err.max():.3e} at ref {ref[err.argmax()]:.6g}"


def dense_from(X_csr, mask=None):
    Xn = osvg.normalize_total_log1p(X_csr if mask is None else X_csr[:, mask])
    return np.asarray(Xn.todense(), dtype=np.float32)


def gpu_moran(dense, nbr, geary=False, ld=None, **kw):
    n, g = dense.shape
    ld = dv.dense_ld(g) if ld is None else ld
    Xd = torch.zeros((n, ld), dtype=torch.float32, device="cuda")
    Xd[:, :g] = torch.from_numpy(dense).cuda()
    nb = torch.from_numpy(np.ascontiguousarray(nbr, dtype=np.int32)).cuda()
    I, C, st = dv.moran_dense(Xd, nb, g, geary=geary, **kw)
    torch.cuda.synchronize()
    return I.cpu().numpy(), (C.cpu().numpy() if geary else None), st


@pytest.mark.parametrize("name", ["golden_hex", "golden_random", "golden_square"])
def test_moran_kernel_matches_golden(name, request):
    z = request.getfixturevalue(name)
    m = z["gene_mask"]
    I, C, _ = gpu_moran(dense_from(golden_csr(z), m), z["nbr"], geary=True)
    assert_close(I, z["morans"][m])
    assert_close(C, z["geary"][m], what="C")


@pytest.mark.parametrize("name", ["svg_hex_k6", "svg_random_k8", "svg_square_k8"])
def test_moran_kernel_matches_reference_run_vectors(name):
    """CUDA Moran's I against numbers the reference itself produced (``get_moransI``,
    /root/reference/chrysalis/functions.py:15-35, executed by tests/golden/make_reference_vectors.py) - no oracle
    arithmetic between the two: inputs are built the same way, the statistic is compared directly."""
    from conftest import load_golden
    ref = load_golden("reference_get_moransI")
    z = load_golden(name)
    m = z["gene_mask"]
    nbr = ow.knn_neighbors(ow.reference_points(z["xy"]), int(z["k"]))
    I, _, _ = gpu_moran(dense_from(golden_csr(z), m), nbr)
    cols = np.searchsorted(np.flatnonzero(m), ref[f"{name}_genes"])
    assert_close(I[cols], ref[f"{name}_I_f64in"], atol=2e-7 + 5e-9)


@pytest.mark.parametrize("variant", [0, 4])
def test_all_kernel_variants_agree(variant, golden_random):
    """The tile kernel over the quad plan (0, default) and the per-spot stream kernel (4): same statistic."""
    z = golden_random
    m = z["gene_mask"]
    I, C, _ = gpu_moran(dense_from(golden_csr(z), m), z["nbr"], geary=True, variant=variant)
    assert_close(I, z["morans"][m])
    assert_close(C, z["geary"][m], what="C")


def test_preshifted_input_and_odd_spot_count():
    xy = synth.make_coords("hex", 2501, seed=9)                       # odd n: last pair has one spot
    X, _ = synth.make_counts_csr(xy, 150, seed=9)
    dense = dense_from(X)
    dense = dense[:, dense.std(0) > 0]
    nbr = ow.knn_neighbors(xy, 6)
    ref = om.morans_I_batched(dense, nbr)
    I, _, _ = gpu_moran(dense, nbr)
    assert_close(I, ref)
    centred = dense - dense.mean(0, keepdims=True).astype(np.float32)
    I2, _, _ = gpu_moran(centred, nbr, preshifted=True)
    assert_close(I2, ref)


def test_duplicate_and_self_neighbours_take_the_generic_path():
    """Neighbour lists with repeated entries / self loops are legal input for the C ABI."""
    rng = np.random.default_rng(4)
    n, g, k = 1001, 40, 8
    dense = rng.normal(1.0, 1.0, (n, g)).astype(np.float32)
    nbr = rng.integers(0, n, (n, k))
    nbr[::3, 1] = nbr[::3, 0]                                          # duplicates
    nbr[::5, 2] = np.arange(n)[::5]                                    # self loops
    I, _, _ = gpu_moran(dense, nbr)
    assert_close(I, om.morans_I_batched(dense, nbr))


@pytest.mark.parametrize("k", [1, 3, 4, 6, 8, 18, 36, 40, 64, 70])
def test_moran_kernel_any_k(k):
    xy = synth.make_coords("random", 3000, seed=k)
    X, _ = synth.make_counts_csr(xy, 200, seed=k)
    dense = dense_from(X)
    dense = dense[:, dense.std(0) > 0]
    nbr = ow.knn_neighbors(xy, k)
    I, C, _ = gpu_moran(dense, nbr, geary=True)
    assert_close(I, om.morans_I_batched(dense, nbr))
    assert_close(C, om.gearys_C_batched(dense, nbr), what="C")


@pytest.mark.parametrize("n,g", [(33, 1), (513, 3), (1000, 127), (1025, 129), (4035, 260)])
def test_ragged_shapes_and_strides(n, g):
    rng = np.random.default_rng(n + g)
    xy = rng.uniform(0, 100, (n, 2))
    dense = np.log1p(rng.poisson(0.7, (n, g))).astype(np.float32) + rng.normal(0, 0.1, (n, g)).astype(np.float32)
    nbr = ow.knn_neighbors(xy, 6)
    ref = om.morans_I_batched(dense, nbr)
    for ld in (dv.dense_ld(g), (g + 3) // 4 * 4, dv.dense_ld(g) + 64):
        I, _, _ = gpu_moran(dense, nbr, ld=ld)
        assert_close(I, ref)


def test_constant_and_highly_expressed_genes():
    """Cancellation guard: large mean / tiny variance genes, and a constant gene -> NaN."""
    rng = np.random.default_rng(3)
    n = 4035
    xy = synth.make_coords("hex", n, seed=1)
    nbr = ow.knn_neighbors(xy, 6)
    base = rng.normal(0, 1, n)
    dense = np.stack([7.0 + 0.01 * base, 100.0 + rng.normal(0, 0.5, n), np.full(n, 2.5), 5.0 + np.sin(xy[:, 0] / 300.0)],
                     1).astype(np.float32)
    I, _, _ = gpu_moran(dense, nbr)
    ref64 = om.morans_I_batched(dense, nbr, dtype=np.float64)
    assert np.isnan(I[2])
    keep = [0, 1, 3]
    # judged against the float64 statistic of the same float32 inputs
    assert_close(I[keep], ref64[keep], rtol=1e-5, atol=2e-7)
    # and the float32-faithful oracle is no closer to the float64 truth than the kernel is
    ref32 = om.morans_I_loop(dense[:, keep], nbr)
    assert np.abs(I[keep] - ref64[keep]).max() <= max(3 * np.abs(ref32 - ref64[keep]).max(), 2e-7)


def test_known_answers_on_torus_lattice():
    rows, cols = 64, 48
    r, c = np.divmod(np.arange(rows * cols), cols)
    def lat(offs):
        return np.stack([((r + dr) % rows) * cols + (c + dc) % cols for dr, dc in offs], 1)
    rook = [(-1, 0), (1, 0), (0, -1), (0, 1)]
    queen = rook + [(-1, -1), (-1, 1), (1, -1), (1, 1)]
    dense = np.stack([(r + c) % 2, r % 2, c % 2], 1).astype(np.float32)
    I4, C4, _ = gpu_moran(dense, lat(rook), geary=True)
    I8, _, _ = gpu_moran(dense, lat(queen))
    assert np.allclose(I4, [-1.0, 0.0, 0.0], atol=1e-12)
    assert np.allclose(I8, [0.0, -0.5, -0.5], atol=1e-12)
    assert C4[0] == pytest.approx(2.0 * (rows * cols - 1) / (rows * cols), rel=1e-12)


def test_invariances_full_spot_count():
    """BASELINE cfg4 spot count (700k, k=8) with a thin gene slab: size-independent properties.
    I is invariant to affine maps of expression and to relabelling spots; the oracle is run on
    a column subsample."""
    n, g, k = 700_000, 96, 8
    side = int(np.ceil(np.sqrt(n)))
    rr, cc = np.divmod(np.arange(n), side)
    # exact square lattice -> queen stencil for interior bins, wrap-free clamp at the border
    offs = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]
    nbr = np.empty((n, k), dtype=np.int32)
    for j, (dr, dc) in enumerate(offs):
        r2 = np.clip(rr + dr, 0, (n - 1) // side)
        c2 = np.clip(cc + dc, 0, side - 1)
        idx = r2 * side + c2
        idx = np.where(idx >= n, np.arange(n) - 1, idx)
        idx = np.where(idx == np.arange(n), (np.arange(n) + 1) % n, idx)
        nbr[:, j] = idx
    gen = torch.Generator(device="cuda").manual_seed(5)
    xs = torch.linspace(0, 40, side, device="cuda")
    field = (torch.sin(xs)[None, :] * torch.cos(xs * 0.7)[:, None]).reshape(-1)[:n]
    amp = torch.linspace(0, 2, g, device="cuda")
    X = torch.poisson(torch.exp(0.3 + amp[None, :] * field[:, None]).clamp_max(50.0), generator=gen)
    X = torch.log1p(X)
    Xd = torch.zeros((n, dv.dense_ld(g)), device="cuda")
    Xd[:, :g] = X
    nb = torch.from_numpy(nbr).cuda()
    I, _, _ = dv.moran_dense(Xd, nb, g)
    I2, _, _ = dv.moran_dense(Xd * 3.0 + 1.5, nb, g)             # affine invariance
    assert_close(I2.cpu().numpy(), I.cpu().numpy(), rtol=2e-6, atol=2e-8)
    perm = torch.randperm(n, device="cuda", generator=gen)        # relabel spots
    inv = torch.empty_like(perm); inv[perm] = torch.arange(n, device="cuda")
    nb_p = inv[nb[perm].long()].to(torch.int32).contiguous()
    I3, _, _ = dv.moran_dense(Xd[perm].contiguous(), nb_p, g)
    assert_close(I3.cpu().numpy(), I.cpu().numpy(), rtol=2e-6, atol=2e-8)
    I4, _, _ = dv.moran_dense(Xd, nb, g)                          # run-to-run bit reproducibility
    assert torch.equal(I4, I)
    cols = [0, 1, 17, 48, 95]
    ref = om.morans_I_batched(X[:, cols].cpu().numpy(), nbr)
    assert_close(I.cpu().numpy()[cols], ref)
    assert I[-1] > 0.3 and abs(float(I[0])) < 0.01                # spatial signal grows with amp


def test_gene_sharding_is_bit_identical():
    """Multi-GPU contract: any contiguous gene shard reproduces the 1-GPU bits."""
    xy = synth.make_coords("square", 20000, seed=3)
    X, _ = synth.make_counts_csr(xy, 600, seed=3)
    dense = dense_from(X)
    nbr = ow.knn_neighbors(xy, 8)
    I, _, _ = gpu_moran(dense, nbr)
    for lo, hi in [(0, 150), (150, 413), (413, 600)]:
        Is, _, _ = gpu_moran(np.ascontiguousarray(dense[:, lo:hi]), nbr)
        assert np.array_equal(Is, I[lo:hi])


@pytest.mark.parametrize("kind,n,k", [("hex", 4035, 6), ("random", 20000, 8), ("square", 30000, 8), ("random", 5000, 36)])
def test_knn_builder_matches_ckdtree(kind, n, k):
    xy = synth.make_coords(kind, n, seed=11)
    xy[:, 1] *= -1
    nbr = dv.knn_build(dv.to_device(xy, torch.float64), k).cpu().numpy()
    ref = ow.knn_neighbors(xy, k)
    assert np.array_equal(np.sort(nbr, 1), np.sort(ref, 1))       # jittered coords: unique kNN sets
    d = np.linalg.norm(xy[nbr] - xy[:, None, :], axis=2)
    assert (np.diff(d, axis=1) >= 0).all()                         # increasing distance order


def test_knn_builder_exact_lattice_ties():
    """Exact lattice: ties everywhere.  Neighbour *distances* must equal cKDTree's."""
    xy = synth.make_coords("square", 10000, seed=0, jitter=0.0)
    nbr = dv.knn_build(dv.to_device(xy, torch.float64), 8).cpu().numpy()
    ref = ow.knn_neighbors(xy, 8)
    dg = np.sort(np.linalg.norm(xy[nbr] - xy[:, None, :], axis=2), 1)
    dr = np.sort(np.linalg.norm(xy[ref] - xy[:, None, :], axis=2), 1)
    assert np.allclose(dg, dr)
    assert not (nbr == np.arange(len(xy))[:, None]).any()


def test_hilbert_order_is_a_permutation_and_local():
    xy = synth.make_coords("random", 50000, seed=2)
    order = dv.hilbert_order(dv.to_device(xy, torch.float64)).cpu().numpy()
    assert np.array_equal(np.sort(order), np.arange(len(xy)))
    step = np.linalg.norm(np.diff(xy[order], axis=0), axis=1)
    assert np.median(step) < 3 * np.sqrt(np.ptp(xy[:, 0]) * np.ptp(xy[:, 1]) / len(xy))


def test_spatial_order_aligns_quads_with_a_square_lattice():
    """On a (jittered) square lattice the row order is a Hilbert curve whose grid is snapped to the lattice: groups of 4
    consecutive spots are 2 x 2 blocks.  Hex lattices and random beads keep the plain curve."""
    xy = synth.make_coords("square", 40000, seed=1)
    xyd = dv.to_device(xy, torch.float64)
    bbox = dv.spatial_bbox(xyd)
    assert abs(dv.square_lattice_pitch(xyd, bbox) - 100.0) < 0.1
    order = dv.spatial_order(xyd, bbox).cpu().numpy()
    assert np.array_equal(np.sort(order), np.arange(len(xy)))
    quads = xy[order].reshape(-1, 4, 2)
    span = quads.max(1) - quads.min(1)
    assert (np.abs(span - 100.0) < 5.0).all(1).mean() > 0.95                  # exact 2 x 2 blocks
    plain = xy[dv.hilbert_order(xyd, bbox).cpu().numpy()].reshape(-1, 4, 2)
    assert (np.abs(plain.max(1) - plain.min(1) - 100.0) < 5.0).all(1).mean() < 0.9
    for kind in ("hex", "random"):
        other = dv.to_device(synth.make_coords(kind, 20000, seed=3), torch.float64)
        assert dv.square_lattice_pitch(other, dv.spatial_bbox(other)) is None
        assert torch.equal(dv.spatial_order(other), dv.hilbert_order(other))


def test_detect_svgs_on_a_square_lattice_matches_the_oracle():
    """The lattice-snapped row order only relabels the spots: Moran's I, Geary's C and the selection follow the oracle."""
    xy = synth.make_coords("square", 3600, seed=5)
    X, _ = synth.make_counts_csr(xy, 400, n_zones=5, seed=5)
    ad = ch.AnnDataLite(X, obsm={"spatial": xy})
    ch.detect_svgs(ad, neighbors=8, top_svg=60, geary=True)
    ref = osvg.detect_svgs(X, xy, neighbors=8, top_svg=60)
    m = ref["gene_mask"]
    got = ad.var["Moran's I"].to_numpy()
    assert np.array_equal(np.isnan(got), ~m)
    assert np.allclose(got[m], ref["morans"][m], rtol=1e-5, atol=2e-7)
    assert np.array_equal(ad.var["spatially_variable"].to_numpy(), ref["spatially_variable"])


def test_preprocessing_kernels_match_scanpy_semantics(golden_hex):
    z = golden_hex
    X = golden_csr(z)
    csr = dv.upload_csr(X)
    nnz = dv.csr_gene_nnz(csr).cpu().numpy()
    assert np.array_equal(nnz, np.asarray((X > 0).sum(0)).ravel())
    keep = torch.from_numpy(z["gene_mask"]).cuda()
    gene_col = torch.where(keep, torch.cumsum(keep.int(), 0, dtype=torch.int32) - 1, torch.tensor(-1, dtype=torch.int32, device="cuda"))
    tot = dv.csr_spot_totals(csr, gene_col)
    assert np.array_equal(tot.cpu().numpy(), np.asarray(X[:, z["gene_mask"]].sum(1)).ravel().astype(np.float64))
    dense = dv.csr_densify(csr, gene_col, int(z["gene_mask"].sum()), dv.median_scale(tot), None, True)
    ref = dense_from(X, z["gene_mask"])
    got = dense[:, :ref.shape[1]].cpu().numpy()
    assert np.array_equal(got == 0, ref == 0)
    assert np.allclose(got, ref, rtol=3e-7, atol=0)
    assert float(dense[:, ref.shape[1]:].abs().sum()) == 0.0


@pytest.mark.parametrize("n,g", [(301, 500), (64, 51300), (1000, 96)])
def test_densify_through_a_row_permutation_matches_the_host(n, g):
    """csr_densify (scale in float64, log1p, scatter through a row permutation) against scipy's todense, and with a
    caller-provided pre-zeroed buffer."""
    rng = np.random.default_rng(n + g)
    X = sp.random(n, g, density=0.03, format="csr", random_state=rng, data_rvs=lambda k: rng.integers(1, 40, k).astype(np.float32))
    X = sp.csr_matrix(X, dtype=np.float32)
    csr = dv.upload_csr(X)
    keep = torch.from_numpy(rng.random(g) < 0.7).cuda()
    gene_col = torch.where(keep, torch.cumsum(keep.int(), 0, dtype=torch.int32) - 1, torch.tensor(-1, dtype=torch.int32, device="cuda"))
    n_keep = int(keep.sum())
    perm = torch.from_numpy(rng.permutation(n)).cuda()
    scale = torch.from_numpy(rng.uniform(0.5, 2.0, n)).cuda()
    dense = dv.csr_densify(csr, gene_col, n_keep, scale, perm, True)
    ref = np.log1p((X[:, keep.cpu().numpy()].toarray().astype(np.float64) * scale.cpu().numpy()[:, None]).astype(np.float32))
    got = dense.cpu().numpy()
    assert np.array_equal(got[perm.cpu().numpy(), :n_keep] == 0, ref == 0)
    assert np.allclose(got[perm.cpu().numpy(), :n_keep], ref, rtol=3e-7, atol=0)
    assert float(np.abs(got[:, n_keep:]).sum()) == 0.0
    again = dv.csr_densify(csr, gene_col, n_keep, scale, perm, True, out=torch.zeros_like(dense))
    assert torch.equal(again, dense)


@pytest.mark.parametrize("name", ["golden_hex", "golden_random", "golden_square"])
def test_detect_svgs_end_to_end(name, request):
    z = request.getfixturevalue(name)
    ad = ch.AnnDataLite(golden_csr(z), obsm={"spatial": z["xy"]})
    ch.detect_svgs(ad, neighbors=int(z["k"]), top_svg=int(z["top_svg"]), min_morans=float(z["min_morans"]), geary=True)
    m = z["gene_mask"]
    got = ad.var["Moran's I"].to_numpy()
    assert np.array_equal(np.isnan(got), ~m)
    assert_close(got[m], z["morans"][m])
    assert_close(ad.var["Geary's C"].to_numpy()[m], z["geary"][m], what="C")
    assert np.array_equal(ad.var["spatially_variable"].to_numpy(), z["spatially_variable"])


def test_detect_svgs_with_empty_spots_and_with_no_gene_passing_the_filter():
    """Spots without any count (scanpy leaves their rows untouched: total 0 is not scaled) and the degenerate call in
    which the min_spots filter removes every gene (all NaN, nothing selected) follow the oracle."""
    xy = synth.make_coords("hex", 900, seed=11)
    X, _ = synth.make_counts_csr(xy, 300, n_zones=4, seed=11)
    X = X.tolil()
    for i in (0, 17, 450, 899):
        X[i, :] = 0
    X = sp.csr_matrix(X.tocsr(), dtype=np.float32)
    X.eliminate_zeros()
    ad = ch.AnnDataLite(X, obsm={"spatial": xy})
    ch.detect_svgs(ad, neighbors=6, top_svg=40)
    ref = osvg.detect_svgs(X, xy, neighbors=6, top_svg=40)
    m = ref["gene_mask"]
    got = ad.var["Moran's I"].to_numpy()
    assert m.any() and np.array_equal(np.isnan(got), ~m)
    assert np.allclose(got[m], ref["morans"][m], rtol=1e-5, atol=2e-7)
    assert np.array_equal(ad.var["spatially_variable"].to_numpy(), ref["spatially_variable"])
    ad2 = ch.AnnDataLite(X, obsm={"spatial": xy})
    ch.detect_svgs(ad2, min_spots=0.999, neighbors=6)
    assert np.isnan(ad2.var["Moran's I"].to_numpy()).all()
    assert not ad2.var["spatially_variable"].to_numpy().any()


def test_detect_svgs_respects_existing_log1p(golden_hex):
    z = golden_hex
    Xn = osvg.normalize_total_log1p(golden_csr(z))
    ad = ch.AnnDataLite(Xn, obsm={"spatial": z["xy"]}, uns={"log1p": {"base": None}})
    ch.detect_svgs(ad, neighbors=6)
    ref = osvg.detect_svgs(Xn, z["xy"], neighbors=6, already_log1p=True)
    m = ref["gene_mask"]
    assert_close(ad.var["Moran's I"].to_numpy()[m], ref["morans"][m])
    assert np.array_equal(ad.var["spatially_variable"].to_numpy(), ref["spatially_variable"])


def test_cfg1_shape_selection_identical():
    """BASELINE cfg1 spot count (4,035 hex spots, k=6) x 3,000 genes: identical SVG set."""
    xy = synth.make_coords("hex", 4035, seed=42)
    X, _ = synth.make_counts_csr(xy, 3000, seed=42)
    ad = ch.AnnDataLite(X, obsm={"spatial": xy})
    ch.detect_svgs(ad, neighbors=6, top_svg=200)
    ref = osvg.detect_svgs(X, xy, neighbors=6, top_svg=200)
    m = ref["gene_mask"]
    assert_close(ad.var["Moran's I"].to_numpy()[m], ref["morans"][m])
    assert np.array_equal(ad.var["spatially_variable"].to_numpy(), ref["spatially_variable"])


def test_morans_I_does_not_depend_on_the_geary_flag(golden_random):
    z = golden_random
    m = z["gene_mask"]
    d = dense_from(golden_csr(z), m)
    I0, _, _ = gpu_moran(d, z["nbr"])
    I1, _, _ = gpu_moran(d, z["nbr"], geary=True)
    assert np.array_equal(I0, I1)


def test_permutation_inference_matches_esda_semantics():
    """esda.moran.Moran(y, w, permutations=P): I exact, simulated moments within Monte-Carlo error
    (draws cannot be bit-matched: esda uses NumPy's global RNG per gene, the kernel shares a draw across genes)."""
    xy = synth.make_coords("hex", 1500, seed=5)
    X, _ = synth.make_counts_csr(xy, 120, n_zones=5, seed=5)
    dense = dense_from(X)
    dense = dense[:, dense.std(0) > 0]
    nbr = ow.knn_neighbors(xy, 6)
    order = np.argsort(-om.morans_I_batched(dense, nbr))
    dense = np.ascontiguousarray(dense[:, np.r_[order[:24], order[-24:]]])     # 24 most spatial + 24 least spatial genes
    w = ow.row_standardised_csr(nbr)
    n, g = dense.shape
    Xd = torch.zeros((n, dv.dense_ld(g)), dtype=torch.float32, device="cuda")
    Xd[:, :g] = torch.from_numpy(dense).cuda()
    nb = torch.from_numpy(nbr.astype(np.int32)).cuda()
    P = 499
    res = {k: v.cpu().numpy() for k, v in dv.moran_permutation_test(Xd, nb, g, permutations=P, seed=123).items()}
    ref = [om.moran_permutation(dense[:, j], w, P, np.random.default_rng(j)) for j in range(g)]
    assert_close(res["I"], [r["I"] for r in ref])
    an = [om.moran_analytic(dense[:, j], w) for j in range(g)]
    # the permutation mean / variance estimate the analytic randomisation moments E[I] = -1/(n-1), VI_rand
    ei = np.array([a["EI"] for a in an]); vi = np.array([a["VI_rand"] for a in an])
    se_mean = np.sqrt(vi / P)
    assert (np.abs(res["EI_sim"] - ei) < 5 * se_mean).all()
    ratio = res["VI_sim"] / vi                                          # a 499-draw variance: ~6 % sd, more for sparse (kurtotic) genes
    assert abs(np.median(ratio) - 1.0) < 0.05 and (np.abs(ratio - 1.0) < 0.6).all()
    # same decisions as the reference-style simulation: strong genes get the minimum p, z-scores agree
    ref_p = np.array([r["p_sim"] for r in ref]); ref_z = np.array([r["z_sim"] for r in ref])
    strong = ref_z > 6
    assert strong.sum() > 3 and (res["p_sim"][strong] == 1.0 / (P + 1)).all()
    weak = np.abs(ref_z) < 1
    assert weak.sum() > 3 and (res["p_sim"][weak] > 0.05).all()
    assert (np.abs(res["z_sim"] - ref_z) < 0.2 * np.abs(ref_z) + 0.5).all()
    assert (res["p_sim"] >= 1.0 / (P + 1)).all() and (res["p_sim"] <= 0.5 + 1e-12).all()
    # a draw is a relabelling of the spots: permuting the rows by hand gives the same statistic (the pilot shift
    # samples different rows, so the float32 rounding differs)
    perm = torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7)).to(torch.int32)
    I_map, _, _ = dv.moran_dense(Xd, nb, g, row_map=perm)
    I_hand, _, _ = dv.moran_dense(Xd[perm.long()].contiguous(), nb, g)
    assert_close(I_map.cpu().numpy(), I_hand.cpu().numpy())


def test_morans_i_permutation_fields(golden_hex):
    z = golden_hex
    ad = ch.AnnDataLite(golden_csr(z), obsm={"spatial": z["xy"]})
    ch.morans_i_permutation(ad, neighbors=6, permutations=99, seed=1)
    m = z["gene_mask"]
    assert_close(ad.var["Moran's I"].to_numpy()[m], z["morans"][m])
    p = ad.var["Moran's I p_sim"].to_numpy()
    assert np.array_equal(np.isnan(p), ~m) and (p[m] >= 0.01 - 1e-12).all() and (p[m] <= 0.5 + 1e-12).all()
    zs = ad.var["Moran's I z_sim"].to_numpy()
    assert (zs[m][z["morans"][m] > 0.2] > 4).all()                     # clearly spatial genes are far out in the null


def test_local_moran_matches_esda_semantics():
    """esda.moran.Moran_Local: Is / lag / quadrants exact, conditional-permutation p_sim within Monte-Carlo error;
    the local statistics sum to (n - 1) times the global Moran's I."""
    n, k, P = 600, 6, 499
    rs = np.random.RandomState(5)
    xy = rs.rand(n, 2)
    nbr = ow.knn_neighbors(xy, k)
    w = ow.row_standardised_csr(nbr)
    y = (np.sin(6 * xy[:, 0]) + 0.5 * rs.randn(n)).astype(np.float32)     # spatially structured + noise
    ref = om.local_moran(y, w, P, np.random.default_rng(1))
    res = {key: v.cpu().numpy() for key, v in
           dv.local_moran(torch.from_numpy(y).cuda(), torch.from_numpy(nbr.astype(np.int32)).cuda(), permutations=P, seed=9).items()}
    np.testing.assert_allclose(res["Is"], ref["Is"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(res["lag"], ref["lag"], rtol=2e-5, atol=2e-6)
    clear = (np.abs(ref["z"]) > 1e-4) & (np.abs(ref["lag"]) > 1e-4)
    assert np.array_equal(res["q"][clear], ref["q"][clear])
    assert abs(res["Is"].astype(np.float64).sum() / (n - 1) - om.morans_I_f64(y, w)) < 1e-5
    # conditional randomisation: E[sim_i] = -Is-scale * z_i / (n - 1) ... compare with the oracle's draws statistically
    assert np.abs(res["EI_sim"] - ref["EI_sim"]).max() < 6 * np.sqrt(ref["VI_sim"].max() / P) + 1e-6
    np.testing.assert_allclose(res["VI_sim"], ref["VI_sim"], rtol=0.35, atol=1e-6)
    # p_sim: binomial noise of two independent P-draw estimates
    assert np.abs(res["p_sim"] - ref["p_sim"]).max() < 6 * np.sqrt(2 * 0.25 / P)
    assert np.corrcoef(res["p_sim"], ref["p_sim"])[0, 1] > 0.97
    assert res["p_sim"].min() >= 1.0 / (P + 1) and res["p_sim"].max() <= 0.5 + 1.0 / (P + 1)
    # same seed -> same draws; different launch has no effect
    again = dv.local_moran(torch.from_numpy(y).cuda(), torch.from_numpy(nbr.astype(np.int32)).cuda(), permutations=P, seed=9)
    assert np.array_equal(again["p_sim"].cpu().numpy(), res["p_sim"])


def test_local_morans_i_fields(golden_hex):
    X, spatial = golden_csr(golden_hex), golden_hex["xy"]
    busy = np.argsort(-np.asarray((X > 0).sum(0)).ravel())[:3]
    ad = ch.AnnDataLite(osvg.normalize_total_log1p(X)[:, busy].tocsr(), obsm={"spatial": spatial})
    names = list(ad.var_names[:3])
    ch.local_morans_i(ad, names, neighbors=6, permutations=99, seed=3)
    n = X.shape[0]
    assert ad.obsm["local_morans_i"].shape == (n, 3) and ad.obsm["local_morans_p_sim"].shape == (n, 3)
    assert set(np.unique(ad.obsm["local_morans_q"])) <= {1, 2, 3, 4}
    nbr = ow.knn_neighbors(ow.reference_points(spatial), 6)
    w = ow.row_standardised_csr(nbr)
    dense = np.asarray(ad.X[:, :3].todense(), dtype=np.float32)
    for j in range(3):
        if dense[:, j].std() > 0:
            np.testing.assert_allclose(ad.obsm["local_morans_i"][:, j], om.local_moran(dense[:, j], w)["Is"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("n,g,density", [(300, 700, 0.2), (500, 60000, 0.002), (64, 70000, 0.01), (2000, 3000, 0.05)])
def test_delta_coded_columns_decode_to_the_scipy_arrays(n, g, density):
    """chr_csr_decode_compact_delta: device CSR arrays of a delta-coded CompactCSR (escapes for gaps >= 255 and counts >= 255,
    empty rows) against the scipy matrix it was made from."""
    rng = np.random.default_rng(n + g)
    X = sp.random(n, g, density=density, format="csr", random_state=n, data_rvs=lambda k: rng.integers(1, 300, k)).astype(np.float32)
    X.sort_indices()
    c = dv.CompactCSR.from_scipy(X, delta=True)
    d = dv.upload_csr(c)
    assert np.array_equal(d.indptr.cpu().numpy(), X.indptr) and np.array_equal(d.indices.cpu().numpy(), X.indices)
    assert np.array_equal(d.data.cpu().numpy(), X.data)


def test_detect_svgs_on_compact_counts_and_uint8_scipy(golden_square):
    """The loader formats of the counts: device.CompactCSR (u16 or delta-coded u8 columns + u8 values + overflow lists) and a
    scipy CSR with uint8 / uint16 data give bit-identical results to the float32 scipy matrix."""
    z = golden_square
    X = golden_csr(z).astype(np.float32)
    X.data[::997] += 300.0                                              # a few counts beyond the u8 range
    res = []
    for form in ("f32", "compact", "compact_delta", "u16"):
        Xi = {"f32": lambda: X, "compact": lambda: dv.CompactCSR.from_scipy(X, delta=False),
              "compact_delta": lambda: dv.CompactCSR.from_scipy(X, delta=True), "u16": lambda: X.astype(np.uint16)}[form]()
        ad = ch.AnnDataLite(Xi, obsm={"spatial": z["xy"]})
        ch.detect_svgs(ad, neighbors=8, top_svg=40, min_morans=0.05, geary=True)
        res.append((ad.var["Moran's I"].to_numpy(), ad.var["Geary's C"].to_numpy(), ad.var["spatially_variable"].to_numpy()))
    for r in res[1:]:
        assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(res[0], r))
    ref = osvg.detect_svgs(X, z["xy"], neighbors=8, top_svg=40, min_morans=0.05)
    m = ref["gene_mask"]
    assert_close(res[1][0][m], ref["morans"][m])
