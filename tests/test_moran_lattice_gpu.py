"""GPU parity of the square-lattice stencil kernel (csrc/moran_lattice.cu, chr_moran_lattice_f32) against the CPU
oracle, the reference-run vectors and the general quad-plan kernel - all calls go through the C ABI.

W is always the oracle's cKDTree kNN (oracle.weights.knn_neighbors), so these tests compare the statistic, not the
neighbour search; tolerance as in tests/test_moran_gpu.py (1e-5 relative + 2e-7)."""
import numpy as np
import pytest
import torch

from chrysalis_b200 import device as dv
from chrysalis_b200 import synth
from conftest import golden_csr, load_golden
from oracle import moran as om
from oracle import svg as osvg
from oracle import weights as ow

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 2e-7


def assert_close(got, ref, rtol=RTOL, atol=ATOL, what="I"):
    got, ref = np.asarray(got), np.asarray(ref)
    err = np.abs(got - ref)
    bad = ~(err <= rtol * np.abs(ref) + atol)
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} off; worst abs {np.nanmax(err):.3e} at ref {ref[np.nanargmax(err)]:.6g}"


def lattice_case(xy, nbr, require=True):
    """(plan, device nbr) for host coordinates ``xy`` (y already negated or not - the lattice does not care)."""
    xt = torch.from_numpy(np.ascontiguousarray(xy, dtype=np.float64)).cuda()
    nb = torch.from_numpy(np.ascontiguousarray(nbr, dtype=np.int32)).cuda()
    plan = dv.lattice_plan(xt, nb)
    if require:
        assert plan is not None, "square lattice not recognised"
    return plan, nb


def gpu_lattice(dense, plan, geary=False, ld=None, **kw):
    n, g = dense.shape
    ld = dv.dense_ld(g) if ld is None else ld
    Xd = torch.zeros((plan.n_rows, ld), dtype=torch.float32, device="cuda")
    Xd[plan.row_of_spot, :g] = torch.from_numpy(dense).cuda()
    I, C, st = dv.moran_lattice(Xd, plan, g, geary=geary, **kw)
    torch.cuda.synchronize()
    return I.cpu().numpy(), (C.cpu().numpy() if geary else None), st


def square_blob(n_side, seed, jitter=1e-3, holes=True):
    """Square lattice cut to a wobbly disc with rectangular holes: ragged band ranges, absent nodes, concave edges."""
    xy = synth.make_coords("square", n_side * n_side, seed=seed, jitter=jitter)
    c = xy.mean(0)
    d = xy - c
    ang = np.arctan2(d[:, 1], d[:, 0])
    rad = np.hypot(d[:, 0], d[:, 1]) / (1.0 + 0.2 * np.sin(3 * ang + 0.5))
    keep = rad < 0.42 * n_side * 100.0
    if holes:
        keep &= ~((np.abs(d[:, 0] - 900) < 450) & (np.abs(d[:, 1] + 300) < 250))
        keep &= ~((np.abs(d[:, 0] + 1500) < 150) & (np.abs(d[:, 1] - 1100) < 650))
    return xy[keep]


def test_golden_square_fixture_and_reference_run_vectors(golden_square):
    z = golden_square
    m = z["gene_mask"]
    dense = np.asarray(osvg.normalize_total_log1p(golden_csr(z)[:, m]).todense(), dtype=np.float32)
    nbr = ow.knn_neighbors(ow.reference_points(z["xy"]), 8)
    plan, _ = lattice_case(ow.reference_points(z["xy"]), nbr)
    I, C, _ = gpu_lattice(dense, plan, geary=True)
    assert_close(I, z["morans"][m])
    assert_close(C, z["geary"][m], what="C")
    ref = load_golden("reference_get_moransI")          # outputs of the reference's own get_moransI (functions.py:15-35)
    cols = np.searchsorted(np.flatnonzero(m), ref["svg_square_k8_genes"])
    assert_close(I[cols], ref["svg_square_k8_I_f64in"], atol=ATOL + 5e-9)


@pytest.mark.parametrize("w,h,g", [(300, 70, 130), (64, 61, 5), (45, 45, 128), (520, 121, 257), (10, 200, 33)])
def test_bands_chunks_and_ragged_gene_counts(w, h, g):
    """Several bands (the last one short: masked rows), several chunks per band (> 256 columns), gene counts that are
    not multiples of 4 / 128, every row stride the ABI allows."""
    rng = np.random.default_rng(w * h + g)
    yy, xx = np.divmod(np.arange(w * h), w)
    xy = np.stack([xx * 55.0, yy * 55.0], 1) + rng.normal(0, 0.02, (w * h, 2))
    dense = (np.log1p(rng.poisson(0.6, (w * h, g))) + 0.3 * np.sin(xx / 9.0)[:, None] * rng.uniform(0, 1, g)[None]).astype(np.float32)
    nbr = ow.knn_neighbors(xy, 8)
    plan, _ = lattice_case(xy, nbr)
    assert plan.bands.shape[0] == (h + 59) // 60
    ref = om.morans_I_batched(dense, nbr)
    refC = om.gearys_C_batched(dense, nbr)
    for ld in (dv.dense_ld(g), (g + 3) // 4 * 4, dv.dense_ld(g) + 96):
        I, C, _ = gpu_lattice(dense, plan, geary=True, ld=ld)
        assert_close(I, ref)
        assert_close(C, refC, what="C")


@pytest.mark.parametrize("jitter", [1e-3, 0.0])
def test_tissue_outline_with_holes(jitter):
    """Absent nodes inside the stored range, ragged per-band column ranges, kNN reaching across holes (corrections);
    un-jittered: the oracle's own tie choices at every edge are taken as W."""
    xy = square_blob(150, 5, jitter=jitter)
    n = len(xy)
    rng = np.random.default_rng(1)
    g = 96
    dense = (np.log1p(rng.poisson(0.8, (n, g))) + (np.sin(xy[:, :1] / 700.0) * rng.uniform(0, 1.5, g)[None])).astype(np.float32)
    nbr = ow.knn_neighbors(xy, 8)
    plan, nb = lattice_case(xy, nbr)
    assert plan.absent.numel() > 0 and plan.corr.shape[0] > 0 and plan.n_rows > n
    I, C, _ = gpu_lattice(dense, plan, geary=True)
    assert_close(I, om.morans_I_batched(dense, nbr))
    assert_close(C, om.gearys_C_batched(dense, nbr), what="C")
    # and the general kernel on the same W agrees (two independent device paths)
    Xd = torch.zeros((n, dv.dense_ld(g)), dtype=torch.float32, device="cuda")
    Xd[:, :g] = torch.from_numpy(dense).cuda()
    I2, _, _ = dv.moran_dense(Xd, nb, g)
    assert_close(I, I2.cpu().numpy())


def test_gene_sharding_is_bit_identical_and_runs_reproduce():
    xy = square_blob(120, 2, holes=False)
    n = len(xy)
    rng = np.random.default_rng(7)
    g = 300
    dense = np.log1p(rng.poisson(0.5, (n, g))).astype(np.float32)
    nbr = ow.knn_neighbors(xy, 8)
    plan, _ = lattice_case(xy, nbr)
    I, _, _ = gpu_lattice(dense, plan)
    I_again, _, _ = gpu_lattice(dense, plan)
    assert np.array_equal(I, I_again)
    for lo, hi in ((0, 128), (128, 300), (100, 229)):
        Is, _, _ = gpu_lattice(np.ascontiguousarray(dense[:, lo:hi]), plan)
        assert np.array_equal(Is, I[lo:hi]), f"gene shard [{lo}, {hi}) differs bitwise"


def test_cancellation_constant_gene_and_preshifted():
    xy = synth.make_coords("square", 90 * 90, seed=3)
    n = len(xy)
    rng = np.random.default_rng(3)
    base = rng.normal(0, 1, n)
    dense = np.stack([7.0 + 0.01 * base, 100.0 + rng.normal(0, 0.5, n), np.full(n, 2.5), 5.0 + np.sin(xy[:, 0] / 300.0)], 1).astype(np.float32)
    nbr = ow.knn_neighbors(xy, 8)
    plan, _ = lattice_case(xy, nbr)
    I, _, _ = gpu_lattice(dense, plan)
    ref64 = om.morans_I_batched(dense, nbr, dtype=np.float64)
    assert np.isnan(I[2])
    assert_close(I[[0, 1, 3]], ref64[[0, 1, 3]])
    centred = dense[:, [0, 1, 3]] - dense[:, [0, 1, 3]].mean(0, keepdims=True).astype(np.float32)
    I2, _, _ = gpu_lattice(centred, plan, preshifted=True)
    assert_close(I2, ref64[[0, 1, 3]])


def test_non_lattices_are_declined():
    rng = np.random.default_rng(0)
    xy = rng.uniform(0, 1000, (3000, 2))
    assert lattice_case(xy, ow.knn_neighbors(xy, 8), require=False)[0] is None            # random beads
    hexy = synth.make_coords("hex", 3000, seed=1)
    assert lattice_case(hexy, ow.knn_neighbors(hexy, 8), require=False)[0] is None        # hex lattice
    sq = synth.make_coords("square", 2500, seed=1)
    assert lattice_case(sq, ow.knn_neighbors(sq, 6), require=False)[0] is None            # k != 8


def test_detect_svgs_takes_the_lattice_path_and_matches_the_oracle(monkeypatch):
    import chrysalis_b200 as ch
    xy = square_blob(70, 11)
    X, _ = synth.make_counts_csr(xy, 260, n_zones=5, seed=11)
    ref = osvg.detect_svgs(X, xy, neighbors=8, top_svg=50, min_morans=0.05, geary=True)
    calls = []
    real, real_csr = dv.moran_lattice, dv.moran_lattice_csr
    monkeypatch.setattr(dv, "moran_lattice", lambda *a, **k: (calls.append("dense"), real(*a, **k))[1])
    monkeypatch.setattr(dv, "moran_lattice_csr", lambda *a, **k: (calls.append("csr"), real_csr(*a, **k))[1])
    monkeypatch.setenv("CHR_MORAN_CSR", "0")                              # the dense-matrix flavour of the stencil kernel
    ad = ch.AnnDataLite(X, obsm={"spatial": xy})
    ch.detect_svgs(ad, neighbors=8, top_svg=50, min_morans=0.05, geary=True)
    assert calls == ["dense"], "the stencil kernel was not used"
    m = ref["gene_mask"]
    got = ad.var["Moran's I"].to_numpy()
    assert np.array_equal(np.isnan(got), ~m)
    assert_close(got[m], ref["morans"][m])
    assert_close(ad.var["Geary's C"].to_numpy()[m], ref["geary"][m], what="C")
    assert np.array_equal(ad.var["spatially_variable"].to_numpy(), ref["spatially_variable"])


# ---- the same kernel fed from the CSR counts (chr_moran_lattice_csr_f32): no dense matrix ------------------------------

def _csr_case(xy, g, seed, density=0.08):
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    n = len(xy)
    lam = density * (1.0 + 2.0 * (np.sin(xy[:, :1] / 600.0) > 0) * (rng.uniform(0, 1, g)[None] > 0.6))
    counts = rng.poisson(lam * rng.gamma(2.0, 1.0, (1, g))).astype(np.float32)
    counts[:, ::17] = 0                                                   # genes the filter drops
    counts[rng.integers(0, n, 30), rng.integers(0, g, 30)] += 400          # a few large counts
    return sp.csr_matrix(counts)


@pytest.mark.parametrize("shape", ["rect", "blob"])
@pytest.mark.parametrize("logged", [False, True])
def test_csr_fed_kernel_matches_oracle_and_dense_path(shape, logged):
    xy = square_blob(110, 4) if shape == "blob" else synth.make_coords("square", 75 * 131, seed=2)      # 131 rows: 3 bands, last short
    n, g = len(xy), 300
    X = _csr_case(xy, g, 5)
    nbr = ow.knn_neighbors(xy, 8)
    plan, _ = lattice_case(xy, nbr)
    csr = dv.upload_csr(X)
    min_cells = int(0.03 * n)
    gene_nnz = dv.csr_gene_nnz(csr)
    keep = gene_nnz >= min_cells
    gene_col = torch.where(keep, torch.cumsum(keep.to(torch.int32), 0, dtype=torch.int32) - 1, torch.full_like(gene_nnz, -1))
    n_keep = int(keep.sum())
    assert 0 < n_keep < g
    scale = None if logged else dv.median_scale(dv.csr_spot_totals(csr, gene_col))
    if logged:                                                           # the caller already normalised: values are used as they are
        Xl = osvg.normalize_total_log1p(X[:, keep.cpu().numpy()])
        Xfull = X.copy().astype(np.float32)
        csr = dv.upload_csr(osvg.normalize_total_log1p(X))
        dense_ref = np.asarray(osvg.normalize_total_log1p(X)[:, keep.cpu().numpy()].todense(), dtype=np.float32)
    else:
        dense_ref = np.asarray(osvg.normalize_total_log1p(X[:, keep.cpu().numpy()]).todense(), dtype=np.float32)
    I, C, st, unsorted = dv.moran_lattice_csr(csr, gene_col, n_keep, scale, plan, apply_log1p=not logged, geary=True, want_stats=True)
    assert int(unsorted.item()) == 0
    I, C = I.cpu().numpy(), C.cpu().numpy()
    assert_close(I, om.morans_I_batched(dense_ref, nbr))
    assert_close(C, om.gearys_C_batched(dense_ref, nbr), what="C")
    assert np.allclose(st[0].cpu().numpy(), dense_ref.astype(np.float64).mean(0), rtol=1e-6, atol=1e-9)
    # the dense path on the same counts: the two differ only in the pilot sample's summation order
    dense = torch.zeros((plan.n_rows, dv.dense_ld(n_keep)), dtype=torch.float32, device="cuda")
    dv.csr_densify(csr, gene_col, n_keep, scale, plan.row_of_spot, apply_log1p=not logged, out=dense)
    I2, C2, _ = dv.moran_lattice(dense, plan, n_keep, geary=True)
    assert_close(I, I2.cpu().numpy())                                     # two float32 evaluations with different shift vectors
    assert_close(C, C2.cpu().numpy(), what="C")
    again = dv.moran_lattice_csr(csr, gene_col, n_keep, scale, plan, apply_log1p=not logged, geary=True)
    assert np.array_equal(again[0].cpu().numpy(), I)                      # run-to-run bit reproducible


def test_csr_fed_kernel_many_slabs_ragged_and_gene_shards():
    """More than one gene slab, a ragged last slab, and gene shards of the same matrix give the same bits."""
    xy = synth.make_coords("square", 64 * 70, seed=6)
    n, g = len(xy), 450
    X = _csr_case(xy, g, 9, density=0.12)
    nbr = ow.knn_neighbors(xy, 8)
    plan, _ = lattice_case(xy, nbr)
    dense_all = np.asarray(X.todense(), dtype=np.float32)

    def run(lo, hi):
        csr = dv.upload_csr(X[:, lo:hi].tocsr())
        gene_col = torch.arange(hi - lo, dtype=torch.int32, device="cuda")
        I, _, _, uns = dv.moran_lattice_csr(csr, gene_col, hi - lo, None, plan, apply_log1p=True)
        assert int(uns.item()) == 0
        return I.cpu().numpy()

    keep = np.flatnonzero(np.log1p(dense_all).std(0) > 0)
    I = run(0, g)
    assert_close(I[keep], om.morans_I_batched(np.log1p(dense_all[:, keep]), nbr))
    for lo, hi in ((0, 128), (128, 450), (200, 333)):
        assert np.array_equal(run(lo, hi), I[lo:hi], equal_nan=True), f"gene shard [{lo}, {hi}) differs bitwise"


def test_detect_svgs_is_fed_from_csr_on_lattices_and_falls_back_when_unsorted(monkeypatch):
    import chrysalis_b200 as ch
    import scipy.sparse as sp
    xy = square_blob(70, 11)
    X, _ = synth.make_counts_csr(xy, 260, n_zones=5, seed=11)
    ref = osvg.detect_svgs(X, xy, neighbors=8, top_svg=50, min_morans=0.05, geary=True)
    calls = {"csr": 0, "dense": 0}
    real_csr, real_dense = dv.moran_lattice_csr, dv.moran_lattice
    monkeypatch.setattr(dv, "moran_lattice_csr", lambda *a, **k: (calls.__setitem__("csr", calls["csr"] + 1), real_csr(*a, **k))[1])
    monkeypatch.setattr(dv, "moran_lattice", lambda *a, **k: (calls.__setitem__("dense", calls["dense"] + 1), real_dense(*a, **k))[1])

    def check(Xin, expect):
        before = dict(calls)
        ad = ch.AnnDataLite(Xin, obsm={"spatial": xy})
        ch.detect_svgs(ad, neighbors=8, top_svg=50, min_morans=0.05, geary=True)
        assert {k: calls[k] - before[k] for k in calls} == expect
        m = ref["gene_mask"]
        got = ad.var["Moran's I"].to_numpy()
        assert np.array_equal(np.isnan(got), ~m)
        assert_close(got[m], ref["morans"][m])
        assert_close(ad.var["Geary's C"].to_numpy()[m], ref["geary"][m], what="C")
        assert np.array_equal(ad.var["spatially_variable"].to_numpy(), ref["spatially_variable"])

    check(X, {"csr": 1, "dense": 0})
    # the same matrix with the entries of every row reversed: a legal CSR, not sorted -> flagged on the device, dense path
    rev = np.concatenate([np.arange(X.indptr[i + 1] - 1, X.indptr[i] - 1, -1) for i in range(X.shape[0])])
    Xu = sp.csr_matrix((X.data[rev], X.indices[rev], X.indptr), shape=X.shape)
    assert not Xu.has_sorted_indices
    check(Xu, {"csr": 1, "dense": 1})


@pytest.mark.parametrize("shape", ["rect_short_last_band", "blob", "tiny_single_band"])
def test_tma_producer_variant_is_bit_identical_to_the_cp_async_producers(shape):
    """The default TMA producer (one cp.async.bulk.tensor box per staged column) and kernel variant 2 (cp.async producer
    warps) deliver the same tiles: identical bits."""
    if shape == "blob":
        xy = square_blob(130, 8)
    elif shape == "tiny_single_band":
        xy = synth.make_coords("square", 30 * 30, seed=1)
    else:
        xy = synth.make_coords("square", 300 * 131, seed=2)[: 300 * 131]
    n = len(xy)
    rng = np.random.default_rng(5)
    g = 200
    dense = (np.log1p(rng.poisson(0.7, (n, g))) + np.sin(xy[:, :1] / 800.0) * rng.uniform(0, 1, g)[None]).astype(np.float32)
    nbr = ow.knn_neighbors(xy, 8)
    plan, _ = lattice_case(xy, nbr)
    I0, C0, _ = gpu_lattice(dense, plan, geary=True, variant=2)
    for ld in (dv.dense_ld(g), dv.dense_ld(g) + 96):
        I1, C1, _ = gpu_lattice(dense, plan, geary=True, ld=ld, variant=0)
        assert np.array_equal(I0, I1) and np.array_equal(C0, C1)
    assert_close(I0, om.morans_I_batched(dense, nbr))
