"""Permutation inference on the device (csrc/moran_perm.cu, csrc/perm.cuh): the counter-based permutations against their
numpy statement, every simulated I against the oracle formula evaluated on the SAME permutation (exact parity, draw by
draw), the statistical parity esda's global-RNG draws allow, and BASELINE cfg3 at full shape (120k bins x 25k genes from
CSR integer counts, P = 999)."""
import time

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import chrysalis_b200 as ch
from chrysalis_b200 import device as dv
from chrysalis_b200 import perm as hperm
from chrysalis_b200 import synth
from oracle import moran as om
from oracle import svg as osvg
from oracle import weights as ow

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [2, 3, 5, 64, 900, 4035, 120_000])
def test_device_permutations_equal_the_numpy_statement(n):
    for seed, draw in ((1, 0), (0xDEADBEEF12345678, 3), (42, 998)):
        got = dv.perm_row_map(n, seed, draw, "cuda").cpu().numpy()
        assert np.array_equal(got, hperm.permutation(seed, draw, n))
        assert np.array_equal(np.sort(got), np.arange(n))


def _sim_reference(dense, nbr, key, P):
    """I of every column under the device's own permutations, from the oracle formula (float64): spot i takes z[pi(i)]."""
    n = dense.shape[0]
    w = ow.row_standardised_csr(nbr)
    Z = dense.astype(np.float64) - dense.astype(np.float64).mean(0)
    zz = np.einsum("ij,ij->j", Z, Z)
    sims = np.empty((P, dense.shape[1]))
    for d in range(P):
        Zp = Z[hperm.permutation(key, d, n)]
        sims[d] = np.einsum("ij,ij->j", Zp, w @ Zp) / zz            # n / s0 = 1 for row-standardised kNN weights
    return sims


def _blob(n_side, seed):
    xy = synth.make_coords("square", n_side * n_side, seed=seed)
    c = xy.mean(0)
    d = xy - c
    keep = (np.hypot(d[:, 0], d[:, 1]) < 0.45 * n_side * 100) & ~((np.abs(d[:, 0] - 600) < 350) & (np.abs(d[:, 1] + 200) < 250))
    return xy[keep]


@pytest.mark.parametrize("shape", ["rect_two_bands", "blob_with_hole"])
def test_lattice_permutation_kernel_is_exact_draw_by_draw(shape):
    if shape == "rect_two_bands":
        w_, h_ = 70, 75                                               # two bands (60 + 15 rows: masked rows), one chunk
        yy, xx = np.divmod(np.arange(w_ * h_), w_)
        xy = np.stack([xx * 80.0, yy * 80.0], 1) + np.random.default_rng(0).normal(0, 0.02, (w_ * h_, 2))
    else:
        xy = _blob(90, 4)
    n = len(xy)
    rng = np.random.default_rng(1)
    g, P, seed = 40, 24, 77
    dense = (np.log1p(rng.poisson(0.9, (n, g))) + np.sin(xy[:, :1] / 500.0) * rng.uniform(0, 1.2, g)[None]).astype(np.float32)
    nbr = ow.knn_neighbors(xy, 8)
    plan = dv.lattice_plan(torch.from_numpy(xy).cuda(), torch.from_numpy(nbr.astype(np.int32)).cuda())
    assert plan is not None
    Xd = torch.zeros((plan.n_rows, dv.dense_ld(g)), dtype=torch.float32, device="cuda")
    Xd[plan.row_of_spot, :g] = torch.from_numpy(dense).cuda()
    res = {k: v.cpu().numpy() for k, v in dv.moran_perm_lattice(Xd, plan, g, permutations=P, seed=seed).items()}
    I_ref = om.morans_I_batched(dense, nbr)
    assert np.all(np.abs(res["I"] - I_ref) <= 1e-5 * np.abs(I_ref) + 2e-7)
    sims = _sim_reference(dense, nbr, dv.perm_seed(seed), P)
    ei, vi = sims.mean(0), sims.var(0)
    assert np.abs(res["EI_sim"] - ei).max() < 2e-7                   # every draw reproduced: the moments agree to rounding
    assert np.all(np.abs(res["VI_sim"] - vi) <= 1e-4 * vi + 1e-12)
    larger = (sims >= I_ref[None]).sum(0)
    larger = np.minimum(larger, P - larger)
    assert np.array_equal(res["p_sim"], (larger + 1.0) / (P + 1.0))


@pytest.mark.parametrize("g", [40, 200, 256])
def test_bulk_copy_producer_is_bit_identical_to_the_cp_async_producers(g):
    """Default producer (one 512-byte bulk copy per gathered row) against kernel variant 2 (16-byte cp.async): same staged
    tiles, hence the same sums bit for bit; 200 genes leave a last slab narrower than 128 columns."""
    xy = _blob(120, 9)
    n = len(xy)
    nbr = ow.knn_neighbors(xy, 8)
    plan = dv.lattice_plan(torch.from_numpy(xy).cuda(), torch.from_numpy(nbr.astype(np.int32)).cuda())
    assert plan is not None
    Xd = torch.zeros((plan.n_rows, dv.dense_ld(g)), dtype=torch.float32, device="cuda")
    Xd[plan.row_of_spot, :g] = torch.from_numpy(np.log1p(np.random.default_rng(2).poisson(1.1, (n, g))).astype(np.float32)).cuda()
    a = dv.moran_perm_lattice(Xd.clone(), plan, g, permutations=19, seed=5, variant=0)
    b = dv.moran_perm_lattice(Xd.clone(), plan, g, permutations=19, seed=5, variant=2)
    for key in ("I", "p_sim", "EI_sim", "VI_sim", "z_sim"):
        assert torch.equal(a[key], b[key]), key


def test_general_path_uses_the_same_draws():
    """Hex lattice, k = 6 (quad-plan kernel + chr_perm_row_map): the same exact draw-by-draw parity."""
    xy = synth.make_coords("hex", 1300, seed=8)
    n = len(xy)
    rng = np.random.default_rng(2)
    g, P, seed = 30, 16, 5
    dense = (np.log1p(rng.poisson(0.8, (n, g))) + np.cos(xy[:, 1:] / 400.0) * rng.uniform(0, 1, g)[None]).astype(np.float32)
    nbr = ow.knn_neighbors(xy, 6)
    Xd = torch.zeros((n, dv.dense_ld(g)), dtype=torch.float32, device="cuda")
    Xd[:, :g] = torch.from_numpy(dense).cuda()
    res = {k: v.cpu().numpy() for k, v in dv.moran_permutation_test(Xd, torch.from_numpy(nbr.astype(np.int32)).cuda(), g, permutations=P, seed=seed).items()}
    sims = _sim_reference(dense, nbr, dv.perm_seed(seed), P)
    assert np.abs(res["EI_sim"] - sims.mean(0)).max() < 2e-7
    assert np.all(np.abs(res["VI_sim"] - sims.var(0)) <= 1e-4 * sims.var(0) + 1e-12)


def test_lattice_null_distribution_matches_the_analytic_moments():
    """P = 999 draws on a lattice: E[I] = -1/(n-1) and the randomisation variance VI_rand of esda's analytic moments."""
    xy = synth.make_coords("square", 64 * 64, seed=2)
    X, _ = synth.make_counts_csr(xy, 64, n_zones=5, seed=2)
    dense = np.asarray(osvg.normalize_total_log1p(X).todense(), dtype=np.float32)
    dense = dense[:, dense.std(0) > 0]
    n, g = dense.shape
    nbr = ow.knn_neighbors(xy, 8)
    plan = dv.lattice_plan(torch.from_numpy(xy).cuda(), torch.from_numpy(nbr.astype(np.int32)).cuda())
    Xd = torch.zeros((plan.n_rows, dv.dense_ld(g)), dtype=torch.float32, device="cuda")
    Xd[plan.row_of_spot, :g] = torch.from_numpy(dense).cuda()
    P = 999
    res = {k: v.cpu().numpy() for k, v in dv.moran_perm_lattice(Xd, plan, g, permutations=P, seed=11).items()}
    w = ow.row_standardised_csr(nbr)
    an = [om.moran_analytic(dense[:, j], w) for j in range(g)]
    ei = np.array([a["EI"] for a in an])
    vi = np.array([a["VI_rand"] for a in an])
    assert (np.abs(res["EI_sim"] - ei) < 5 * np.sqrt(vi / P)).all()
    ratio = res["VI_sim"] / vi
    assert abs(np.median(ratio) - 1.0) < 0.04 and (np.abs(ratio - 1.0) < 0.5).all()
    again = dv.moran_perm_lattice(Xd, plan, g, permutations=P, seed=11)
    assert np.array_equal(again["p_sim"].cpu().numpy(), res["p_sim"]) and np.array_equal(again["z_sim"].cpu().numpy(), res["z_sim"])


def test_cfg3_stereoseq_shape_csr_int_counts_999_permutations(capsys):
    """BASELINE cfg3: 120,000 bins x 25,000 genes, CSR integer counts, kNN k = 8, Moran's I with 999 permutations through
    ``ch.morans_i_permutation``.  I against the oracle on a gene subsample; p_sim / z_sim against the reference-style
    simulation (oracle.moran.moran_permutation, esda semantics) within Monte-Carlo error."""
    n, g, P = 120_000, 25_000, 999
    xy = synth.make_coords("square", n, seed=3)
    X, _ = synth.make_counts_csr(xy, g, n_zones=8, seed=3, device="cuda", log_rate_mean=-2.8, log_rate_sd=1.2)
    assert X.dtype == np.int32
    ad = ch.AnnDataLite(X, obsm={"spatial": xy})
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ch.morans_i_permutation(ad, neighbors=8, permutations=P, seed=1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    mask = osvg.filter_genes_mask(X, int(n * 0.05))
    got = ad.var["Moran's I"].to_numpy()
    assert np.array_equal(np.isnan(got), ~mask)
    kept = np.flatnonzero(mask)
    with capsys.disabled():
        print(f"\n[cfg3] morans_i_permutation: {n} bins x {g} genes ({len(kept)} scored), P = {P}: {dt:.2f} s end to end from host CSR")
    Xn = osvg.normalize_total_log1p(X[:, mask]).tocsc()
    nbr = ow.knn_neighbors(ow.reference_points(xy), 8)
    sub = np.unique(np.linspace(0, len(kept) - 1, 1536).astype(np.int64))
    dense = np.asarray(Xn[:, sub].todense(), dtype=np.float32)
    I_ref = om.morans_I_batched(dense, nbr)
    assert np.all(np.abs(got[kept[sub]] - I_ref) <= 1e-5 * np.abs(I_ref) + 2e-7)
    p_sim, z_sim = ad.var["Moran's I p_sim"].to_numpy()[kept], ad.var["Moran's I z_sim"].to_numpy()[kept]
    assert (p_sim >= 1.0 / (P + 1)).all() and (p_sim <= 0.5 + 1e-12).all()
    # reference-style simulation for a few genes across the range of I (one CSR.vector per draw: ~1 s per gene)
    w = ow.row_standardised_csr(nbr)
    order = np.argsort(I_ref)
    probe = order[[0, len(order) // 4, len(order) // 2, -len(order) // 8, -1]]
    for j in probe:
        ref = om.moran_permutation(dense[:, j], w, P, np.random.default_rng(int(j)))
        pj, zj = p_sim[sub[j]], z_sim[sub[j]]
        # both are 999-draw estimates of the same tail probability: binomial noise on each side
        se = np.sqrt(2 * max(ref["p_sim"], pj) * (1 - min(ref["p_sim"], pj)) / P) + 1.0 / (P + 1)
        assert abs(pj - ref["p_sim"]) <= 4 * se, (j, pj, ref["p_sim"])
        assert abs(zj - ref["z_sim"]) <= 0.15 * abs(ref["z_sim"]) + 0.6, (j, zj, ref["z_sim"])
