"""Host-side logic of detect_svgs (field contract, name handling, selection rule) with the device
stage replaced by the oracle's numbers - runs on the CPU box."""
import numpy as np
import pytest

import chrysalis_b200 as ch
from chrysalis_b200 import core
from conftest import golden_csr


def _oracle_stage(z):
    def stage(X, spatial, min_spots, neighbors, geary, already_log1p, timing=False, **kw):
        m = z["gene_mask"]
        return m, z["morans"][m], (z["geary"][m] if geary else None)
    return stage


@pytest.mark.parametrize("name", ["golden_hex", "golden_random", "golden_square"])
def test_fields_and_selection_match_oracle(name, request, monkeypatch):
    z = request.getfixturevalue(name)
    monkeypatch.setattr(core, "svg_stage", _oracle_stage(z))
    ad = ch.AnnDataLite(golden_csr(z), obsm={"spatial": z["xy"]})
    X_before = ad.X.copy()
    assert ch.detect_svgs(ad, neighbors=int(z["k"]), top_svg=int(z["top_svg"]), min_morans=float(z["min_morans"]),
                          geary=True) is None
    got = ad.var["Moran's I"].to_numpy()
    assert np.array_equal(np.isnan(got), ~z["gene_mask"])          # filtered genes -> NaN (core.py:80)
    assert np.array_equal(got[z["gene_mask"]], z["morans"][z["gene_mask"]])
    assert np.array_equal(ad.var["Geary's C"].to_numpy()[z["gene_mask"]], z["geary"][z["gene_mask"]])
    assert ad.var["spatially_variable"].dtype == bool
    assert np.array_equal(ad.var["spatially_variable"].to_numpy(), z["spatially_variable"])
    assert (ad.X != X_before).nnz == 0                              # caller's X untouched
    assert "log1p" not in ad.uns


def test_min_spots_assertion():
    z = None
    ad = ch.AnnDataLite(np.zeros((4, 3), np.float32), obsm={"spatial": np.zeros((4, 2))})
    for bad in (0, 1, -0.1, 1.5):
        with pytest.raises(AssertionError):
            ch.detect_svgs(ad, min_spots=bad)


def test_make_unique_names():
    out = core._make_unique(["a", "b", "a", "a", "b", "a-1"])
    assert out.is_unique and out[0] == "a" and out[1] == "b" and out[5] == "a-1"
    assert set(out[[2, 3]]) == {"a-2", "a-3"} and out[4] == "b-1"


def test_duplicate_var_names_follow_pandas_alignment(golden_hex, monkeypatch):
    z = golden_hex
    monkeypatch.setattr(core, "svg_stage", _oracle_stage(z))
    names = [f"g{j}" for j in range(int(z["shape"][1]))]
    kept = np.flatnonzero(z["gene_mask"])
    names[kept[1]] = names[kept[0]]                                  # duplicate among scored genes
    ad = ch.AnnDataLite(golden_csr(z), obsm={"spatial": z["xy"]}, var_names=names)
    ch.detect_svgs(ad, top_svg=int(z["top_svg"]))
    v = ad.var["Moran's I"].to_numpy()
    assert v[kept[0]] == v[kept[1]] == z["morans"][kept[0]]          # both rows align to the first name


def test_top_svg_rule_branches(golden_hex, monkeypatch):
    z = golden_hex
    monkeypatch.setattr(core, "svg_stage", _oracle_stage(z))
    from oracle import svg as osvg
    for top, thr in [(5, 0.2), (1000, 0.2), (10, -1.0), (50, 0.05)]:
        ad = ch.AnnDataLite(golden_csr(z), obsm={"spatial": z["xy"]})
        ch.detect_svgs(ad, top_svg=top, min_morans=thr)
        exp = osvg.select_svgs(z["morans"], z["gene_mask"], top, thr)
        assert np.array_equal(ad.var["spatially_variable"].to_numpy(), exp), (top, thr)


def test_aa_argument_errors_precede_any_device_work():
    ad = ch.AnnDataLite(np.zeros((4, 3), np.float32), obsm={"spatial": np.zeros((4, 2))})
    with pytest.raises(TypeError):
        ch.aa(ad, n_archetypes=8.0)                                  # core.py:176-177
    with pytest.raises(ValueError, match="n_archetypes cannot be less than 2"):
        ch.aa(ad, n_archetypes=1)                                    # core.py:178-179
    with pytest.raises(KeyError):
        ch.aa(ad, n_archetypes=3)                                    # obsm['chr_X_pca'] missing


def test_furthest_sum_matches_oracle_draws():
    from oracle import aa as oaa
    rng = np.random.default_rng(5)
    X = rng.normal(size=(500, 7))
    for a in (3, 8, 15):
        r1, r2 = np.random.RandomState(42), np.random.RandomState(42)
        assert np.array_equal(core._furthest_sum(X, a, r1), oaa.furthest_sum(X, a, r2))
        assert r1.randint(1 << 30) == r2.randint(1 << 30)           # same number of draws consumed


def test_integrate_adatas_union_and_varm(golden_hex, golden_random, monkeypatch):
    """utils.py:161-250 with the device stage replaced by the oracle's numbers: per-sample columns, union, inner join."""
    from oracle import svg as osvg
    samples, refs = [], []
    for z, drop in ((golden_hex, 0), (golden_random, 3)):
        X = golden_csr(z)
        g = min(X.shape[1], 300) - drop                              # different gene sets -> inner join
        names = [f"gene{j}" for j in range(g)]
        ad = ch.AnnDataLite(X[:, :g], obsm={"spatial": z["xy"]}, var_names=names)
        ref = osvg.detect_svgs(X[:, :g], z["xy"], neighbors=int(z["k"]), top_svg=30)
        refs.append(ref)
        samples.append(ad)

    it = iter(refs)

    def stage(samples, min_spots, neighbors, geary):
        out = []
        for _ in samples:                                            # samples with different gene lists arrive in separate calls
            r = next(it)
            m = r["gene_mask"]
            out.append((m, r["morans"][m], None))
        return out
    monkeypatch.setattr(core, "svg_stage_multi", stage)
    # each sample keeps its own neighbour count through kwargs of the per-sample call (here the same for both)
    out = ch.integrate_adatas(samples, sample_names=["a", "b"], calculate_svgs=True, top_svg=30, neighbors=6)
    g_common = 297
    assert out.shape == (len(samples[0]) + len(samples[1]), g_common)
    assert list(out.obs["sample"].astype(str).unique()) == ["a", "b"]
    assert out.obs_names[0].endswith("-a") and out.obs_names[-1].endswith("-b")
    sv = out.varm["spatially_variable"]
    assert list(sv.columns[:2]) == ["spatially_variable_a", "spatially_variable_b"]
    union = sv["spatially_variable_a"].to_numpy(bool) | sv["spatially_variable_b"].to_numpy(bool)
    assert np.array_equal(out.var["spatially_variable"].to_numpy(), union)
    assert np.array_equal(sv["spatially_variable_a"].to_numpy(bool), refs[0]["spatially_variable"][:g_common])
    mi = out.varm["Moran's I"]
    m0 = refs[0]["gene_mask"][:g_common]
    assert np.array_equal(mi["Moran's I_a"].to_numpy()[m0], refs[0]["morans"][:g_common][m0])
    assert "spatially_variable_a" not in out.var.columns and "Moran's I_a" not in out.var.columns
    assert out.obsm["spatial"].shape == (out.shape[0], 2)
    with pytest.raises(Exception, match="sample_id_col is already present"):
        ch.integrate_adatas(samples, sample_names=["a", "b"])


def test_estimate_compartments_fields(monkeypatch):
    n = 60
    ad = ch.AnnDataLite(np.zeros((n, 5), np.float32), obsm={"spatial": np.zeros((n, 2)), "chr_X_pca": np.random.default_rng(0).normal(size=(n, 8))})

    def fake_sweep(X, counts, max_iter=10, **kw):
        return {a: {"rss": float(100 - a), "entropy": np.full(X.shape[0], np.log(a)), "n_iter": max_iter} for a in counts}
    monkeypatch.setattr(core, "aa_sweep", fake_sweep)
    ch.estimate_compartments(ad, n_pcs=6, range_archetypes=(3, 7), max_iter=4)
    assert ad.obsm["entropy"].shape == (n, 4)
    assert np.allclose(ad.obsm["entropy"][:, 0], np.log(3))
    assert ad.uns["chr_aa"]["RSSs"] == {3: 97.0, 4: 96.0, 5: 95.0, 6: 94.0}
    ad2 = ch.AnnDataLite(np.zeros((4, 3), np.float32), obsm={"spatial": np.zeros((4, 2))})
    with pytest.raises(ValueError, match="chr_X_pca"):
        ch.estimate_compartments(ad2)


@pytest.mark.parametrize("kind,n,expect", [("square", 90000, 100.0), ("square", 4000, 100.0), ("hex", 20000, None), ("random", 20000, None)])
def test_square_lattice_detection(kind, n, expect):
    """The row-order heuristic of device.spatial_order: a (jittered) square lattice yields its pitch to 1e-4 even though
    the nearest-neighbour distance is biased low by the jitter; hex lattices and random beads are left alone."""
    import torch
    from scipy.spatial import cKDTree
    from chrysalis_b200 import device as dv, synth
    xy = synth.make_coords(kind, n, seed=3)
    d1 = cKDTree(xy).query(xy, k=2)[0][:, 1]
    t = torch.from_numpy(xy)
    bbox = torch.cat([t.min(0).values, t.max(0).values])
    p = dv.lattice_pitch_from_nn(t, bbox, torch.from_numpy(d1))
    if expect is None:
        assert p is None
    else:
        assert np.median(d1) < expect * (1 - 5e-4)          # the raw estimate drifts by more than a node over the lattice
        assert abs(p - expect) < 1e-2
    # exact lattice, duplicated coordinates (median distance 0): no lattice claimed, no division blow-up
    dup = torch.zeros((100, 2), dtype=torch.float64)
    assert dv.lattice_pitch_from_nn(dup, torch.zeros(4, dtype=torch.float64), torch.zeros(100, dtype=torch.float64)) is None



@pytest.mark.parametrize("shape", ["full", "blob", "exact"])
def test_lattice_plan_splits_the_knn_adjacency_exactly(shape):
    """device.lattice_plan (plain tensor arithmetic, runs on the CPU): the kNN adjacency A equals the 8-neighbour
    stencil B between present nodes plus the correction list D, and the (indeg - k) weights give A's column sums -
    the identity the stencil kernel (csrc/moran_lattice.cu) rests on.  W is the oracle's cKDTree kNN."""
    import torch
    from chrysalis_b200 import device as dv
    from chrysalis_b200 import synth
    from oracle import weights as ow

    xy = synth.make_coords("square", 70 * 70 if shape == "blob" else 2000, seed=3, jitter=0.0 if shape == "exact" else 1e-3)
    if shape == "blob":
        c = xy.mean(0)
        keep = (np.hypot(*(xy - c).T) < 3000) & ~((np.abs(xy[:, 0] - c[0]) < 400) & (np.abs(xy[:, 1] - c[1]) < 300))
        xy = xy[keep]
    n = len(xy)
    pts = ow.reference_points(xy)
    nbr = ow.knn_neighbors(pts, 8)
    xt = torch.from_numpy(pts)
    bbox = torch.tensor([pts[:, 0].min(), pts[:, 1].min(), pts[:, 0].max(), pts[:, 1].max()])
    pitch = dv.lattice_pitch_from_nn(xt, bbox, torch.from_numpy(np.linalg.norm(pts[nbr[:, 0]] - pts, axis=1)))
    plan = dv.lattice_plan(xt, torch.from_numpy(nbr.astype(np.int32)), bbox, pitch)
    assert plan is not None and plan.n_rows >= n and plan.absent.numel() == plan.n_rows - n
    rows = plan.row_of_spot.numpy()
    assert len(np.unique(rows)) == n and rows.max() < plan.n_rows
    # every chunk lies inside its band's stored range, and the chunks tile the bands
    bands, chunks = plan.bands.numpy(), plan.chunks.numpy()
    assert (bands[:, 3] <= dv.LATTICE_BAND_ROWS).all() and int((bands[:, 2] * bands[:, 3]).sum()) == plan.n_rows
    assert all(chunks[chunks[:, 0] == b, 2].sum() == bands[b, 2] for b in range(len(bands)))
    # x'Ax' = stencil over the layout + corrections;  1'A x = k sum x + weights
    x = np.random.default_rng(0).normal(size=n)
    A = np.zeros((n, n))
    np.add.at(A, (np.repeat(np.arange(n), 8), nbr.ravel()), 1.0)
    ix, iy = (t.numpy() for t in dv.lattice_nodes(xt, bbox, pitch))
    Wd, Hd = plan.lattice_shape
    g = np.zeros((Hd + 2, Wd + 2))
    g[iy + 1, ix + 1] = x
    Q = sum((g[1:-1, 1:-1] * g[1 + dy:Hd + 1 + dy, 1 + dx:Wd + 1 + dx]).sum() for dy in (-1, 0, 1) for dx in (-1, 0, 1) if dx or dy)
    xr = np.zeros(plan.n_rows)
    xr[rows] = x
    corr = plan.corr.numpy()
    coef = corr[:, 2].copy().view(np.float32).astype(np.float64)
    quad = corr[:, 3] == 0
    assert Q + (coef[quad] * xr[corr[quad, 0]] * xr[corr[quad, 1]]).sum() == pytest.approx(x @ A @ x, rel=1e-12, abs=1e-9)
    assert 8 * x.sum() + (coef[~quad] * xr[corr[~quad, 0]]).sum() == pytest.approx(A.sum(0) @ x, rel=1e-12, abs=1e-9)


def test_compact_csr_round_trip_and_limits():
    """device.CompactCSR (the 2-3-bytes-per-entry loader format of detect_svgs): exact round trip incl. counts >= 255, uint16
    columns and delta-coded columns (with and without escapes), automatic choice of the smaller form."""
    import scipy.sparse as sp
    from chrysalis_b200 import device as dv
    rng = np.random.default_rng(0)
    dense = rng.poisson(0.3, (200, 700)).astype(np.float32)
    dense[rng.integers(0, 200, 40), rng.integers(0, 700, 40)] = rng.integers(255, 70000, 40)
    dense[7] = 0                                                                   # an empty row
    X = sp.csr_matrix(dense)
    c = dv.CompactCSR.from_scipy(X, pin=False, delta=False)
    assert c.cols.dtype == torch_dtype("uint16") and c.vals.dtype == torch_dtype("uint8") and c.over_pos.numel() == int((X.data >= 255).sum())
    assert (c.to_scipy() != X).nnz == 0 and c.shape == X.shape
    assert c.nbytes < 0.5 * (X.data.nbytes + X.indices.nbytes + X.indptr.nbytes) + 8 * c.over_pos.numel() + X.indptr.nbytes
    d = dv.CompactCSR.from_scipy(X, pin=False)                                     # dense rows: deltas, no escapes
    assert d.deltas is not None and d.deltas.dtype == torch_dtype("uint8") and d.gap_pos.numel() == 0 and d.nbytes < c.nbytes
    assert (d.to_scipy() != X).nnz == 0 and np.array_equal(d.columns(), X.indices)
    # sparse rows over many genes: gaps >= 255 are escaped; forced deltas still round-trip, the automatic choice is uint16
    S = sp.random(300, 60000, density=0.002, format="csr", random_state=3, data_rvs=lambda k: rng.integers(1, 9, k)).astype(np.float32)
    e = dv.CompactCSR.from_scipy(S, pin=False, delta=True)
    assert e.gap_pos.numel() > 0 and (e.to_scipy() != S).nnz == 0 and np.array_equal(e.columns(), S.indices)
    assert dv.CompactCSR.from_scipy(S, pin=False).deltas is None
    # unsorted rows are sorted for the delta coding; more genes than uint16 holds need it
    U = sp.csr_matrix((np.array([1, 2, 3], dtype=np.float32), np.array([5, 1, 3]), np.array([0, 3])), shape=(1, 70000))
    u = dv.CompactCSR.from_scipy(U, pin=False)
    assert u.deltas is not None and (u.to_scipy() != U).nnz == 0
    with pytest.raises(ValueError):
        dv.CompactCSR.from_scipy(sp.csr_matrix(dense * 0.5), pin=False)            # not counts
    with pytest.raises(ValueError):
        dv.CompactCSR.from_scipy(sp.csr_matrix((3, 70000), dtype=np.float32), pin=False, delta=False)   # too many genes for uint16


def torch_dtype(name):
    import torch
    return getattr(torch, name)


def test_lattice_layout_band_tables_against_brute_force():
    """device.lattice_layout: per band the first / last occupied column, rows and row numbering - against a numpy loop; an empty
    band in the middle, a short last band, and two spots on one node (no layout)."""
    import torch
    from chrysalis_b200 import device as dv

    rng = np.random.default_rng(4)
    W, H = 83, 3 * dv.LATTICE_BAND_ROWS + 17
    keep = rng.random((H, W)) < 0.55
    keep[dv.LATTICE_BAND_ROWS:2 * dv.LATTICE_BAND_ROWS] = False                    # band 1 is empty
    keep[:, :5] &= rng.random((H, 5)) < 0.1
    keep[0, 0] = keep[H - 1, W - 1] = True                                          # the grid spans (W, H)
    iy, ix = np.nonzero(keep)
    order = rng.permutation(len(ix))
    ix, iy = ix[order], iy[order]
    row_of_spot, n_rows, bands, chunks, shape = dv.lattice_layout(torch.from_numpy(ix), torch.from_numpy(iy))
    assert shape == (W, H)
    bands = bands.numpy()
    exp_rows, row0 = np.empty(len(ix), dtype=np.int64), 0
    for b in range(len(bands)):
        inb = iy // dv.LATTICE_BAND_ROWS == b
        bh = min(dv.LATTICE_BAND_ROWS, H - b * dv.LATTICE_BAND_ROWS)
        if not inb.any():
            assert tuple(bands[b]) == (row0, 0, 0, bh)
            continue
        x0, x1 = ix[inb].min(), ix[inb].max()
        assert tuple(bands[b]) == (row0, x0, x1 - x0 + 1, bh)
        exp_rows[inb] = row0 + (ix[inb] - x0) * bh + (iy[inb] - b * dv.LATTICE_BAND_ROWS)
        row0 += (x1 - x0 + 1) * bh
    assert n_rows == row0 and np.array_equal(row_of_spot.numpy(), exp_rows)
    ch = chunks.numpy()
    assert all(ch[ch[:, 0] == b, 2].sum() == bands[b, 2] for b in range(len(bands))) and (ch[:, 2] <= dv.LATTICE_CHUNK_COLS).all()
    # two spots on one node
    out = dv.lattice_layout(torch.from_numpy(np.append(ix, ix[0])), torch.from_numpy(np.append(iy, iy[0])))
    assert out[0] is None
