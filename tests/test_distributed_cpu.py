"""World-size-2 checks of the multi-GPU plumbing on the CPU (gloo): shard ranges, ragged
all-gather, all-reduce, and the gene-sharded detect_svgs host flow with the device kernels
replaced by the oracle's arithmetic (the CUDA kernels themselves are covered by the -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, GOLDEN


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, size, port, fn_name, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        globals()[fn_name](rank, size, out_dir)
    finally:
        dist.destroy_process_group()


def _run(fn_name, tmp_path, size=2):
    mp.spawn(_worker, args=(size, _free_port(), fn_name, str(tmp_path)), nprocs=size, join=True)


# ---- helpers -------------------------------------------------------------------------------
def _check_helpers(rank, size, out_dir):
    from chrysalis_b200 import distributed as dd
    assert dd.world() == (rank, size)
    # ranges tile [0, n) in rank order, balanced to within one item
    for n in (0, 1, 7, 10, 18000, 36601):
        ranges = [dd.shard_range(n, r, size) for r in range(size)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert max(hi - lo for lo, hi in ranges) - min(hi - lo for lo, hi in ranges) <= 1
    # ragged all-gather: rank r contributes r + 2 rows
    t = torch.full((rank + 2, 3), float(rank))
    g = dd.allgather_ragged(t, n_total=sum(r + 2 for r in range(size)))
    exp = torch.cat([torch.full((r + 2, 3), float(r)) for r in range(size)])
    assert torch.equal(g, exp)
    e = dd.allgather_ragged(torch.empty(0, dtype=torch.float64) if rank == 0 else torch.ones(2, dtype=torch.float64))
    assert e.tolist() == [1.0, 1.0] * (size - 1)
    s = dd.allreduce_sum_(torch.arange(4, dtype=torch.float64) * (rank + 1))
    assert torch.equal(s, torch.arange(4, dtype=torch.float64) * sum(r + 1 for r in range(size)))


def test_helpers_world2(tmp_path):
    _run("_check_helpers", tmp_path)


def test_helpers_are_identity_without_a_group():
    from chrysalis_b200 import distributed as dd
    assert dd.world() == (0, 1) and dd.shard_range(10) == (0, 10)
    t = torch.arange(3.0)
    assert dd.allgather_ragged(t) is t and dd.allreduce_sum_(t) is t


# ---- gene-sharded detect_svgs flow -------------------------------------------------------------
class _FakeDevice:
    """Stands in for chrysalis_b200.device on the CPU: same call surface as svg_stage uses, the
    arithmetic done by numpy / the oracle.  Only the orchestration (sharding, collectives) is under test."""

    def __init__(self):
        from oracle import moran as om, weights as ow
        self.om, self.ow = om, ow

    from chrysalis_b200.device import CompactCSR, DeviceCSR, PinnedCSR      # svg_stage only asks "is the input already resident / pinned?"

    def upload_csr(self, X):
        import scipy.sparse as sp
        return sp.csr_matrix(X)

    def host_arrays_pinned(self, X):
        return False

    def upload_csr_selected(self, X, col_map, row_lo=0, row_hi=None):
        """What chr_host_csr_select_* deliver: rows [row_lo, row_hi), the columns with col_map >= 0, renumbered."""
        import scipy.sparse as sp
        X = sp.csr_matrix(X)[row_lo:row_hi]
        cols = np.flatnonzero(col_map >= 0)
        assert np.array_equal(col_map[cols], np.arange(len(cols)))
        return X[:, cols]

    def lattice_plan(self, xy, nbr, bbox=None, **kw):
        return None

    def to_host(self, t):
        return t.cpu().numpy()

    def csr_gene_nnz(self, csr):
        return torch.from_numpy(np.asarray((csr > 0).sum(0)).ravel().astype(np.int32))

    def csr_spot_totals(self, csr, gene_col):
        keep = (gene_col >= 0).numpy()
        return torch.from_numpy(np.asarray(csr[:, keep].sum(1)).ravel().astype(np.float64))

    def median_scale(self, totals):
        from chrysalis_b200 import device as real
        return real.median_scale(totals)

    def to_device(self, a, dtype=None):
        return torch.from_numpy(np.ascontiguousarray(a))

    def upload_array(self, a, dtype=None):
        t = torch.from_numpy(np.array(a, copy=True))
        return t if dtype is None else t.to(dtype)

    def spatial_bbox(self, xy):
        return None

    def hilbert_order(self, xy, bbox):
        return torch.arange(xy.shape[0])

    spatial_order = hilbert_order

    def knn_build(self, xy, k, bbox):
        return torch.from_numpy(self.ow.knn_neighbors(xy.numpy(), k).astype(np.int32))

    def moran_plan(self, nbr):
        return None

    def dense_ld(self, n_cols):
        return n_cols

    def csr_densify(self, csr, gene_col, n_keep, scale, rank_of, apply_log1p, out=None):
        keep = (gene_col >= 0).numpy()
        d = np.asarray(csr[:, keep].todense(), dtype=np.float64)
        if scale is not None:
            d = d * scale.numpy()[:, None]
        d = d.astype(np.float32)
        return torch.from_numpy(np.log1p(d) if apply_log1p else d)

    def moran_dense(self, dense, nbr, n_keep, geary=False, timing=False, plan=None):
        I = torch.from_numpy(self.om.morans_I_batched(dense.numpy(), nbr.numpy()))
        C = torch.from_numpy(self.om.gearys_C_batched(dense.numpy(), nbr.numpy())) if geary else None
        return I, C, None


    # ---- PCA surface (numpy stand-ins) ----
    def column_sums(self, P, s):
        return P[:, :s].to(torch.float64).sum(0)

    def gram_centered(self, P, s, mean, timing=False):
        Z = P[:, :s].to(torch.float64) - mean
        return Z.T @ Z

    def eigh_topk(self, C, r):
        w, v = np.linalg.eigh(C.numpy())
        return torch.from_numpy(w[::-1][:r].copy()), torch.from_numpy(v[:, ::-1][:, :r].T.copy()), 0.0

    def pca_scores(self, P, s, mean, V):
        return ((P[:, :s].to(torch.float64) - mean) @ V.T).to(torch.float32)


def _check_sharded_pca(rank, size, out_dir):
    import scipy.sparse as sp
    import chrysalis_b200 as ch
    from chrysalis_b200 import core
    from oracle import pca as opca
    z = np.load(os.path.join(GOLDEN, "pca_aa_hex.npz"), allow_pickle=False)
    pcs = z["pcs"]
    n, s = pcs.shape
    fake = _FakeDevice()
    fake.csr_densify = lambda csr, gene_col, n_keep, scale, rank_of, apply_log1p, out=None: torch.from_numpy(
        np.asarray(csr[:, (gene_col >= 0).numpy()].todense(), dtype=np.float32))
    core.dv = fake
    ad = ch.AnnDataLite(sp.csr_matrix(pcs), obsm={"spatial": np.zeros((n, 2))})
    ad.var["spatially_variable"] = np.ones(s, dtype=bool)
    ch.pca(ad, n_pcs=12)
    L = ad.uns["chr_pca"]["loadings"]
    assert ad.obsm["chr_X_pca"].shape == (n, 12)                    # every rank holds all spots
    assert opca.subspace_cosines(z["loadings"], L).min() >= 0.9999
    assert np.allclose(ad.uns["chr_pca"]["variance_ratio"], z["variance_ratio"], rtol=2e-4)
    sc = ad.obsm["chr_X_pca"][:, :6]
    assert np.abs(sc - z["X_pca"][:, :6]).max() < 2e-3 * np.abs(z["X_pca"][:, :6]).max()


def test_spot_sharded_pca_world2(tmp_path):
    _run("_check_sharded_pca", tmp_path)


def _check_sharded_svgs(rank, size, out_dir):
    from conftest import golden_csr
    import chrysalis_b200 as ch
    from chrysalis_b200 import core
    z = np.load(os.path.join(GOLDEN, "svg_hex_k6.npz"), allow_pickle=False)
    core.dv = _FakeDevice()
    ad = ch.AnnDataLite(golden_csr(z), obsm={"spatial": z["xy"]})
    ch.detect_svgs(ad, neighbors=int(z["k"]), top_svg=int(z["top_svg"]), min_morans=float(z["min_morans"]), geary=True)
    m = z["gene_mask"]
    got = ad.var["Moran's I"].to_numpy()
    assert np.array_equal(np.isnan(got), ~m)
    # the all-reduced totals make the sharded normalisation identical to the single-process one
    assert np.allclose(got[m], z["morans"][m], rtol=1e-5, atol=2e-7)
    assert np.allclose(ad.var["Geary's C"].to_numpy()[m], z["geary"][m], rtol=1e-5, atol=2e-7)
    assert np.array_equal(ad.var["spatially_variable"].to_numpy(), z["spatially_variable"])
    np.save(os.path.join(out_dir, f"I_rank{rank}.npy"), got)


def test_gene_sharded_detect_svgs_world2(tmp_path):
    _run("_check_sharded_svgs", tmp_path)
    a, b = (np.load(os.path.join(tmp_path, f"I_rank{r}.npy")) for r in range(2))
    assert np.array_equal(a, b, equal_nan=True)                     # every rank holds the same full result


def _check_gene_shard_names(rank, size, out_dir):
    """A gene-sharded AnnData per rank (.uns['chr_gene_shard']): unique names take the hash exchange and the array path, a
    name that occurs on both ranks falls back to the names and to the reference's pandas statements."""
    from conftest import golden_csr
    import chrysalis_b200 as ch
    from chrysalis_b200 import core
    from chrysalis_b200 import distributed as dd
    z = np.load(os.path.join(GOLDEN, "svg_hex_k6.npz"), allow_pickle=False)
    core.dv = _FakeDevice()
    X = golden_csr(z)
    g = X.shape[1]
    lo, hi = dd.shard_range(g, rank, size)
    for case in ("unique", "duplicated"):
        names = [f"g{j}" for j in range(g)]
        if case == "duplicated":
            names[g - 1] = names[0]                                   # the same name on the first and on the last rank
        ad = ch.AnnDataLite(X[:, lo:hi].tocsr(), obsm={"spatial": z["xy"]}, var_names=names[lo:hi], uns={"chr_gene_shard": True})
        ch.detect_svgs(ad, neighbors=int(z["k"]), top_svg=int(z["top_svg"]), min_morans=float(z["min_morans"]))
        np.save(os.path.join(out_dir, f"{case}_I_rank{rank}.npy"), ad.var["Moran's I"].to_numpy())
        np.save(os.path.join(out_dir, f"{case}_sv_rank{rank}.npy"), ad.var["spatially_variable"].to_numpy())


def test_gene_sharded_anndata_names_world2(tmp_path):
    _run("_check_gene_shard_names", tmp_path)
    z = np.load(os.path.join(GOLDEN, "svg_hex_k6.npz"), allow_pickle=False)
    m = z["gene_mask"]
    I = np.concatenate([np.load(os.path.join(tmp_path, f"unique_I_rank{r}.npy")) for r in range(2)])
    sv = np.concatenate([np.load(os.path.join(tmp_path, f"unique_sv_rank{r}.npy")) for r in range(2)])
    assert np.array_equal(np.isnan(I), ~m) and np.allclose(I[m], z["morans"][m], rtol=1e-5, atol=2e-7)
    assert np.array_equal(sv, z["spatially_variable"])
    # duplicated name: the statistics are the same; the fields follow the reference's index alignment (var_names_make_unique
    # renames the second occurrence, so the row labelled with the duplicate name reads the first occurrence's value)
    Id = np.concatenate([np.load(os.path.join(tmp_path, f"duplicated_I_rank{r}.npy")) for r in range(2)])
    same = np.ones(len(I), dtype=bool)
    same[[0, len(I) - 1]] = False
    assert np.allclose(Id[same], I[same], rtol=1e-12, atol=0, equal_nan=True)
