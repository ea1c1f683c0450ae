"""Two ranks driving the real CUDA kernels (-m gpu): gene-sharded ``detect_svgs`` (both flavours: every rank holds the
whole AnnData / every rank holds its gene shard) and spot-sharded ``pca`` against the single-process results.

With two or more GPUs the ranks take one device each over NCCL; on a one-GPU box both ranks share cuda:0 and the
collectives go through gloo (host-side exchanges only - no kernel of one rank ever waits for the other)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, GOLDEN

pytestmark = pytest.mark.gpu


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
    multi = torch.cuda.device_count() >= size
    torch.cuda.set_device(rank if multi else 0)
    if multi:
        dist.init_process_group("nccl", rank=rank, world_size=size, device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        globals()[fn_name](rank, size, out_dir)
    finally:
        dist.destroy_process_group()


def _run(fn_name, tmp_path, size=2):
    mp.spawn(_worker, args=(size, _free_port(), fn_name, str(tmp_path)), nprocs=size, join=True)


def _square_case():
    from chrysalis_b200 import synth
    xy = synth.make_coords("square", 60 * 60, seed=5)
    X, _ = synth.make_counts_csr(xy, 333, n_zones=5, seed=5)
    return xy, X


def _check_gene_sharded_svgs(rank, size, out_dir):
    import chrysalis_b200 as ch
    from chrysalis_b200 import distributed as dd
    from conftest import golden_csr
    # (a) hex fixture, every rank holds the whole AnnData (columns selected while uploading), quad-plan kernel
    z = np.load(os.path.join(GOLDEN, "svg_hex_k6.npz"), allow_pickle=False)
    ad = ch.AnnDataLite(golden_csr(z), obsm={"spatial": z["xy"]})
    ch.detect_svgs(ad, neighbors=int(z["k"]), top_svg=int(z["top_svg"]), min_morans=float(z["min_morans"]), geary=True)
    m = z["gene_mask"]
    got = ad.var["Moran's I"].to_numpy()
    assert np.array_equal(np.isnan(got), ~m)
    assert np.allclose(got[m], z["morans"][m], rtol=1e-5, atol=2e-7)
    assert np.allclose(ad.var["Geary's C"].to_numpy()[m], z["geary"][m], rtol=1e-5, atol=2e-7)
    assert np.array_equal(ad.var["spatially_variable"].to_numpy(), z["spatially_variable"])
    np.save(os.path.join(out_dir, f"hex_rank{rank}.npy"), got)
    # (b) square lattice (stencil kernel), every rank holds ITS gene shard (.uns['chr_gene_shard'])
    xy, X = _square_case()
    lo, hi = dd.shard_range(X.shape[1], rank, size)
    names = [f"g{j}" for j in range(X.shape[1])]
    ad = ch.AnnDataLite(X[:, lo:hi].tocsr(), obsm={"spatial": xy}, var_names=names[lo:hi], uns={"chr_gene_shard": True})
    ch.detect_svgs(ad, neighbors=8, top_svg=40, min_morans=0.05)
    np.save(os.path.join(out_dir, f"sq_I_rank{rank}.npy"), ad.var["Moran's I"].to_numpy())
    np.save(os.path.join(out_dir, f"sq_sv_rank{rank}.npy"), ad.var["spatially_variable"].to_numpy())


def test_gene_sharded_detect_svgs_two_ranks(tmp_path):
    import chrysalis_b200 as ch
    from chrysalis_b200 import distributed as dd
    _run("_check_gene_sharded_svgs", tmp_path)
    a, b = (np.load(os.path.join(tmp_path, f"hex_rank{r}.npy")) for r in range(2))
    assert np.array_equal(a, b, equal_nan=True)                                 # every rank holds the same full result
    # the gene-sharded AnnDatas together reproduce the single-process call bit for bit (the summation tree of a gene
    # does not depend on the sharding) and make the same global top_svg / min_morans choice
    xy, X = _square_case()
    ad = ch.AnnDataLite(X, obsm={"spatial": xy}, var_names=[f"g{j}" for j in range(X.shape[1])])
    ch.detect_svgs(ad, neighbors=8, top_svg=40, min_morans=0.05)
    I1, sv1 = ad.var["Moran's I"].to_numpy(), ad.var["spatially_variable"].to_numpy()
    I2 = np.concatenate([np.load(os.path.join(tmp_path, f"sq_I_rank{r}.npy")) for r in range(2)])
    sv2 = np.concatenate([np.load(os.path.join(tmp_path, f"sq_sv_rank{r}.npy")) for r in range(2)])
    assert np.array_equal(np.isnan(I1), np.isnan(I2))
    assert np.allclose(I1, I2, rtol=1e-6, atol=1e-9, equal_nan=True)            # normalize_total: all-reduced totals in another order
    assert np.array_equal(sv1, sv2)


def _check_spot_sharded_pca(rank, size, out_dir):
    import scipy.sparse as sp
    import chrysalis_b200 as ch
    z = np.load(os.path.join(GOLDEN, "pca_aa_hex.npz"), allow_pickle=False)
    pcs = z["pcs"]
    n, s = pcs.shape
    ad = ch.AnnDataLite(sp.csr_matrix(pcs), obsm={"spatial": np.zeros((n, 2))})
    ad.var["spatially_variable"] = np.ones(s, dtype=bool)
    ch.pca(ad, n_pcs=12)
    assert ad.obsm["chr_X_pca"].shape == (n, 12)                                # every rank holds all spots
    np.save(os.path.join(out_dir, f"load_rank{rank}.npy"), ad.uns["chr_pca"]["loadings"])
    np.save(os.path.join(out_dir, f"scores_rank{rank}.npy"), ad.obsm["chr_X_pca"])


def test_spot_sharded_pca_two_ranks(tmp_path):
    from oracle import pca as opca
    _run("_check_spot_sharded_pca", tmp_path)
    z = np.load(os.path.join(GOLDEN, "pca_aa_hex.npz"), allow_pickle=False)
    L = [np.load(os.path.join(tmp_path, f"load_rank{r}.npy")) for r in range(2)]
    Sc = [np.load(os.path.join(tmp_path, f"scores_rank{r}.npy")) for r in range(2)]
    assert np.array_equal(L[0], L[1]) and np.array_equal(Sc[0], Sc[1])
    assert opca.subspace_cosines(z["loadings"], L[0]).min() >= 0.9999
    assert opca.component_cosines(z["loadings"][:6], L[0][:6]).min() >= 0.9999
    assert np.abs(Sc[0][:, :6] - z["X_pca"][:, :6]).max() < 2e-3 * np.abs(z["X_pca"][:, :6]).max()
