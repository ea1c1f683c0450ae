"""Multi-GPU parity under NCCL (torchrun, one process per GPU): gene-sharded detect_svgs and
spot-sharded pca on the golden fixtures must reproduce the single-process oracle values, and the
gene-sharded Moran's I must be BIT-identical to the single-GPU run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import chrysalis_b200 as ch  # noqa: E402
from chrysalis_b200 import core, distributed as dd  # noqa: E402
from oracle import pca as opca  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, size = dd.world()
    z = np.load(os.path.join(ROOT, "tests", "golden", "svg_hex_k6.npz"), allow_pickle=False)
    X = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))
    ad = ch.AnnDataLite(X, obsm={"spatial": z["xy"]})
    ch.detect_svgs(ad, neighbors=int(z["k"]), top_svg=int(z["top_svg"]), min_morans=float(z["min_morans"]), geary=True)
    m = z["gene_mask"]
    got = ad.var["Moran's I"].to_numpy()
    assert np.array_equal(np.isnan(got), ~m)
    assert np.all(np.abs(got[m] - z["morans"][m]) <= 1e-5 * np.abs(z["morans"][m]) + 2e-7)
    assert np.array_equal(ad.var["spatially_variable"].to_numpy(), z["spatially_variable"])
    # bit identity with the single-GPU result: rank 0 recomputes without the group's sharding
    full = torch.from_numpy(got.copy()).cuda()
    ref = full.clone()
    if rank == 0:
        saved = dd.world
        dd.world = lambda: (0, 1)
        try:
            ad1 = ch.AnnDataLite(X, obsm={"spatial": z["xy"]})
            ch.detect_svgs(ad1, neighbors=int(z["k"]), top_svg=int(z["top_svg"]), min_morans=float(z["min_morans"]), geary=True)
            ref = torch.from_numpy(ad1.var["Moran's I"].to_numpy().copy()).cuda()
        finally:
            dd.world = saved
    dist.broadcast(ref, 0)
    same = torch.equal(torch.nan_to_num(full, nan=-9.0), torch.nan_to_num(ref, nan=-9.0))
    assert same, "gene-sharded Moran's I differs from the single-GPU bits"

    zp = np.load(os.path.join(ROOT, "tests", "golden", "pca_aa_hex.npz"), allow_pickle=False)
    pcs = zp["pcs"]
    n, s = pcs.shape
    adp = ch.AnnDataLite(sp.csr_matrix(pcs), obsm={"spatial": np.zeros((n, 2))})
    adp.var["spatially_variable"] = np.ones(s, dtype=bool)
    ch.pca(adp, n_pcs=12)
    L = adp.uns["chr_pca"]["loadings"]
    cos = opca.subspace_cosines(zp["loadings"], L).min()
    assert cos >= 0.9999, cos
    assert adp.obsm["chr_X_pca"].shape == (n, 12)
    if rank == 0:
        print(f"dist_check ok: world {size}, Moran bit-identical to 1 GPU, PCA subspace cosine {cos:.7f}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
