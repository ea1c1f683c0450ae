"""Where the end-to-end SVG step (core.svg_stage from pinned host CSR) spends its time at cfg4.
usage: python scripts/e2e_breakdown.py [N G K]   -> one JSON line (ms per stage, each bracketed by a device sync)"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from chrysalis_b200 import core, synth, device as dv

n, g, k = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (700000, 18000, 8))]
MIN_SPOTS = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-4     # bench.py's filter (keeps every synthetic gene)
class Args: pass
a = Args(); a.n, a.g, a.k = n, g, k
torch.cuda.set_device(0)
xy_s, nbr, csr = bench.build_device_workload(a, 0, torch, dv, synth)
host = dv.PinnedCSR(csr.indptr.cpu().pin_memory(), csr.indices.cpu().pin_memory(), csr.data.cpu().pin_memory(), csr.shape)
xy_host = xy_s.cpu().numpy(); xy_host[:, 1] *= -1
del csr, nbr
torch.cuda.empty_cache()
def sync(): torch.cuda.synchronize(); return time.perf_counter()
out = {"n": n, "g": g, "k": k, "h2d_bytes": int(host.nbytes)}
for rep in range(3):
    t = {}
    t0 = sync()
    c = dv.upload_csr(host)
    t1 = sync(); t["h2d"] = t1 - t0
    pts = np.array(xy_host, dtype=np.float64, copy=True); pts[:, 1] *= -1
    xy = dv.to_device(pts, torch.float64)
    bbox = dv.spatial_bbox(xy); order = dv.spatial_order(xy, bbox)
    rank_of = torch.empty_like(order); rank_of[order] = torch.arange(n, device=order.device)
    nb = dv.knn_build(xy[order].contiguous(), k, bbox)
    plan = dv.moran_plan(nb)
    t2 = sync(); t["weights_plan"] = t2 - t1
    dense_buf = torch.empty((n, dv.dense_ld(g)), dtype=torch.float32, device="cuda"); dense_buf.zero_()
    t3 = sync(); t["zero_fill"] = t3 - t2
    gene_nnz = dv.csr_gene_nnz(c)
    t4 = sync(); t["gene_nnz"] = t4 - t3
    keep = gene_nnz >= int(n * MIN_SPOTS)
    gene_col = torch.where(keep, torch.cumsum(keep.to(torch.int32), 0, dtype=torch.int32) - 1, torch.full_like(gene_nnz, -1))
    n_keep = int(keep.sum().item())
    t5 = sync(); t["keep"] = t5 - t4
    scale = dv.median_scale(dv.csr_spot_totals(c, gene_col))
    t6 = sync(); t["totals_median"] = t6 - t5
    dense = dv.csr_densify(c, gene_col, n_keep, scale, rank_of, apply_log1p=True, out=dense_buf)
    t7 = sync(); t["densify"] = t7 - t6
    I, _, _ = dv.moran_dense(dense, nb, n_keep, plan=plan)
    t8 = sync(); t["moran"] = t8 - t7
    I_h = I.cpu().numpy(); kh = keep.cpu().numpy()
    t9 = sync(); t["d2h"] = t9 - t8
    t["sum"] = t9 - t0
    del c, dense, dense_buf
    t0 = sync()
    core.svg_stage(host, xy_host, MIN_SPOTS, k, False, already_log1p=False, shard=False)
    t["svg_stage_call"] = sync() - t0
    out[f"rep{rep}_ms"] = {kk: round(v * 1e3, 2) for kk, v in t.items()}
print(json.dumps(out))
