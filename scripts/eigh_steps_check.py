"""Lanczos depth needed by the PCA eigensolver on the bench's own covariance (cfg4: 2000 SVGs, 30 PCs).
usage: python scripts/eigh_steps_check.py [m ...]"""
import ctypes, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from chrysalis_b200 import _lib, device as dv, synth
from chrysalis_b200.device import _ptr, _stream, _workspace

n, g, S, r = 700000, 18000, 2000, 30
ms = [int(x) for x in sys.argv[1:]] or [62, 92, 122, 160, 244]
torch.cuda.set_device(0)
class Args: pass
a = Args(); a.n, a.g, a.k = n, g, 8
xy_s, nbr, csr = bench.build_device_workload(a, 0, torch, dv, synth)
Xm, g_scored = bench.dense_from_csr(csr, 1e-4, torch, dv)
I, _, _ = dv.moran_dense(Xm, nbr, g_scored, plan=dv.moran_plan(nbr))
del Xm
keep = dv.csr_gene_nnz(csr) >= int(n * 1e-4)
kept_idx = torch.nonzero(keep).squeeze(1)[:g_scored]
top = torch.topk(torch.nan_to_num(I, nan=-2.0), S).indices
sel = torch.zeros(g, dtype=torch.bool, device="cuda"); sel[kept_idx[top]] = True
scale = dv.median_scale(dv.csr_spot_totals(csr, torch.arange(g, dtype=torch.int32, device="cuda")))
gene_col = torch.where(sel, torch.cumsum(sel.int(), 0, dtype=torch.int32) - 1, torch.full((g,), -1, dtype=torch.int32, device="cuda"))
P = dv.csr_densify(csr, gene_col, S, scale, None, apply_log1p=True)
mean = dv.column_sums(P, S) / n
C = dv.gram_centered(P, S, mean) / (n - 1)
ref = torch.linalg.eigvalsh(C).flip(0)[:r]
for m in ms:
    ws_bytes = _lib.query("chr_eigh_workspace_bytes", S, r, m)
    ws = _workspace(ws_bytes, C.device)
    evals = torch.empty(r, dtype=torch.float64, device="cuda"); evecs = torch.empty((r, S), dtype=torch.float64, device="cuda")
    res = torch.empty(r, dtype=torch.float64, device="cuda")
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _lib.call("chr_eigh_topk_f64", _ptr(C), S, r, m, _ptr(evals), _ptr(evecs), _ptr(res), _ptr(ws), ws_bytes, _stream())
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ev = torch.sort(evals, descending=True).values
    print(json.dumps({"m": m, "ms": round(dt * 1e3, 2), "worst_residual_rel": float((res / evals.abs().max()).max()),
                      "max_eval_rel_err": float(((ev - ref).abs() / ref[0]).max()), "gap_29_30_rel": float((ref[28] - ref[29]) / ref[0])}))
