"""Stage timings of detect_svgs -> pca -> aa on a device-resident synthetic workload (cfg-shaped).
usage: python scripts/pipeline_check.py N G K S R A [max_iter]"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from chrysalis_b200 import _lib, core, synth, device as dv

n, g, k, S, r, A = [int(x) for x in sys.argv[1:7]]
max_iter = int(sys.argv[7]) if len(sys.argv) > 7 else 200
class Args: pass
a = Args(); a.n, a.g, a.k = n, g, k
torch.cuda.set_device(0)
xy_s, nbr, csr = bench.build_device_workload(a, 0, torch, dv, synth)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
out = {"n": n, "g": g, "k": k, "S": S, "r": r, "A": A}
t0 = sync()
X, g_scored = bench.dense_from_csr(csr, 0.05, torch, dv)
plan = dv.moran_plan(nbr)
t1 = sync()
I, _, _ = dv.moran_dense(X, nbr, g_scored, plan=plan)
t2 = sync()
out["svg"] = {"prep_s": t1 - t0, "moran_s": t2 - t1, "genes_scored": g_scored}
del X
# PCA on the top-S genes: normalise over all genes + log1p (README flow), densify the SVG columns
gene_nnz = dv.csr_gene_nnz(csr)
keep = gene_nnz >= int(n * 0.05)
top = torch.topk(I, min(S, g_scored)).indices
kept_idx = torch.nonzero(keep).squeeze(1)
sel = torch.zeros(g, dtype=torch.bool, device="cuda"); sel[kept_idx[top]] = True
s = int(sel.sum())
all_col = torch.arange(g, dtype=torch.int32, device="cuda")
scale = dv.median_scale(dv.csr_spot_totals(csr, all_col))
gene_col = torch.where(sel, torch.cumsum(sel.int(), 0, dtype=torch.int32) - 1, torch.full((g,), -1, dtype=torch.int32, device="cuda"))
t0 = sync()
P = dv.csr_densify(csr, gene_col, s, scale, None, apply_log1p=True)
t1 = sync()
mean = dv.column_sums(P, s) / n
t2 = sync()
C = dv.gram_centered(P, s, mean, timing=True) / (n - 1)
t3 = sync()
gram_ms = dv.last_kernel_ms(_lib.FAMILY_GRAM)
evals, V, worst = dv.eigh_topk(C, r)
t4 = sync()
scores = dv.pca_scores(P, s, mean, V)
t5 = sync()
flops = 2.0 * n * s * s
out["pca"] = {"densify_s": t1 - t0, "colsum_s": t2 - t1, "gram_s": t3 - t2, "gram_kernel_ms": gram_ms,
              "gram_tflops_useful_full": flops / (gram_ms * 1e-3) / 1e12, "eigh_s": t4 - t3, "eigh_residual": worst, "scores_s": t5 - t4,
              "total_s": t5 - t0}
emb = scores[:, :r].cpu().numpy()
t0 = sync()
fit = core.aa_stage(emb, A, max_iter=max_iter)
t1 = sync()
out["aa"] = {"total_s": t1 - t0, "n_iter_best": fit["n_iter"], "rss": fit["rss"]}
print(json.dumps(out))
