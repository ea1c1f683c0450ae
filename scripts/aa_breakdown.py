"""Per-kernel time of the AA stage at a cfg-shaped size (torch.profiler / CUPTI sees the ctypes-launched kernels).
usage: python scripts/aa_breakdown.py [N d A iters]"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chrysalis_b200 import core, device as dv

real = "--real" in sys.argv
argv = [a for a in sys.argv[1:] if a != "--real"]
n, d, A, iters = [int(x) for x in (argv[:4] if len(argv) > 3 else (700000, 30, 12, 40))]
torch.cuda.set_device(0)
if real:
    # the bench's own embedding: cfg4-shaped counts -> Moran -> top 2000 SVGs -> PCA scores (bench.pipeline_stages)
    import bench
    from chrysalis_b200 import synth
    class Args: pass
    a = Args(); a.n, a.g, a.k = n, 18000, 8
    xy_s, nbr, csr = bench.build_device_workload(a, 0, torch, dv, synth)
    Xm, g_scored = bench.dense_from_csr(csr, 1e-4, torch, dv)
    I, _, _ = dv.moran_dense(Xm, nbr, g_scored, plan=dv.moran_plan(nbr))
    del Xm
    g = a.g
    keep = dv.csr_gene_nnz(csr) >= int(n * 1e-4)
    kept_idx = torch.nonzero(keep).squeeze(1)[:g_scored]
    S = 2000
    top = torch.topk(torch.nan_to_num(I, nan=-2.0), S).indices
    sel = torch.zeros(g, dtype=torch.bool, device="cuda"); sel[kept_idx[top]] = True
    scale = dv.median_scale(dv.csr_spot_totals(csr, torch.arange(g, dtype=torch.int32, device="cuda")))
    gene_col = torch.where(sel, torch.cumsum(sel.int(), 0, dtype=torch.int32) - 1, torch.full((g,), -1, dtype=torch.int32, device="cuda"))
    P = dv.csr_densify(csr, gene_col, S, scale, None, apply_log1p=True)
    mean = dv.column_sums(P, S) / n
    C = dv.gram_centered(P, S, mean) / (n - 1)
    evals, V, worst = dv.eigh_topk(C, d)
    X = dv.pca_scores(P, S, mean, V)[:, :d].cpu().numpy().astype(np.float64)
    del P, csr, C
    torch.cuda.empty_cache()
else:
    rng = np.random.default_rng(0)
    # PCA-score-like data: mixtures of A latent archetypes + decaying-variance noise
    Zt = rng.normal(size=(A, d)) * (3.0 / np.sqrt(1 + np.arange(d)))
    al = rng.dirichlet(np.full(A, 0.3), size=n)
    X = al @ Zt + rng.normal(size=(n, d)) * (0.5 / np.sqrt(1 + np.arange(d)))
Xd = dv.to_device(np.ascontiguousarray(X), torch.float64)
rows = dv.aa_furthest_sum(Xd, A, 123)
st = dv.AAState(Xd, A); st.init_from_rows(rows)
import ctypes
from chrysalis_b200 import _lib
lib = _lib.load()
def stats(reset=True):
    if not hasattr(lib, "chr_debug_aa_stats"): return None
    buf = (ctypes.c_ulonglong * 4)()
    lib.chr_debug_aa_stats(buf, int(reset))
    return list(buf)
stats()
hist = []
for _ in range(int(os.environ.get("AA_PRE_ITERS", "3"))):
    st.iterate(); hist.append(stats())
torch.cuda.synchronize()
print(json.dumps({"stats_pre_iters [spots, general-path spots, warps hit, warps]": hist}))
sup = (st.alphas > 0).sum(1)
print(json.dumps({"support_size_hist": torch.bincount(sup, minlength=A + 1).tolist()}))
from torch.profiler import profile, ProfilerActivity
t0 = time.perf_counter()
passes = []
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(iters):
        st.iterate(); passes.append(st.beta_passes); hist.append(stats())
    torch.cuda.synchronize()
wall = time.perf_counter() - t0
rowsk = {}
for e in prof.key_averages():
    rowsk[e.key] = {"calls": e.count, "total_ms": round(e.device_time_total / 1e3, 3) if hasattr(e, "device_time_total") else round(e.cuda_time_total / 1e3, 3)}
tot = sum(v["total_ms"] for v in rowsk.values())
print(json.dumps({"n": n, "d": d, "A": A, "iters": iters, "wall_ms_per_iter_profiled": round(wall * 1e3 / iters, 3),
                  "kernel_ms_per_iter": round(tot / iters, 3), "beta_passes": passes, "stats_last": hist[-5:],
                  "kernels": dict(sorted(rowsk.items(), key=lambda kv: -kv[1]["total_ms"]))}))
t0 = time.perf_counter()
for _ in range(iters): st.iterate()
torch.cuda.synchronize()
print(json.dumps({"wall_ms_per_iter_unprofiled": round((time.perf_counter() - t0) * 1e3 / iters, 3)}))
