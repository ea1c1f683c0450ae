"""cfg3-shaped permutation inference timing: N bins x G genes, k = 8, P draws (one Moran pass per draw).
usage: python scripts/perm_check.py [N G P]   -> one JSON line (seconds per draw, extrapolation to 999 draws)"""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from chrysalis_b200 import synth, device as dv
n, g, P = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (120000, 25000, 25)))
class A: pass
a = A(); a.n, a.g, a.k = n, g, 8
torch.cuda.set_device(0)
xy_s, nbr, csr = bench.build_device_workload(a, 0, torch, dv, synth)
X, g_scored = bench.dense_from_csr(csr, 1e-4, torch, dv)
plan = dv.moran_plan(nbr)
dv.moran_permutation_test(X, nbr, g_scored, permutations=2, seed=0, plan=plan)
torch.cuda.synchronize(); t0 = time.perf_counter()
res = dv.moran_permutation_test(X, nbr, g_scored, permutations=P, seed=1, plan=plan)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
per = dt / (P + 1)
print(json.dumps({"workload": f"cfg3_stereoseq-shaped: {n} bins x {g_scored} genes, k=8, {P} draws timed", "seconds": round(dt, 3),
                  "s_per_draw_all_genes": round(per, 5), "gene_draws_per_s": round(g_scored / per, 0),
                  "extrapolated_999_draws_s": round(per * 1000, 2),
                  "achieved_GBs_per_draw": round(4.0 * n * g_scored / per / 1e9, 1),
                  "min_p_sim": float(res["p_sim"].min()), "max_abs_z_sim": float(res["z_sim"].abs().max())}))
