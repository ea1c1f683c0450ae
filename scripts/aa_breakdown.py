"""Per-kernel time of the AA stage at a cfg-shaped size (torch.profiler / CUPTI sees the ctypes-launched kernels).
usage: python scripts/aa_breakdown.py [N d A iters]"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chrysalis_b200 import core, device as dv

n, d, A, iters = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (700000, 30, 12, 40))]
torch.cuda.set_device(0)
rng = np.random.default_rng(0)
# PCA-score-like data: mixtures of A latent archetypes + decaying-variance noise
Zt = rng.normal(size=(A, d)) * (3.0 / np.sqrt(1 + np.arange(d)))
al = rng.dirichlet(np.full(A, 0.3), size=n)
X = al @ Zt + rng.normal(size=(n, d)) * (0.5 / np.sqrt(1 + np.arange(d)))
Xd = dv.to_device(np.ascontiguousarray(X), torch.float64)
rows = dv.aa_furthest_sum(Xd, A, 123)
st = dv.AAState(Xd, A); st.init_from_rows(rows)
for _ in range(3): st.iterate()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
t0 = time.perf_counter()
passes = []
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(iters):
        st.iterate(); passes.append(st.beta_passes)
    torch.cuda.synchronize()
wall = time.perf_counter() - t0
rowsk = {}
for e in prof.key_averages():
    rowsk[e.key] = {"calls": e.count, "total_ms": round(e.device_time_total / 1e3, 3) if hasattr(e, "device_time_total") else round(e.cuda_time_total / 1e3, 3)}
tot = sum(v["total_ms"] for v in rowsk.values())
print(json.dumps({"n": n, "d": d, "A": A, "iters": iters, "wall_ms_per_iter_profiled": round(wall * 1e3 / iters, 3),
                  "kernel_ms_per_iter": round(tot / iters, 3), "beta_passes": passes,
                  "kernels": dict(sorted(rowsk.items(), key=lambda kv: -kv[1]["total_ms"]))}))
t0 = time.perf_counter()
for _ in range(iters): st.iterate()
torch.cuda.synchronize()
print(json.dumps({"wall_ms_per_iter_unprofiled": round((time.perf_counter() - t0) * 1e3 / iters, 3)}))
