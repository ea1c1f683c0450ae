"""Times chr_moran_perm_lattice_f32 on a cfg3-shaped lattice (120k bins, k = 8): ms per draw and the equivalent matrix bytes/s.
usage: python scripts/perm_bench.py [genes] [permutations] [n_spots]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chrysalis_b200 import device as dv
from chrysalis_b200 import synth

g = int(sys.argv[1]) if len(sys.argv) > 1 else 12544
P = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n = int(sys.argv[3]) if len(sys.argv) > 3 else 120_000
xy = synth.make_coords("square", n, seed=3)
xy_d = dv.to_device(xy, torch.float64)
nbr = dv.knn_build(xy_d, 8)
plan = dv.lattice_plan(xy_d, nbr)
X = torch.empty((plan.n_rows, dv.dense_ld(g)), dtype=torch.float32, device="cuda")
X.normal_(1.0, 0.5)
X[:, :64] += torch.sin(torch.arange(plan.n_rows, device="cuda") / 5000.0)[:, None]
res = {}
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = dv.moran_perm_lattice(X, plan, g, permutations=P, seed=1, timing=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ms = dv.last_kernel_ms()
    res = {"n": n, "genes": g, "permutations": P, "chunks": int(plan.chunks.shape[0]), "wall_s": round(dt, 4), "draw_loop_ms": round(ms, 3),
           "ms_per_draw": round(ms / P, 4), "matrix_GB": round(4.0 * n * g / 1e9, 3), "equivalent_GBps": round(4.0 * n * g * P / (ms * 1e-3) / 1e9, 1)}
# the general kernel for comparison: one pass per draw through a row map
order = dv.spatial_order(xy_d)
rank_of = torch.empty_like(order); rank_of[order] = torch.arange(n, device="cuda")
nbr_rows = rank_of.to(torch.int32)[nbr[order].long()].contiguous()
qplan = dv.moran_plan(nbr_rows)
Xg = X[:n].contiguous()
rm = dv.perm_row_map(n, 1, 0, "cuda")
for _ in range(3):
    dv.moran_dense(Xg, nbr_rows, g, plan=qplan, row_map=rm, timing=True)
res["general_kernel_ms_per_draw"] = round(dv.last_kernel_ms(), 4)
print(json.dumps(res))
