"""Experiment: rounds per quad and tile-kernel time as a function of the owner-balancing sweeps (CHR_PLAN_SWEEPS)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from chrysalis_b200 import device as dv, synth
n, g, k = 700_000, 2304, 8
xy = torch.from_numpy(synth.make_coords("square", n, seed=0)).cuda()
bbox = dv.spatial_bbox(xy)
order = dv.spatial_order(xy, bbox)   # lattice-snapped on this square lattice
nbr = dv.knn_build(xy[order].contiguous(), k, bbox)
X = torch.rand((n, dv.dense_ld(g)), device="cuda")
for sweeps in sys.argv[1:]:
    if ":" in sweeps:
        sweeps, comb = sweeps.split(":")
        os.environ["CHR_PLAN_COMBINED"] = comb
    else:
        os.environ.pop("CHR_PLAN_COMBINED", None)
    os.environ["CHR_PLAN_SWEEPS"] = sweeps
    plan = dv.moran_plan(nbr)
    for _ in range(3):
        dv.moran_dense(X, nbr, g, plan=plan)
    ms = []
    for _ in range(8):
        dv.moran_dense(X, nbr, g, plan=plan, timing=True)
        torch.cuda.synchronize()
        ms.append(dv.last_kernel_ms())
    print(json.dumps({"sweeps": int(sweeps), "combined": os.environ.get("CHR_PLAN_COMBINED"), "total_units": plan.total_units, "rounds_per_quad": (plan.total_units / ((n + 3) // 4) - 4) / 2,
                      "kernel_ms_2304": round(float(np.median(ms)), 4)}))
