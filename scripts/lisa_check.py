"""Local Moran's I at the cfg4 spot count: time per variable for P = 999 conditional permutations."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from chrysalis_b200 import device as dv, synth
n, k, P = 700_000, 8, 999
xy = torch.from_numpy(synth.spot_coordinates(n, seed=0)).cuda() if hasattr(synth, "spot_coordinates") else torch.rand(n, 2, dtype=torch.float64, device="cuda")
nbr = dv.knn_build(xy.contiguous(), k, dv.spatial_bbox(xy))
y = (torch.sin(6 * xy[:, 0]) + 0.5 * torch.randn(n, device="cuda", dtype=torch.float64)).float()
for _ in range(2):
    r = dv.local_moran(y, nbr, permutations=P, seed=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); r = dv.local_moran(y, nbr, permutations=P, seed=1); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
sig = float((r["p_sim"] <= 0.05).float().mean())
print(json.dumps({"n": n, "k": k, "permutations": P, "ms_per_variable": round(ms, 2), "draws_per_s": n * P * k / ms * 1e3,
                  "frac_p_le_0.05": sig, "sum_Is_over_n1": float(r["Is"].double().sum() / (n - 1))}))
