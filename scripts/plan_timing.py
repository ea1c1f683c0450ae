"""Host + device time of the W build of detect_svgs on a square lattice (cfg4: 700k bins, k = 8): bbox, kNN, lattice plan steps.
usage: python scripts/plan_timing.py [n]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chrysalis_b200 import device as dv  # noqa: E402
from chrysalis_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 700_000
xy = dv.to_device(synth.make_coords("square", n, seed=42), torch.float64)


def clock(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out


res = {"n": n}
res["bbox_ms"], bbox = clock(lambda: dv.spatial_bbox(xy))
res["knn_ms"], nbr = clock(lambda: dv.knn_build(xy, 8, bbox))
res["pitch_ms"], pitch = clock(lambda: dv.lattice_pitch_from_nn(xy, bbox, (xy[nbr[:, 0].long()] - xy).norm(dim=1)))
res["nodes_ms"], (ix, iy) = clock(lambda: dv.lattice_nodes(xy, bbox, pitch))
res["layout_ms"], lay = clock(lambda: dv.lattice_layout(ix, iy))
res["corrections_ms"], cor = clock(lambda: dv.lattice_corrections(ix, iy, nbr, lay[4]))
res["plan_total_ms"], plan = clock(lambda: dv.lattice_plan(xy, nbr, bbox))
res["n_corr"] = int(plan.corr.shape[0])
print(json.dumps(res))
