"""Times the CSR-fed lattice Moran call (chr_moran_lattice_csr_f32) at cfg4 shape next to densify + the dense kernel.
usage: python scripts/csr_bench.py [genes] [n_spots]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chrysalis_b200 import device as dv  # noqa: E402
from chrysalis_b200 import synth  # noqa: E402

g = int(sys.argv[1]) if len(sys.argv) > 1 else 18000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 700_000
args = argparse.Namespace(n=n, g=g, k=8, coords="square")
xy = bench.make_coords(args)
xy_d = dv.to_device(xy, torch.float64)
bbox = dv.spatial_bbox(xy_d)
nbr = dv.knn_build(xy_d, 8, bbox)
lat = dv.lattice_plan(xy_d, nbr, bbox)
csr = bench.build_counts(args, xy, g, 0, g, 42, torch, dv, synth)
gene_col = torch.arange(g, dtype=torch.int32, device="cuda")
scale = dv.median_scale(dv.csr_spot_totals(csr, gene_col))
ev = lambda: torch.cuda.Event(enable_timing=True)
res = {"n": n, "genes": g, "nnz": int(csr.indices.numel())}
for rep in range(3):
    e0, e1 = ev(), ev()
    e0.record()
    I, _, _, uns = dv.moran_lattice_csr(csr, gene_col, g, scale, lat, apply_log1p=True, timing=True)
    e1.record()
    torch.cuda.synchronize()
    res["csr_call_ms"] = round(e0.elapsed_time(e1), 3)
    res["csr_main_kernel_ms"] = round(dv.last_kernel_ms(), 3)
dense = torch.zeros((lat.n_rows, dv.dense_ld(g)), dtype=torch.float32, device="cuda")
for rep in range(2):
    e0, e1, e2 = ev(), ev(), ev()
    dense.zero_()
    e0.record()
    dv.csr_densify(csr, gene_col, g, scale, lat.row_of_spot, apply_log1p=True, out=dense)
    e1.record()
    I2, _, _ = dv.moran_lattice(dense, lat, g, timing=True)
    e2.record()
    torch.cuda.synchronize()
    res["densify_ms"], res["dense_call_ms"] = round(e0.elapsed_time(e1), 3), round(e1.elapsed_time(e2), 3)
res["max_abs_diff"] = float((I - I2).abs().max())
print(json.dumps(res))
