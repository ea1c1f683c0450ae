#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
python - <<'PY' 2>&1 | tail -20
import os, sys, time, json
os.environ["CHR_SVG_TRACE"] = "1"
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from chrysalis_b200 import core, synth, device as dv
class A: pass
a = A(); a.n, a.g, a.k = 700000, 18000, 8
torch.cuda.set_device(0)
xy_s, nbr, csr = bench.build_device_workload(a, 0, torch, dv, synth)
X, g_scored = bench.dense_from_csr(csr, 1e-4, torch, dv)      # bench keeps the resident matrix and the device CSR alive
host = dv.PinnedCSR(csr.indptr.cpu().pin_memory(), csr.indices.cpu().pin_memory(), csr.data.cpu().pin_memory(), csr.shape)
xy_host = xy_s.cpu().numpy(); xy_host[:, 1] *= -1
def run(tag):
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        core.svg_stage(host, xy_host, 1e-4, 8, False, already_log1p=False, shard=False)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        print(tag, rep, round(dt, 2), core.SVG_TRACE)
run("with resident X + device csr")
del X; torch.cuda.empty_cache()
run("X freed")
del csr, nbr; torch.cuda.empty_cache()
run("csr freed")
PY
