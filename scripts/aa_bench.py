"""Times the archetypal-analysis fit at cfg4 shape (700k spots x 30 PCs, 12 archetypes): restarts sequential vs concurrent.
usage: python scripts/aa_bench.py [max_iter] [n_spots] [workers ...]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chrysalis_b200 import core  # noqa: E402
from chrysalis_b200 import device as dv  # noqa: E402

max_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n = int(sys.argv[2]) if len(sys.argv) > 2 else 700_000
workers = [int(w) for w in sys.argv[3:]] or [3, 3, 1, 1, 3]
d, a = 30, 12
rng = np.random.default_rng(7)
V = rng.normal(size=(a, d)) * 3.0 * (0.85 ** np.arange(d))          # decaying scales like PCA scores
X = np.empty((n, d))
for lo in range(0, n, 100_000):
    hi = min(n, lo + 100_000)
    X[lo:hi] = rng.dirichlet(np.full(a, 0.4), size=hi - lo) @ V + rng.normal(0, 0.05, (hi - lo, d))
Xd = dv.to_device(X, torch.float64)
res = {"n": n, "d": d, "A": a, "max_iter": max_iter}
ref = None
res["runs"] = []
if os.environ.get("AA_EMPTY_CACHE"):
    torch.cuda.empty_cache()
for w in workers:
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    f = core._aa_fit_device(Xd, a, max_iter, 3, 0.001, 42, restart_workers=w)
    torch.cuda.synchronize()
    rec = {"workers": w, "fit_s": round(time.perf_counter() - t0, 4), "n_iter_best": f["n_iter"]}
    if ref is None:
        ref = f
    else:
        rec["identical_to_first"] = bool(torch.equal(ref["alphas"], f["alphas"]) and ref["rss"] == f["rss"])
    res["runs"].append(rec)
print(json.dumps(res))
