"""Wall-clock composition of core.aa_stage at a cfg-shaped size: seeding, first (cold) iteration, warm iterations.
usage: python scripts/aa_stage_timing.py [N d A]"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chrysalis_b200 import core, device as dv

n, d, A = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (700000, 30, 12))]
torch.cuda.set_device(0)
rng = np.random.default_rng(0)
X = (rng.normal(size=(n, d)) * (3.0 / np.sqrt(1 + np.arange(d)))).astype(np.float32)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
out = {}
for rep in range(2):
    t0 = sync()
    Xd = dv.to_device(np.ascontiguousarray(X, dtype=np.float64), torch.float64)
    t1 = sync()
    rows = dv.aa_furthest_sum(Xd, A, 123)
    t2 = sync()
    st = dv.AAState(Xd, A); st.init_from_rows(rows)
    t3 = sync()
    st.iterate(); t4 = sync(); p0 = st.beta_passes
    st.iterate(); t5 = sync()
    for _ in range(50): st.iterate()
    t6 = sync()
    al = st.alphas.cpu().numpy(); t7 = sync()
    out[f"rep{rep}_ms"] = {"upload_f64": (t1 - t0) * 1e3, "furthest_sum": (t2 - t1) * 1e3, "state_init": (t3 - t2) * 1e3,
                           "iter0_cold": (t4 - t3) * 1e3, "iter0_beta_passes": p0, "iter1": (t5 - t4) * 1e3, "warm_iter_avg": (t6 - t5) * 1e3 / 50,
                           "alphas_d2h": (t7 - t6) * 1e3}
print(json.dumps(out))
for rep in range(2):
    t0 = sync()
    fit = core.aa_stage(X, A, max_iter=200)
    print(json.dumps({"aa_stage_200x3_s": round(sync() - t0, 4), "rss": fit["rss"], "n_iter": fit["n_iter"]}))
