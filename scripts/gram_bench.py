"""Times the tcgen05 Gram kernel at cfg4 shape (700k spots x 2000 SVG columns): 256 x 256 super-tile kernel vs 128 x 128 tile kernel.
usage: python scripts/gram_bench.py [n] [s]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chrysalis_b200 import _lib  # noqa: E402
from chrysalis_b200 import device as dv  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 700_000
s = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
P = torch.empty((n, dv.dense_ld(s)), dtype=torch.float32, device="cuda")
P.normal_(0.4, 0.6).clamp_(min=0.0)
P[:, :64] += torch.sin(torch.arange(n, device="cuda") / 9000.0)[:, None]
mean = dv.column_sums(P, s) / n
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
res = {"n": n, "s": s}
out = {}
for v in (0, 2, 0, 2):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    out[v] = dv.gram_centered(P, s, mean, timing=True, variant=v)
    ev1.record()
    torch.cuda.synchronize()
    ms = dv.last_kernel_ms(_lib.FAMILY_GRAM)
    res[f"variant{v}"] = {"kernel_ms": round(ms, 3), "stage_ms": round(ev0.elapsed_time(ev1), 3),
                          "useful_tflops": round(2.0 * n * s * s / (ms * 1e-3) / 1e12, 1),
                          "useful_frac_of_burst": round(2.0 * n * s * s / (ms * 1e-3) / 1e12 / peaks["bf16_tflops"], 4)}
res["bit_identical"] = bool(torch.equal(out[0], out[2]))
print(json.dumps(res))
