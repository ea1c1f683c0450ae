#!/bin/bash
# quick AA check: parity tests + per-kernel breakdown at the cfg4 shape.  usage: gpu_aa_quick.sh tag
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-aaq}
timeout 900 python -m pytest tests/test_aa_gpu.py tests/test_multisample_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -5
AA_PRE_ITERS=30 timeout 900 python scripts/aa_breakdown.py --real 700000 30 12 30 > $OUT/aa_breakdown_$TAG.json 2> $OUT/aa_breakdown_$TAG.err; echo "aa rc=$?"
python - <<PY
import json
for line in open("$OUT/aa_breakdown_$TAG.json"):
    d = json.loads(line)
    if "kernels" in d:
        print("kernel ms/iter", d["kernel_ms_per_iter"], "passes", sum(d["beta_passes"]) / len(d["beta_passes"]))
        for k, v in d["kernels"].items():
            if v["total_ms"] > 0.05: print(f'  {v["total_ms"] / d["iters"]:.3f} ms/iter  {v["calls"]:4d} calls  {k[:70]}')
    elif "wall_ms_per_iter_unprofiled" in d: print(d)
PY
tail -2 $OUT/aa_breakdown_$TAG.err
