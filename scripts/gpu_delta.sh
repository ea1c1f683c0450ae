#!/bin/bash
TAG=${1:-r9q}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_moran_gpu.py -m gpu -q --timeout 600 --tb=short -p no:cacheprovider -k "compact or delta" 2>&1 | cut -c1-600 | tail -8
CHR_SVG_TRACE=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-pipeline --no-perm --no-cpu --no-pageable > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/bench_$TAG.json").readline())
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"])
for k in ("e2e","e2e_scipy_pinned"):
    print(k, d[k]["ms_per_step"], d[k]["value"], d[k]["h2d_bytes_per_step"], d[k]["max_abs_diff_to_resident_run"], d[k]["path"][:120])
PY
tail -3 $OUT/bench_$TAG.err | cut -c1-900
