#!/bin/bash
TAG=${1:-r7e}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_moran_lattice_gpu.py -m gpu -q --timeout 600 --tb=short -p no:cacheprovider 2>&1 | cut -c1-500 > $OUT/pytest_$TAG.log; echo "pytest rc=${PIPESTATUS[0]}" | tee -a $OUT/summary_$TAG.txt
tail -40 $OUT/pytest_$TAG.log
CHR_SVG_TRACE=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-pipeline --no-perm --no-cpu --no-pageable > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?" | tee -a $OUT/summary_$TAG.txt
python - <<PY
import json
d=json.loads(open("$OUT/bench_$TAG.json").readline())
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"])
for k in ("e2e","e2e_scipy_pinned"):
    print(k, d[k]["ms_per_step"], d[k]["value"], d[k]["h2d_bytes_per_step"], d[k]["max_abs_diff_to_resident_run"])
PY
tail -4 $OUT/bench_$TAG.err | cut -c1-900
