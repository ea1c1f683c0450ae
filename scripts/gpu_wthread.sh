#!/bin/bash
TAG=${1:-r9y}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_moran_gpu.py tests/test_moran_lattice_gpu.py tests/test_multirank_gpu.py tests/test_multisample_gpu.py -m gpu -q --timeout 600 --tb=short -p no:cacheprovider -x 2>&1 | cut -c1-500 | tail -6
for G in 2250 18000; do
CHR_SVG_TRACE=1 timeout 500 python bench.py --g $G --steps 5 --warmup 3 --no-pipeline --no-perm --no-cpu --no-pageable > $OUT/bench_${TAG}_g$G.json 2> $OUT/bench_${TAG}_g$G.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_${TAG}_g$G.json").readline())
print("g=$G value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "scipy pinned", d["e2e_scipy_pinned"]["ms_per_step"], d["e2e"]["max_abs_diff_to_resident_run"])
PY
grep "svg_stage trace" $OUT/bench_${TAG}_g$G.err | tail -1 | cut -c1-700
done
