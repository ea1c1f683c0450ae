#!/bin/bash
# bench lines at the other BASELINE shapes with the current kernels (cfg1 hex k=6, cfg2 random k=8): parity-test cases, not the headline
TAG=${1:-r7l}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python bench.py --n 4035 --g 36601 --k 6 --coords hex --steps 30 --warmup 5 --svgs 1000 --pcs 20 --archetypes 8 --no-perm --cpu-seconds 6 > $OUT/bench_cfg1_$TAG.json 2> $OUT/bench_cfg1_$TAG.err; echo "cfg1 rc=$?" | tee -a $OUT/summary_$TAG.txt
timeout 900 python bench.py --n 40000 --g 20000 --k 8 --coords random --steps 30 --warmup 5 --svgs 1000 --pcs 20 --archetypes 8 --no-perm --cpu-seconds 6 > $OUT/bench_cfg2_$TAG.json 2> $OUT/bench_cfg2_$TAG.err; echo "cfg2 rc=$?" | tee -a $OUT/summary_$TAG.txt
for c in cfg1 cfg2; do python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${c}_$TAG.json").readline())
    print("$c", d["config"]["workload"], "| value", d["value"], "ms", d["ms_per_step"], "kernel", d["roofline"]["kernel"], d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "| e2e", d["e2e"]["ms_per_step"], "| pipeline", d["pipeline"]["pinned"], "| parity", d["parity"]["max_rel_err_vs_oracle"], d["pipeline"]["pca_parity"]["min_subspace_cosine_vs_sklearn_arpack"], d["pipeline"]["aa_parity"]["max_abs_alpha_err_vs_oracle"])
except Exception as e: print("$c parse failed", e)
PY
tail -3 $OUT/bench_${c}_$TAG.err | cut -c1-400
done
