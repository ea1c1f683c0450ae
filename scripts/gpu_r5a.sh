#!/bin/bash
# session r5a: tests + smoke, default bench, ncu of the committed Moran kernel, e2e and AA breakdowns
OUT=gpurun_out; mkdir -p $OUT
T0=$(date +%s)
bash scripts/gpu_tests.sh r5a
echo "tests took $(( $(date +%s) - T0 )) s" | tee -a $OUT/summary_r5a.txt
python bench.py > $OUT/bench_r5a.json 2> $OUT/bench_r5a.err; echo "bench rc=$?" | tee -a $OUT/summary_r5a.txt
cut -c1-1500 $OUT/bench_r5a.json
timeout 600 python scripts/e2e_breakdown.py > $OUT/e2e_breakdown_r5a.json 2> $OUT/e2e_breakdown_r5a.err; echo "e2e rc=$?" | tee -a $OUT/summary_r5a.txt
cat $OUT/e2e_breakdown_r5a.json
timeout 600 python scripts/aa_breakdown.py > $OUT/aa_breakdown_r5a.json 2> $OUT/aa_breakdown_r5a.err; echo "aa rc=$?" | tee -a $OUT/summary_r5a.txt
cat $OUT/aa_breakdown_r5a.json; tail -3 $OUT/aa_breakdown_r5a.err
bash scripts/gpu_prof.sh r5a moran_tile
echo "total $(( $(date +%s) - T0 )) s" | tee -a $OUT/summary_r5a.txt
