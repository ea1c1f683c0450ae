#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r5i}
T0=$(date +%s)
bash scripts/gpu_tests.sh $TAG
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?" | tee -a $OUT/summary_$TAG.txt
cut -c1-3800 $OUT/bench_$TAG.json
python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench ref rc=$?" | tee -a $OUT/summary_$TAG.txt
cut -c1-800 $OUT/bench_ref_$TAG.json
bash scripts/gpu_prof.sh $TAG moran_tile
echo "total $(( $(date +%s) - T0 )) s" | tee -a $OUT/summary_$TAG.txt
