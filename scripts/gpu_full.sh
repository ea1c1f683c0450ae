#!/bin/bash
# all GPU tests + smoke + a reduced-size bench + the default bench.  usage: scripts/gpu_full.sh [tag]
TAG=${1:-r6c}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --tb=short -p no:cacheprovider 2>&1 | cut -c1-400 > $OUT/pytest_$TAG.log; echo "pytest rc=${PIPESTATUS[0]}" | tee -a $OUT/summary_$TAG.txt
tail -40 $OUT/pytest_$TAG.log
timeout 600 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary_$TAG.txt
tail -6 $OUT/smoke_$TAG.log
timeout 600 python bench.py --n 60000 --g 3000 --steps 5 --warmup 3 --svgs 500 --pcs 20 --archetypes 8 --aa-max-iter 30 --cpu-seconds 3 > $OUT/bench_small_$TAG.json 2> $OUT/bench_small_$TAG.err; echo "bench small rc=$?" | tee -a $OUT/summary_$TAG.txt
cat $OUT/bench_small_$TAG.json; tail -5 $OUT/bench_small_$TAG.err
CHR_SVG_TRACE=1 timeout 1200 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?" | tee -a $OUT/summary_$TAG.txt
cat $OUT/bench_$TAG.json; tail -12 $OUT/bench_$TAG.err
