#!/bin/bash
# round record: every GPU test, smoke, the default bench (both arms).  usage: scripts/gpu_round.sh tag
TAG=${1:-r7a}
OUT=gpurun_out
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 --tb=short -p no:cacheprovider --durations=8 2>&1 | cut -c1-500 > $OUT/pytest_$TAG.log; echo "pytest rc=${PIPESTATUS[0]}" | tee -a $OUT/summary_$TAG.txt
tail -40 $OUT/pytest_$TAG.log
timeout 600 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary_$TAG.txt
tail -5 $OUT/smoke_$TAG.log
timeout 1500 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?" | tee -a $OUT/summary_$TAG.txt
cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench ref rc=$?" | tee -a $OUT/summary_$TAG.txt
cat $OUT/bench_ref_$TAG.json
