#!/bin/bash
# Moran parity tests + kernel sweep.  usage: scripts/gpu_moran.sh [tag] [sweep]
TAG=${1:-r1f}
SWEEP=${2:-0,4,5}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_moran_gpu.py -m gpu -q -x --timeout 600 --tb=short -p no:cacheprovider 2>&1 | cut -c1-300 > $OUT/pytest_$TAG.log; echo "pytest rc=${PIPESTATUS[0]}" | tee -a $OUT/summary_$TAG.txt
tail -25 $OUT/pytest_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 --sweep $SWEEP --no-e2e --cpu-seconds 6 > $OUT/bench_sweep_$TAG.json 2> $OUT/bench_sweep_$TAG.err; echo "bench sweep rc=$?" | tee -a $OUT/summary_$TAG.txt
cat $OUT/bench_sweep_$TAG.json; tail -5 $OUT/bench_sweep_$TAG.err
