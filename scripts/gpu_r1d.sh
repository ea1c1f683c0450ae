#!/bin/bash
TAG=${1:-r1d}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -x > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary_$TAG.txt
tail -30 $OUT/pytest_$TAG.log
timeout 600 python -m pytest tests/test_pca_gpu.py -m gpu -q --timeout 300 > $OUT/pytest_pca_$TAG.log 2>&1; echo "pytest pca rc=$?" | tee -a $OUT/summary_$TAG.txt
tail -40 $OUT/pytest_pca_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 --sweep 0,6 --no-e2e --no-cpu > $OUT/bench_sweep_$TAG.json 2> $OUT/bench_sweep_$TAG.err; echo "bench sweep rc=$?" | tee -a $OUT/summary_$TAG.txt
cat $OUT/bench_sweep_$TAG.json; tail -3 $OUT/bench_sweep_$TAG.err
