#!/bin/bash
# lattice-kernel parity tests + a short resident bench.  usage: scripts/gpu_lat.sh [tag]
TAG=${1:-r6a}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_moran_lattice_gpu.py tests/test_moran_gpu.py -m gpu -q --timeout 600 --tb=short -p no:cacheprovider 2>&1 | cut -c1-400 > $OUT/pytest_$TAG.log; echo "pytest rc=${PIPESTATUS[0]}" | tee -a $OUT/summary_$TAG.txt
tail -40 $OUT/pytest_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-pipeline --cpu-seconds 6 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?" | tee -a $OUT/summary_$TAG.txt
cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
