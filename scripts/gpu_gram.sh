#!/bin/bash
TAG=${1:-r9e}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_pca_gpu.py -m gpu -q --timeout 300 --tb=short -p no:cacheprovider -x 2>&1 | cut -c1-600 | tail -15
timeout 300 python scripts/gram_bench.py > $OUT/grambench_$TAG.json 2> $OUT/grambench_$TAG.err; echo "grambench rc=$?"
cat $OUT/grambench_$TAG.json; tail -3 $OUT/grambench_$TAG.err
