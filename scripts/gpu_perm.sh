#!/bin/bash
TAG=${1:-r6h}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_moran_perm_gpu.py tests/test_moran_gpu.py -m gpu -q --timeout 900 --tb=short -p no:cacheprovider -x 2>&1 | cut -c1-600 > $OUT/pytest_$TAG.log; echo "pytest rc=${PIPESTATUS[0]}" | tee -a $OUT/summary_$TAG.txt
tail -40 $OUT/pytest_$TAG.log
