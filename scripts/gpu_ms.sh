#!/bin/bash
TAG=${1:-r6f}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_multisample_gpu.py tests/test_multirank_gpu.py tests/test_aa_gpu.py -m gpu -q --timeout 600 --tb=short -p no:cacheprovider 2>&1 | cut -c1-400 > $OUT/pytest_$TAG.log; echo "pytest rc=${PIPESTATUS[0]}" | tee -a $OUT/summary_$TAG.txt
tail -40 $OUT/pytest_$TAG.log
