#!/bin/bash
# GPU parity tests + smoke only.  usage: scripts/gpu_tests.sh [tag]
TAG=${1:-r1e}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --tb=short -p no:cacheprovider 2>&1 | cut -c1-400 > $OUT/pytest_$TAG.log; echo "pytest rc=${PIPESTATUS[0]}" | tee -a $OUT/summary_$TAG.txt
tail -60 $OUT/pytest_$TAG.log
timeout 600 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary_$TAG.txt
tail -8 $OUT/smoke_$TAG.log
