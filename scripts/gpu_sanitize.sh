#!/bin/bash
# compute-sanitizer memcheck (ONE tool per call) over a small run of every kernel family.  usage: scripts/gpu_sanitize.sh tag
TAG=${1:-r7n}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python scripts/sanitize_small.py > $OUT/sanitize_plain_$TAG.log 2>&1; echo "plain rc=$?" | tee -a $OUT/summary_$TAG.txt
tail -2 $OUT/sanitize_plain_$TAG.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python scripts/sanitize_small.py > $OUT/sanitize_memcheck_$TAG.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/summary_$TAG.txt
tail -12 $OUT/sanitize_memcheck_$TAG.log | cut -c1-300
