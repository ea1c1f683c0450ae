#!/bin/bash
TAG=${1:-r9w}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python scripts/gram_bench.py > $OUT/plain_gram_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_tcgen05_super_kernel -s 1 -c 1 -f -o $OUT/gram_$TAG python scripts/gram_bench.py > $OUT/ncu_gram_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/ncu_gram_$TAG.log; cat $OUT/plain_gram_$TAG.log | tail -1
