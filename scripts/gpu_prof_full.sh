#!/bin/bash
# one ncu --set full capture of the Moran tile kernel on a 2304-gene slab.  usage: scripts/gpu_prof_full.sh [tag] [kernel-regex]
TAG=${1:-r4}
KRE=${2:-moran_tile}
OUT=gpurun_out
mkdir -p $OUT
NCUCMD="python bench.py --g 2304 --steps 3 --warmup 3 --no-e2e --no-cpu --no-pipeline"
$NCUCMD > $OUT/plain_$TAG.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 4 -c 1 -f -o $OUT/moran_$TAG $NCUCMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?" | tee -a $OUT/summary_$TAG.txt
tail -3 $OUT/ncu_full_$TAG.log
