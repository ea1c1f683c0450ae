#!/bin/bash
# ncu launch list + full capture of the Moran kernel on a 2304-gene slab.  usage: scripts/gpu_prof.sh [tag] [kernel-regex] [extra bench args]
TAG=${1:-r1g}
KRE=${2:-moran_sym}
EXTRA=${3:-}
OUT=gpurun_out
mkdir -p $OUT
NCUCMD="python bench.py --g 2304 --steps 3 --warmup 3 --no-e2e --no-cpu --no-pipeline $EXTRA"
$NCUCMD > $OUT/plain_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'moran|sym_' -c 200 --csv --log-file $OUT/launches_$TAG.csv $NCUCMD > $OUT/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?" | tee -a $OUT/summary_$TAG.txt
$NCUCMD > $OUT/plain2_$TAG.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 4 -c 1 -f -o $OUT/moran_$TAG $NCUCMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?" | tee -a $OUT/summary_$TAG.txt
tail -3 $OUT/ncu_full_$TAG.log
