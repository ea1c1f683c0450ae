#!/bin/bash
# smoke + launch list + one ncu --set full capture of the lattice kernel on a 2304-gene slab of cfg4.  usage: scripts/gpu_prof_lat.sh [tag]
TAG=${1:-r6d}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary_$TAG.txt
tail -6 $OUT/smoke_$TAG.log
NCUCMD="python bench.py --g 2304 --steps 3 --warmup 3 --no-e2e --no-cpu --no-pipeline"
$NCUCMD > $OUT/plain_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_$TAG.csv $NCUCMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?" | tee -a $OUT/summary_$TAG.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:moran_lattice_kernel -s 4 -c 1 -f -o $OUT/moran_$TAG $NCUCMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?" | tee -a $OUT/summary_$TAG.txt
tail -3 $OUT/ncu_full_$TAG.log
