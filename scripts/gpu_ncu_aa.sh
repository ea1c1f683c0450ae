#!/bin/bash
# ncu --set full of one AA kernel at the cfg4 shape.  usage: gpu_ncu_aa.sh tag kernel-regex [launch-skip]
OUT=gpurun_out; mkdir -p $OUT
TAG=$1; KRE=$2; SKIP=${3:-20}
AA_PRE_ITERS=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -f -o $OUT/aa_${TAG} python scripts/aa_breakdown.py --real 700000 30 12 20 > $OUT/ncu_aa_$TAG.log 2>&1; echo "ncu rc=$?"
tail -2 $OUT/ncu_aa_$TAG.log
