#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r5d}
T0=$(date +%s)
timeout 900 python -m pytest tests/test_aa_gpu.py tests/test_multisample_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -8
AA_PRE_ITERS=30 timeout 900 python scripts/aa_breakdown.py --real 700000 30 12 30 > $OUT/aa_breakdown_$TAG.json 2> $OUT/aa_breakdown_$TAG.err; echo "aa rc=$?"
cut -c1-3000 $OUT/aa_breakdown_$TAG.json | tail -2; tail -3 $OUT/aa_breakdown_$TAG.err
AA_PRE_ITERS=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:aa_alpha_fixed -s 20 -c 1 -f -o $OUT/aa_alpha_$TAG python scripts/aa_breakdown.py --real 700000 30 12 20 > $OUT/ncu_aa_$TAG.log 2>&1; echo "ncu rc=$?"
echo "total $(( $(date +%s) - T0 )) s"
