#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r5c}
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -15
timeout 600 python scripts/e2e_breakdown.py > $OUT/e2e_breakdown_$TAG.json 2> $OUT/e2e_breakdown_$TAG.err; echo "e2e rc=$?"
cat $OUT/e2e_breakdown_$TAG.json
AA_PRE_ITERS=30 timeout 900 python scripts/aa_breakdown.py --real 700000 30 12 30 > $OUT/aa_breakdown_$TAG.json 2> $OUT/aa_breakdown_$TAG.err; echo "aa rc=$?"
cat $OUT/aa_breakdown_$TAG.json; tail -3 $OUT/aa_breakdown_$TAG.err
echo "total $(( $(date +%s) - T0 )) s"
