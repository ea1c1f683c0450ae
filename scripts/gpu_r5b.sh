#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
T0=$(date +%s)
timeout 300 python -m pytest tests/test_moran_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "svg or detect or filter or prep or csr" 2>&1 | tail -3
timeout 600 python scripts/e2e_breakdown.py > $OUT/e2e_breakdown_r5b.json 2> $OUT/e2e_breakdown_r5b.err; echo "e2e rc=$?"
cat $OUT/e2e_breakdown_r5b.json
CHRYSALIS_B200_LIB=$PWD/ab_libs/aastats.so AA_PRE_ITERS=30 timeout 900 python scripts/aa_breakdown.py --real 700000 30 12 30 > $OUT/aa_breakdown_r5b.json 2> $OUT/aa_breakdown_r5b.err; echo "aa rc=$?"
cat $OUT/aa_breakdown_r5b.json; tail -3 $OUT/aa_breakdown_r5b.err
AA_PRE_ITERS=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:aa_alpha -s 20 -c 1 -f -o $OUT/aa_alpha_r5b python scripts/aa_breakdown.py --real 700000 30 12 20 > $OUT/ncu_aa_r5b.log 2>&1; echo "ncu rc=$?"
tail -3 $OUT/ncu_aa_r5b.log
echo "total $(( $(date +%s) - T0 )) s"
