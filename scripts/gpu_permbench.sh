#!/bin/bash
TAG=${1:-r6i}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_moran_perm_gpu.py -m gpu -q --timeout 900 --tb=short -p no:cacheprovider -x -k "not cfg3" 2>&1 | cut -c1-400 | tail -8
timeout 600 python scripts/perm_bench.py 12544 64 > $OUT/permbench_$TAG.json 2> $OUT/permbench_$TAG.err; echo "permbench rc=$?" | tee -a $OUT/summary_$TAG.txt
cat $OUT/permbench_$TAG.json; tail -3 $OUT/permbench_$TAG.err
if [ "$2" = ncu ]; then
NCUCMD="python scripts/perm_bench.py 2048 16"
$NCUCMD > $OUT/plain_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:moran_lattice_perm_kernel -s 1 -c 1 -f -o $OUT/perm_$TAG $NCUCMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu rc=$?" | tee -a $OUT/summary_$TAG.txt; tail -2 $OUT/ncu_full_$TAG.log
fi
