#!/bin/bash
TAG=${1:-r9c}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_moran_perm_gpu.py -m gpu -q --timeout 900 --tb=short -p no:cacheprovider -x -k "not cfg3" 2>&1 | cut -c1-400 | tail -8
for v in 0 2 0 2; do
CHR_PERM_VARIANT=$v timeout 600 python scripts/perm_bench.py 12544 64 > $OUT/permbench_${TAG}_v$v.json 2> $OUT/permbench_${TAG}_v$v.err; echo "variant $v rc=$?"
cat $OUT/permbench_${TAG}_v$v.json; tail -2 $OUT/permbench_${TAG}_v$v.err
done
CHR_PERM_VARIANT=0 timeout 600 python scripts/perm_bench.py 25000 64 > $OUT/permbench_${TAG}_25k.json 2>&1; cat $OUT/permbench_${TAG}_25k.json
