#!/bin/bash
TAG=${1:-r7o}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_moran_lattice_gpu.py -m gpu -q --timeout 600 --tb=short -p no:cacheprovider -k "tma" 2>&1 | cut -c1-400 | tail -12
for V in 0 1 0 1; do
CHR_LATTICE_VARIANT=$V timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-pipeline --no-perm 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('variant $V', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done
NCUCMD="python bench.py --g 2304 --steps 3 --warmup 3 --no-e2e --no-cpu --no-pipeline --no-perm"
CHR_LATTICE_VARIANT=1 $NCUCMD > $OUT/plain_$TAG.log 2>&1 && \
CHR_LATTICE_VARIANT=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:moran_lattice_kernel -s 4 -c 1 -f -o $OUT/moran_tma_$TAG $NCUCMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 $OUT/ncu_full_$TAG.log
