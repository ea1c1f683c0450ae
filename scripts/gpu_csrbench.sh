#!/bin/bash
TAG=${1:-r7f}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_moran_lattice_gpu.py -m gpu -q --timeout 600 --tb=short -p no:cacheprovider -k "csr" 2>&1 | cut -c1-400 | tail -12
timeout 600 python scripts/csr_bench.py > $OUT/csrbench_$TAG.json 2> $OUT/csrbench_$TAG.err; echo "csrbench rc=$?"
cat $OUT/csrbench_$TAG.json; tail -3 $OUT/csrbench_$TAG.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lattice|finalize|densify" -c 14 --csv --log-file $OUT/launches_$TAG.csv python scripts/csr_bench.py 4608 > $OUT/ncu_$TAG.log 2>&1
grep -v "^==" $OUT/launches_$TAG.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin): print(r['Kernel Name'][:70], r['Grid Size'], r['Metric Value'])" | head -14
