#!/bin/bash
# lattice tests + launch list of the resident step on a 2250-gene shard (what a rank of the 8-GPU run executes)
TAG=${1:-r7c}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_moran_lattice_gpu.py -m gpu -q --timeout 600 --tb=short -p no:cacheprovider 2>&1 | cut -c1-300 | tail -5
for G in 2250 18000; do
timeout 600 python bench.py --g $G --steps 20 --warmup 5 --no-e2e --no-cpu --no-pipeline --no-perm > $OUT/bench_g${G}_$TAG.json 2> $OUT/bench_g${G}_$TAG.err; echo "bench g=$G rc=$?" | tee -a $OUT/summary_$TAG.txt
python -c "
import json; d=json.loads(open('$OUT/bench_g${G}_$TAG.json').readline()); print('g=$G ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'])"
done
NCUCMD="python bench.py --g 2250 --steps 3 --warmup 3 --no-e2e --no-cpu --no-pipeline --no-perm"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"moran|fill|finalize" -s 30 -c 20 --csv --log-file $OUT/launches_$TAG.csv $NCUCMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?" | tee -a $OUT/summary_$TAG.txt
grep -v "^==" $OUT/launches_$TAG.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin): print(r['Kernel Name'][:60], r['Grid Size'], r['Metric Value'])" | head -24
