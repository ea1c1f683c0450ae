#!/bin/bash
TAG=${1:-r1c}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary_$TAG.txt
tail -6 $OUT/pytest_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 --sweep 1,4,6,7 --no-e2e --no-cpu > $OUT/bench_sweep_$TAG.json 2> $OUT/bench_sweep_$TAG.err; echo "bench sweep rc=$?" | tee -a $OUT/summary_$TAG.txt
cat $OUT/bench_sweep_$TAG.json; tail -3 $OUT/bench_sweep_$TAG.err
timeout 600 python bench.py --n 4035 --g 36601 --k 6 --steps 10 --warmup 3 --sweep 1,4,6 --no-e2e --no-cpu > $OUT/bench_cfg1_$TAG.json 2> $OUT/bench_cfg1_$TAG.err; echo "bench cfg1 rc=$?" | tee -a $OUT/summary_$TAG.txt
cat $OUT/bench_cfg1_$TAG.json
NCUCMD="python bench.py --g 2304 --steps 3 --warmup 3 --no-e2e --no-cpu"
$NCUCMD > $OUT/plain_$TAG.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:moran_pair_kernel -s 4 -c 2 -f -o $OUT/moran_$TAG $NCUCMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?" | tee -a $OUT/summary_$TAG.txt
