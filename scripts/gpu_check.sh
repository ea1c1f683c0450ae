#!/bin/bash
# One GPU-box visit: parity tests, smoke, benches, ncu evidence.  Outputs land in gpurun_out/.
# usage: scripts/gpu_check.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt; free -g | head -2 >> $OUT/gpu_$TAG.txt
echo "== pytest" ; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary_$TAG.txt
tail -5 $OUT/pytest_gpu_$TAG.log
echo "== smoke" ; timeout 600 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary_$TAG.txt
tail -3 $OUT/smoke_$TAG.log
echo "== bench small" ; timeout 600 python bench.py --n 100000 --g 4096 --steps 5 --warmup 3 --sweep 0,1,2,3 > $OUT/bench_small_$TAG.json 2> $OUT/bench_small_$TAG.err; echo "bench_small rc=$?" | tee -a $OUT/summary_$TAG.txt
cat $OUT/bench_small_$TAG.json; tail -3 $OUT/bench_small_$TAG.err
echo "== bench full" ; timeout 1200 python bench.py --steps 20 --warmup 5 --sweep 0,1,2,3 > $OUT/bench_full_$TAG.json 2> $OUT/bench_full_$TAG.err; echo "bench_full rc=$?" | tee -a $OUT/summary_$TAG.txt
cat $OUT/bench_full_$TAG.json; tail -3 $OUT/bench_full_$TAG.err
echo "== reference arm" ; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench_ref rc=$?" | tee -a $OUT/summary_$TAG.txt
cat $OUT/bench_ref_$TAG.json
if [ "${NCU:-1}" = "1" ]; then
  NCUCMD="python bench.py --g 2304 --steps 3 --warmup 3 --no-e2e --no-cpu"
  echo "== ncu launch list"
  $NCUCMD > $OUT/plain_$TAG.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'moran|knn|csr|hilbert|bbox|scan' -c 60 --csv --log-file $OUT/launches_$TAG.csv $NCUCMD > $OUT/ncu_list_$TAG.log 2>&1
  echo "ncu list rc=$?" | tee -a $OUT/summary_$TAG.txt
  echo "== ncu full"
  $NCUCMD > $OUT/plain2_$TAG.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:moran_ell -s 4 -c 2 -f -o $OUT/moran_$TAG $NCUCMD > $OUT/ncu_full_$TAG.log 2>&1
  echo "ncu full rc=$?" | tee -a $OUT/summary_$TAG.txt
fi
cat $OUT/summary_$TAG.txt
