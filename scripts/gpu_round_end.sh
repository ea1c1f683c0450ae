#!/bin/bash
# what the driver runs at round end, on one GPU: tests + smoke, both bench arms.  usage: scripts/gpu_round_end.sh [tag]
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-final}
bash scripts/gpu_tests.sh $TAG
python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench ref rc=$?" | tee -a $OUT/summary_$TAG.txt
cut -c1-300 $OUT/bench_ref_$TAG.json
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?" | tee -a $OUT/summary_$TAG.txt
python -c "
import json; d=json.load(open('$OUT/bench_$TAG.json')); print(d['value'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'], d['cpu_baseline']['value'], {k:v for k,v in d['pipeline'].items() if k!='cpu_baseline' and k!='gram_roofline'})"
