#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r5r}
bash scripts/gpu_tests.sh $TAG
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$OUT/bench_$TAG.json')); print(d['value'], d['roofline'], d['e2e']['ms_per_step'], d['e2e']['value'], d['parity'], {k:v for k,v in d['pipeline'].items() if k!='cpu_baseline' and k!='gram_roofline'})"
