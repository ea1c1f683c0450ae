#!/bin/bash
TAG=${1:-r9a}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_aa_gpu.py -m gpu -q --timeout 600 --tb=short -p no:cacheprovider 2>&1 | cut -c1-400 | tail -8
timeout 600 python scripts/aa_bench.py > $OUT/aabench_$TAG.json 2> $OUT/aabench_$TAG.err; echo "aabench rc=$?"
cat $OUT/aabench_$TAG.json; tail -3 $OUT/aabench_$TAG.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/aa_launches_$TAG.csv python scripts/aa_bench.py 12 700000 1 > $OUT/ncu_$TAG.log 2>&1
grep -v "^==" $OUT/aa_launches_$TAG.csv | python -c "
import csv,sys,collections
t=collections.OrderedDict()
for r in csv.DictReader(sys.stdin):
    k=r['Kernel Name'][:60]; v=float(r['Metric Value'].replace(',',''))
    c=t.setdefault(k,[0,0.0]); c[0]+=1; c[1]+=v
for k,(c,v) in sorted(t.items(), key=lambda x:-x[1][1]): print(f'{k:60s} {c:5d} {v/1e3:10.1f} us total {v/c/1e3:9.1f} us avg')
" | head -30
