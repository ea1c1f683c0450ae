#!/bin/bash
TAG=${1:-r9n}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/aa_launches_$TAG.csv python scripts/aa_bench.py 40 700000 1 > $OUT/ncu_$TAG.log 2>&1
grep -v "^==" $OUT/aa_launches_$TAG.csv | python -c "
import csv,sys,collections
rows=list(csv.DictReader(sys.stdin))
# steady state: the last third of the launches
rows=rows[len(rows)*2//3:]
t=collections.OrderedDict()
for r in rows:
    k=r['Kernel Name'][:60]; v=float(r['Metric Value'].replace(',',''))
    c=t.setdefault(k,[0,0.0]); c[0]+=1; c[1]+=v
for k,(c,v) in sorted(t.items(), key=lambda x:-x[1][1]): print(f'{k:60s} {c:5d} {v/1e3:10.1f} us total {v/c/1e3:9.1f} us avg')
" | head -30
