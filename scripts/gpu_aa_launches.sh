#!/bin/bash
# per-kernel times of steady-state AA iterations (ncu launch list; skips the first launches).  usage: gpu_aa_launches.sh tag [skip] [count]
TAG=${1:-r9n}; SKIP=${2:-2000}; CNT=${3:-600}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python scripts/aa_bench.py 120 700000 1 > $OUT/plain_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $CNT --csv --log-file $OUT/aa_launches_$TAG.csv python scripts/aa_bench.py 120 700000 1 > $OUT/ncu_$TAG.log 2>&1
grep -v "^==" $OUT/aa_launches_$TAG.csv | python -c "
import csv,sys,collections
rows=list(csv.DictReader(sys.stdin))
t=collections.OrderedDict()
for r in rows:
    k=r['Kernel Name'][:60]; v=float(r['Metric Value'].replace(',',''))
    c=t.setdefault(k,[0,0.0,[]]); c[0]+=1; c[1]+=v; c[2].append(v)
n_it=max(1,t.get([k for k in t if 'aa_pinv' in k][0],[1])[0])
print('iterations in window:', n_it)
for k,(c,v,l) in sorted(t.items(), key=lambda x:-x[1][1]): print(f'{k:60s} {c:5d} calls {v/1e3/n_it:9.1f} us/iter  avg {v/c/1e3:8.1f} us  min {min(l)/1e3:7.1f} max {max(l)/1e3:7.1f}')
" | head -30
