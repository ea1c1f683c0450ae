#!/bin/bash
# ncu launch list of the default bench command's timed step (cfg4, 18,000 genes resident): kernel shares of the step
TAG=${1:-r9B}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-pipeline --no-perm"
timeout 600 $CMD > $OUT/plain_ll_$TAG.json 2> $OUT/plain_ll_$TAG.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"moran|lattice|finalize|p2p" -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
python - <<PY
import csv, json
rows=[r for r in csv.DictReader(l for l in open("$OUT/launches_$TAG.csv") if not l.startswith("=="))]
d=json.loads(open("$OUT/plain_ll_$TAG.json").readline())
print("plain run: ms_per_step", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"], "kernels_per_step", d.get("kernels_per_step"))
# the last step: the launches after the last finalize but one
names=[r["Kernel Name"] for r in rows]
fin=[i for i,nm in enumerate(names) if "moran_finalize" in nm]
lo=fin[-2]+1 if len(fin)>1 else 0
step=rows[lo:fin[-1]+1]
tot=sum(float(r["Metric Value"].replace(",","")) for r in step)
for r in step: print(f'{r["Kernel Name"][:70]:70s} grid {r["Grid Size"]:>14s} {float(r["Metric Value"].replace(",",""))/1e3:9.1f} us  {100*float(r["Metric Value"].replace(",",""))/tot:5.1f} %')
print("step under ncu (cold caches, serialised):", round(tot/1e6,3), "ms")
PY
