#!/bin/bash
# the driver's scaling run: bench at N = 1, 2, 4, 8 back to back on one box.  usage: scripts/gpu_scale.sh tag "1 2 4 8"
TAG=${1:-r7b}; NS=${2:-"1 2 4 8"}
OUT=gpurun_out
mkdir -p $OUT
for N in $NS; do
  if [ "$N" = 1 ]; then
    timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-pipeline --no-perm --no-pageable --cpu-seconds 4 > $OUT/scale_n${N}_$TAG.json 2> $OUT/scale_n${N}_$TAG.err
  else
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > $OUT/scale_n${N}_$TAG.json 2> $OUT/scale_n${N}_$TAG.err
  fi
  echo "bench N=$N rc=$?" | tee -a $OUT/summary_$TAG.txt
  tail -2 $OUT/scale_n${N}_$TAG.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/scale_n${N}_$TAG.json").readline())
    print("N=$N value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d.get("e2e",{}).get("value"), d.get("e2e",{}).get("ms_per_step"), "weak", d.get("weak",{}).get("value"), "pca_sharded", {k:v for k,v in d.get("pca_sharded",{}).items() if k.endswith("_ms")})
except Exception as e: print("parse failed", e)
PY
done
