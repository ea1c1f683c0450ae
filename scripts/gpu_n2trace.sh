#!/bin/bash
# N-rank bench with the host timeline of the e2e step (rank 0).  usage: scripts/gpu_n2trace.sh tag N
TAG=$1; N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
CHR_SVG_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-weak --no-perm --no-cpu > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err; echo "bench N=$N rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/bench_n${N}_$TAG.json").readline())
print("value", d["value"], "ms", d["ms_per_step"], d["config"]["parallelism"])
for k in ("e2e","e2e_scipy_pinned"):
    print(k, d[k]["ms_per_step"], d[k]["h2d_bytes_per_step"])
print("pca_sharded", json.dumps(d.get("pca_sharded"))[:700])
PY
grep "svg_stage trace" $OUT/bench_n${N}_$TAG.err | cut -c1-1200; tail -2 $OUT/bench_n${N}_$TAG.err | cut -c1-300
