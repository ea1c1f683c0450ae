#!/bin/bash
# N-GPU check: multi-rank tests + bench at N ranks (peer-memory gather, then NCCL for comparison).  usage: scripts/gpu_multi2.sh tag N
TAG=$1; N=$2
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_multirank_gpu.py tests/test_p2p_gpu.py -m gpu -q --timeout 600 --tb=short -p no:cacheprovider 2>&1 | cut -c1-600 > $OUT/pytest_$TAG.log; echo "pytest rc=${PIPESTATUS[0]}" | tee -a $OUT/summary_$TAG.txt
tail -30 $OUT/pytest_$TAG.log
for MODE in 0 1 0 1; do
CHR_BENCH_NCCL=$MODE timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$MODE bench.py --gpus $N --steps 30 --warmup 5 --no-weak --no-e2e --no-pipeline > $OUT/bench_n${N}_nccl${MODE}_$TAG.json 2> $OUT/bench_n${N}_nccl${MODE}_$TAG.err; echo "bench N=$N nccl=$MODE rc=$?" | tee -a $OUT/summary_$TAG.txt
python -c "
import json; d=json.loads(open('$OUT/bench_n${N}_nccl${MODE}_$TAG.json').readline()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['parallelism'])"
tail -3 $OUT/bench_n${N}_nccl${MODE}_$TAG.err | cut -c1-300
done
