#!/bin/bash
# N-GPU check: multi-rank tests + bench at N ranks.  usage: scripts/gpu_multi2.sh tag N
TAG=$1; N=$2
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_multirank_gpu.py -m gpu -q --timeout 600 --tb=short -p no:cacheprovider 2>&1 | cut -c1-400 > $OUT/pytest_$TAG.log; echo "pytest rc=${PIPESTATUS[0]}" | tee -a $OUT/summary_$TAG.txt
tail -30 $OUT/pytest_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err; echo "bench N=$N rc=$?" | tee -a $OUT/summary_$TAG.txt
cat $OUT/bench_n${N}_$TAG.json; tail -8 $OUT/bench_n${N}_$TAG.err
