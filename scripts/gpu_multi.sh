#!/bin/bash
# 2-GPU checks: NCCL parity script + bench at N=1,2.  usage: scripts/gpu_multi.sh [tag] [ngpus]
TAG=${1:-r1t}
NG=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > $OUT/dist_check_$TAG.log 2>&1; echo "dist_check rc=$?" | tee -a $OUT/summary_$TAG.txt
tail -3 $OUT/dist_check_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 30 --warmup 5 > $OUT/bench_n${NG}_$TAG.json 2> $OUT/bench_n${NG}_$TAG.err; echo "bench N=$NG rc=$?" | tee -a $OUT/summary_$TAG.txt
tail -2 $OUT/bench_n${NG}_$TAG.err; cut -c1-600 $OUT/bench_n${NG}_$TAG.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $NG --steps 2 --warmup 1 > $OUT/bench_ref_n${NG}_$TAG.json 2> $OUT/bench_ref_n${NG}_$TAG.err; echo "bench ref N=$NG rc=$?" | tee -a $OUT/summary_$TAG.txt
cut -c1-400 $OUT/bench_ref_n${NG}_$TAG.json
