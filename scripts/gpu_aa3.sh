#!/bin/bash
TAG=${1:-r9u}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_aa_gpu.py tests/test_full_shapes_gpu.py -m gpu -q --timeout 600 --tb=short -p no:cacheprovider -x -k "not cfg5 and not ties" 2>&1 | cut -c1-700 | tail -12
for ENVV in "" "CHR_AA_RSS_DIRECT=1" ""; do
env $ENVV timeout 600 python scripts/aa_bench.py 200 700000 3 3 1 2> $OUT/aabench_$TAG.err | tail -1; echo "[$ENVV] rc=$?"
done
