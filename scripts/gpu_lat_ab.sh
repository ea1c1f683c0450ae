#!/bin/bash
# lattice parity tests, A/B of library builds, then one ncu --set full capture.  usage: scripts/gpu_lat_ab.sh tag lib1 lib2 ...
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_moran_lattice_gpu.py -m gpu -q --timeout 600 --tb=short -p no:cacheprovider 2>&1 | cut -c1-300 > $OUT/pytest_$TAG.log; echo "pytest rc=${PIPESTATUS[0]}" | tee -a $OUT/summary_$TAG.txt
tail -15 $OUT/pytest_$TAG.log
cp chrysalis_b200/libchrysalis_b200.so /tmp/keep.so
bash scripts/ab_kernels.sh "$@" 2>&1 | tee $OUT/ab_$TAG.txt
cp /tmp/keep.so chrysalis_b200/libchrysalis_b200.so
bash scripts/gpu_prof_full.sh $TAG moran_lattice_kernel
