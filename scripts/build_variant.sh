#!/bin/bash
# builds ab_libs/<name>.so = the library with ONE source compiled with extra flags (experiment macros).
# usage: build_variant.sh name "-DEXP_X=1" [source-stem, default moran]; scripts/ab_kernels.sh copies it over the in-tree library on the GPU box
set -e
NAME=$1; EXTRA=$2; SRC=${3:-moran}
cd "$(dirname "$0")/../chrysalis_b200/csrc"
mkdir -p ../../ab_libs build
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC -O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr -Xptxas -v $EXTRA -c $SRC.cu -o build/var_${SRC}_$NAME.o 2> build/${SRC}_$NAME.log
grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" build/${SRC}_$NAME.log | sort | uniq -c | tr "\n" " "; echo
OBJS=$(ls build/*.o | grep -v "build/$SRC.o\|build/var_")
$NVCC $ARCH -shared -o ../../ab_libs/$NAME.so $OBJS build/var_${SRC}_$NAME.o -lcudart_static -ldl -lrt -lpthread
