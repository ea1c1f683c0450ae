#!/bin/bash
# builds ab_libs/<name>.so = the library with moran.cu compiled with extra flags (experiment macros).  usage: build_variant.sh name "-DEXP_X=1"
set -e
NAME=$1; EXTRA=$2
cd "$(dirname "$0")/../chrysalis_b200/csrc"
mkdir -p ../../ab_libs build
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC -O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr -Xptxas -v $EXTRA -c moran.cu -o build/var_moran_$NAME.o 2> build/moran_$NAME.log
grep -A1 "moran_tile_kernelILb0ELb1ELb0E" build/moran_$NAME.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo
OBJS=$(ls build/*.o | grep -v "build/moran.o\|build/var_")
$NVCC $ARCH -shared -o ../../ab_libs/$NAME.so $OBJS build/var_moran_$NAME.o -lcudart_static -ldl -lrt -lpthread
