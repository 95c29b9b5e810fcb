#!/bin/bash
# A/B builds of the CUDA library with extra preprocessor flags, next to the shipped one:
#   tools/build_variant.sh NAME -DMCB_TURN_REPLICAS=1 ...   ->  montecarlocuda_b200/lib_exp/NAME/libmcb200.so
# Time it with  MCB200_LIBRARY=montecarlocuda_b200/lib_exp/NAME/libmcb200.so python bench.py ...
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
OUT=montecarlocuda_b200/lib_exp/$NAME
mkdir -p $OUT
FLAGS="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -I include"
for u in kernels_vanilla kernels_basket kernels_cva kernels_debug engine; do
  nvcc $FLAGS "$@" -c montecarlocuda_b200/csrc/$u.cu -o $OUT/$u.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $OUT/libmcb200.so $OUT/*.o
rm -f $OUT/*.o
ls -la $OUT
