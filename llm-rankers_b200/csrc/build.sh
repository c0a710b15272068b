#!/bin/bash
# Builds libb200rank.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=../libb200rank.so
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
     -Xcompiler -fPIC -Xcompiler -Wall -shared -cudart static \
     ${B200RANK_NVCC_EXTRA} -o "$OUT" engine.cu
echo "built $(realpath $OUT)"
