#!/bin/sh
# Build libstacb.so (sm_100a) in-tree. -fmad=false: fused multiply-adds are explicit in the source.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
  -Xcompiler -fPIC -shared -I ../../include ${STACB_NVCC_EXTRA} \
  stacb_kernels.cu -o ../libstacb.so
echo "built $(cd .. && pwd)/libstacb.so"
