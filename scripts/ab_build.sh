#!/bin/bash
# Builds an alternative libb200_bev_ops with extra nvcc defines for A/B runs:  scripts/ab_build.sh <tag> -DFOO=1 ...
# -> bevformer_tensorrt_b200/lib/libb200_bev_ops_<tag>.so ; select it with B200_BEV_OPS_LIB=<path>.
set -e
TAG=$1; shift
cd "$(dirname "$0")/.."
OUT=bevformer_tensorrt_b200/lib/libb200_bev_ops_${TAG}.so
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -shared \
  -I include -I bevformer_tensorrt_b200/csrc "$@" -o $OUT bevformer_tensorrt_b200/csrc/*.cu
echo $OUT
