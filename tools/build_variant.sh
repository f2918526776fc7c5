#!/bin/bash
# Builds a variant of the library for A/B timing on the GPU box: tools/build_variant.sh NAME [-DFLAG=VALUE ...]
#   -> sfm-nerf_b200/variants/libnerf_b200_NAME.so   (run anything with NERF_B200_LIB=<that path>)
set -e
cd "$(dirname "$0")/.."
mkdir -p sfm-nerf_b200/variants
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" \
  -o sfm-nerf_b200/variants/libnerf_b200_$name.so sfm-nerf_b200/csrc/api.cu
echo sfm-nerf_b200/variants/libnerf_b200_$name.so
