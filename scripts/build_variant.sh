#!/bin/bash
# experimental build of libpbgpu.so with extra compile-time definitions, for A/B runs through PBGPU_LIB:
#   scripts/build_variant.sh nofull_bin -DPBGPU_BIN_NOFULL   ->  build_variants/libpbgpu_nofull_bin.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build_variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-O3 -shared -cudart static \
  "$@" -I include -o build_variants/libpbgpu_$name.so polars_bio_b200/csrc/arrow_bridge.cpp polars_bio_b200/csrc/pbgpu.cu -lpthread
ls -la build_variants/libpbgpu_$name.so
