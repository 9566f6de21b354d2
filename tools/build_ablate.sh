#!/bin/bash
# Build the ablation/profiling variant of the library (never shipped): same sources, -DCA_ABLATE enables the
# CA_ABLATE environment switches of the row-block kernel (bit 0 skip histogram, bit 1 skip gather, bit 2 skip
# table zeroing + size staging, bit 3 skip label tile loads, bits 8.. cap on partners per phase).
set -e
cd "$(dirname "$0")/.."
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -DCA_ABLATE -Xcompiler -fPIC,-O3 --threads 0 -shared \
    -o tools/libclustassess_b200_ablate.so clustassess_b200/csrc/*.cu clustassess_b200/csrc/ca_plan.cpp -lpthread
echo built tools/libclustassess_b200_ablate.so
