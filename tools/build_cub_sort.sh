#!/bin/sh
# builds tools/build/libcubsort.so (git-ignored, shipped to the GPU box by gpurun); see tools/cub_sort.cu
set -e
cd "$(dirname "$0")"
mkdir -p build
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC cub_sort.cu -o build/libcubsort.so
echo built tools/build/libcubsort.so
