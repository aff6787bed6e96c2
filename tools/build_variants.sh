#!/bin/bash
# Compile-time variants of the library for a same-box A/B (run HERE before gpurun: nvcc cross-compiles without a GPU; build/ travels with the snapshot).
#   LIINIT_GPU_LIB=build/variants/lib_plane_prefetch.so python tools/probe_variant.py      (brick search + plane pass timings)
#   LIINIT_GPU_LIB=build/variants/lib_plane_prefetch.so python -m pytest tests -m gpu -x -q
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -Xcompiler -fPIC -shared"
nvcc $FLAGS -DLI_PLANE_PREFETCH=1 -o build/variants/lib_plane_prefetch.so lidar_imu_init_b200/csrc/liinit_gpu.cu &
nvcc $FLAGS -DLI_GROUP_BOUND=1 -o build/variants/lib_gb.so lidar_imu_init_b200/csrc/liinit_gpu.cu &
wait
ls -la build/variants
