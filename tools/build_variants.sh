#!/bin/bash
# Compile-time variants of the library for a same-box A/B (run HERE before gpurun: nvcc cross-compiles without a GPU; build/ travels with the snapshot).
#   LIINIT_GPU_LIB=build/variants/lib_x.so python tools/probe_knn.py --variants 1:0:3:0
# usage: tools/build_variants.sh name1="-DFLAG=1 -DX=2" name2="..."
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -Xcompiler -fPIC -shared"
for spec in "$@"; do
  name="${spec%%=*}"; defs="${spec#*=}"
  nvcc $FLAGS $defs -o build/variants/lib_$name.so lidar_imu_init_b200/csrc/liinit_gpu.cu &
done
wait
ls -la build/variants
