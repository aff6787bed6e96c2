#!/bin/bash
# same-box A/B of compile-time variants under build/variants against the in-tree library
set +e
mkdir -p gpurun_out
: > gpurun_out/probe_variants.log
timeout 100 python tools/probe_knn.py --variants 1:4:3:0,1:8:3:0 2>&1 | grep -v "^gen" | tee -a gpurun_out/probe_variants.log
for lib in build/variants/lib_*.so; do LIINIT_GPU_LIB=$lib timeout 100 python tools/probe_knn.py --variants 1:4:3:0,1:8:3:0 --check 0 2>&1 | grep -v "^gen" | tee -a gpurun_out/probe_variants.log; done
for lib in $SMALL_LIBS; do LIINIT_GPU_LIB=$lib timeout 300 python tools/probe_small.py --sizes 8000,20000,30000,45000,60000,90000,130000,170000 --variants 1:4,1:8,1:16,1:32 2>&1 | grep -v "^gen" | sed "s|^|$(basename $lib) |" | tee -a gpurun_out/probe_variants.log; done
