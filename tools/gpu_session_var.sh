#!/bin/bash
# same-box A/B of compile-time variants under build/variants against the in-tree library
set +e
mkdir -p gpurun_out
: > gpurun_out/probe_variants.log
timeout 100 python tools/probe_knn.py --variants 1:4:3:0,1:8:3:0 2>&1 | grep -v "^gen" | tee -a gpurun_out/probe_variants.log
for lib in build/variants/lib_*.so; do LIINIT_GPU_LIB=$lib timeout 100 python tools/probe_knn.py --variants 1:4:3:0,1:8:3:0 2>&1 | grep -v "^gen" | tee -a gpurun_out/probe_variants.log; done
for lib in build/variants/lib_*.so; do LIINIT_GPU_LIB=$lib timeout 200 python tools/probe_small.py --sizes 2000,30000,130000 --variants 1:0 2>&1 | grep -v "^gen" | sed "s|^|$(basename $lib) |" | tee -a gpurun_out/probe_variants.log; done
timeout 200 python tools/probe_small.py --sizes 2000,30000,130000 --variants 1:0 2>&1 | grep -v "^gen" | sed "s|^|default |" | tee -a gpurun_out/probe_variants.log
