#!/bin/bash
# same-box A/B of compile-time variants under build/variants against the in-tree library: bench.py lines (L2 flushed between steps) + warm probes
set +e
mkdir -p gpurun_out
: > gpurun_out/probe_variants.log
timeout 100 python tools/probe_knn.py --variants 1:4:3:0 2>&1 | grep -v "^gen" | tee -a gpurun_out/probe_variants.log
timeout 200 python bench.py --no-cpu --steps 30 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default bench ms', d['ms_per_step'], 'knn', d['roofline']['kernel_ms'], 'plane', d['roofline']['plane_kernel_ms'])" | tee -a gpurun_out/probe_variants.log
for lib in build/variants/lib_*.so; do
  LIINIT_GPU_LIB=$lib timeout 100 python tools/probe_knn.py --variants 1:4:3:0 --check 0 2>&1 | grep -v "^gen" | tee -a gpurun_out/probe_variants.log
  LIINIT_GPU_LIB=$lib timeout 200 python bench.py --no-cpu --steps 30 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$(basename $lib) bench ms', d['ms_per_step'], 'knn', d['roofline']['kernel_ms'], 'plane', d['roofline']['plane_kernel_ms'])" | tee -a gpurun_out/probe_variants.log
done
