#!/bin/bash
# round 2, session 5 (one GPU): compile-time variants of the lockstep search (probe-all-then-stream shape, occupancy), e2e parity test
set +e
mkdir -p gpurun_out
T0=$SECONDS
: > gpurun_out/probe_variants.log
timeout 100 python tools/probe_knn.py --variants 1:4:3:0,1:8:3:0 2>&1 | grep -v "^gen" | tee -a gpurun_out/probe_variants.log
for lib in build/variants/lib_*.so; do LIINIT_GPU_LIB=$lib timeout 100 python tools/probe_knn.py --variants 1:4:3:0,1:8:3:0 2>&1 | grep -v "^gen" | tee -a gpurun_out/probe_variants.log; done
echo "probe variants t=$((SECONDS-T0))"
for lib in build/variants/lib_v2_pb2.so build/variants/lib_v2_pb4.so; do LIINIT_GPU_LIB=$lib timeout 200 python tools/probe_small.py --sizes 2000,30000,130000 --variants 1:0 2>&1 | grep -v "^gen" | sed "s|^|$(basename $lib) |" | tee -a gpurun_out/probe_variants.log; done
echo "probe small t=$((SECONDS-T0))"
LIINIT_GPU_LIB=build/variants/lib_v2_pb2.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/t_v2.log 2>&1; echo "v2 parity tests rc=$? t=$((SECONDS-T0))"
tail -3 gpurun_out/t_v2.log
timeout 900 python -m pytest tests/test_gpu_e2e_parity.py -m gpu -x -q -s > gpurun_out/t_e2e.log 2>&1; echo "e2e rc=$? t=$((SECONDS-T0))"
grep -E "e2e parity|passed|failed|Error" gpurun_out/t_e2e.log | cut -c1-1800
