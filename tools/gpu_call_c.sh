#!/bin/bash
# GPU session C (last of the round, ~3.5 minutes of run time available): default configuration first, then the cells experiment.
set +e
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/c_t_default.log 2>&1; echo "default suite rc=$? t=$((SECONDS-T0))"
tail -2 gpurun_out/c_t_default.log
timeout 200 python bench.py > gpurun_out/c_bench_default.json 2> gpurun_out/c_bench_default.err; echo "bench default rc=$? t=$((SECONDS-T0))"
cut -c1-330 gpurun_out/c_bench_default.json
LIINIT_KNN_INDEX=2 timeout 120 python -m pytest tests/test_gpu_cells.py tests/test_gpu_parity.py -x -q > gpurun_out/c_t_cells.log 2>&1; echo "cells tests rc=$? t=$((SECONDS-T0))"
tail -1 gpurun_out/c_t_cells.log
timeout 150 python tools/quick_ab.py --variants 1:0:2:0,2:6:2:3,2:4:2:3,2:8:2:3 > gpurun_out/c_ab.log 2>&1; echo "ab rc=$? t=$((SECONDS-T0))"
grep -v "^gen" gpurun_out/c_ab.log
timeout 120 ncu --set full --clock-control none --import-source on -k regex:"k_knn_cells_scan" -s 2 -c 1 -f -o gpurun_out/c_cells_full python tools/prof_run.py --index 2 --passes 3 > gpurun_out/c_ncu_full_cells.log 2>&1; echo "ncu full cells rc=$? t=$((SECONDS-T0))"
LIINIT_GPU_LIB=build/variants/libliinit_gpu_u2.so timeout 100 python tools/quick_ab.py --variants 2:6:2:3,2:8:2:3 > gpurun_out/c_ab_u2.log 2>&1; echo "ab u2 rc=$? t=$((SECONDS-T0))"
grep -v "^gen" gpurun_out/c_ab_u2.log
timeout 100 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/c_bench_reference.json 2> gpurun_out/c_bench_reference.err; echo "reference arm rc=$? t=$((SECONDS-T0))"
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/c_ncu_launches.log 2>&1; echo "ncu launches default rc=$? t=$((SECONDS-T0))"
