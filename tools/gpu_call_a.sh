#!/bin/bash
# One-shot GPU session: cells index through the whole GPU suite, same-box A/B against the brick search, bench line,
# ncu launch list + full capture, sanitizer on the smoke. Everything lands in gpurun_out/.
set +e
mkdir -p gpurun_out
T0=$SECONDS
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv > gpurun_out/gpu.txt 2>&1
LIINIT_KNN_INDEX=2 timeout 600 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/t_cells.log 2>&1; echo "cells suite rc=$? t=$((SECONDS-T0))"
tail -3 gpurun_out/t_cells.log
timeout 240 python tools/quick_ab.py > gpurun_out/ab.log 2>&1; echo "ab rc=$? t=$((SECONDS-T0))"
cat gpurun_out/ab.log | grep -v "^gen"
timeout 300 python bench.py --knn-index 2 > gpurun_out/bench_cells.json 2> gpurun_out/bench_cells.err; echo "bench rc=$? t=$((SECONDS-T0))"
cut -c1-700 gpurun_out/bench_cells.json
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cells.csv python bench.py --knn-index 2 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$? t=$((SECONDS-T0))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_knn_cells_scan|k_icp_plane" -s 4 -c 3 -f -o gpurun_out/cells_full python tools/prof_run.py --index 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$? t=$((SECONDS-T0))"
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cells.py -x -q > gpurun_out/t_default.log 2>&1; echo "default-index parity rc=$? t=$((SECONDS-T0))"
tail -2 gpurun_out/t_default.log
LIINIT_GPU_LIB=build/variants/libliinit_gpu_t64.so timeout 120 python tools/quick_ab.py --variants 2:6:2,2:8:2 > gpurun_out/ab_t64.log 2>&1; echo "ab t64 rc=$? t=$((SECONDS-T0))"
grep -v "^gen" gpurun_out/ab_t64.log
LIINIT_KNN_INDEX=2 timeout 240 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_cells.log 2>&1; echo "sanitizer rc=$? t=$((SECONDS-T0))"
tail -4 gpurun_out/sanitizer_cells.log
