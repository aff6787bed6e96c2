#!/bin/bash
# round 2, session 1: first measurement of the warp-per-scan-point search (knn_index 5) against the lockstep brick search
set +e
mkdir -p gpurun_out
T0=$SECONDS
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv > gpurun_out/gpu.txt 2>&1
timeout 200 python tools/probe_knn.py --variants 1:4:3:0,5:0:3:0,5:0:3:19,5:0:3:18,1:4:3:19 2>&1 | grep -v "^gen" | tee gpurun_out/probe_wq1.log
echo "probe default t=$((SECONDS-T0))"
for lib in build/variants/lib_wq_*.so; do LIINIT_GPU_LIB=$lib timeout 100 python tools/probe_knn.py --variants 5:0:3:19 --check 0 2>&1 | grep -v "^gen" | tee -a gpurun_out/probe_wq1.log; done
echo "probe variants t=$((SECONDS-T0))"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"k_knn_wq" -s 2 -c 1 -f -o gpurun_out/wq_full python tools/prof_run.py --index 5 > gpurun_out/ncu_full_wq.log 2>&1; echo "ncu full wq rc=$? t=$((SECONDS-T0))"
LIINIT_KNN_INDEX=5 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_wq.log 2>&1; echo "wq suite rc=$? t=$((SECONDS-T0))"
tail -3 gpurun_out/t_wq.log
LIINIT_KNN_INDEX=5 timeout 200 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_wq_racecheck.log 2>&1; echo "racecheck rc=$? t=$((SECONDS-T0))"
tail -2 gpurun_out/sanitizer_wq_racecheck.log
