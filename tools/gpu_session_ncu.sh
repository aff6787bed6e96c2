#!/bin/bash
# ncu --set full of the two kernels of a first search pass + the reuse pass (C2), with source lines
set +e
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_knn_scan|k_icp_plane" -s 4 -c 3 -f -o gpurun_out/pass_full python tools/prof_run.py --index 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/pass_full.ncu-rep --page raw --csv 2>/dev/null | cut -d, -f5 | head -5
