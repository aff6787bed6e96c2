#!/bin/bash
# round 2, session 4 (one GPU): e2e parity (dead-reckoning LIO leg), automatic lane groups, seed-radius sweep, G=2 at large frames, bench lines
set +e
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests/test_gpu_e2e_parity.py tests/test_gpu_parity.py -m gpu -x -q -s > gpurun_out/t_e2e.log 2>&1; echo "e2e+parity rc=$? t=$((SECONDS-T0))"
grep -E "e2e parity|passed|failed|Error" gpurun_out/t_e2e.log | cut -c1-1800
timeout 300 python tools/probe_small.py --sizes 2000,30000,130000,240000 --variants 1:0,1:2,1:4,1:8 2>&1 | grep -v "^gen" | tee gpurun_out/probe_auto.log
echo "probe auto t=$((SECONDS-T0))"
for rho in 1 2 3 4 6; do timeout 100 python - <<PY 2>&1 | grep -v "^gen" | tee -a gpurun_out/probe_rho.log
import sys; sys.path.insert(0, '.')
import numpy as np
from lidar_imu_init_b200 import capi
z = np.load('/tmp/C2_probe_240000_5000000.npz'); I, zero = np.eye(3), np.zeros(3)
g = capi.LiInitGpu(0.15, max_map_points=6_000_000, max_scan_points=250_000, knn_seed_radius_cells=float($rho))
g.map_build(z["map"]); g.scan_upload(z["body"])
out = []
for pose in ("init", "gt"):
    ks = []
    for it in range(12):
        g.icp_iterate(z[pose + "_R"], z[pose + "_p"], I, zero, False, True); ks.append(g.last_pass_kernel_times())
    out.append(f"{pose} knn {np.median(np.array(ks)[3:, 0]):.4f}")
print("seed radius cells $rho |", " | ".join(out))
PY
done
echo "probe rho t=$((SECONDS-T0))"
timeout 400 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench C2 rc=$? t=$((SECONDS-T0))"
cut -c1-300 gpurun_out/bench_c2.json
timeout 400 python bench.py --config C3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench C3 rc=$? t=$((SECONDS-T0))"
cut -c1-300 gpurun_out/bench_c3.json
