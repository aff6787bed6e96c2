#!/bin/bash
# round 2, session 3 (one GPU): e2e parity test, small-frame latency of the search variants, compaction soak
set +e
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests/test_gpu_e2e_parity.py -m gpu -x -q -s > gpurun_out/t_e2e.log 2>&1; echo "e2e rc=$? t=$((SECONDS-T0))"
grep -E "e2e parity|passed|failed|Error" gpurun_out/t_e2e.log | cut -c1-1500
timeout 600 python tools/probe_small.py 2>&1 | grep -v "^gen" | tee gpurun_out/probe_small.log
echo "probe small t=$((SECONDS-T0))"
