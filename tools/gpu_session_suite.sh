#!/bin/bash
# the -m gpu suite (default index), slowest tests listed
set +e
mkdir -p gpurun_out
T0=$SECONDS
timeout 2400 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/t_default.log 2>&1; echo "default suite rc=$? t=$((SECONDS-T0))"
tail -25 gpurun_out/t_default.log
