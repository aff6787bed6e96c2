#!/bin/bash
# one GPU: the default bench line with the launch thread / pinned frame bound to the GPU's NUMA node and without (e2e leg), plus the
# parity tests of the search path on the rebuilt library
set +e
mkdir -p gpurun_out
T0=$SECONDS
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
cat /sys/devices/system/node/node*/cpulist >> gpurun_out/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_edge.py -m gpu -x -q > gpurun_out/t_bind_parity.log 2>&1; echo "parity rc=$? t=$((SECONDS-T0))"; tail -2 gpurun_out/t_bind_parity.log
for i in 1 2; do
  timeout 300 python bench.py > gpurun_out/bench_bind_$i.json 2> gpurun_out/bench_bind_$i.err; echo "bench bind $i rc=$? t=$((SECONDS-T0))"
  timeout 300 python bench.py --no-bind --no-cpu > gpurun_out/bench_nobind_$i.json 2> gpurun_out/bench_nobind_$i.err; echo "bench nobind $i rc=$? t=$((SECONDS-T0))"
done
for f in bind_1 nobind_1 bind_2 nobind_2; do
  python -c "import json; d=json.load(open('gpurun_out/bench_$f.json')); e=d['e2e']; print('$f', d['ms_per_step'], e['ms_per_step'], e['kernel_ms'], e['plane_kernel_ms'], e['ms_per_step_staged_copy'], e['host_binding'], (d.get('cpu_baseline') or {}).get('value'))"
done
