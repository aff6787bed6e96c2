#!/bin/bash
# closing check of the tree as committed (one GPU): whole -m gpu suite, smoke, the default bench line and the reference arm
set +e
mkdir -p gpurun_out
T0=$SECONDS
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/t_close.log 2>&1; echo "suite rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/t_close.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_close.log 2>&1; echo "smoke rc=$? t=$((SECONDS-T0))"; tail -1 gpurun_out/smoke_close.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_close.json 2> gpurun_out/bench_close.err; echo "bench rc=$? t=$((SECONDS-T0))"
python -c "import json; d=json.load(open('gpurun_out/bench_close.json')); print(d['ms_per_step'], d['details']['step_ms'], d['e2e']['ms_per_step'], d['e2e']['kernel_ms'], d['e2e']['host_binding'], d['roofline']['frac'], d.get('parity'), d['cpu_baseline']['value'])"
timeout 300 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/bench_close_ref.json 2> gpurun_out/bench_close_ref.err; echo "ref rc=$? t=$((SECONDS-T0))"; cut -c1-300 gpurun_out/bench_close_ref.json
