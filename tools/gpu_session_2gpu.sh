#!/bin/bash
# two GPUs of one box: the multi-GPU path behind the C-ABI -- hardware parity test, weak-scaled C2, strong-scaled C5
set +e
mkdir -p gpurun_out
T0=$SECONDS
nvidia-smi --query-gpu=name --format=csv,noheader | head -4
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/t_multi.log 2>&1; echo "multi test rc=$? t=$((SECONDS-T0))"
tail -5 gpurun_out/t_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; echo "bench C2 N=2 rc=$? t=$((SECONDS-T0))"
cut -c1-400 gpurun_out/bench_c2_n2.json; python -c "import json; d=json.load(open('gpurun_out/bench_c2_n2.json')); print(d['multi_gpu_check'], d['details'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --config C5 > gpurun_out/bench_c5_n2.json 2> gpurun_out/bench_c5_n2.err; echo "bench C5 N=2 rc=$? t=$((SECONDS-T0))"
cut -c1-400 gpurun_out/bench_c5_n2.json; python -c "import json; d=json.load(open('gpurun_out/bench_c5_n2.json')); print(d['multi_gpu_check'], d['details'])"
tail -3 gpurun_out/bench_c5_n2.err
