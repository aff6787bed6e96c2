#!/bin/bash
# two GPUs of one box: the multi-GPU path behind the C-ABI -- hardware parity test (peer-memory and NCCL modes), weak-scaled C2, strong-scaled C5
set +e
mkdir -p gpurun_out
T0=$SECONDS
nvidia-smi --query-gpu=name --format=csv,noheader | head -4
nvidia-smi topo -m 2>/dev/null | head -6
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/t_multi.log 2>&1; echo "multi test rc=$? t=$((SECONDS-T0))"
tail -5 gpurun_out/t_multi.log
for mode in p2p nccl; do
  [ $mode = nccl ] && export LIINIT_COMM_MODE=nccl || unset LIINIT_COMM_MODE
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 > gpurun_out/bench_c2_n2_$mode.json 2> gpurun_out/bench_c2_n2_$mode.err; echo "bench C2 N=2 $mode rc=$? t=$((SECONDS-T0))"
  python -c "import json; d=json.load(open('gpurun_out/bench_c2_n2_$mode.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['multi_gpu_check'])"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --config C5 > gpurun_out/bench_c5_n2_$mode.json 2> gpurun_out/bench_c5_n2_$mode.err; echo "bench C5 N=2 $mode rc=$? t=$((SECONDS-T0))"
  python -c "import json; d=json.load(open('gpurun_out/bench_c5_n2_$mode.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['multi_gpu_check'])"
done
tail -3 gpurun_out/bench_c2_n2_p2p.err
