#!/bin/bash
# two GPUs of one box: the multi-GPU path behind the C-ABI -- hardware parity tests (Python-launched ranks in both exchange modes, and the
# C++-only two-rank drop-in), weak-scaled C2, strong-scaled C5
set +e
mkdir -p gpurun_out
T0=$SECONDS
nvidia-smi --query-gpu=name --format=csv,noheader | head -4
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_dropin_cpp.py -m gpu -q > gpurun_out/t_multi.log 2>&1; echo "multi tests rc=$? t=$((SECONDS-T0))"
tail -5 gpurun_out/t_multi.log
if [ -z "$LIINIT_SESSION_QUICK" ]; then
for cfg in C2 C5; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --config $cfg > gpurun_out/bench_${cfg}_n2_p2p.json 2> gpurun_out/bench_${cfg}_n2_p2p.err; echo "bench $cfg N=2 rc=$? t=$((SECONDS-T0))"
  python -c "import json; d=json.load(open('gpurun_out/bench_${cfg}_n2_p2p.json')); print(d['ms_per_step'], d['e2e'], d['multi_gpu_check'])"
done
fi
