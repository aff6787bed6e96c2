#!/bin/bash
# eight GPUs of one box: weak-scaled C2 and strong-scaled C5 at N = 4 and 8 (sum over the ranks fused in the plane kernel, peer memory)
set +e
mkdir -p gpurun_out
T0=$SECONDS
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
for n in 4 8; do
  for cfg in C2 C5; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --config $cfg > gpurun_out/bench_${cfg}_n${n}_p2p.json 2> gpurun_out/bench_${cfg}_n${n}_p2p.err; echo "bench $cfg N=$n rc=$? t=$((SECONDS-T0))"
    python -c "import json; d=json.load(open('gpurun_out/bench_${cfg}_n${n}_p2p.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['multi_gpu_check'])"
  done
done
LIINIT_COMM_MODE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29619 bench.py --gpus 8 > gpurun_out/bench_C2_n8_nccl.json 2> gpurun_out/bench_C2_n8_nccl.err; echo "bench C2 N=8 nccl rc=$? t=$((SECONDS-T0))"
python -c "import json; d=json.load(open('gpurun_out/bench_C2_n8_nccl.json')); print(d['value'], d['ms_per_step'], d['multi_gpu_check']['collective'])"
