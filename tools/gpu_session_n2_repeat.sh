#!/bin/bash
# two GPUs: the default C2 bench line repeated (run-to-run spread of the weak-scaled step, per-step min / median / max), both exchange modes
set +e
mkdir -p gpurun_out
T0=$SECONDS
for i in 1 2 3 4; do
  mode=p2p; [ $i = 4 ] && mode=nccl
  LIINIT_COMM_MODE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29620+i)) bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_C2_n2_rep$i.json 2> gpurun_out/bench_C2_n2_rep$i.err
  echo "rep $i ($mode) rc=$? t=$((SECONDS-T0))"
  python -c "import json; d=json.load(open('gpurun_out/bench_C2_n2_rep$i.json')); print(d['ms_per_step'], d['details']['step_ms'], 1e3/d['iters_per_s_l2_warm'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['plane_kernel_ms'])"
done
