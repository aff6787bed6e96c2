#!/bin/bash
# round 2, session 2 (one GPU): suite with the new config tests, near_xyz plane pass, bench lines for C2 / C3 / C5 + reference arm
set +e
mkdir -p gpurun_out
T0=$SECONDS
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/t_default.log 2>&1; echo "default suite rc=$? t=$((SECONDS-T0))"
tail -14 gpurun_out/t_default.log
timeout 200 python tools/probe_knn.py --variants 1:4:3:0,5:0:3:0 2>&1 | grep -v "^gen" | tee gpurun_out/probe_r2b.log
echo "probe t=$((SECONDS-T0))"
timeout 400 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench C2 rc=$? t=$((SECONDS-T0))"
cut -c1-1500 gpurun_out/bench_c2.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_c2_ref.json 2> gpurun_out/bench_c2_ref.err; echo "bench ref rc=$? t=$((SECONDS-T0))"
cut -c1-900 gpurun_out/bench_c2_ref.json
timeout 400 python bench.py --config C3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench C3 rc=$? t=$((SECONDS-T0))"
cut -c1-600 gpurun_out/bench_c3.json
timeout 400 python bench.py --config C5 --no-cpu > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err; echo "bench C5 rc=$? t=$((SECONDS-T0))"
cut -c1-600 gpurun_out/bench_c5_n1.json
