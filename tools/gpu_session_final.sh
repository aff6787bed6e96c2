#!/bin/bash
# round 2, closing session (one GPU): whole -m gpu suite, bench lines of every configuration + reference arm, ncu launch list and
# ncu --set full capture of the two kernels of a pass, compute-sanitizer on the smoke.  LIINIT_SESSION_QUICK=1: suite + C2 + ncu only
set +e
mkdir -p gpurun_out
T0=$SECONDS
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/t_default.log 2>&1; echo "default suite rc=$? t=$((SECONDS-T0))"
tail -4 gpurun_out/t_default.log
timeout 400 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench C2 rc=$? t=$((SECONDS-T0))"
cut -c1-260 gpurun_out/bench_c2.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_c2_ref.json 2> gpurun_out/bench_c2_ref.err; echo "bench ref rc=$? t=$((SECONDS-T0))"
if [ -z "$LIINIT_SESSION_QUICK" ]; then
timeout 400 python bench.py --config C3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench C3 rc=$? t=$((SECONDS-T0))"
timeout 400 python bench.py --config C5 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench C5 rc=$? t=$((SECONDS-T0))"
timeout 900 python bench.py --config C4 --no-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench C4 rc=$? t=$((SECONDS-T0))"
fi
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$? t=$((SECONDS-T0))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_knn_scan|k_icp_plane" -s 4 -c 3 -f -o gpurun_out/pass_full python tools/prof_run.py --index 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$? t=$((SECONDS-T0))"
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$? t=$((SECONDS-T0))"
tail -2 gpurun_out/sanitizer_memcheck.log
timeout 300 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$? t=$((SECONDS-T0))"
tail -2 gpurun_out/sanitizer_racecheck.log
