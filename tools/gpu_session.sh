#!/bin/bash
# One gpurun call that covers a development cycle (about 2.5 minutes of box time on one B200):
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/gpu_session.sh'
# GPU suite with the cells index forced (the default index is covered by the plain suite), same-box A/B of the search variants
# (tools/quick_ab.py), bench lines + ncu launch lists + ncu --set full captures for BOTH indexes, sanitizer on the smoke.
# Everything lands in gpurun_out/; tools/ncu_metrics.py and tools/ncu_lines.py turn the .ncu-rep files into the text kept under profiles/.
set +e
mkdir -p gpurun_out
T0=$SECONDS
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_default.log 2>&1; echo "default suite rc=$? t=$((SECONDS-T0))"
tail -2 gpurun_out/t_default.log
LIINIT_KNN_INDEX=2 timeout 600 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/t_cells.log 2>&1; echo "cells suite rc=$? t=$((SECONDS-T0))"
tail -2 gpurun_out/t_cells.log
LIINIT_KNN_INDEX=2 LIINIT_CELLS_REFRESH=thread timeout 300 python -m pytest tests/test_gpu_cells.py tests/test_gpu_parity.py -x -q > gpurun_out/t_cells_threadrefresh.log 2>&1; echo "cells thread-refresh rc=$? t=$((SECONDS-T0))"
tail -1 gpurun_out/t_cells_threadrefresh.log
for s in 2 1; do LIINIT_KNN_INDEX=2 LIINIT_CELLS_SEARCH=$s timeout 300 python -m pytest tests/test_gpu_cells.py -x -q > gpurun_out/t_cells_search$s.log 2>&1; echo "cells search=$s rc=$? t=$((SECONDS-T0))"; done
timeout 300 python tools/quick_ab.py > gpurun_out/ab.log 2>&1; echo "ab rc=$? t=$((SECONDS-T0))"
grep -v "^gen" gpurun_out/ab.log
# compile-time variants: tools/build_variants.sh (run before gpurun) builds them into build/variants/; each is probed here when present
for lib in build/variants/*.so; do [ -f "$lib" ] && LIINIT_GPU_LIB=$lib timeout 100 python tools/probe_variant.py 2>/dev/null | tee -a gpurun_out/probe_variants.log; done
timeout 100 python tools/probe_variant.py 2>/dev/null | tee -a gpurun_out/probe_variants.log
echo "variant probes t=$((SECONDS-T0))"
timeout 300 python bench.py --knn-index 2 > gpurun_out/bench_cells.json 2> gpurun_out/bench_cells.err; echo "bench cells rc=$? t=$((SECONDS-T0))"
cut -c1-400 gpurun_out/bench_cells.json
timeout 300 python bench.py --knn-index 1 > gpurun_out/bench_bricks.json 2> gpurun_out/bench_bricks.err; echo "bench bricks rc=$? t=$((SECONDS-T0))"
cut -c1-400 gpurun_out/bench_bricks.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cells.csv python bench.py --knn-index 2 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches_cells.log 2>&1; echo "ncu launches cells rc=$? t=$((SECONDS-T0))"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bricks.csv python bench.py --knn-index 1 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches_bricks.log 2>&1; echo "ncu launches bricks rc=$? t=$((SECONDS-T0))"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"k_knn_cells_scan|k_icp_plane" -s 4 -c 3 -f -o gpurun_out/cells_full python tools/prof_run.py --index 2 > gpurun_out/ncu_full_cells.log 2>&1; echo "ncu full cells rc=$? t=$((SECONDS-T0))"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"k_knn_scan|k_icp_plane" -s 4 -c 3 -f -o gpurun_out/bricks_full python tools/prof_run.py --index 1 > gpurun_out/ncu_full_bricks.log 2>&1; echo "ncu full bricks rc=$? t=$((SECONDS-T0))"
LIINIT_KNN_INDEX=2 timeout 200 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_cells_memcheck.log 2>&1; echo "memcheck rc=$? t=$((SECONDS-T0))"
tail -2 gpurun_out/sanitizer_cells_memcheck.log
LIINIT_KNN_INDEX=2 timeout 240 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_cells_racecheck.log 2>&1; echo "racecheck rc=$? t=$((SECONDS-T0))"
tail -2 gpurun_out/sanitizer_cells_racecheck.log
