#!/bin/bash
# last GPU minutes of the round: the final commit's smoke, both indexes pinned explicitly, the default bench once more
set +e
mkdir -p gpurun_out
T0=$SECONDS
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/d_smoke.log 2>&1; echo "smoke rc=$? t=$((SECONDS-T0))"; tail -1 gpurun_out/d_smoke.log
timeout 100 python -m pytest tests/test_gpu_cells.py -x -q > gpurun_out/d_t_cells.log 2>&1; echo "test_gpu_cells rc=$? t=$((SECONDS-T0))"; tail -1 gpurun_out/d_t_cells.log
timeout 100 python bench.py --no-cpu > gpurun_out/d_bench_default.json 2> gpurun_out/d_bench_default.err; echo "bench rc=$? t=$((SECONDS-T0))"; cut -c1-260 gpurun_out/d_bench_default.json
