set +e
mkdir -p gpurun_out/r02
T0=$SECONDS
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv > gpurun_out/r02/gpu.txt 2>&1
timeout 300 python tools/quick_ab.py --variants "1:0:2:0,3:6:2:3,2:6:2:3:d,2:6:2:3" > gpurun_out/r02/ab_existing.log 2>&1; echo "ab rc=$? t=$((SECONDS-T0))"
grep -v "^gen" gpurun_out/r02/ab_existing.log
for lib in build/variants/*.so; do [ -f "$lib" ] && LIINIT_GPU_LIB=$lib timeout 100 python tools/probe_variant.py 2>/dev/null | tee -a gpurun_out/r02/probe_variants.log; done
timeout 100 python tools/probe_variant.py 2>/dev/null | tee -a gpurun_out/r02/probe_variants.log
echo "t=$((SECONDS-T0))"
