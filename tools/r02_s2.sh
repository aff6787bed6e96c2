set +e
mkdir -p gpurun_out/r02
export LIINIT_GPU_LIB=build/dev/libliinit_gpu.so
T0=$SECONDS
timeout 300 python tools/probe_fused.py > gpurun_out/r02/probe_fused_v1.log 2>&1; echo "probe rc=$? t=$((SECONDS-T0))"
grep -v "^gen" gpurun_out/r02/probe_fused_v1.log
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"k_icp_fused" -s 3 -c 1 -f -o gpurun_out/r02/fused_v1 python tools/prof_fused.py 4:4:3 > gpurun_out/r02/ncu_fused_v1.log 2>&1; echo "ncu rc=$? t=$((SECONDS-T0))"
