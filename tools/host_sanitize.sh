#!/bin/bash
# AddressSanitizer + UndefinedBehaviorSanitizer over the host-side libraries (csrc/host/liinit_host.cpp, csrc/calib/li_calib.cpp): builds
# them sanitized IN PLACE, runs their CPU tests with libasan preloaded, restores the regular builds. (The kernels and the C-ABI layer
# have their own driver: tools/emul_asan.py.)
set -e
cd "$(dirname "$0")/.."
ASAN=$(gcc -print-file-name=libasan.so)
P=lidar_imu_init_b200
SAN="-O1 -g -std=c++17 -fPIC -shared -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer -I include"
cp $P/libliinit_calib.so /tmp/libliinit_calib_bak.so; cp $P/libliinit_host.so /tmp/libliinit_host_bak.so
trap 'cp /tmp/libliinit_calib_bak.so $P/libliinit_calib.so; cp /tmp/libliinit_host_bak.so $P/libliinit_host.so' EXIT
g++ $SAN -ffp-contract=off -o $P/libliinit_calib.so $P/csrc/calib/li_calib.cpp
g++ $SAN -o $P/libliinit_host.so $P/csrc/host/liinit_host.cpp -L $P -lliinit_gpu -Wl,-rpath,'$ORIGIN'
LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 python -m pytest tests/test_calib.py tests/test_host.py -x -q
