"""Build the in-tree native libraries (nvcc cross-compiles sm_100a without a GPU).

  libliinit_gpu.so   CUDA kernels + C-ABI (include/liinit_gpu.h)       -- nvcc
  libliinit_host.so  host-side IESKF / per-scan driver over the C-ABI   -- g++ (see csrc/host)
  libliinit_calib.so LI-Init batch initialisation, host only (include/liinit_calib.h) -- g++ (see csrc/calib)
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
GPU_LIB = os.path.join(HERE, "libliinit_gpu.so")
HOST_LIB = os.path.join(HERE, "libliinit_host.so")
CALIB_LIB = os.path.join(HERE, "libliinit_calib.so")

# --fmad=false: the fp32 distances and the fp64 plane / Jacobian arithmetic follow the reference's x86-64
# build (-O3, no -march => no FMA contraction, CMakeLists.txt:8), so f32 outputs are bit-comparable with the oracle.
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--fmad=false",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    n = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(n):
        raise RuntimeError("nvcc not found")
    return n


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def gpu_sources():
    inc = os.path.join(HERE, "..", "include", "liinit_gpu.h")
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + [inc]


def build_gpu(force: bool = False, verbose: bool = False) -> str:
    srcs = gpu_sources()
    if force or _newer(GPU_LIB, srcs):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", GPU_LIB, os.path.join(CSRC, "liinit_gpu.cu"), "-ldl"]
        subprocess.check_call(cmd)
    return GPU_LIB


def host_sources():
    d = os.path.join(CSRC, "host")
    if not os.path.isdir(d):
        return []
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith((".cpp", ".hpp", ".h"))]


def build_host(force: bool = False) -> str | None:
    srcs = host_sources()
    cpps = [s for s in srcs if s.endswith(".cpp")]
    if not cpps:
        return None
    if force or _newer(HOST_LIB, srcs):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        cmd = [gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(HERE, "..", "include"), "-o", HOST_LIB] + cpps + \
              ["-L", HERE, "-lliinit_gpu", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    return HOST_LIB


def build_calib(force: bool = False) -> str:
    src = os.path.join(CSRC, "calib", "li_calib.cpp")
    inc = os.path.join(HERE, "..", "include")
    if force or _newer(CALIB_LIB, [src, os.path.join(inc, "liinit_calib.h")]):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        # no FMA contraction: the filters follow the reference's x86-64 arithmetic operation by operation
        subprocess.check_call([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I", inc, "-o", CALIB_LIB, src])
    return CALIB_LIB


def build_all(force: bool = False):
    build_gpu(force)
    build_host(force)
    build_calib(force)


if __name__ == "__main__":
    build_all(force=True)
    print("built", GPU_LIB)
