"""ctypes binding of the CPU checker of the cell-directory 5-NN (tests/emul/cells_emul.cpp). Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "cells_emul.cpp")
LIB = os.path.join(HERE, "emul", "libcells_emul.so")
CSRC = os.path.join(HERE, "..", "lidar_imu_init_b200", "csrc")
CUDA_INC = "/usr/local/cuda/include"

_f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def available() -> bool:
    return shutil.which("g++") is not None and os.path.exists(os.path.join(CUDA_INC, "vector_types.h"))


def build(force: bool = False) -> str:
    deps = [SRC, os.path.join(CSRC, "cells.cuh"), os.path.join(CSRC, "common.cuh")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        # -ffp-contract=off: the library is built with --fmad=false, the checker must not fuse either
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-I", CUDA_INC, SRC, "-o", LIB])
    return LIB


_L = None


def load():
    global _L
    if _L is None:
        L = C.CDLL(build())
        L.emul_create.restype = C.c_void_p
        L.emul_create.argtypes = [_f32, C.c_int, C.c_float, C.c_int]
        L.emul_add.argtypes = [C.c_void_p, _f32, C.c_int]
        L.emul_destroy.argtypes = [C.c_void_p]
        L.emul_num_bricks.argtypes = [C.c_void_p]
        L.emul_check_directory.argtypes = [C.c_void_p]
        L.emul_knn.argtypes = [C.c_void_p, _f32, C.c_int, C.c_float, C.c_int, _f32, _f32, _i32, C.c_void_p]
        L.emul_trace.restype = C.c_longlong
        L.emul_trace.argtypes = [C.c_void_p, _f32, C.c_int, C.c_float, _i32, C.c_longlong, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")]
        _L = L
    return _L


STAT_NAMES = ("rounds", "supers", "probes", "found", "cells", "cells_scanned", "points", "inserts")


class CellsEmul:
    def __init__(self, map_xyz, ds=0.15, hash_log2=None):
        self.L = load()
        xyz = np.ascontiguousarray(map_xyz, np.float32).reshape(-1, 3)
        if hash_log2 is None:
            hash_log2 = 12
            while (1 << hash_log2) < max(len(xyz) // 2, 1024):
                hash_log2 += 1
        self.h = C.c_void_p(self.L.emul_create(xyz, len(xyz), ds, hash_log2))

    def add(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        self.L.emul_add(self.h, xyz, len(xyz))

    def check_directory(self) -> int:
        return self.L.emul_check_directory(self.h)

    def num_bricks(self) -> int:
        return self.L.emul_num_bricks(self.h)

    def knn(self, q, rho=0.3, stats=False, variant=0):
        q = np.ascontiguousarray(q, np.float32).reshape(-1, 3)
        n = len(q)
        xyz = np.zeros((n, 5, 3), np.float32)
        d2 = np.zeros((n, 5), np.float32)
        cnt = np.zeros(n, np.int32)
        st = np.zeros((n, 8), np.int32) if stats else None
        self.L.emul_knn(self.h, q, n, np.float32(rho) * np.float32(rho), variant, xyz, d2, cnt, st.ctypes.data_as(C.c_void_p) if stats else None)
        return (xyz, d2, cnt, st) if stats else (xyz, d2, cnt)

    def trace(self, q, rho=0.3, cap_per_query=400):
        """Structure traces of the growing-boxes search (for the lockstep cost model): (tokens, offsets)."""
        q = np.ascontiguousarray(q, np.float32).reshape(-1, 3)
        n = len(q)
        tr = np.zeros(n * cap_per_query, np.int32)
        off = np.zeros(n + 1, np.int64)
        used = self.L.emul_trace(self.h, q, n, np.float32(rho) * np.float32(rho), tr, len(tr), off)
        return tr[:used], off

    def close(self):
        if self.h:
            self.L.emul_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
