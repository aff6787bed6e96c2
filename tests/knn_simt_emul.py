"""ctypes binding of the CPU run of the product's lockstep 5-NN kernel (tests/emul/knn_simt_emul.cpp + simt_shim.h).
Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "knn_simt_emul.cpp")
LIB = os.path.join(HERE, "emul", "libknn_simt_emul.so")
CSRC = os.path.join(HERE, "..", "lidar_imu_init_b200", "csrc")
CUDA_INC = "/usr/local/cuda/include"
_f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def available() -> bool:
    return shutil.which("g++") is not None and os.path.exists(os.path.join(CUDA_INC, "vector_types.h"))


def build(force: bool = False) -> str:
    deps = [SRC, os.path.join(HERE, "emul", "simt_shim.h"), os.path.join(HERE, "emul", "emul_map.h")] + \
           [os.path.join(CSRC, f) for f in ("knn_kernels.cuh", "icp_kernels.cuh", "common.cuh", "cells.cuh")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-attributes", "-Wno-unknown-pragmas",
                               "-I", CUDA_INC, SRC, "-o", LIB])
    return LIB


_L = None


def load():
    global _L
    if _L is None:
        L = C.CDLL(build())
        L.simt_map_create.restype = C.c_void_p
        L.simt_map_create.argtypes = [_f32, C.c_int, C.c_float, C.c_int]
        L.simt_map_destroy.argtypes = [C.c_void_p]
        L.simt_knn_scan.restype = C.c_longlong
        L.simt_knn_scan.argtypes = [C.c_void_p, _f32, C.c_int, _f64, C.c_float, C.c_int, _f32, _f32, _i32]
        L.simt_icp_pass.argtypes = [C.c_void_p, _f32, C.c_int, _f64, C.c_void_p, C.c_float, C.c_int, C.c_int, _f64, C.c_void_p, _f32, _f32, _i32,
                                    np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS"), _f32]
        _L = L
    return _L


def _pose_vec(pose):
    return np.ascontiguousarray(np.concatenate([np.asarray(pose.rot_end, np.float64).ravel(), np.asarray(pose.pos_end, np.float64).ravel(),
                                                np.asarray(pose.R_LI, np.float64).ravel(), np.asarray(pose.T_LI, np.float64).ravel()]))


class SimtMap:
    def __init__(self, map_xyz, ds=0.15):
        self.L = load()
        xyz = np.ascontiguousarray(map_xyz, np.float32).reshape(-1, 3)
        hl = 12
        while (1 << hl) < max(len(xyz) // 2, 1024):
            hl += 1
        self.h = C.c_void_p(self.L.simt_map_create(xyz, len(xyz), ds, hl))

    def knn_scan(self, body, pose, rho=0.3, G=4):
        """One search pass of k_knn_scan<G> (body -> world transform + exact 5-NN): world [n,3], near_xyz [n,5,3], near_cnt [n]."""
        body = np.ascontiguousarray(body, np.float32).reshape(-1, 3)
        n = len(body)
        P = np.concatenate([np.asarray(pose.rot_end, np.float64).ravel(), np.asarray(pose.pos_end, np.float64).ravel(),
                            np.asarray(pose.R_LI, np.float64).ravel(), np.asarray(pose.T_LI, np.float64).ravel()])
        world = np.zeros((n, 3), np.float32)
        near = np.zeros((n, 5, 3), np.float32)
        cnt = np.zeros(n, np.int32)
        self.rendezvous = self.L.simt_knn_scan(self.h, body, n, np.ascontiguousarray(P), np.float32(rho) * np.float32(rho), G, world, near, cnt)
        return world, near, cnt

    def icp_pass(self, body, pose, imu_en, pose2=None, rho=0.3, G=4):
        """Search pass (k_knn_scan + k_icp_plane) and, with pose2, a reuse pass behind it. Returns dict(H, b, m, res_sq, [H2, b2, m2],
        world, near_xyz, near_cnt, selected, normvec) with H [12,12], b [12] as liinit_icp_iterate delivers them."""
        body = np.ascontiguousarray(body, np.float32).reshape(-1, 3)
        n = len(body)
        out, out2 = np.zeros(160), np.zeros(160)
        world = np.zeros((n, 3), np.float32)
        near = np.zeros((n, 5, 3), np.float32)
        cnt = np.zeros(n, np.int32)
        sel = np.zeros(n, np.uint8)
        nv = np.zeros((n, 4), np.float32)
        P2 = _pose_vec(pose2) if pose2 is not None else None
        self.L.simt_icp_pass(self.h, body, n, _pose_vec(pose), P2.ctypes.data_as(C.c_void_p) if P2 is not None else None,
                             np.float32(rho) * np.float32(rho), G, int(bool(imu_en)), out,
                             out2.ctypes.data_as(C.c_void_p) if P2 is not None else None, world, near, cnt, sel, nv)
        r = dict(H=out[:144].reshape(12, 12).copy(), b=out[144:156].copy(), res_sq=out[156], m=int(round(out[157])), world=world, near_xyz=near,
                 near_cnt=cnt, selected=sel, normvec=nv)
        if P2 is not None:
            r.update(H2=out2[:144].reshape(12, 12).copy(), b2=out2[144:156].copy(), m2=int(round(out2[157])))
        return r

    def close(self):
        if self.h:
            self.L.simt_map_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
