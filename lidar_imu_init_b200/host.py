"""ctypes binding of the host-side library (csrc/host/liinit_host.h): StatesGroup algebra, the IESKF update in
information form and the per-scan driver loop of laserMapping.cpp:936-1134 over the C-ABI."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build, capi

STATE_DOUBLES = 612  # rot_end9 pos3 R_LI9 T_LI3 vel3 bg3 ba3 grav3 cov576 (struct liinit_state)
_LIB = None
_f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


class ScanStats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("search_passes", C.c_int), ("effect_feat_num", C.c_int), ("converged", C.c_int),
                ("last_rot_deg", C.c_double), ("last_trans_cm", C.c_double), ("res_sq", C.c_double)]


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    capi.load()  # libliinit_gpu.so first (the host library links against it)
    if not os.path.exists(_build.HOST_LIB):
        raise FileNotFoundError(f"{_build.HOST_LIB} is missing: run __graft_entry__.build()")
    L = C.CDLL(_build.HOST_LIB)
    L.liinit_state_init.argtypes = [_f64]
    L.liinit_state_boxplus.argtypes = [_f64, _f64]
    L.liinit_state_boxminus.argtypes = [_f64, _f64, _f64]
    L.liinit_so3_exp.argtypes = [_f64, _f64]
    L.liinit_so3_log.argtypes = [_f64, _f64]
    L.liinit_ieskf_update.restype = C.c_int
    L.liinit_ieskf_update.argtypes = [_f64, _f64, _f64, _f64, _f64, C.c_void_p]
    L.liinit_scan_update.restype = C.c_int
    L.liinit_scan_update.argtypes = [C.c_void_p, _f64, C.c_int, C.c_int, C.POINTER(ScanStats)]
    L.liinit_propagate_cv.restype = None
    L.liinit_propagate_cv.argtypes = [_f64, C.c_double, _f64, _f64]
    L.liinit_fov_segment.restype = C.c_int
    L.liinit_fov_segment.argtypes = [_f64, C.c_double, C.c_double, np.ctypeslib.ndpointer(dtype=np.float32, flags='C_CONTIGUOUS'), C.POINTER(C.c_int),
                                     np.ctypeslib.ndpointer(dtype=np.float32, flags='C_CONTIGUOUS')]
    _LIB = L
    return L


def state_init():
    s = np.zeros(STATE_DOUBLES)
    load().liinit_state_init(s)
    return s


def state_from_pose(rot_end, pos_end, R_LI, T_LI):
    s = state_init()
    s[0:9] = np.asarray(rot_end, float).reshape(9)
    s[9:12] = pos_end
    s[12:21] = np.asarray(R_LI, float).reshape(9)
    s[21:24] = T_LI
    return s


def state_pose(s):
    return s[0:9].reshape(3, 3).copy(), s[9:12].copy(), s[12:21].reshape(3, 3).copy(), s[21:24].copy()


def boxplus(s, d):
    s = np.ascontiguousarray(s, np.float64).copy()
    load().liinit_state_boxplus(s, np.ascontiguousarray(d, np.float64))
    return s


def boxminus(a, b):
    o = np.zeros(24)
    load().liinit_state_boxminus(np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64), o)
    return o


def ieskf_update(state, state_prop, HtH, Htr):
    s = np.ascontiguousarray(state, np.float64).copy()
    sol = np.zeros(24)
    KH = np.zeros((24, 12))
    rc = load().liinit_ieskf_update(s, np.ascontiguousarray(state_prop, np.float64), np.ascontiguousarray(HtH, np.float64).reshape(144),
                                    np.ascontiguousarray(Htr, np.float64), sol, KH.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("liinit_ieskf_update failed (singular covariance?)")
    return s, sol, KH


def propagate_cv(state, dt, cov_gyr_scale=0.1, cov_acc_scale=0.1):
    """Constant-velocity propagation of the LiDAR-only mode (IMU_Processing.hpp:212-243); covariance scales as
    mapping/gyr_cov, mapping/acc_cov (laserMapping.cpp:776-777,833-834)."""
    s = np.ascontiguousarray(state, np.float64).copy()
    load().liinit_propagate_cv(s, float(dt), np.full(3, cov_gyr_scale, np.float64) if np.isscalar(cov_gyr_scale) else np.ascontiguousarray(cov_gyr_scale, np.float64),
                               np.full(3, cov_acc_scale, np.float64) if np.isscalar(cov_acc_scale) else np.ascontiguousarray(cov_acc_scale, np.float64))
    return s


def scan_update(gpu: capi.LiInitGpu, state, max_iteration=5, imu_en=False):
    """Per-scan ICP + IESKF update on the scan already uploaded to `gpu`. Returns (posterior state, stats dict)."""
    s = np.ascontiguousarray(state, np.float64).copy()
    st = ScanStats()
    rc = load().liinit_scan_update(gpu.h, s, int(max_iteration), int(imu_en), C.byref(st))
    if rc != 0:
        raise capi.LiInitError(rc, (gpu.L.liinit_last_error(gpu.h) or b"").decode())
    return s, {k: getattr(st, k) for k, _ in ScanStats._fields_}


class FovSegmenter:
    """lasermap_fov_segment (laserMapping.cpp:260-305) with its two globals (LocalMap_Points, Localmap_Initialized)."""

    def __init__(self, cube_len: float, det_range: float):
        self.cube_len, self.det_range = float(cube_len), float(det_range)
        self.box = np.zeros(6, np.float32)
        self.init = C.c_int(0)

    def update(self, pos_lidar):
        out = np.zeros(18, np.float32)
        n = load().liinit_fov_segment(np.ascontiguousarray(pos_lidar, np.float64), self.cube_len, self.det_range, self.box, C.byref(self.init), out)
        return out.reshape(3, 6)[:n].copy()
