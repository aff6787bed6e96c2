"""tests-only ctypes binding of tests/emul/libliinit_emul.so: the product's C-ABI layer + kernels compiled for the HOST from their own
source (tests/emul/make_liinit_emul.py). A checker of logic for `-m "not gpu"` tests -- the package never loads it and has no CPU path."""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
import shutil

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_liinit_emul", os.path.join(HERE, "emul", "make_liinit_emul.py"))
_mk = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mk)


def available() -> bool:
    return shutil.which("g++") is not None and os.path.exists(os.path.join(_mk.CUDA_INC, "vector_types.h"))


class Config(C.Structure):   # include/liinit_gpu.h: liinit_config
    _fields_ = [("filter_size_map", C.c_float), ("max_map_points", C.c_int), ("max_scan_points", C.c_int), ("device_id", C.c_int),
                ("brick_cells_log2", C.c_int), ("hash_capacity_log2", C.c_int), ("knn_group_lanes", C.c_int), ("knn_seed_radius_cells", C.c_float),
                ("knn_index", C.c_int), ("reserved", C.c_int * 6)]


_L = None
vp = C.c_void_p
_f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def load():
    global _L
    if _L is None:
        L = C.CDLL(_mk.build())
        L.liinit_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
        L.liinit_destroy.argtypes = [vp]
        L.liinit_last_error.restype = C.c_char_p
        L.liinit_last_error.argtypes = [vp]
        L.liinit_map_build.argtypes = [vp, vp, C.c_int, C.c_int]
        L.liinit_map_add_points.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.liinit_map_delete_boxes.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int)]
        L.liinit_map_compact.argtypes = [vp]
        L.liinit_map_stats.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        L.liinit_map_validnum.argtypes = [vp, C.POINTER(C.c_int)]
        L.liinit_map_download.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int)]
        L.liinit_map_nearest_search.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double, vp, vp, vp]
        L.liinit_scan_upload.argtypes = [vp, vp, C.c_int, C.c_int]
        L.liinit_scan_attach_host.argtypes = [vp, vp, C.c_int, C.c_int]
        L.liinit_scan_upload_raw.argtypes = [vp, vp, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_int)]
        L.liinit_scan_download_body.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int)]
        L.liinit_icp_iterate.argtypes = [vp, _f64, _f64, _f64, _f64, C.c_int, C.c_int, _f64, _f64, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.liinit_scan_download_state.argtypes = [vp, vp, vp, vp, vp, vp]
        L.liinit_scan_download_effect.argtypes = [vp, vp, vp, C.c_int, C.POINTER(C.c_int)]
        L.liinit_map_incremental.argtypes = [vp, _f64, _f64, _f64, _f64, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.liinit_knn_index.argtypes = [vp, C.POINTER(C.c_int)]
        L.liinit_raw_upload.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
        L.liinit_raw_undistort_cv.argtypes = [vp, _f64, _f64, _f64]
        L.liinit_raw_undistort_imu.argtypes = [vp, _f64, C.c_int, _f64, _f64, _f64, _f64]
        L.liinit_raw_download.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int)]
        L.liinit_raw_downsample.argtypes = [vp, C.c_float, C.POINTER(C.c_int)]
        L.liinit_debug_esti_plane.argtypes = [vp, vp, C.c_int, vp, vp]
        L.liinit_set_reseed.argtypes = [vp, C.c_int]
        _L = L
    return _L


class EmulError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"liinit (emulated) error {code}: {msg}")
        self.code = code


def _pts(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] in (3, 4, 12)
    return a


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class EmulGpu:
    """Same calls as lidar_imu_init_b200.capi.LiInitGpu for the entry points the CPU tests use."""

    def __init__(self, filter_size_map=0.15, max_map_points=200_000, max_scan_points=20_000, knn_group_lanes=0, knn_seed_radius_cells=0.0,
                 knn_index=0, brick_cells_log2=0, hash_capacity_log2=0):
        self.L = load()
        cfg = Config(filter_size_map, max_map_points, max_scan_points, 0, brick_cells_log2, hash_capacity_log2, knn_group_lanes, knn_seed_radius_cells,
                     knn_index)
        h = vp()
        rc = self.L.liinit_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise EmulError(rc, (self.L.liinit_last_error(None) or b"").decode())
        self.h = h
        self.scan_n = 0

    def _ck(self, rc):
        if rc != 0:
            raise EmulError(rc, (self.L.liinit_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.liinit_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def knn_index(self):
        v = C.c_int()
        self._ck(self.L.liinit_knn_index(self.h, C.byref(v)))
        return v.value

    def map_build(self, xyz):
        a = _pts(xyz)
        self._ck(self.L.liinit_map_build(self.h, a.ctypes.data_as(vp), a.shape[1], len(a)))

    def map_add_points(self, xyz, downsample_on):
        a = _pts(xyz)
        n = C.c_int()
        self._ck(self.L.liinit_map_add_points(self.h, a.ctypes.data_as(vp), a.shape[1], len(a), int(downsample_on), C.byref(n)))
        return n.value

    def map_delete_boxes(self, boxes):
        b = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
        n = C.c_int()
        self._ck(self.L.liinit_map_delete_boxes(self.h, b.ctypes.data_as(vp), len(b), C.byref(n)))
        return n.value

    def map_validnum(self):
        n = C.c_int()
        self._ck(self.L.liinit_map_validnum(self.h, C.byref(n)))
        return n.value

    def map_download(self):
        n = self.map_validnum()
        out = np.zeros((max(n, 1), 3), np.float32)
        m = C.c_int()
        self._ck(self.L.liinit_map_download(self.h, out.ctypes.data_as(vp), len(out), C.byref(m)))
        return out[:m.value]

    def map_compact(self):
        self._ck(self.L.liinit_map_compact(self.h))

    def map_stats(self):
        b, sl, pu, pc = C.c_int(0), C.c_int(0), C.c_longlong(0), C.c_longlong(0)
        self._ck(self.L.liinit_map_stats(self.h, C.byref(b), C.byref(sl), C.byref(pu), C.byref(pc)))
        return dict(bricks=b.value, hash_slots=sl.value, pool_used=pu.value, pool_cap=pc.value)

    def set_reseed(self, enabled):
        self._ck(self.L.liinit_set_reseed(self.h, int(enabled)))

    def debug_esti_plane(self, nb):
        a = np.ascontiguousarray(nb, np.float32).reshape(-1, 15)
        out = np.zeros((len(a), 4))
        ok = np.zeros(len(a), np.uint8)
        self._ck(self.L.liinit_debug_esti_plane(self.h, a.ctypes.data_as(vp), len(a), out.ctypes.data_as(vp), ok.ctypes.data_as(vp)))
        return out, ok.astype(bool)

    def nearest_search(self, q):
        q = _pts(q)
        n = len(q)
        xyz = np.zeros((n, 5, 3), np.float32)
        d2 = np.zeros((n, 5), np.float32)
        cnt = np.zeros(n, np.int32)
        self._ck(self.L.liinit_map_nearest_search(self.h, q.ctypes.data_as(vp), q.shape[1], n, 5.0, xyz.ctypes.data_as(vp), d2.ctypes.data_as(vp),
                                                  cnt.ctypes.data_as(vp)))
        return xyz, d2, cnt

    def scan_upload(self, body):
        a = _pts(body)
        self._ck(self.L.liinit_scan_upload(self.h, a.ctypes.data_as(vp), a.shape[1], len(a)))
        self.scan_n = len(a)

    def scan_attach(self, body):
        """liinit_scan_attach_host (every host buffer counts as page-locked here); keeps the array alive"""
        self._attached = _pts(body)
        self._ck(self.L.liinit_scan_attach_host(self.h, self._attached.ctypes.data_as(vp), self._attached.shape[1], len(self._attached)))
        self.scan_n = len(self._attached)

    def scan_upload_raw(self, pts, leaf):
        a = _pts(pts)
        n = C.c_int()
        self._ck(self.L.liinit_scan_upload_raw(self.h, a.ctypes.data_as(vp), a.shape[1], len(a), leaf, C.byref(n)))
        self.scan_n = n.value
        return n.value

    def raw_upload(self, pts, time_index=-1):
        a = np.ascontiguousarray(pts, np.float32)
        self._ck(self.L.liinit_raw_upload(self.h, a.ctypes.data_as(vp), a.shape[1], time_index, len(a)))
        self.raw_n = len(a)

    def raw_undistort_cv(self, omega, rot_end, vel_end):
        self._ck(self.L.liinit_raw_undistort_cv(self.h, _c64(omega), _c64(rot_end), _c64(vel_end)))

    def raw_undistort_imu(self, poses, rot_end, pos_end, R_LI, T_LI):
        P = _c64(poses)
        self._ck(self.L.liinit_raw_undistort_imu(self.h, P.reshape(-1), len(P), _c64(rot_end), _c64(pos_end), _c64(R_LI), _c64(T_LI)))

    def raw_points(self):
        out = np.zeros((self.raw_n, 3), np.float32)
        n = C.c_int()
        self._ck(self.L.liinit_raw_download(self.h, out.ctypes.data_as(vp), len(out), C.byref(n)))
        return out[:n.value]

    def raw_downsample(self, leaf):
        n = C.c_int()
        self._ck(self.L.liinit_raw_downsample(self.h, leaf, C.byref(n)))
        self.scan_n = n.value
        return n.value

    def scan_body(self):
        out = np.zeros((max(self.scan_n, 1), 3), np.float32)
        n = C.c_int()
        self._ck(self.L.liinit_scan_download_body(self.h, out.ctypes.data_as(vp), len(out), C.byref(n)))
        return out[:n.value]

    def icp_iterate(self, rot_end, pos_end, R_LI, T_LI, imu_en, search):
        H = np.zeros((12, 12))
        b = np.zeros(12)
        m = C.c_int()
        rs = C.c_double()
        self._ck(self.L.liinit_icp_iterate(self.h, _c64(rot_end), _c64(pos_end), _c64(R_LI), _c64(T_LI), int(imu_en), int(search), H.reshape(-1), b,
                                           C.byref(m), C.byref(rs)))
        return H, b, m.value, rs.value

    def scan_state(self):
        n = self.scan_n
        world = np.zeros((n, 3), np.float32)
        near = np.zeros((n, 5, 3), np.float32)
        cnt = np.zeros(n, np.int32)
        sel = np.zeros(n, np.uint8)
        nv = np.zeros((n, 4), np.float32)
        self._ck(self.L.liinit_scan_download_state(self.h, world.ctypes.data_as(vp), near.ctypes.data_as(vp), cnt.ctypes.data_as(vp), sel.ctypes.data_as(vp),
                                                   nv.ctypes.data_as(vp)))
        return dict(world=world, near_xyz=near, near_cnt=cnt, selected=sel, normvec=nv)

    def scan_effect(self):
        n = self.scan_n
        ori = np.zeros((n, 3), np.float32)
        nv = np.zeros((n, 4), np.float32)
        m = C.c_int()
        self._ck(self.L.liinit_scan_download_effect(self.h, ori.ctypes.data_as(vp), nv.ctypes.data_as(vp), n, C.byref(m)))
        return ori[:m.value], nv[:m.value]

    def map_incremental(self, rot_end, pos_end, R_LI, T_LI, ds, flg_EKF_inited=True):
        a, b = C.c_int(), C.c_int()
        self._ck(self.L.liinit_map_incremental(self.h, _c64(rot_end), _c64(pos_end), _c64(R_LI), _c64(T_LI), float(ds), int(flg_EKF_inited), C.byref(a),
                                               C.byref(b)))
        return a.value, b.value


# ---- the product's host-side C++ (csrc/host/liinit_host.cpp: IESKF, per-scan driver) linked against the emulated library ----------------
_H = None


class ScanStats(C.Structure):   # csrc/host/liinit_host.h: liinit_scan_stats
    _fields_ = [("iterations", C.c_int), ("search_passes", C.c_int), ("effect_feat_num", C.c_int), ("converged", C.c_int),
                ("last_rot_deg", C.c_double), ("last_trans_cm", C.c_double), ("res_sq", C.c_double)]


def host_lib():
    """liinit_host.cpp compiled as it is and linked against libliinit_emul.so instead of libliinit_gpu.so: liinit_scan_update then drives
    the emulated kernels through the same C-ABI calls -- the per-scan driver on the CPU, for -m "not gpu" runs"""
    global _H
    if _H is not None:
        return _H
    import subprocess
    emul = _mk.build()
    src = os.path.join(_mk.CSRC, "host", "liinit_host.cpp")
    out = os.path.join(_mk.GEN, "libliinit_host_emul.so")
    deps = [src, os.path.join(_mk.CSRC, "host", "liinit_host.h"), emul]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.normpath(os.path.join(_mk.CSRC, "..", "..", "include")),
                               "-o", out, src, "-L", os.path.dirname(emul), "-lliinit_emul", "-Wl,-rpath," + os.path.dirname(emul)])
    load()
    H = C.CDLL(out)
    H.liinit_scan_update.restype = C.c_int
    H.liinit_scan_update.argtypes = [vp, _f64, C.c_int, C.c_int, C.POINTER(ScanStats)]
    H.liinit_state_init.argtypes = [_f64]
    _H = H
    return H


def scan_update(g: "EmulGpu", state, max_iteration=5, imu_en=False):
    s = np.ascontiguousarray(state, np.float64).copy()
    st = ScanStats()
    rc = host_lib().liinit_scan_update(g.h, s, int(max_iteration), int(imu_en), C.byref(st))
    if rc != 0:
        raise EmulError(rc, (g.L.liinit_last_error(g.h) or b"").decode())
    return s, {k: getattr(st, k) for k, _ in ScanStats._fields_}


def state_from_pose(rot_end, pos_end, R_LI, T_LI):
    """liinit_state (csrc/host/liinit_host.h): rot_end 9 | pos_end 3 | offset_R_L_I 9 | offset_T_L_I 3 | vel 3 | bias_g 3 | bias_a 3 | gravity 3 | cov 24x24"""
    s = np.zeros(36 + 576)
    host_lib().liinit_state_init(s)
    s[0:9] = np.asarray(rot_end, float).reshape(9)
    s[9:12] = pos_end
    s[12:21] = np.asarray(R_LI, float).reshape(9)
    s[21:24] = T_LI
    return s
