"""ctypes binding of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module. The product package
(lidar_imu_init_b200) never does.

Two shared libraries (built by oracle/Makefile):
  oracle/liboracle.so            restatement only (backend 0)
  oracle/_ref/liboracle_ref.so   restatement + the reference's verbatim ikd-Tree
                                 (backend 1 available); preferred when present.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")

STATE_DOUBLES = 612  # rot_end9 pos3 R_LI9 T_LI3 vel3 bg3 ba3 grav3 cov576


def build(ref: bool = True) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", _HERE])
    if ref and os.path.isdir("/root/reference/include/ikd-Tree"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def lib_path(prefer_ref: bool = True) -> str:
    ref = os.path.join(_HERE, "_ref", "liboracle_ref.so")
    if prefer_ref and os.path.exists(ref):
        return ref
    return os.path.join(_HERE, "liboracle.so")


def _opt(arr):
    return None if arr is None else arr.ctypes.data_as(C.c_void_p)


def load(prefer_ref: bool = True):
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path(prefer_ref)
    if not os.path.exists(path):
        build()
        path = lib_path(prefer_ref)
    L = C.CDLL(path)
    L.oracle_has_ikd.restype = C.c_int
    L.oracle_map_create.restype = C.c_void_p
    L.oracle_map_create.argtypes = [C.c_int, C.c_float]
    L.oracle_map_destroy.argtypes = [C.c_void_p]
    L.oracle_map_build.argtypes = [C.c_void_p, _f32p, C.c_int]
    L.oracle_map_add_points.restype = C.c_int
    L.oracle_map_add_points.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int]
    L.oracle_map_delete_boxes.restype = C.c_int
    L.oracle_map_delete_boxes.argtypes = [C.c_void_p, _f32p, C.c_int]
    L.oracle_map_size.restype = C.c_int
    L.oracle_map_size.argtypes = [C.c_void_p]
    L.oracle_map_validnum.restype = C.c_int
    L.oracle_map_validnum.argtypes = [C.c_void_p]
    L.oracle_map_flatten.restype = C.c_int
    L.oracle_map_flatten.argtypes = [C.c_void_p, _f32p, C.c_int]
    L.oracle_map_knn.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_double, _f32p, _f32p, _i32p, C.c_void_p, C.c_int]
    L.oracle_knn_bruteforce.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_int, C.c_double, _f32p, _f32p, _i32p, C.c_void_p, C.c_int]
    L.oracle_scan_create.restype = C.c_void_p
    L.oracle_scan_create.argtypes = [_f32p, C.c_int]
    L.oracle_scan_destroy.argtypes = [C.c_void_p]
    L.oracle_icp_iterate.argtypes = [C.c_void_p, C.c_void_p, _f64p, _f64p, _f64p, _f64p, C.c_int, C.c_int, C.c_int, _f64p, _f64p, C.POINTER(C.c_int)]
    L.oracle_scan_get.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    L.oracle_scan_get_H.restype = C.c_int
    L.oracle_scan_get_H.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.oracle_state_doubles.restype = C.c_int
    L.oracle_ieskf_update.argtypes = [C.c_void_p, _f64p, _f64p, _f64p, _f64p]
    L.oracle_scan_update.restype = C.c_int
    L.oracle_scan_update.argtypes = [C.c_void_p, C.c_void_p, _f64p, C.c_int, C.c_int, C.c_int, _i32p]
    L.oracle_map_incremental.restype = C.c_int
    L.oracle_map_incremental.argtypes = [C.c_void_p, C.c_void_p, _f64p, _f64p, _f64p, _f64p, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
    L.oracle_undistort_cv.argtypes = [_f32p, _f32p, C.c_int, _f64p, _f64p, _f64p]
    L.oracle_undistort_imu.argtypes = [_f32p, _f32p, C.c_int, _f64p, C.c_int, _f64p, _f64p, _f64p, _f64p]
    L.oracle_voxel_grid.restype = C.c_int
    L.oracle_voxel_grid.argtypes = [_f32p, C.c_int, C.c_float, _f32p]
    L.oracle_esti_plane.restype = C.c_int
    L.oracle_esti_plane.argtypes = [_f32p, _f64p]
    L.oracle_so3_exp.argtypes = [_f64p, _f64p]
    L.oracle_so3_log.argtypes = [_f64p, _f64p]
    L.oracle_mat_inverse.restype = C.c_int
    L.oracle_mat_inverse.argtypes = [_f64p, _f64p, C.c_int]
    assert L.oracle_state_doubles() == STATE_DOUBLES
    _LIB = L
    return L


def has_ikd() -> bool:
    return bool(load().oracle_has_ikd())


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleMap:
    """backend 0 = restated kd-tree/Add_Points; backend 1 = verbatim reference KD_TREE."""

    def __init__(self, ds: float, backend: int = 0):
        self.L = load()
        if backend == 1 and not has_ikd():
            raise RuntimeError("oracle/_ref/liboracle_ref.so (verbatim ikd-Tree) is not built")
        self.h = self.L.oracle_map_create(backend, float(ds))
        self.backend = backend
        self.ds = float(ds)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oracle_map_destroy(self.h)
            self.h = None

    def build(self, xyz):
        xyz = _c32(xyz).reshape(-1, 3)
        self.L.oracle_map_build(self.h, xyz, len(xyz))

    def add_points(self, xyz, downsample_on: bool) -> int:
        xyz = _c32(xyz).reshape(-1, 3)
        if len(xyz) == 0:
            return 0
        return self.L.oracle_map_add_points(self.h, xyz, len(xyz), int(downsample_on))

    def delete_boxes(self, boxes) -> int:
        b = _c32(boxes).reshape(-1, 6)
        return self.L.oracle_map_delete_boxes(self.h, b, len(b)) if len(b) else 0

    def _settled(self, fn):
        # the verbatim ikd-Tree answers -1 while its rebuild thread holds the tree (ikd_Tree.cpp:71-88,120-137): ask again
        import time
        for _ in range(2000):
            v = fn(self.h)
            if v >= 0:
                return v
            time.sleep(0.002)
        return v

    def size(self):
        return self._settled(self.L.oracle_map_size)

    def validnum(self):
        return self._settled(self.L.oracle_map_validnum)

    def flatten(self):
        n = self.L.oracle_map_flatten(self.h, np.zeros((1, 3), np.float32), 0)
        out = np.zeros((max(n, 1), 3), np.float32)
        n2 = self.L.oracle_map_flatten(self.h, out, n)
        return out[:n2]

    def knn(self, q, k=5, max_dist=5.0, nthreads=8):
        q = _c32(q).reshape(-1, 3)
        nq = len(q)
        xyz = np.zeros((nq, k, 3), np.float32)
        d2 = np.zeros((nq, k), np.float32)
        cnt = np.zeros(nq, np.int32)
        ids = np.zeros((nq, k), np.int32)
        self.L.oracle_map_knn(self.h, q, nq, k, max_dist, xyz, d2, cnt, _opt(ids), nthreads)
        return xyz, d2, cnt, ids


def knn_bruteforce(map_xyz, q, k=5, max_dist=5.0, nthreads=8):
    L = load()
    map_xyz = _c32(map_xyz).reshape(-1, 3)
    q = _c32(q).reshape(-1, 3)
    nq = len(q)
    xyz = np.zeros((nq, k, 3), np.float32)
    d2 = np.zeros((nq, k), np.float32)
    cnt = np.zeros(nq, np.int32)
    ids = np.zeros((nq, k), np.int32)
    L.oracle_knn_bruteforce(map_xyz, len(map_xyz), q, nq, k, max_dist, xyz, d2, cnt, _opt(ids), nthreads)
    return xyz, d2, cnt, ids


class OracleScan:
    def __init__(self, body_xyz):
        self.L = load()
        body_xyz = _c32(body_xyz).reshape(-1, 3)
        self.N = len(body_xyz)
        self.h = self.L.oracle_scan_create(body_xyz, self.N)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oracle_scan_destroy(self.h)
            self.h = None

    def iterate(self, omap: OracleMap, rot_end, pos_end, R_LI, T_LI, imu_en, search, nthreads=8):
        HtH = np.zeros((12, 12), np.float64)
        Htr = np.zeros(12, np.float64)
        m = C.c_int(0)
        self.L.oracle_icp_iterate(self.h, omap.h, _c64(rot_end).reshape(9), _c64(pos_end), _c64(R_LI).reshape(9), _c64(T_LI),
                                  int(imu_en), int(search), nthreads, HtH, Htr, C.byref(m))
        return HtH, Htr, m.value

    def get(self):
        N = self.N
        out = dict(
            world=np.zeros((N, 3), np.float32), near_cnt=np.zeros(N, np.int32), near_xyz=np.zeros((N, 5, 3), np.float32),
            near_d2=np.zeros((N, 5), np.float32), selected=np.zeros(N, np.uint8), normvec=np.zeros((N, 4), np.float32),
            res_last=np.zeros(N, np.float32))
        self.L.oracle_scan_get(self.h, *[_opt(out[k]) for k in ("world", "near_cnt", "near_xyz", "near_d2", "selected", "normvec", "res_last")])
        return out

    def get_H(self):
        m = self.L.oracle_scan_get_H(self.h, None, None, None)
        H = np.zeros((max(m, 1), 12), np.float64)
        meas = np.zeros(max(m, 1), np.float64)
        idx = np.zeros(max(m, 1), np.int32)
        self.L.oracle_scan_get_H(self.h, _opt(H), _opt(meas), _opt(idx))
        return H[:m], meas[:m], idx[:m]

    def ieskf_update(self, state, state_prop):
        state = _c64(state).copy()
        sol = np.zeros(24)
        KH = np.zeros((24, 12))
        self.L.oracle_ieskf_update(self.h, state, _c64(state_prop), sol, KH)
        return state, sol, KH

    def scan_update(self, omap: OracleMap, state, max_iter=5, imu_en=False, nthreads=8):
        state = _c64(state).copy()
        stats = np.zeros(2, np.int32)
        iters = self.L.oracle_scan_update(self.h, omap.h, state, max_iter, int(imu_en), nthreads, stats)
        return state, iters, int(stats[0]), int(stats[1])

    def map_incremental(self, omap: OracleMap, rot_end, pos_end, R_LI, T_LI, ds, flg_EKF_inited=True):
        n_add, n_nod = C.c_int(0), C.c_int(0)
        flags = np.zeros(self.N, np.int32)
        c = self.L.oracle_map_incremental(self.h, omap.h, _c64(rot_end).reshape(9), _c64(pos_end), _c64(R_LI).reshape(9), _c64(T_LI),
                                          float(ds), int(flg_EKF_inited), C.byref(n_add), C.byref(n_nod), _opt(flags))
        return c, n_add.value, n_nod.value, flags


def undistort_cv(xyz, t_ms, omega, rot_end, vel_end):
    out = _c32(xyz).reshape(-1, 3).copy()
    load().oracle_undistort_cv(out, _c32(t_ms), len(out), _c64(omega), _c64(rot_end).reshape(9), _c64(vel_end))
    return out


def undistort_imu(xyz, t_ms, poses, rot_end, pos_end, R_LI, T_LI):
    out = _c32(xyz).reshape(-1, 3).copy()
    P = _c64(poses).reshape(-1, 22)
    load().oracle_undistort_imu(out, _c32(t_ms), len(out), P.reshape(-1), len(P), _c64(rot_end).reshape(9), _c64(pos_end), _c64(R_LI).reshape(9), _c64(T_LI))
    return out


def voxel_grid(xyz, leaf):
    """PCL VoxelGrid restatement (xyz centroids, ascending leaf index). Returns [m,3] float32."""
    xyz = _c32(xyz).reshape(-1, 3)
    out = np.zeros((max(len(xyz), 1), 3), np.float32)
    m = load().oracle_voxel_grid(xyz, len(xyz), float(leaf), out)
    if m < 0:
        raise ValueError("leaf size too small (index overflow) or no finite point")
    return out[:m]


def esti_plane(nb):
    L = load()
    out = np.zeros(4)
    ok = L.oracle_esti_plane(_c32(nb).reshape(15), out)
    return bool(ok), out


def so3_exp(v):
    R = np.zeros(9)
    load().oracle_so3_exp(_c64(v), R)
    return R.reshape(3, 3)


def so3_log(R):
    v = np.zeros(3)
    load().oracle_so3_log(_c64(R).reshape(9), v)
    return v


# ---- flat state helpers (layout of struct State in liinit_oracle.cpp) ----
def state_pack(rot_end=None, pos_end=None, R_LI=None, T_LI=None, vel=None, bg=None, ba=None, grav=None, cov=None):
    s = np.zeros(STATE_DOUBLES)
    s[0:9] = np.eye(3).reshape(9) if rot_end is None else np.asarray(rot_end, float).reshape(9)
    s[9:12] = 0 if pos_end is None else pos_end
    s[12:21] = np.eye(3).reshape(9) if R_LI is None else np.asarray(R_LI, float).reshape(9)
    s[21:24] = 0 if T_LI is None else T_LI
    s[24:27] = 0 if vel is None else vel
    s[27:30] = 0 if bg is None else bg
    s[30:33] = 0 if ba is None else ba
    s[33:36] = 0 if grav is None else grav
    if cov is None:  # StatesGroup() default (common_lib.h:79-80)
        cov = np.eye(24)
        cov[15:24, 15:24] = np.eye(9) * 1e-5
    s[36:] = np.asarray(cov, float).reshape(576)
    return s


def state_unpack(s):
    return dict(rot_end=s[0:9].reshape(3, 3).copy(), pos_end=s[9:12].copy(), R_LI=s[12:21].reshape(3, 3).copy(), T_LI=s[21:24].copy(),
                vel=s[24:27].copy(), bg=s[27:30].copy(), ba=s[30:33].copy(), grav=s[33:36].copy(), cov=s[36:].reshape(24, 24).copy())
