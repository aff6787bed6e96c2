"""Host-side library (IESKF in information form, state algebra): CPU tests against the oracle's literal restatement
of laserMapping.cpp:1080-1087 and common_lib.h:109-151. No GPU compute is invoked."""
import numpy as np
import pytest

from lidar_imu_init_b200 import scenes


@pytest.fixture(scope="module")
def host_lib(gpu_lib):
    from lidar_imu_init_b200 import _build, host
    _build.build_host()
    host.load()
    return host


def test_state_algebra_matches_oracle(host_lib, oracle_mod):
    rng = np.random.default_rng(11)
    assert np.array_equal(host_lib.state_init(), oracle_mod.state_pack())
    for _ in range(20):
        v = rng.normal(size=3)
        v *= rng.uniform(1e-4, 3.0) / np.linalg.norm(v)
        R = np.zeros(9)
        host_lib.load().liinit_so3_exp(v, R)
        assert np.array_equal(R.reshape(3, 3), oracle_mod.so3_exp(v))
        w = np.zeros(3)
        host_lib.load().liinit_so3_log(R, w)
        assert np.array_equal(w, oracle_mod.so3_log(R.reshape(3, 3)))
    s = host_lib.state_init()
    d = rng.normal(size=24) * 0.01
    s2 = host_lib.boxplus(s, d)
    back = host_lib.boxminus(s2, s)
    assert np.allclose(back, d, atol=1e-12)


@pytest.mark.parametrize("imu_en", [False, True])
def test_ieskf_information_form_matches_literal(host_lib, oracle_mod, imu_en):
    c = scenes.make_config("C2", N=2500, M=30000, open_air_frac=0.0, imu_en=imu_en)
    p = c["pose_init"]
    om = oracle_mod.OracleMap(c["ds"], 0)
    om.build(c["map_xyz"])
    sc = oracle_mod.OracleScan(c["body_xyz"])
    HtH, Htr, m = sc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, imu_en, True)
    st0 = host_lib.state_from_pose(p.rot_end, p.pos_end, p.R_LI, p.T_LI)
    prop = st0.copy()
    prop[9:12] += [0.01, -0.02, 0.005]
    s_lit, sol_lit, KH_lit = sc.ieskf_update(st0, prop)
    s_inf, sol_inf, KH_inf = host_lib.ieskf_update(st0, prop, HtH, Htr)
    assert np.allclose(sol_inf, sol_lit, rtol=1e-5, atol=1e-9)
    assert np.allclose(s_inf[:36], s_lit[:36], rtol=0, atol=1e-8)      # pose part: far inside the 1e-3 bar
    assert np.allclose(KH_inf, KH_lit, rtol=1e-4, atol=1e-5)


def test_fov_segment_follows_reference_logic(host_lib):
    """lasermap_fov_segment (laserMapping.cpp:260-305): first call centres the cube, later calls shift it when the LiDAR
    comes within 1.5*det_range of a face and report the slabs that fell out."""
    cube, det = 1000.0, 100.0
    seg = host_lib.FovSegmenter(cube, det)
    assert len(seg.update([0.0, 0.0, 0.0])) == 0 and np.allclose(seg.box, [-500] * 3 + [500] * 3)
    assert len(seg.update([100.0, 0.0, 0.0])) == 0                      # 400 m from the +x face > 150 m
    boxes = seg.update([360.0, 0.0, 0.0])                                # 140 m from the +x face -> move
    mov = max((cube - 2 * 1.5 * det) * 0.5 * 0.9, det * 0.5)             # = 315
    assert len(boxes) == 1
    assert np.allclose(boxes[0], [-500, -500, -500, -500 + mov, 500, 500])   # the slab left behind on the -x side
    assert np.allclose(seg.box, [-500 + mov, -500, -500, 500 + mov, 500, 500])
    b2 = seg.update([360.0, -370.0, 0.0])                                # now near the -y face
    assert len(b2) == 1 and np.allclose(b2[0], [-185, 500 - mov, -500, 815, 500, 500])


def test_propagate_cv_matches_dense_formula():
    """liinit_propagate_cv vs F cov F^T + Q written out densely (IMU_Processing.hpp:225-243)."""
    from lidar_imu_init_b200 import host
    from lidar_imu_init_b200.scenes import so3_exp
    rng = np.random.default_rng(3)
    s = host.state_init()
    A = rng.standard_normal((24, 24))
    cov = A @ A.T * 1e-2 + np.eye(24) * 1e-3
    s[36:] = cov.reshape(-1)
    s[0:9] = so3_exp(np.array([0.2, -0.1, 0.4])).reshape(9)
    s[9:12] = [1.0, 2.0, 3.0]
    s[24:27] = [0.3, -0.2, 0.1]       # vel_end
    s[27:30] = [0.2, -0.4, 0.7]       # bias_g = angular velocity in the CV model
    dt = 0.07
    F = np.eye(24)
    F[0:3, 0:3] = so3_exp(-s[27:30] * dt)
    F[0:3, 15:18] = np.eye(3) * dt
    F[3:6, 12:15] = np.eye(3) * dt
    Q = np.zeros((24, 24))
    Q[15:18, 15:18] = np.diag([0.1, 0.2, 0.3]) * dt * dt
    Q[12:15, 12:15] = np.diag([0.4, 0.5, 0.6]) * dt * dt
    o = host.propagate_cv(s, dt, np.array([0.1, 0.2, 0.3]), np.array([0.4, 0.5, 0.6]))
    assert np.allclose(o[36:].reshape(24, 24), F @ cov @ F.T + Q, rtol=1e-13, atol=1e-15)
    assert np.allclose(o[0:9].reshape(3, 3), s[0:9].reshape(3, 3) @ so3_exp(s[27:30] * dt), atol=1e-15)
    assert np.allclose(o[9:12], s[9:12] + s[24:27] * dt)
    assert np.array_equal(o[12:36], s[12:36])


def test_propagate_cv_small_angular_velocity_follows_exp_of_velocity():
    """Exp(ang_vel, dt) (so3_math.h:39-59) is gated on |ang_vel| > 1e-7, not on the angle: with |bias_g| = 5e-5 rad/s and dt = 0.02 s the
    angle 1e-6 is below the 1e-5 gate of Exp(v1, v2, v3) but the reference still rotates (IMU_Processing.hpp:225-226,239)."""
    from lidar_imu_init_b200 import _build, host
    _build.build_host()
    s = host.state_init()
    w = np.array([3e-5, -4e-5, 0.0])
    s[27:30] = w
    dt = 0.02
    out = host.propagate_cv(s, dt)
    R = out[0:9].reshape(3, 3)
    ang = np.linalg.norm(w) * dt
    k = w / np.linalg.norm(w)
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    want = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    assert not np.array_equal(R, np.eye(3))
    assert np.abs(R - want).max() < 1e-15
    # below the velocity gate: identity
    s[27:30] = np.array([5e-8, 0.0, 0.0])
    assert np.array_equal(host.propagate_cv(s, dt)[0:9].reshape(3, 3), np.eye(3))
