"""Row N3: per-point motion compensation on the device vs the oracle's restatement of the two
back-propagation loops of IMU_Processing.hpp, and the whole raw front end (undistort -> voxel grid -> ICP pass)."""
import numpy as np
import pytest

from lidar_imu_init_b200 import scenes

pytestmark = pytest.mark.gpu


def _raw(n, seed):
    rng = np.random.default_rng(seed)
    pts = np.zeros((n, 12), np.float32)                 # pcl::PointXYZINormal layout, curvature (index 9) = time in ms
    pts[:, :3] = rng.uniform(-30, 30, (n, 3))
    pts[:, 9] = rng.uniform(0, 100.0, n).astype(np.float32)
    pts[7, 9] = 0.0                                     # a point at the scan start
    return pts


def _ulp_close(a, b, k=1):
    return np.all(np.abs(a - b) <= k * np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32)))


def test_cv_undistortion_matches_oracle(gpu_lib, oracle_mod):
    raw = _raw(50000, 1)
    omega = np.array([0.3, -0.2, 0.5])
    R = scenes.rot_from_rpy(0.1, -0.2, 0.7)
    v = np.array([2.0, -1.0, 0.3])
    g = gpu_lib.LiInitGpu(0.15, max_map_points=1000, max_scan_points=60000)
    g.raw_upload(raw, time_index=9)
    g.raw_undistort_cv(omega, R, v)
    got = g.raw_points()
    want = oracle_mod.undistort_cv(raw[:, :3], raw[:, 9], omega, R, v)
    assert _ulp_close(got, want) and (got != want).mean() < 1e-4      # fp64 sin/cos of the two libms, float store
    first = int(np.argmin(raw[:, 9]))
    assert np.array_equal(got[first], raw[first, :3])                   # the earliest point is never touched (:250)
    assert np.abs(got - raw[:, :3]).max() > 0.1
    g.close()


def test_imu_undistortion_matches_oracle(gpu_lib, oracle_mod):
    raw = _raw(50000, 2)
    rng = np.random.default_rng(3)
    npose = 22                                           # 200 Hz IMU over a 100 ms scan
    poses = np.zeros((npose, 22))
    Rk = scenes.rot_from_rpy(0.02, 0.01, 0.3)
    pos = np.array([5.0, 2.0, 1.0])
    vel = np.array([1.5, 0.2, -0.1])
    for k in range(npose):
        poses[k, 0] = 0.005 * k
        poses[k, 1:4] = rng.normal(0, 0.5, 3)            # acc
        poses[k, 4:7] = rng.normal(0, 0.3, 3)            # gyr
        poses[k, 7:10] = vel
        poses[k, 10:13] = pos
        poses[k, 13:22] = Rk.reshape(9)
        pos = pos + vel * 0.005
        vel = vel + poses[k, 1:4] * 0.005
        Rk = Rk @ scenes.so3_exp(poses[k, 4:7] * 0.005)
    R_LI, T_LI = scenes.sample_extrinsic()
    g = gpu_lib.LiInitGpu(0.15, max_map_points=1000, max_scan_points=60000)
    g.raw_upload(raw, time_index=9)
    g.raw_undistort_imu(poses, Rk, pos, R_LI, T_LI)
    got = g.raw_points()
    want = oracle_mod.undistort_imu(raw[:, :3], raw[:, 9], poses, Rk, pos, R_LI, T_LI)
    assert _ulp_close(got, want) and (got != want).mean() < 1e-4
    untouched = raw[:, 9] <= 0.0
    assert np.array_equal(got[untouched], raw[untouched, :3])           # t_j <= first pose offset: not compensated
    g.close()


def test_raw_front_end_to_icp_pass(gpu_lib, oracle_mod):
    """raw points -> CV undistortion -> voxel grid -> search pass, every stage against the oracle."""
    c = scenes.make_config("C2", N=50000, M=150000, open_air_frac=0.0)
    p = c["pose_init"]
    raw = np.zeros((len(c["body_xyz"]), 12), np.float32)
    raw[:, :3] = c["body_xyz"]
    raw[:, 9] = np.random.default_rng(5).uniform(0, 20.0, len(raw)).astype(np.float32)
    omega, vel = np.array([0.02, -0.01, 0.05]), np.array([0.5, 0.1, 0.0])
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=300000, max_scan_points=60000)
    g.map_build(c["map_xyz"])
    g.raw_upload(raw, time_index=9)
    g.raw_undistort_cv(omega, p.rot_end, vel)
    und = g.raw_points()
    n = g.raw_downsample(0.2)
    body = g.scan_body()
    want_body = oracle_mod.voxel_grid(und, 0.2)          # voxel grid of the DEVICE-undistorted cloud: isolates stage 2
    assert n == len(want_body) and np.array_equal(body, want_body)
    H, b, m, _ = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    om = oracle_mod.OracleMap(c["ds"], 1 if oracle_mod.has_ikd() else 0)
    om.build(c["map_xyz"])
    sc = oracle_mod.OracleScan(body)
    Ho, bo, mo = sc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    assert m == mo and np.abs(H - Ho).max() <= 1e-9 * np.abs(Ho).max() and np.abs(b - bo).max() <= 1e-9 * np.abs(bo).max()
    g.close()
