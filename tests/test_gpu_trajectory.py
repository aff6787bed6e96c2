"""End-to-end per-scan parity on the GPU: host C++ driver (liinit_scan_update) + device map updates vs the oracle's
restatement of laserMapping.cpp:936-1134 + :516-559 over a multi-scan synthetic trajectory.
Bar (north star): per-scan poses within 1e-3 m / 1e-3 rad; we assert 1e-6."""
import numpy as np
import pytest

from lidar_imu_init_b200 import scenes

pytestmark = pytest.mark.gpu


def _angle(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1) / 2
    return float(np.arccos(np.clip(c, -1, 1)))


@pytest.mark.parametrize("imu_en", [False, True])
def test_scan_sequence_matches_oracle(gpu_lib, oracle_mod, imu_en):
    from lidar_imu_init_b200 import _build, host
    _build.build_host()
    ds = 0.15
    scene = scenes.box_scene(40.0, 25.0, 6.0, n_slabs_x=2, n_slabs_y=1)
    R_LI, T_LI = scenes.sample_extrinsic() if imu_en else scenes.identity_extrinsic()
    rng = np.random.default_rng(7)
    # trajectory: 12 scans, ~0.25 m and ~1.5 deg per scan
    poses = []
    pos = np.array([12.0, 9.0, 1.4])
    yaw = 0.3
    for k in range(12):
        R = scenes.rot_from_rpy(0.02 * np.sin(k), 0.015 * np.cos(k), yaw)
        poses.append(scenes.Pose(R, pos.copy(), R_LI, T_LI))
        pos = pos + np.array([0.22, 0.11, 0.01])
        yaw += np.deg2rad(1.5)
    scans = [scenes.scan_points(scene, p, 3500, seed=100 + k, det_range=60.0, sigma=0.01, open_air_frac=0.01) for k, p in enumerate(poses)]

    g = gpu_lib.LiInitGpu(ds, max_map_points=500000, max_scan_points=8000)
    bk = 1 if oracle_mod.has_ikd() else 0
    om = oracle_mod.OracleMap(ds, bk)
    # first scan initialises the map (laserMapping.cpp:921-931)
    w0 = (poses[0].rot_end @ (R_LI @ scans[0].T.astype(np.float64) + T_LI[:, None]) + poses[0].pos_end[:, None]).T.astype(np.float32)
    g.map_build(w0)
    om.build(w0)
    st_g = host.state_from_pose(poses[0].rot_end, poses[0].pos_end, R_LI, T_LI)
    cov = np.eye(24) * 1e-5
    cov[0:3, 0:3] = np.eye(3) * 1e-3      # attitude prior
    cov[3:6, 3:6] = np.eye(3) * 1e-2      # position prior
    cov[6:9, 6:9] = np.eye(3) * 5e-5      # Rot_LI_cov (config/avia.yaml:18)
    cov[9:12, 9:12] = np.eye(3) * 1e-5    # Trans_LI_cov (config/avia.yaml:19)
    st_g[36:] = cov.reshape(-1)
    st_o = st_g.copy()
    max_dp = max_dr = 0.0
    for k in range(1, len(poses)):
        # motion prior (stands in for IMU / constant-velocity propagation, IMU_Processing.hpp): the true pose
        # disturbed by 0.2 deg / 3 cm; extrinsic and covariance carried over, pose covariance re-inflated
        prior = scenes.perturb_pose(poses[k], 500 + k, dtheta_deg=0.2, dpos=0.03)
        for st in (st_g, st_o):
            st[0:9] = prior.rot_end.reshape(9)
            st[9:12] = prior.pos_end
            c = st[36:].reshape(24, 24)
            c[0:3, 0:3] += np.eye(3) * 1e-4
            c[3:6, 3:6] += np.eye(3) * 1e-3
        g.scan_upload(scans[k])
        st_g, stats = host.scan_update(g, st_g, 5, imu_en)
        sc = oracle_mod.OracleScan(scans[k])
        st_o, iters, searches, m = sc.scan_update(om, st_o, 5, imu_en)
        assert stats["iterations"] == iters and stats["search_passes"] == searches and stats["effect_feat_num"] == m
        Rg, pg, RLg, TLg = host.state_pose(st_g)
        Ro, po, RLo, TLo = host.state_pose(st_o)
        max_dp = max(max_dp, float(np.abs(pg - po).max()), float(np.abs(TLg - TLo).max()))
        max_dr = max(max_dr, _angle(Rg, Ro), _angle(RLg, RLo))
        na, nn = g.map_incremental(Rg, pg, RLg, TLg, ds)
        _, oa, on, _ = sc.map_incremental(om, Ro, po, RLo, TLo, ds)
        assert (na, nn) == (oa, on)
        assert g.map_validnum() == om.validnum()
        # the estimate tracks the ground truth: LiDAR pose in the world = (R_end R_LI, R_end T_LI + pos)
        gt_R, gt_p = poses[k].rot_end @ R_LI, poses[k].rot_end @ T_LI + poses[k].pos_end
        assert np.abs(Rg @ TLg + pg - gt_p).max() < 0.02 and _angle(Rg @ RLg, gt_R) < 0.002
    assert max_dp < 1e-6 and max_dr < 1e-6, (max_dp, max_dr)
    assert np.allclose(st_g[36:], st_o[36:], rtol=1e-3, atol=1e-7)   # covariance
    g.close()
