"""End-to-end parity over a 60 s run (north star: "same per-scan poses and final calibrated extrinsic / time offset within 1e-3 m /
1e-3 rad on identical synthetic input"): the node's LiDAR-only loop (laserMapping.cpp:893-1234 -- constant-velocity propagation,
ICP / IESKF update, map_incremental, LI-Init data accumulation, LI_Initialization, hand-over to the 12-column LIO mode) is driven
TWICE over the same 2999 scans and the same IMU stream, in lockstep:

  * product:  lidar_imu_init_b200.odometry.LidarOdometry -- liinit_scan_update (C++ IESKF loop over the C-ABI) on the device map,
              liinit_map_incremental on the device;
  * oracle:   the same loop with the three hot-path calls replaced by the CPU oracle (verbatim ikd-Tree + the restated
              laserMapping.cpp:936-1134 / :516-559, literal m-wide IESKF).

Both feed the same LI-Init library (host side, row N4). Asserted per scan: pose difference <= 1e-3 m / 1e-3 rad (measured: ~1e-9);
the data-sufficiency trigger fires at the same scan; the two calibration results (R_LI, T_LI, time lag, gravity, biases) agree
within 1e-3; after the hand-over the 12-column passes (imu_en = 1) agree likewise, including the online-refined extrinsic."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _angle(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)))


def make_oracle_odometry(orc, ds, nthreads, **kw):
    from lidar_imu_init_b200.odometry import LidarOdometry

    class OracleOdometry(LidarOdometry):
        """LidarOdometry with the hot path on the CPU oracle (test infrastructure)."""

        def __init__(self):
            super().__init__(None, ds, **kw)
            self.om = orc.OracleMap(ds, 1 if orc.has_ikd() else 0)
            self.sc = None

        def _map_build(self, world):
            self.om.build(world)

        def _scan_update(self, body_xyz, state):
            self.sc = orc.OracleScan(body_xyz)
            st, iters, searches, m = self.sc.scan_update(self.om, state, self.max_iteration, self.imu_en, nthreads=nthreads)
            return st, dict(iterations=iters, search_passes=searches, effect_feat_num=m)

        def _map_incremental(self, R, p, RLI, TLI):
            self.sc.map_incremental(self.om, R, p, RLI, TLI, self.ds)

    return OracleOdometry()


def run_lockstep(gpu_lib, orc, seconds=60.0, points=4000, seed=0, t_off=0.013, ds=0.15, lio_scans=None, verbose=False, odometry_cls=None):
    import calib_sim
    from lidar_imu_init_b200 import host, scenes
    from lidar_imu_init_b200.odometry import LidarOdometry

    S = calib_sim.make_streams(seed=seed, duration=seconds, t_off=t_off)
    tr = calib_sim.Trajectory(seed)
    scene = scenes.box_scene(40.0, 25.0, 6.0, n_slabs_x=2, n_slabs_y=1)
    p0 = np.array([15.0, 12.0, 2.2])
    eye, zero = np.eye(3), np.zeros(3)
    kw = dict(max_iteration=4, orig_odom_freq=10, cut_frame_num=5)
    g = gpu_lib.LiInitGpu(ds, max_map_points=3_000_000, max_scan_points=points + 16) if gpu_lib is not None else None
    lo_g = (odometry_cls or LidarOdometry)(g, ds, **kw) if g is not None else None
    lo_o = make_oracle_odometry(orc, ds, min(16, os.cpu_count() or 1), **kw)
    ti, wi, ai = S["imu"]
    tl = S["lidar"][0]
    t0 = 100.0
    k_imu = 0
    out = dict(max_dp=0.0, max_dr=0.0, scans=0, init_scan=None, lio=dict(max_dp=0.0, max_dr=0.0, max_dT=0.0, max_dRLI=0.0, scans=0))
    res_g = res_o = None
    prev_true = None
    for j, t_end in enumerate(tl):
        while k_imu < len(ti) and ti[k_imu] <= t_end:
            for lo in (lo_g, lo_o):
                if lo is not None:
                    lo.push_imu(wi[k_imu], ai[k_imu], ti[k_imu])
            k_imu += 1
        Rt, pt = tr.R(t_end - t0), tr.pos(t_end - t0)
        pose = scenes.Pose(Rt, p0 + pt, eye, zero)                 # the LiDAR's true pose in the scene
        body = scenes.scan_points(scene, pose, points, seed=1000 + j, det_range=60.0, sigma=0.01, open_air_frac=0.0)
        if res_o is None:
            # ---- LiDAR-only mode: the node's own propagation + update -------------------------------------------
            st_o = lo_o.process_scan(body, t_end - 0.02, t_end)
            if lo_g is not None:
                st_g = lo_g.process_scan(body, t_end - 0.02, t_end)
                Rg, pg, _, _ = host.state_pose(st_g)
                Ro, po, _, _ = host.state_pose(st_o)
                out["max_dp"] = max(out["max_dp"], float(np.abs(pg - po).max()))
                out["max_dr"] = max(out["max_dr"], _angle(Rg, Ro))
                assert out["max_dp"] <= 1e-3 and out["max_dr"] <= 1e-3, (j, out)
                if lo_g.stats is not None:
                    # the two runs differ at round-off level (~1e-9 in the pose): a point sitting exactly on a gate (s > 0.9, the 0.1 m plane
                    # test, the convergence thresholds) may fall on either side -- counted, bounded, not required to be zero
                    out["dm_max"] = max(out.get("dm_max", 0), abs(lo_g.stats["effect_feat_num"] - lo_o.stats["effect_feat_num"]))
                    out["iter_diff_scans"] = out.get("iter_diff_scans", 0) + int(lo_g.stats["iterations"] != lo_o.stats["iterations"])
                assert lo_g.data_accum_finished == lo_o.data_accum_finished, j      # the trigger fires at the same scan
            out["scans"] += 1
            if lo_o.data_accum_finished:
                # ---- LI_Initialization + hand-over (laserMapping.cpp:1197-1222) ---------------------------------------
                res_o = lo_o.initialize(0.0)
                lo_o.hand_over(res_o)
                if lo_g is not None:
                    res_g = lo_g.initialize(0.0)
                    lo_g.hand_over(res_g)
                out["init_scan"] = j
        else:
            # ---- LIO mode (imu_en = 1, 12-column Jacobian). The reference's prior comes from IMU propagation (host side, out of the
            # path). Its stand-in here, identical on both sides: each side's OWN previous posterior moved by the true motion increment of
            # the IMU frame between the two scans (dead reckoning, what an ideal IMU integrates to), pose covariance re-inflated a little;
            # extrinsic + its covariance carried by the filter (online refinement, config/avia.yaml:18-19). Closed loop, like the LO leg.
            # (A prior re-drawn around the truth every scan -- 0.2 deg / 3 cm -- leaves the 4 iterations unconverged and makes the loop
            # chaotic at the 0.5 mm level: the oracle run against ITSELF with a 1e-11 m disturbance then drifts apart by 5e-4 m.) ------
            if lio_scans is not None and out["lio"]["scans"] >= lio_scans:
                break
            R_I = pose.rot_end @ S["R_LI"].T
            p_I = pose.pos_end - R_I @ S["T_LI"]
            dR, dp = prev_true[0].T @ R_I, prev_true[0].T @ (p_I - prev_true[1])
            for lo in (lo_g, lo_o):
                if lo is None:
                    continue
                st = lo.state
                Rp, pp = st[0:9].reshape(3, 3).copy(), st[9:12].copy()
                st[0:9] = (Rp @ dR).reshape(9)
                st[9:12] = pp + Rp @ dp
                c = st[36:].reshape(24, 24)
                c[0:3, 0:3] += np.eye(3) * 1e-6
                c[3:6, 3:6] += np.eye(3) * 1e-5
                lo.state, lo.stats = lo._scan_update(body, st)
                R, p, RLI, TLI = host.state_pose(lo.state)
                lo._map_incremental(R, p, RLI, TLI)
            if lo_g is not None:
                Rg, pg, RLg, TLg = host.state_pose(lo_g.state)
                Ro, po, RLo, TLo = host.state_pose(lo_o.state)
                L = out["lio"]
                L["max_dp"] = max(L["max_dp"], float(np.abs(pg - po).max()))
                L["max_dr"] = max(L["max_dr"], _angle(Rg, Ro))
                L["max_dT"] = max(L["max_dT"], float(np.abs(TLg - TLo).max()))
                L["max_dRLI"] = max(L["max_dRLI"], _angle(RLg, RLo))
                assert max(L["max_dp"], L["max_dr"], L["max_dT"], L["max_dRLI"]) <= 1e-3, (j, L)
                L["dm_max"] = max(L.get("dm_max", 0), abs(lo_g.stats["effect_feat_num"] - lo_o.stats["effect_feat_num"]))
            out["lio"]["scans"] += 1
        prev_true = (pose.rot_end @ S["R_LI"].T, pose.pos_end - (pose.rot_end @ S["R_LI"].T) @ S["T_LI"])
        if verbose and j % 200 == 0:
            print(j, out["max_dp"], out["max_dr"], out["lio"], flush=True)
    out.update(res_g=res_g, res_o=res_o, truth=S, map_points_oracle=lo_o.om.validnum(), map_points_gpu=g.map_validnum() if g is not None else None)
    if g is not None:
        g.close()
    return out


@pytest.mark.gpu
def test_sixty_second_run_matches_oracle_end_to_end(gpu_lib, oracle_mod):
    from lidar_imu_init_b200 import _build
    _build.build_host()
    _build.build_calib()
    out = run_lockstep(gpu_lib, oracle_mod, seconds=60.0, points=4000)
    assert out["init_scan"] is not None and out["scans"] > 200            # LI-Init fired after a few seconds of motion; the rest of the 60 s runs in LIO mode
    assert out["max_dp"] <= 1e-3 and out["max_dr"] <= 1e-3                 # per-scan poses, LiDAR-only mode (measured ~1e-9)
    assert out.get("dm_max", 0) <= 3 and out.get("iter_diff_scans", 0) <= out["scans"] // 100   # gate flips at round-off level only
    rg, ro = out["res_g"], out["res_o"]
    assert _angle(rg["R_LI"], ro["R_LI"]) <= 1e-3                          # final calibrated extrinsic ...
    assert np.abs(rg["T_LI"] - ro["T_LI"]).max() <= 1e-3
    assert abs((rg["time_lag_1"] + rg["time_lag_2"]) - (ro["time_lag_1"] + ro["time_lag_2"])) <= 1e-3    # ... and time offset
    assert np.abs(rg["grav_L0"] - ro["grav_L0"]).max() <= 1e-3 and np.abs(rg["gyro_bias"] - ro["gyro_bias"]).max() <= 1e-3
    assert np.abs(rg["acc_bias"] - ro["acc_bias"]).max() <= 1e-3
    L = out["lio"]
    assert L["scans"] == 2999 - out["init_scan"] - 1 and L["scans"] > 2000 and max(L["max_dp"], L["max_dr"], L["max_dT"], L["max_dRLI"]) <= 1e-3   # 12-column leg incl. refined extrinsic
    assert L.get("dm_max", 0) <= 3
    print("e2e parity:", {k: v for k, v in out.items() if k not in ("res_g", "res_o", "truth")})
    # (one gate flip in 3000 scans moves a couple of map points; the maps are otherwise the same)
    assert abs(out["map_points_gpu"] - out["map_points_oracle"]) <= 1e-3 * out["map_points_oracle"]
    # and the calibration is the one the simulated rig has (coarse: the constant-velocity odometry lags the motion)
    S = out["truth"]
    assert _angle(rg["R_LI"], S["R_LI"]) < 2e-2 and abs(rg["time_lag_1"] + rg["time_lag_2"] - S["t_off"]) < 0.03


def test_oracle_side_of_the_loop_runs_on_cpu(oracle_mod):
    """The oracle half of the lockstep harness (no GPU): a short run reaches the hand-over and keeps tracking in LIO mode."""
    from lidar_imu_init_b200 import _build
    _build.build_gpu()
    _build.build_host()     # (libliinit_host.so provides propagate_cv / state algebra on the host; no device call is made)
    _build.build_calib()
    out = run_lockstep(None, oracle_mod, seconds=12.0, points=1500, lio_scans=20)
    assert out["init_scan"] is not None and out["lio"]["scans"] == 20


def test_product_side_of_the_loop_on_the_cpu_build(oracle_mod):
    """The PRODUCT half of the lockstep harness without a GPU: LidarOdometry (constant-velocity propagation, liinit_scan_update -- the C++
    IESKF loop --, map_incremental) on the CPU build of the library (tests/emul: the same kernels and host code compiled for the host,
    liinit_host.cpp linked against it) against the oracle-driven loop, scan after scan. A checker of logic for -m "not gpu" runs; short
    (the emulated kernels are 10^3 times slower), so it stays in the LiDAR-only leg."""
    import liinit_emul as le
    if not le.available():
        pytest.skip("g++ or the CUDA vector-type headers are missing")
    from lidar_imu_init_b200 import _build
    from lidar_imu_init_b200.odometry import LidarOdometry
    _build.build_gpu()
    _build.build_host()
    _build.build_calib()

    class EmulOdometry(LidarOdometry):
        def _scan_update(self, body_xyz, state):
            self.g.scan_upload(body_xyz)
            return le.scan_update(self.g, state, self.max_iteration, self.imu_en)

    class Lib:
        @staticmethod
        def LiInitGpu(ds, max_map_points, max_scan_points):
            return le.EmulGpu(ds, max_map_points=200000, max_scan_points=max_scan_points)

    out = run_lockstep(Lib, oracle_mod, seconds=0.8, points=300, odometry_cls=EmulOdometry)
    assert out["scans"] >= 35 and out["init_scan"] is None
    assert out["max_dp"] <= 1e-6 and out["max_dr"] <= 1e-6, out
    assert out.get("dm_max", 0) <= 1 and out["map_points_gpu"] == out["map_points_oracle"]
