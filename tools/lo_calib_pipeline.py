"""End-to-end replay without ROS: synthetic room + moving sensor -> LiDAR-only odometry on the GPU map (hot path) ->
LI-Init batch initialisation on the host -> extrinsic / time offset / gravity against the ground truth.

  python tools/lo_calib_pipeline.py [--seconds 30] [--points 4000]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(seconds=30.0, points=4000, seed=0, t_off=0.013, ds=0.15, verbose=True, stop_when_sufficient=False):
    import calib_sim
    from lidar_imu_init_b200 import capi, scenes
    from lidar_imu_init_b200.odometry import LidarOdometry

    S = calib_sim.make_streams(seed=seed, duration=seconds, t_off=t_off)
    tr = calib_sim.Trajectory(seed)
    scene = scenes.box_scene(40.0, 25.0, 6.0, n_slabs_x=2, n_slabs_y=1)
    p0 = np.array([15.0, 12.0, 2.2])
    eye, zero = np.eye(3), np.zeros(3)
    g = capi.LiInitGpu(ds, max_map_points=3_000_000, max_scan_points=points + 16)
    lo = LidarOdometry(g, ds, max_iteration=4, orig_odom_freq=10, cut_frame_num=5)
    ti, wi, ai = S["imu"]
    tl = S["lidar"][0]
    t0 = 100.0
    k_imu = 0
    err_p = err_r = 0.0
    wall = time.time()
    n_scans = 0
    for j, t_end in enumerate(tl):
        while k_imu < len(ti) and ti[k_imu] <= t_end:
            lo.push_imu(wi[k_imu], ai[k_imu], ti[k_imu])
            k_imu += 1
        Rt, pt = tr.R(t_end - t0), tr.pos(t_end - t0)
        pose = scenes.Pose(Rt, p0 + pt, eye, zero)
        body = scenes.scan_points(scene, pose, points, seed=1000 + j, det_range=60.0, sigma=0.01, open_air_frac=0.0)
        st = lo.process_scan(body, t_end - 0.02, t_end)
        n_scans += 1
        R, p = st[0:9].reshape(3, 3), st[9:12]
        err_p = max(err_p, float(np.abs(p - pt).max()))
        err_r = max(err_r, float(np.arccos(np.clip((np.trace(R.T @ Rt) - 1) / 2, -1, 1))))
        if stop_when_sufficient and lo.data_accum_finished:
            break
    wall = time.time() - wall
    res = lo.initialize(0.0)
    dR = res["R_LI"] @ S["R_LI"].T
    out = dict(scans=n_scans, wall_s=wall, map_points=g.map_validnum(), odo_pos_err=err_p, odo_rot_err=err_r,
               sufficient=lo.data_accum_finished, n_samples=res["n_samples"],
               rot_err=float(np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1))), T_err=res["T_LI"] - S["T_LI"],
               time_err=res["time_lag_1"] + res["time_lag_2"] - S["t_off"], bg_err=res["gyro_bias"] - S["b_g"],
               g_angle=float(np.arccos(np.clip(res["grav_L0"] @ S["g_W"] / 9.81 ** 2, -1, 1))), result=res, truth=S)
    if verbose:
        for k in ("scans", "wall_s", "map_points", "odo_pos_err", "odo_rot_err", "sufficient", "n_samples", "rot_err", "T_err", "time_err", "bg_err",
                  "g_angle"):
            print(f"{k:14s} {out[k]}")
    g.close()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--points", type=int, default=4000)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    run(a.seconds, a.points, a.seed)
