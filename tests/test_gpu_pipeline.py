"""End to end without ROS: LiDAR-only odometry on the device map (constant-velocity propagation -> hot path ->
map_incremental, 1500 scans at 50 Hz in a synthetic room) feeding the LI-Init batch initialisation; the whole chain is
compared with the ground truth of the simulated rig (tools/lo_calib_pipeline.py)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lidar_odometry_feeds_li_init(gpu_lib):
    from lidar_imu_init_b200 import _build
    _build.build_host()
    _build.build_calib()
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import lo_calib_pipeline
    out = lo_calib_pipeline.run(seconds=30.0, points=4000, seed=0, t_off=0.013, verbose=False)
    assert out["scans"] == 1499
    # the odometry follows the simulated sensor (peak 2 rad/s, 1.5 m/s) scan by scan
    assert out["odo_pos_err"] < 0.10 and out["odo_rot_err"] < 5e-3
    # excitation about all three axes was recognised, initialisation ran on what had been gathered until then
    assert out["sufficient"] and out["n_samples"] > 150
    # calibration from ~5 s of odometry whose angular-velocity state lags the motion (constant-velocity model): coarse,
    # as in the reference, refined online afterwards
    assert out["rot_err"] < 2e-2
    assert abs(out["time_err"]) < 0.03
    assert out["g_angle"] < 0.03
    assert np.abs(out["bg_err"]).max() < 5e-3
    r = out["result"]
    assert np.abs(r["R_LI"].T @ r["acc_bias"]).max() <= 0.01 + 1e-12
