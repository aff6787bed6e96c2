"""LI-Init stage (include/liinit_calib.h): replay of the reference's own logs + synthetic recovery.

tests/golden/li_init_log.npz holds what the reference's LI_Init object wrote during one real run (Log/*.txt,
result/Initialization_result.txt; tools/make_calib_golden.py). The two *_before_filter files are the inputs of the
temporal + rotational stages; IMU_meas / LiDAR_meas / Lidar_omg_after_rot / the time-stamp columns of acc_cost and the
printed rotation / gyro bias are their outputs. The translation stage cannot be replayed (the odometry's rotation and
velocity were not logged): it is covered by the synthetic streams with known ground truth.
"""
import ctypes
import os
import re

import numpy as np
import pytest

import calib_sim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "li_init_log.npz")


@pytest.fixture(scope="module")
def calib_mod():
    from lidar_imu_init_b200 import _build, calib
    _build.build_calib()
    return calib


def _replay(calib_mod, converge_fully=False):
    g = np.load(GOLD)
    imu, lid = g["imu_before"], g["lidar_before"]
    c = calib_mod.LiCalib(converge_fully=converge_fully)
    eye = np.eye(3)
    for r in imu:
        c.push_imu(r[0:3], r[4:7], r[7])
    for r in lid:
        c.push_lidar(eye, r[0:3], np.zeros(3), r[4])
    # fout_before_filter (LI_init.cpp:43-52) stops one element short of the groups: restore a last element. Its value
    # only reaches the final samples through the filter's mirrored extension (checked below with a looser bound).
    c.push_imu(imu[-1, 0:3], imu[-1, 4:7], imu[-1, 7] + 0.02)
    c.push_lidar(eye, lid[-1, 0:3], np.zeros(3), lid[-1, 4] + 0.02)
    res = c.initialize(10, 5, 0.0, 0.0, from_groups=True)   # orig_odom_freq 10, cut_frame_num 5: 50 Hz rows in the logs
    return g, c, res


def test_header_symbols_exported(calib_mod):
    h = open(os.path.join(ROOT, "include", "liinit_calib.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    names = sorted(set(re.findall(r"\b(li_calib_[a-z0-9_]+)\s*\(", h)))
    assert set(names) == set(calib_mod.SYMBOLS)
    lib = ctypes.CDLL(calib_mod._build.CALIB_LIB)
    for n in names:
        assert hasattr(lib, n)


def test_replay_temporal_stage_matches_reference_logs(calib_mod):
    g, c, res = _replay(calib_mod)
    im, lm = c.log_rows("IMU_meas"), c.log_rows("LiDAR_meas")
    # identical row counts: every pop / cut / alignment of the reference happened at the same sample
    assert im.shape == g["imu_meas"].shape and lm.shape == g["lidar_meas"].shape
    assert res["lag_frames"] == -4 and res["time_lag_1"] == pytest.approx(-0.08, abs=1e-15)
    head = slice(0, 1000)
    # the logs carry 12 significant digits (setprecision(12)): 1e-11 on O(1) rates, 1e-9 on O(10) accelerations /
    # derivatives, 1e-8 on the 3106.x s time stamps
    assert np.abs(im[head, 0:4] - g["imu_meas"][head, 0:4]).max() < 2e-11
    assert np.abs(im[head, 4:7] - g["imu_meas"][head, 4:7]).max() < 1e-10
    assert np.abs(im[head, 7:10] - g["imu_meas"][head, 7:10]).max() < 5e-9
    assert np.abs(im[:, 10] - g["imu_meas"][:, 10]).max() < 1e-8
    assert np.abs(lm[head, 0:4] - g["lidar_meas"][head, 0:4]).max() < 2e-11
    assert np.abs(lm[head, 7:10] - g["lidar_meas"][head, 7:10]).max() < 1e-8
    assert np.abs(lm[:, 10] - g["lidar_meas"][:, 10]).max() < 1e-8
    # tail: influenced by the one group element the reference did not log
    assert np.abs(im[:, 0:10] - g["imu_meas"][:, 0:10]).max() < 1e-5
    assert np.abs(lm[:, [0, 1, 2, 3, 7, 8, 9]] - g["lidar_meas"][:, [0, 1, 2, 3, 7, 8, 9]]).max() < 1e-5


def test_replay_rotation_stage_matches_reference_result(calib_mod):
    g, c, res = _replay(calib_mod)
    # printed with 6 decimals (LI_init.cpp:634-650, result/Initialization_result.txt)
    assert np.abs(res["euler_deg"] - g["printed_euler_deg"]).max() < 6e-7
    assert np.abs(res["gyro_bias"] - g["printed_gyro_bias"]).max() < 6e-7
    assert np.abs(res["R_LI"] - g["printed_T"][:3, :3]).max() < 6e-7
    ar = c.log_rows("Lidar_omg_after_rot")
    assert ar.shape == g["after_rot"].shape
    assert np.abs(ar[:1000, :3] - g["after_rot"][:1000, :3]).max() < 1e-8
    assert np.abs(ar[:, :3] - g["after_rot"][:, :3]).max() < 1e-6
    assert np.abs(ar[:, 3] - g["after_rot"][:, 3]).max() < 1e-8
    # both time shifts (cross-correlation + optimised lag) as they show in the stamps the reference wrote at the very end
    ac = c.log_rows("acc_cost")
    assert ac.shape[0] == g["acc_cost_times"].shape[0]
    assert np.abs(ac[:, 6:8] - g["acc_cost_times"]).max() < 1e-6   # setprecision(10) on 3106.x


def test_reference_schedule_stops_within_tolerance_of_the_optimum(calib_mod):
    _, _, ref = _replay(calib_mod, converge_fully=False)
    _, _, opt = _replay(calib_mod, converge_fully=True)
    for k in ("cost_rot", "cost_rot_bias"):
        assert opt[k] <= ref[k] * (1 + 1e-12)
        assert ref[k] - opt[k] <= 1e-6 * ref[k]            # ceres' function_tolerance
    dR = ref["R_LI"] @ opt["R_LI"].T
    assert np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1)) < 5e-4
    assert np.abs(ref["gyro_bias"] - opt["gyro_bias"]).max() < 5e-5
    assert abs(ref["time_lag_2"] - opt["time_lag_2"]) < 1e-4


def _run_synthetic(calib_mod, S, converge_fully):
    c = calib_mod.LiCalib(converge_fully=converge_fully)
    ti, wi, ai = S["imu"]
    tl, Rl, wl, vl = S["lidar"]
    for k in range(len(ti)):
        c.push_imu_all(wi[k], ai[k], ti[k], 9.81)
    for k in range(len(tl)):
        c.push_lidar(Rl[k], wl[k], vl[k], tl[k])
    return c.initialize(10, 5, 0.0, S["t_move"], from_groups=False)


@pytest.mark.parametrize("converge_fully", [False, True])
@pytest.mark.parametrize("seed,t_off", [(0, 0.013), (3, -0.031)])
def test_synthetic_streams_recover_ground_truth(calib_mod, seed, t_off, converge_fully):
    S = calib_sim.make_streams(seed=seed, t_off=t_off)
    r = _run_synthetic(calib_mod, S, converge_fully)
    dR = r["R_LI"] @ S["R_LI"].T
    assert np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1)) < 1e-3                      # extrinsic rotation [rad]
    assert abs(r["time_lag_1"] + r["time_lag_2"] - t_off) < 1e-3                         # time offset [s]
    assert np.abs(r["gyro_bias"] - S["b_g"]).max() < 5e-4
    assert np.abs(r["T_LI"] - S["T_LI"]).max() < 0.03                                    # extrinsic translation [m]
    g_hat, g_true = r["grav_L0"], S["g_W"]
    assert np.linalg.norm(g_hat) == pytest.approx(9.81, abs=1e-9)
    assert np.arccos(np.clip(g_hat @ g_true / 9.81 ** 2, -1, 1)) < 5e-3                  # gravity direction [rad]
    R_LI_body = r["R_LI"].T @ r["acc_bias"]
    assert np.abs(R_LI_body).max() <= 0.01 + 1e-12                                       # the box of LI_init.cpp:448-451


def test_too_few_samples_is_an_error(calib_mod):
    c = calib_mod.LiCalib()
    for k in range(100):
        c.push_imu(np.zeros(3), np.array([0, 0, 9.81]), 0.02 * k)
        c.push_lidar(np.eye(3), np.zeros(3), np.zeros(3), 0.02 * k)
    with pytest.raises(calib_mod.CalibError):
        c.initialize(10, 5, 0.0, 0.0, from_groups=True)


def test_data_sufficiency_needs_all_three_axes(calib_mod):
    # data_sufficiency_assess (LI_init.cpp:506-563): products of the scaled eigenvalues of sum([w]x^T [w]x)
    c = calib_mod.LiCalib(data_accum_length=300)
    rng = np.random.default_rng(0)
    done_at = None
    for f in range(1, 4000):
        ok, pr = c.data_sufficiency(f, np.array([0.0, 0.0, 1.0]) * rng.uniform(0.5, 1.0), 10, 5)
        assert not ok                                   # rotation about one axis never suffices
    c = calib_mod.LiCalib(data_accum_length=300)
    for f in range(1, 4000):
        ok, pr = c.data_sufficiency(f, rng.uniform(-1, 1, 3), 10, 5)
        if f % 10 != 0:
            assert not ok and (pr < 0).all()            # assessed only when (frame_num % orig_odom_freq) * cut == 0
        if ok:
            done_at = f
            break
    assert done_at is not None and done_at % 10 == 0
    # closed form: H = sum |w|^2 I - w w^T
    c2 = calib_mod.LiCalib(data_accum_length=300)
    H = np.zeros((3, 3))
    W = rng.uniform(-1, 1, (50, 3))
    for f, w in enumerate(W, start=1):
        ok, pr = c2.data_sufficiency(f, w, 10, 5)
        H += (w @ w) * np.eye(3) - np.outer(w, w)
    ev = np.sort(np.linalg.eigvalsh(H) / 300)
    assert np.allclose(np.sort(pr), np.sort([ev[1] * ev[2], ev[0] * ev[2], ev[0] * ev[1]]), rtol=1e-10)
