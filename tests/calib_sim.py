"""Synthetic LiDAR-odometry + IMU streams with known extrinsic, time offset, biases and gravity (test infrastructure
for the LI-Init stage, row N4). Rigid-body kinematics only; noise-free unless asked."""
import numpy as np


def _rot(axis, a):
    c, s = np.cos(a), np.sin(a)
    if axis == 0:
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    if axis == 1:
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def _log(R):
    c = np.clip((np.trace(R) - 1) / 2, -1, 1)
    th = np.arccos(c)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return v * (0.5 if th < 1e-9 else th / (2 * np.sin(th)))


def exp_so3(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


class Trajectory:
    """LiDAR pose in the odometry frame W (= first LiDAR frame): still for t < t_move, then sums of sinusoids."""

    def __init__(self, seed=0, t_move=4.0, rot_amp=0.6, pos_amp=0.6):
        r = np.random.default_rng(seed)
        self.t_move = t_move
        self.fa = r.uniform(0.25, 0.9, (3, 3)); self.pa = r.uniform(0, 2 * np.pi, (3, 3)); self.aa = r.uniform(0.4, 1.0, (3, 3)) * rot_amp / 3
        self.fp = r.uniform(0.2, 0.8, (3, 3)); self.pp = r.uniform(0, 2 * np.pi, (3, 3)); self.ap = r.uniform(0.4, 1.0, (3, 3)) * pos_amp / 3

    def _ramp(self, t):   # C^3 onset of the motion over 2 s
        u = np.clip((t - self.t_move) / 2.0, 0, 1)
        return u ** 4 * (35 - 84 * u + 70 * u ** 2 - 20 * u ** 3)

    def _angles(self, t):
        s = t - self.t_move
        return self._ramp(t) * np.sum(self.aa * (np.sin(2 * np.pi * self.fa * s + self.pa) - np.sin(self.pa)), axis=1)

    def pos(self, t):
        s = t - self.t_move
        return self._ramp(t) * np.sum(self.ap * (np.sin(2 * np.pi * self.fp * s + self.pp) - np.sin(self.pp)), axis=1)

    def R(self, t):
        a = self._angles(t)
        return _rot(2, a[2]) @ _rot(1, a[1]) @ _rot(0, a[0])

    def omega_body(self, t, h=1e-5):
        return _log(self.R(t - h).T @ self.R(t + h)) / (2 * h)

    def alpha_body(self, t, h=1e-4):
        return (self.omega_body(t + h) - self.omega_body(t - h)) / (2 * h)

    def vel(self, t, h=1e-5):
        return (self.pos(t + h) - self.pos(t - h)) / (2 * h)

    def acc(self, t, h=1e-4):
        return (self.pos(t + h) - 2 * self.pos(t) + self.pos(t - h)) / (h * h)


def make_streams(seed=0, duration=34.0, imu_hz=200.0, lidar_hz=50.0, t_off=0.013, R_LI=None, T_LI=None, b_g=None, b_a=None,
                 tilt=(0.05, -0.08), gyro_noise=0.0, acc_noise=0.0):
    """Returns dict(truth..., imu=(t, omg, acc), lidar=(t, R, omg, vel), t_move). IMU stamps = true time + t_off."""
    rng = np.random.default_rng(seed + 1)
    tr = Trajectory(seed)
    R_LI = _rot(2, 1.5) @ _rot(1, -0.03) @ _rot(0, 0.02) if R_LI is None else R_LI
    T_LI = np.array([-0.02, 0.03, 0.15]) if T_LI is None else np.asarray(T_LI, float)   # LiDAR origin in the IMU frame
    T_IL = -R_LI.T @ T_LI                                                                # IMU origin in the LiDAR frame
    b_g = np.array([0.002, -0.001, 0.0015]) if b_g is None else np.asarray(b_g, float)
    b_a = np.array([0.004, -0.006, 0.005]) if b_a is None else np.asarray(b_a, float)     # IMU frame
    R_GL0 = _rot(1, tilt[1]) @ _rot(0, tilt[0])
    g_W = R_GL0 @ np.array([0, 0, -9.81])
    t0 = 100.0
    ti = t0 + np.arange(int(duration * imu_hz)) / imu_hz
    omg_i, acc_i = np.zeros((len(ti), 3)), np.zeros((len(ti), 3))
    for k, t in enumerate(ti):
        tau = t - t0 - t_off
        w, al, R = tr.omega_body(tau), tr.alpha_body(tau), tr.R(tau)
        Wx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        Ax = np.array([[0, -al[2], al[1]], [al[2], 0, -al[0]], [-al[1], al[0], 0]])
        a_imu_W = tr.acc(tau) + R @ (Wx @ Wx + Ax) @ T_IL
        omg_i[k] = R_LI @ w + b_g
        acc_i[k] = R_LI @ R.T @ (a_imu_W - g_W) + b_a
    omg_i += gyro_noise * rng.standard_normal(omg_i.shape)
    acc_i += acc_noise * rng.standard_normal(acc_i.shape)
    tl = t0 + 0.5 / lidar_hz + np.arange(int(duration * lidar_hz) - 1) / lidar_hz
    Rl = np.array([tr.R(t - t0) for t in tl])
    wl = np.array([tr.omega_body(t - t0) for t in tl])
    vl = np.array([tr.vel(t - t0) for t in tl])
    return dict(R_LI=R_LI, T_LI=T_LI, b_g=b_g, b_a=b_a, g_W=g_W, t_off=t_off, imu=(ti, omg_i, acc_i), lidar=(tl, Rl, wl, vl),
                t_move=t0 + tr.t_move, lidar_hz=lidar_hz)
