"""The LiDAR-only branch of the node's main loop (src/laserMapping.cpp:893-1234) over the C-ABI: per scan
constant-velocity propagation -> ICP / IESKF update on the device map -> map_incremental -> LI-Init data accumulation.

Host-side mirror used by the replay tool and the end-to-end test; every step is one call into the native libraries
(libliinit_host.so: liinit_propagate_cv, liinit_scan_update; libliinit_gpu.so: liinit_scan_upload / liinit_map_*;
libliinit_calib.so: li_calib_*). Nothing is computed in Python.
"""
from __future__ import annotations

import numpy as np

from . import capi, host
from .calib import LiCalib


class LidarOdometry:
    def __init__(self, gpu: capi.LiInitGpu, filter_size_map: float, max_iteration: int = 4, gyr_cov: float = 0.1, acc_cov: float = 0.1,
                 orig_odom_freq: int = 10, cut_frame_num: int = 1, data_accum_length: float = 300.0, calib: LiCalib | None = None):
        self.g = gpu
        self.ds = float(filter_size_map)
        self.max_iteration = int(max_iteration)            # "max_iteration", laserMapping.cpp:767
        self.gyr_cov, self.acc_cov = gyr_cov, acc_cov      # mapping/gyr_cov, mapping/acc_cov (:776-777)
        self.orig_odom_freq, self.cut_frame_num = int(orig_odom_freq), int(cut_frame_num)
        self.state = host.state_init()
        self.calib = calib if calib is not None else LiCalib(data_accum_length)
        self.map_ready = False
        self.first_prop = True
        self.t_last = None
        self.frame_num = 0
        self.data_accum_start = False
        self.data_accum_finished = False
        self.move_start_time = 0.0
        self.stats = None
        self.imu_en = False

    def push_imu(self, omg, acc, t, mean_acc_norm=9.81):
        """imu_cbk in LO mode (:428-430)."""
        if not self.data_accum_finished:
            self.calib.push_imu_all(omg, acc, t, mean_acc_norm)

    def process_scan(self, body_xyz: np.ndarray, t_beg: float, t_end: float):
        """One pass of the loop body for a downsampled, undistorted scan in the LiDAR frame. Returns the posterior state."""
        # p_imu->Process in LO mode: Forward_propagation_without_imu (IMU_Processing.hpp:212-243)
        if self.first_prop:
            dt = 0.1
            self.first_prop = False
        else:
            dt = t_beg - self.t_last
        self.t_last = t_beg
        self.state = host.propagate_cv(self.state, dt, self.gyr_cov, self.acc_cov)
        R, p, RLI, TLI = host.state_pose(self.state)
        if not self.map_ready:                              # :921-931
            if len(body_xyz) > 5:
                world = (R @ (RLI @ body_xyz.T.astype(np.float64) + TLI[:, None]) + p[:, None]).T.astype(np.float32)
                self._map_build(world)
                self.map_ready = True
            return self.state
        self.state, self.stats = self._scan_update(body_xyz, self.state)   # :936-1134
        R, p, RLI, TLI = host.state_pose(self.state)
        self._map_incremental(R, p, RLI, TLI)               # :1140
        if not self.data_accum_start and np.linalg.norm(p) > 0.05:   # :1145-1149
            self.data_accum_start = True
            self.move_start_time = t_end
        self.frame_num += 1                                 # :1160
        if not self.data_accum_finished and self.data_accum_start:   # :1190-1196
            bias_g, vel = self.state[27:30], self.state[24:27]
            self.calib.push_lidar(R, bias_g, vel, t_end)
            ok, _ = self.calib.data_sufficiency(self.frame_num, bias_g, self.orig_odom_freq, self.cut_frame_num)
            self.data_accum_finished = ok
        return self.state

    # the three calls into the hot path; a test substitutes the CPU oracle here to drive the SAME loop from the other side
    def _map_build(self, world):
        self.g.map_build(world)

    def _scan_update(self, body_xyz, state):
        self.g.scan_upload(body_xyz)
        return host.scan_update(self.g, state, self.max_iteration, self.imu_en)

    def _map_incremental(self, R, p, RLI, TLI):
        self.g.map_incremental(R, p, RLI, TLI, self.ds)

    def hand_over(self, res):
        """laserMapping.cpp:1198-1222 after LI_Initialization: the pose is re-expressed in the IMU frame with the calibrated extrinsic,
        gravity / biases are taken over and the filter continues with the 12-column Jacobian (imu_en = true)."""
        R, p, _, _ = host.state_pose(self.state)
        R_LI, T_LI = np.asarray(res["R_LI"], float).reshape(3, 3), np.asarray(res["T_LI"], float)
        s = self.state
        s[12:21] = R_LI.reshape(9)                          # offset_R_L_I = Init_LI->get_R_LI() (:1204)
        s[21:24] = T_LI                                     # offset_T_L_I (:1205)
        s[9:12] = p - R @ R_LI.T @ T_LI                     # pos_end = -rot_end * R_LI^T * T_LI + pos_end (:1206)
        s[0:9] = (R @ R_LI.T).reshape(9)                    # rot_end = rot_end * R_LI^T (:1207)
        s[33:36] = np.asarray(res["grav_L0"], float)        # gravity (:1208)
        s[27:30] = np.asarray(res["gyro_bias"], float)      # bias_g (:1209)
        s[30:33] = np.asarray(res["acc_bias"], float)       # bias_a (:1210)
        self.imu_en = True

    def initialize(self, timediff_imu_wrt_lidar: float = 0.0):
        """LI_Initialization on what has been accumulated (:1198); returns the calibration dict of calib.LiCalib."""
        return self.calib.initialize(self.orig_odom_freq, self.cut_frame_num, timediff_imu_wrt_lidar, self.move_start_time, from_groups=False)
