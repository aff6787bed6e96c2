"""Golden vectors for the LI-Init stage (row N4) from the reference's OWN logs of one real run.

The reference ships, under Log/ and result/, what its LI_Init object wrote during a real initialisation
(include/LI_init/LI_init.cpp:43-52, :135-139, :151-157, :397-400, :477-484, :634-650):
  IMU_before_filter.txt / Lidar_before_filter.txt   the state groups entering the filters (all but their last element)
  IMU_meas.txt / LiDAR_meas.txt                     after filter 1, cross-correlation shift, central differences
  Lidar_omg_after_rot.txt                           R_LI * w_L + b_g after the rotation / gyro-bias / time-lag solve
  acc_cost.txt                                      last two columns: IMU and LiDAR time stamps after both time shifts
  result/Initialization_result.txt                  the printed calibration
These are data, not source: they are stored (float64, compressed) as tests/golden/li_init_log.npz so the CPU suite can
replay the stage anywhere. Run here, where /root/reference is mounted:  python tools/make_calib_golden.py
"""
import os
import re

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "li_init_log.npz")


def main():
    L = lambda n: np.loadtxt(os.path.join(REF, "Log", n))
    txt = open(os.path.join(REF, "result", "Initialization_result.txt")).read()
    init = txt.split("Refinement result")[0]
    num = r"(-?\d+\.\d+)"

    def grab(label, k):
        m = re.search(re.escape(label) + r"\s*=\s*" + r"\s+".join([num] * k), init)
        return np.array([float(x) for x in m.groups()])

    hom = re.search(r"Homogeneous Transformation Matrix from LiDAR to IMU:\s*\n((?:.*\n){4})", init).group(1)
    H = np.array([[float(x) for x in row.split()] for row in hom.strip().splitlines()])
    acc_cost = L("acc_cost.txt")
    np.savez_compressed(
        OUT,
        imu_before=L("IMU_before_filter.txt"), lidar_before=L("Lidar_before_filter.txt"),
        imu_meas=L("IMU_meas.txt"), lidar_meas=L("LiDAR_meas.txt"), after_rot=L("Lidar_omg_after_rot.txt"),
        acc_cost_times=acc_cost[:, 6:8],
        printed_euler_deg=grab("Rotation LiDAR to IMU (degree)", 3), printed_T_LI=grab("Translation LiDAR to IMU (meter)", 3),
        printed_gyro_bias=grab("Bias of Gyroscope  (rad/s)", 3), printed_acc_bias=grab("Bias of Accelerometer (meters/s^2)", 3),
        printed_gravity=grab("Gravity in World Frame(meters/s^2)", 3), printed_time_lag=grab("Time Lag IMU to LiDAR (second)", 1),
        printed_T=H)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
