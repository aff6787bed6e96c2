import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from lidar_imu_init_b200 import scenes, capi
c = scenes.make_config("C2")
p = c["pose_init"]; gt = c["pose_gt"]
g = capi.LiInitGpu(c["ds"], max_map_points=6_000_000, max_scan_points=250_000)
g.map_build(c["map_xyz"]); g.scan_upload(c["body_xyz"])
for k in range(3):
    g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    t = time.perf_counter(); na, nn = g.map_incremental(gt.rot_end, gt.pos_end, gt.R_LI, gt.T_LI, c["ds"]); dt = time.perf_counter() - t
    print(k, f"{dt*1e3:.3f} ms", na, nn, g.map_validnum(), flush=True)
