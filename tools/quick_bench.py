"""Developer timing sweep (not the contract bench): kernel ms of search / reuse passes for a few settings."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from lidar_imu_init_b200 import scenes, capi
import argparse
ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=240000)
ap.add_argument("--M", type=int, default=5000000)
ap.add_argument("--groups", default="1,4")
ap.add_argument("--bricks", default="2")
ap.add_argument("--air", type=float, default=0.01)
ap.add_argument("--order", default="voxel")
ap.add_argument("--imu", default="0")
ap.add_argument("--rho", default="3")
a = ap.parse_args()
t = time.time(); c = scenes.make_config("C2", N=a.N, M=a.M, open_air_frac=a.air, order=a.order); print("gen", round(time.time() - t, 2), flush=True)
for bl in [int(x) for x in a.bricks.split(",")]:
  for grp in [int(x) for x in a.groups.split(",")]:
   for rho in [float(x) for x in a.rho.split(",")]:
    g = capi.LiInitGpu(c["ds"], max_map_points=int(a.M * 1.2), max_scan_points=a.N + 10, knn_group_lanes=grp, brick_cells_log2=bl, knn_seed_radius_cells=rho)
    t = time.time(); g.map_build(c["map_xyz"]); tb = time.time() - t
    g.scan_upload(c["body_xyz"])
    for pose_name in ("init", "gt"):
        p = c["pose_" + pose_name]
        for imu in [bool(int(x)) for x in a.imu.split(",")]:
            for search in (True, False):
                ts = []
                for it in range(7):
                    H, b, m, rs = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, imu, search)
                    ts.append(g.last_pass_timing()[0])
                print(f"brick 2^{bl} grp {grp:2d} rho {rho} pose {pose_name:4s} imu {int(imu)} search {int(search)}: m={m} kernel ms {np.median(ts[2:]):.4f} (min {min(ts):.4f})  build {tb*1e3:.1f} ms", flush=True)
    g.close()
