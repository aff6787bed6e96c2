"""Tiny driver for ncu captures: C2 scene, a few search + reuse passes."""
import sys
sys.path.insert(0, '/root/repo')
import argparse
from lidar_imu_init_b200 import scenes, capi
ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=240000)
ap.add_argument("--M", type=int, default=5000000)
ap.add_argument("--group", type=int, default=0)
ap.add_argument("--brick", type=int, default=0)
ap.add_argument("--imu", type=int, default=0)
ap.add_argument("--passes", type=int, default=4)
ap.add_argument("--rho", type=float, default=0.0)
ap.add_argument("--index", type=int, default=0)
a = ap.parse_args()
c = scenes.make_config("C2", N=a.N, M=a.M)
g = capi.LiInitGpu(c["ds"], max_map_points=int(a.M * 1.2), max_scan_points=a.N + 10, knn_group_lanes=a.group, brick_cells_log2=a.brick, knn_seed_radius_cells=a.rho, knn_index=a.index)
g.set_reseed(False)   # every search pass from scratch: the captures show the FIRST search pass of a scan (the bench metric)
g.map_build(c["map_xyz"])
g.scan_upload(c["body_xyz"])
p = c["pose_init"]
for i in range(a.passes):
    g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, bool(a.imu), True)
    g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, bool(a.imu), False)
print("done")
