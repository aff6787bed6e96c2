"""ncu target: a few search passes of one variant at the C2 initial pose (index:group:bs from argv[1], default 4:4:3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lidar_imu_init_b200 import capi
v = sys.argv[1] if len(sys.argv) > 1 else "4:4:3"
idx, grp, bs = (int(x) for x in v.split(":"))
z = np.load('/tmp/c2_probe_240000_5000000.npz')
g = capi.LiInitGpu(0.15, max_map_points=6_000_000, max_scan_points=250_000, knn_index=idx, knn_group_lanes=grp, brick_cells_log2=bs)
g.map_build(z["map"]); g.scan_upload(z["body"])
I, zero = np.eye(3), np.zeros(3)
for it in range(6):
    g.icp_iterate(z["init_R"], z["init_p"], I, zero, False, True)
g.close()
