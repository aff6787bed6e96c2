"""ncu driver: C2 scene, search passes that read the pinned host scan in place (liinit_scan_attach_host)."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import torch
from lidar_imu_init_b200 import scenes, capi
c = scenes.make_config("C2")
N = len(c["body_xyz"])
g = capi.LiInitGpu(c["ds"], max_map_points=6_000_000, max_scan_points=N + 16)
g.map_build(c["map_xyz"])
host = torch.from_numpy(np.ascontiguousarray(c["body_xyz"])).pin_memory()
p = c["pose_init"]
for i in range(4):
    g.scan_attach_ptr(host.data_ptr(), 3, N)
    g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
print("done")
