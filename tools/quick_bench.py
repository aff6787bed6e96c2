import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from lidar_imu_init_b200 import scenes, capi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 240000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 5000000
t = time.time(); c = scenes.make_config("C2", N=N, M=M); print("gen", time.time() - t, flush=True)
p = c["pose_init"]
for tile in (8, 16, 32):
    g = capi.LiInitGpu(c["ds"], max_map_points=int(M * 1.2), max_scan_points=N + 10, knn_tile=tile)
    t = time.time(); g.map_build(c["map_xyz"]); print("build s", time.time() - t, g.map_stats(), flush=True)
    g.scan_upload(c["body_xyz"])
    for imu in (False, True):
        for search in (True, False):
            ts = []
            for it in range(8):
                H, b, m, rs = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, imu, search)
                ts.append(g.last_pass_timing()[0])
            print(f"tile {tile} imu {imu} search {search}: m={m} kernel ms {np.median(ts[2:]):.4f} (min {min(ts):.4f})", flush=True)
    g.close()
