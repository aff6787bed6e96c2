"""Developer probe on one box: pass latency against frame size for the search variants (C2 scene and map, the frame is a prefix
sample of the 240k-point scan). Prints per N: search kernel / plane kernel ms (events inside the library) and the wall clock of
liinit_icp_iterate (launch + kernels + 1.28 kB back + sync).   python tools/probe_small.py --sizes 2000,8000,30000 --variants 1:4,1:8,1:32,1:0"""
import os, sys, argparse, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lidar_imu_init_b200 import capi
ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="1000,2000,4000,8000,16000,30000,60000,120000,240000")
ap.add_argument("--variants", default="1:2,1:4,1:8,1:16,1:32,1:0")
a = ap.parse_args()
cache = '/tmp/C2_probe_240000_5000000.npz'
if not os.path.exists(cache):
    from lidar_imu_init_b200 import scenes
    c = scenes.make_config("C2")
    np.savez(cache, map=c["map_xyz"], body=c["body_xyz"], init_R=c["pose_init"].rot_end, init_p=c["pose_init"].pos_end,
             gt_R=c["pose_gt"].rot_end, gt_p=c["pose_gt"].pos_end)
z = np.load(cache)
I, zero = np.eye(3), np.zeros(3)
sizes = [int(x) for x in a.sizes.split(",")]
for v in a.variants.split(","):
    idx, grp = (int(x) for x in v.split(":"))
    g = capi.LiInitGpu(0.15, max_map_points=6_000_000, max_scan_points=250_000, knn_index=idx, knn_group_lanes=grp)
    g.map_build(z["map"])
    for n in sizes:
        step = max(1, len(z["body"]) // n)
        body = np.ascontiguousarray(z["body"][::step][:n])
        g.scan_upload(body)
        out = []
        for pose in ("init", "gt"):
            R, p = z[pose + "_R"], z[pose + "_p"]
            ks, ws = [], []
            for it in range(14):
                t = time.perf_counter()
                H, b, m, rs = g.icp_iterate(R, p, I, zero, False, True)
                ws.append((time.perf_counter() - t) * 1e3)
                ks.append(g.last_pass_kernel_times())
            ks = np.array(ks)
            out.append(f"{pose}: knn {np.median(ks[4:, 0]):.4f} plane {np.median(ks[4:, 1]):.4f} wall {np.median(ws[4:]):.4f} m={m}")
        print(f"index {idx} G {grp} N {len(body):6d} | " + " | ".join(out), flush=True)
    g.close()
