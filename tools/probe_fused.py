"""Developer A/B on one box: fused search pass (knn_index 4) against the two-kernel brick search (knn_index 1) on the C2 scene.
  LIINIT_GPU_LIB=build/dev/libliinit_gpu.so python tools/probe_fused.py [--variants 1:4:3,4:4:3,4:8:3,4:4:2]
variant = index:group:brick_cells_log2. Prints pass ms (CUDA events inside the library) at both poses, and checks that every
variant leaves the same per-point state (neighbours, flags, normals) and HtH within 1e-12 of the first one."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lidar_imu_init_b200 import capi
ap = argparse.ArgumentParser()
ap.add_argument("--variants", default="1:4:3,4:4:3,4:8:3,4:4:2,4:8:2,1:4:2")
ap.add_argument("--N", type=int, default=240000)
ap.add_argument("--M", type=int, default=5000000)
ap.add_argument("--imu", type=int, default=0)
ap.add_argument("--check", type=int, default=1)
a = ap.parse_args()
cache = f'/tmp/c2_probe_{a.N}_{a.M}.npz'
if not os.path.exists(cache):
    from lidar_imu_init_b200 import scenes
    c = scenes.make_config("C2", N=a.N, M=a.M)
    np.savez(cache, map=c["map_xyz"], body=c["body_xyz"], init_R=c["pose_init"].rot_end, init_p=c["pose_init"].pos_end,
             gt_R=c["pose_gt"].rot_end, gt_p=c["pose_gt"].pos_end)
z = np.load(cache)
I, zero = np.eye(3), np.zeros(3)
ref = None
for v in a.variants.split(","):
    idx, grp, bs = (int(x) for x in v.split(":"))
    g = capi.LiInitGpu(0.15, max_map_points=int(a.M * 1.2), max_scan_points=a.N + 10, knn_index=idx, knn_group_lanes=grp, brick_cells_log2=bs)
    g.map_build(z["map"]); g.scan_upload(z["body"])
    out = [f"index {idx} G {grp} bs {bs}"]
    for pose in ("init", "gt"):
        R, p = z[pose + "_R"], z[pose + "_p"]
        ts = []
        for it in range(12):
            H, b, m, rs = g.icp_iterate(R, p, I, zero, bool(a.imu), True)
            ts.append(g.last_pass_timing()[0])
        ts = np.array(ts)
        line = f"{pose}: m={m} pass {np.median(ts[3:]):.4f} (min {ts.min():.4f})"
        if pose == "init" and a.check:
            st = g.scan_state()
            if ref is None:
                ref = (H.copy(), b.copy(), m, st)
            else:
                same = all(np.array_equal(st[k], ref[3][k]) for k in ("world", "near_cnt", "near_xyz", "selected"))
                sel = ref[3]["selected"].astype(bool)
                same = same and np.array_equal(st["normvec"][sel], ref[3]["normvec"][sel])
                line += f" [m same {m == ref[2]}, state same {same}, HtH rel {np.abs(H - ref[0]).max() / np.abs(ref[0]).max():.1e}]"
        out.append(line)
    print(" | ".join(out), flush=True)
    g.close()
