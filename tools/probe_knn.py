"""Developer A/B on one box: search-pass variants on the C2 scene (cached in /tmp), warm L2.
  [LIINIT_GPU_LIB=build/variants/x.so] python tools/probe_knn.py --variants 1:4:3:0,1:8:3:0,2:0:3:0
variant = knn_index:group:brick_cells_log2:hash_capacity_log2 (0 = default). Prints the search kernel / plane kernel ms (CUDA events
inside the library) at the initial and at the converged pose and checks that every variant leaves the same per-point state
(neighbours, flags, normals) as the first one and HtH within 1e-12."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lidar_imu_init_b200 import capi
ap = argparse.ArgumentParser()
ap.add_argument("--variants", default="1:0:3:0,1:4:3:0,1:8:3:0")
ap.add_argument("--N", type=int, default=240000)
ap.add_argument("--M", type=int, default=5000000)
ap.add_argument("--config", default="C2")
ap.add_argument("--imu", type=int, default=0)
ap.add_argument("--check", type=int, default=1)
ap.add_argument("--tag", default="")
a = ap.parse_args()
cache = f'/tmp/{a.config}_probe_{a.N}_{a.M}.npz'
if not os.path.exists(cache):
    from lidar_imu_init_b200 import scenes
    c = scenes.make_config(a.config, N=a.N, M=a.M)
    np.savez(cache, map=c["map_xyz"], body=c["body_xyz"], init_R=c["pose_init"].rot_end, init_p=c["pose_init"].pos_end,
             gt_R=c["pose_gt"].rot_end, gt_p=c["pose_gt"].pos_end)
z = np.load(cache)
I, zero = np.eye(3), np.zeros(3)
ref = None
lib = os.path.basename(os.environ.get("LIINIT_GPU_LIB", "default"))
for v in a.variants.split(","):
    idx, grp, bs, hl = (int(x) for x in v.split(":"))
    g = capi.LiInitGpu(0.15, max_map_points=int(a.M * 1.2), max_scan_points=a.N + 10, knn_index=idx, knn_group_lanes=grp, brick_cells_log2=bs,
                       hash_capacity_log2=hl)
    g.map_build(z["map"]); g.scan_upload(z["body"])
    out = [f"{lib} {a.tag} index {idx} G {grp} bs {bs} hash {hl}"]
    for pose in ("init", "gt"):
        R, p = z[pose + "_R"], z[pose + "_p"]
        ks, ts = [], []
        for it in range(12):
            H, b, m, rs = g.icp_iterate(R, p, I, zero, bool(a.imu), True)
            ks.append(g.last_pass_kernel_times()); ts.append(g.last_pass_timing()[0])
        ks, ts = np.array(ks), np.array(ts)
        line = f"{pose}: m={m} knn {np.median(ks[3:, 0]):.4f} (min {ks[:, 0].min():.4f}) plane {np.median(ks[3:, 1]):.4f} pass {np.median(ts[3:]):.4f}"
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
