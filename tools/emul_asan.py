"""AddressSanitizer over EVERY kernel and entry point, on the CPU: builds the emulated library (tests/emul/make_liinit_emul.py) with
-fsanitize=address,undefined and drives build / search + reuse pass / map_incremental / Add_Points / box delete / nearest search / voxel grid /
download through it for both spatial indexes. Complements compute-sanitizer on the GPU (profiles/r01_sanitizer.txt).
usage:  python tools/emul_asan.py          (re-executes itself with libasan preloaded)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
ASAN_LIB = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
OUT = "/tmp/libliinit_emul_asan.so"

if os.environ.get("LI_EMUL_ASAN_CHILD") != "1":
    import liinit_emul as le
    le._mk.build(force=False)            # generates tests/emul/_gen/liinit_gpu_emul.cpp
    gen = os.path.join(le._mk.GEN, "liinit_gpu_emul.cpp")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
                           "-Wno-attributes", "-Wno-unknown-pragmas", "-DLI_SIMT_EMUL=1", "-I", le._mk.HERE, "-I", le._mk.CSRC,
                           "-I", os.path.join(ROOT, "include"), "-I", le._mk.CUDA_INC, gen, "-o", OUT])
    env = dict(os.environ, LD_PRELOAD=ASAN_LIB, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0", LI_EMUL_ASAN_CHILD="1")
    sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)], env=env))

import numpy as np
import liinit_emul as le
from lidar_imu_init_b200 import scenes
le._mk.build = lambda force=False: OUT
c = scenes.make_config("C2", N=1500, M=30000, open_air_frac=0.02)
p, gt = c["pose_init"], c["pose_gt"]
for index in (1, 2):
    for search in ((0,) if index == 1 else (1, 2, 3)):
        os.environ["LIINIT_CELLS_SEARCH"] = str(search or 3)
        g = le.EmulGpu(c["ds"], max_map_points=100000, max_scan_points=3000, knn_index=index, hash_capacity_log2=14)
        g.map_build(c["map_xyz"])
        g.scan_upload(c["body_xyz"])
        H, b, m, rs = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, True, True)
        g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, True, False)
        na, nn = g.map_incremental(gt.rot_end, gt.pos_end, gt.R_LI, gt.T_LI, c["ds"])
        g.map_add_points(c["body_xyz"][:500] + 50.0, True)
        g.map_add_points(c["body_xyz"][500:900] + 50.0, False)
        g.map_delete_boxes(np.array([[-1, -1, -1, 30, 30, 30]], np.float32))
        g.nearest_search(c["body_xyz"][:300])
        n = g.scan_upload_raw(np.repeat(c["body_xyz"][:800], 3, 0) + np.random.default_rng(1).normal(0, .05, (2400, 3)).astype(np.float32), 0.5)
        live = g.map_download()
        print(f"index {index} search {search}: m={m} map_incremental={na},{nn} voxel-grid leaves={n} live={len(live)}", flush=True)
        g.close()
# Add_Points(downsample) on lattice-aligned clouds far from the origin: nearly every box is coupled with a neighbour (one-ulp overlaps of
# the float boxes), the batch is walked by k_ds_coupled
def lattice(rng, n, ds, off):
    k = rng.integers(-40, 40, (n, 3)).astype(np.float32)
    a = (k * np.float32(ds)).astype(np.float32)
    j = rng.integers(0, 3, n)
    a = np.where(j[:, None] == 0, a, np.where(j[:, None] == 1, np.nextafter(a, np.float32(1e9)), np.nextafter(a, np.float32(-1e9))))
    return (a.astype(np.float64) + off).astype(np.float32)
rng = np.random.default_rng(5)
off = np.array([-5200.0, 3640.0, -520.0])
for index in (1, 2):
    g = le.EmulGpu(0.2, max_map_points=60000, max_scan_points=100, knn_index=index, hash_capacity_log2=13)
    g.map_build(lattice(rng, 3000, 0.2, off))
    ch = [g.map_add_points(lattice(rng, 1500, 0.2, off), True) for _ in range(3)]
    print(f"index {index}: lattice downsample batches changed {ch}, live {g.map_validnum()}", flush=True)
    g.close()
# the lockstep search at every group width: first pass, seeded later pass, in-place host read, compaction, and the hollow scene whose
# last shell lists 44 bricks in one probing round (tests/hollow_case.py)
import hollow_case as hc
hmap, hq = hc.hollow_map_and_queries(40)
I3, z3 = np.eye(3), np.zeros(3)
for group in (0, 2, 4, 8, 16, 32):
    g = le.EmulGpu(c["ds"], max_map_points=100000, max_scan_points=3000, knn_group_lanes=group, hash_capacity_log2=14)
    g.map_build(c["map_xyz"])
    g.scan_attach(c["body_xyz"][:1200])
    g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    _, _, m2, _ = g.icp_iterate(gt.rot_end, gt.pos_end, gt.R_LI, gt.T_LI, False, True)     # seeded by the first pass's neighbours
    g.map_delete_boxes(np.array([[-1, -1, -1, 30, 30, 30]], np.float32))
    g.map_compact()
    g.icp_iterate(gt.rot_end, gt.pos_end, gt.R_LI, gt.T_LI, False, True)
    g.close()
    g = le.EmulGpu(hc.DS, max_map_points=80000, max_scan_points=100, knn_group_lanes=group)
    g.map_build(hmap)
    g.scan_upload(hq)
    g.icp_iterate(I3, z3, I3, z3, False, True)
    cnt = g.scan_state()["near_cnt"]
    print(f"group {group}: seeded pass m={m2}, hollow scene neighbours found {int(cnt.min())}..{int(cnt.max())}", flush=True)
    g.close()
print("emul_asan: no AddressSanitizer / UndefinedBehaviorSanitizer report")
