"""Randomised differential test on the CPU: random maps and update sequences (Build, Add_Points with and without downsampling, box
deletes, points ON voxel boundaries and one ulp beside them, duplicates, far coordinates) through the emulated library (tests/emul: the
product's entry points and kernels compiled for the host) against the verbatim ikd-Tree -- live count, live set, exact 5-NN.
usage: python tools/emul_fuzz.py [n_scenarios] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import liinit_emul as le
from oracle import oracle as orc

# EMUL_FUZZ_COUPLED=0: keep new points that lie in two float boxes (one-ulp overlaps) out of downsample batches, as before the coupled
# boxes were walked in batch order (k_ds_coupled); default: let them in -- the lattice clouds then couple nearly every box
COUPLED = os.environ.get("EMUL_FUZZ_COUPLED", "1") != "0"
n_sc = int(sys.argv[1]) if len(sys.argv) > 1 else 20
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bk = 1 if orc.has_ikd() else 0


def cloud(rng, n, ds, kind, off):
    if kind == 0:      # surfaces: jittered planes
        a = rng.uniform(-6, 6, (n, 3))
        a[:, rng.integers(0, 3)] = rng.choice([-3.0, 0.0, 2.4]) + rng.normal(0, 0.01, n)
    elif kind == 1:    # exactly on the voxel lattice and one ulp around it
        k = rng.integers(-40, 40, (n, 3)).astype(np.float32)
        a = (k * np.float32(ds)).astype(np.float32)
        j = rng.integers(0, 3, n)
        a = np.where(j[:, None] == 0, a, np.where(j[:, None] == 1, np.nextafter(a, np.float32(1e9)), np.nextafter(a, np.float32(-1e9))))
    elif kind == 2:    # dense blob (several points per voxel) + duplicates
        a = rng.normal(0, 0.6, (n, 3))
        a[: n // 10] = a[n // 10: 2 * (n // 10)][: n // 10]
    else:              # uniform volume
        a = rng.uniform(-4, 4, (n, 3))
    return (np.asarray(a, np.float64) + off).astype(np.float32)


def multi_box(pts, ds):
    """new points that lie in a float box other than (or besides) the one of their division cell: [fl(k ds), fl(fl(k ds) + ds)) boxes of
    neighbouring k overlap or leave gaps by one ulp. Within ONE downsample batch such a point also competes in the neighbouring box in the
    reference's sequential walk; the device couples such boxes and walks them in batch order (coupled boxes). With
    EMUL_FUZZ_COUPLED=0 the fuzz keeps these points out of downsample batches, as it had to before that."""
    f = np.float32
    d = f(ds)
    c = np.floor(pts / d).astype(np.float32)
    bad = np.zeros(len(pts), bool)
    for a in range(3):
        x = pts[:, a]
        for k in (-1, 0, 1):
            mn = ((c[:, a] + f(k)) * d).astype(f)
            mx = (mn + d).astype(f)
            inside = (x >= mn) & (x < mx)
            bad |= inside if k != 0 else ~inside
    return bad


def same_knn(g, om, q, tag):
    gx, gd, gc = g.nearest_search(q)
    ox6, od6, oc6, _ = om.knn(q, k=6)
    ox, od, oc = ox6[:, :5], od6[:, :5], np.minimum(oc6, 5)
    assert np.array_equal(gc, oc), (tag, "counts")
    assert np.array_equal(gd, np.where(np.arange(5)[None, :] < oc[:, None], od, -1)), (tag, "distances", np.argwhere(gd != od)[:3])
    # every neighbour returned is at the distance reported for it (reference association, f32)
    d = q[:, None, :] - gx
    dd = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    valid = np.arange(5)[None, :] < gc[:, None]
    assert np.array_equal(dd[valid], gd[valid]), (tag, "reported distances")
    # which of several points at EXACTLY the same distance is returned depends on the traversal order of the reference's tree (also
    # between the 5th and the 6th): rows without such a tie must agree point for point
    d6 = np.where(np.arange(6)[None, :] < oc6[:, None], od6, np.inf)
    untied = np.all(np.diff(d6, axis=1) != 0, axis=1) | (oc6 < 2)
    assert np.array_equal(gx[untied], ox[untied]), (tag, "neighbours")


t0 = time.time()
for sc in range(n_sc):
    rng = np.random.default_rng(seed0 * 1000 + sc)
    ds = float(rng.choice([0.15, 0.2, 0.5]))
    off = rng.choice([0.0, 0.0, 900.0, -5200.0]) * np.array([1.0, -0.7, 0.1])
    index = int(rng.integers(1, 3))   # LIINIT_KNN_BRICKS, LIINIT_KNN_CELLS
    g = le.EmulGpu(ds, max_map_points=60000, max_scan_points=4000, knn_index=index, hash_capacity_log2=13)
    # step-by-step comparison against the oracle's restated tree (deterministic); the verbatim ikd-Tree runs beside it and is compared
    # too, but its background rebuild thread makes its own counters timing dependent once in a while (seen: 941 vs 940 changed voxels
    # for the same input on two runs), so a disagreement of the two ORACLES is reported, not failed
    om = orc.OracleMap(ds, 0)
    ik = orc.OracleMap(ds, 1) if bk else None
    first = cloud(rng, int(rng.integers(200, 6000)), ds, int(rng.integers(0, 4)), off)
    g.map_build(first); om.build(first)
    if ik: ik.build(first)
    ref_disagree = 0
    log = [f"scenario {sc}: ds {ds} index {index} off {off[0]:.0f} build {len(first)}"]
    q = cloud(rng, 400, ds, 3, off)
    for step in range(int(rng.integers(2, 6))):
        op = rng.integers(0, 3)
        if op < 2:
            pts = cloud(rng, int(rng.integers(50, 2500)), ds, int(rng.integers(0, 4)), off)
            down = bool(rng.integers(0, 2))
            if down and not COUPLED:
                pts = np.ascontiguousarray(pts[~multi_box(pts, ds)])
                if len(pts) == 0:
                    continue
            a, b = g.map_add_points(pts, down), om.add_points(pts, down)
            if ik: ref_disagree += int(ik.add_points(pts, down) != b and down)
            log.append(f"add {len(pts)} down={down} -> {a}/{b}")
            if down:
                assert a == b, log
        else:
            lo = rng.uniform(-5, 2, 3) + off
            box = np.concatenate([lo, lo + rng.uniform(0.5, 5, 3)]).astype(np.float32)[None]
            a, b = g.map_delete_boxes(box), om.delete_boxes(box)
            if ik: ref_disagree += int(ik.delete_boxes(box) != b)
            log.append(f"delete -> {a}/{b}")
            assert a == b, log
        assert g.map_validnum() == om.validnum(), log
        same_knn(g, om, q, log)
    live_g, live_o = g.map_download(), om.flatten()
    assert set(map(bytes, live_g)) == set(map(bytes, live_o)), log
    if ik and (ref_disagree or set(map(bytes, ik.flatten())) != set(map(bytes, live_o))):
        print("   note: the verbatim ikd-Tree and the restated tree disagreed in this scenario (reference-side timing)", flush=True)
    g.close()
    print(log[0], "|", len(log) - 1, "updates ok | live", len(live_g), flush=True)
print(f"emul_fuzz: {n_sc} scenarios, no mismatch ({time.time() - t0:.0f} s)")
