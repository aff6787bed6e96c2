"""Lockstep cost model of the cell-directory search (developer tool, CPU only).

The free-running thread-per-point kernel measured 5.1 of 32 lanes active (profiles/r01_cells): lanes that leave a loop
iteration early never meet the others again before the loop exit. This model answers what a WARP-UNIFORM version would
cost: every loop runs for the maximum trip count over the warp's lanes, inactive lanes predicated off. Input: the
structure traces of knn5_boxes from the CPU checker (tests/cells_emul.py); instruction costs per block from the SASS
of the built kernel. Output: warp instructions per query for a few loop shapes, to be compared with the 780 warp
instructions per query of the lockstep brick search.
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

C = dict(fixed=150, round=110, sup=45, brick_pre=75, probe=30, found=30, cell=25, pt=13, ins=36, adv=15)


def _c2_cached():
    """C2 map + world-frame queries at both poses (cached under /tmp: the 5M-point scene takes a few seconds to generate)"""
    files = ['/tmp/c2_map.npy', '/tmp/c2_q_pose_init.npy', '/tmp/c2_q_pose_gt.npy']
    if not all(os.path.exists(f) for f in files):
        from lidar_imu_init_b200 import scenes
        c = scenes.make_config("C2")
        np.save(files[0], c["map_xyz"])
        for name, f in (("pose_init", files[1]), ("pose_gt", files[2])):
            p = c[name]
            w = (p.rot_end @ (p.R_LI @ c["body_xyz"].T.astype(np.float64) + p.T_LI[:, None]) + p.pos_end[:, None]).T.astype(np.float32)
            np.save(f, w)
    return np.load(files[0])


def parse(tr):
    """tokens of one query -> [round][super][brick] = (status, [(npts, insbits), ...])"""
    rounds = []
    i, n = 0, len(tr)
    while i < n:
        t = tr[i]
        if t == -1:
            rounds.append([]); i += 1
        elif t == -2:
            rounds[-1].append([]); i += 1
        elif t == -3:
            rounds[-1][-1].append([tr[i + 1], []]); i += 2
        elif t == -4:
            rounds[-1][-1][-1][1].append((tr[i + 1], tr[i + 2])); i += 3
        else:
            raise ValueError(t)
    return rounds


def cost_points_nested(cells):
    """cells: list over lanes of (npts, insbits) for the same cell iteration"""
    mx = max(c[0] for c in cells)
    tot = 0
    for j in range(mx):
        tot += C["pt"] + (C["ins"] if any((c[1] >> j) & 1 for c in cells if j < c[0]) else 0)
    return tot


def cost_brick_nested(bricks):
    """bricks: lanes' (status, cells) for the same brick iteration; cells nested: per cell iteration max points"""
    tot = C["brick_pre"]
    if any(b[0] >= 1 for b in bricks): tot += C["probe"]
    f = [b for b in bricks if b[0] == 2]
    if f:
        tot += C["found"]
        for k in range(max(len(b[1]) for b in f)):
            cs = [b[1][k] for b in f if k < len(b[1])]
            tot += C["cell"] + cost_points_nested(cs)
    return tot


def cost_brick_flat(bricks):
    """cells + points of a brick flattened into one loop: one candidate per lane and iteration, range advance predicated"""
    tot = C["brick_pre"]
    if any(b[0] >= 1 for b in bricks): tot += C["probe"]
    f = [b for b in bricks if b[0] == 2]
    if f:
        tot += C["found"]
        seqs = []
        for b in f:
            s = []
            for (npts, bits) in b[1]:
                s += [(bits >> j) & 1 for j in range(npts)]
            seqs.append(s)
        for j in range(max(len(s) for s in seqs)):
            tot += C["adv"] + C["pt"] + (C["ins"] if any(s[j] for s in seqs if j < len(s)) else 0)
    return tot


def cost_rounds(lanes, brick_cost, first=0, last=None):
    """lanes: parsed queries of one warp. Rounds [first, last) in lockstep."""
    tot = 0
    nr = max(len(q) for q in lanes)
    for r in range(first, nr if last is None else min(nr, last)):
        act = [q[r] for q in lanes if r < len(q)]
        if not act: break
        tot += C["round"]
        for s in range(max(len(a) for a in act)):
            sup = [a[s] for a in act if s < len(a)]
            tot += C["sup"]
            for b in range(max(len(x) for x in sup) if sup else 0):
                br = [x[b] for x in sup if b < len(x)]
                tot += brick_cost(br)
    return tot


def lane_instr(q):
    t = C["fixed"]
    for rnd in q:
        t += C["round"]
        for sup in rnd:
            t += C["sup"]
            for st, cells in sup:
                t += C["brick_pre"] + (C["probe"] if st >= 1 else 0) + (C["found"] if st == 2 else 0)
                for npts, bits in cells:
                    t += C["cell"] + npts * C["pt"] + bin(bits).count("1") * C["ins"]
    return t


def main():
    import cells_emul as ce
    mp = _c2_cached()
    E = ce.CellsEmul(mp, 0.15, hash_log2=22)
    W = int(os.environ.get("WARPS", "1500"))
    for name in ("pose_init", "pose_gt"):
        q = np.load(f'/tmp/c2_q_{name}.npy')
        # sample whole warps (32 consecutive queries) spread over the scan
        nw = len(q) // 32
        pick = np.linspace(0, nw - 1, W).astype(int)
        idx = (pick[:, None] * 32 + np.arange(32)[None, :]).ravel()
        for rho in (0.3, 0.45):
            tr, off = E.trace(q[idx], rho=rho)
            Q = [parse(tr[off[i]:off[i + 1]].tolist()) for i in range(len(idx))]
            lane = np.array([lane_instr(x) for x in Q])
            res = {}
            for label, bc in (("nested", cost_brick_nested), ("flat", cost_brick_flat)):
                tot = sum(C["fixed"] + cost_rounds(Q[w * 32:(w + 1) * 32], bc) for w in range(W))
                res[label] = tot / (W * 32)
            # two-phase: round 0 for everybody; the rest only for unfinished queries, compacted into new warps (in order)
            p1 = sum(C["fixed"] + cost_rounds(Q[w * 32:(w + 1) * 32], cost_brick_flat, 0, 1) for w in range(W))
            rest = [x for x in Q if len(x) > 1]
            p2 = sum(C["fixed"] + cost_rounds(rest[k:k + 32], cost_brick_flat, 1, None) for k in range(0, len(rest), 32))
            print(f"{name} rho {rho}: lane-instr/query {lane.mean():.0f} (ideal warp-instr/query {lane.mean()/32:.0f}) | lockstep nested {res['nested']:.0f} "
                  f"flat {res['flat']:.0f} | two-phase {p1/(W*32):.0f} + {p2/(W*32):.0f} = {(p1+p2)/(W*32):.0f} (unfinished after round 1: {len(rest)/len(Q):.2f})  [brick search: 580 at the initial pose, measured]", flush=True)


if __name__ == "__main__":
    main()


def breakdown():
    """how many lockstep iterations of each level a warp runs (nested shape), per query"""
    import cells_emul as ce
    mp = _c2_cached()
    E = ce.CellsEmul(mp, 0.15, hash_log2=22)
    W = 800
    for name in ("pose_init", "pose_gt"):
        q = np.load(f'/tmp/c2_q_{name}.npy')
        nw = len(q) // 32
        pick = np.linspace(0, nw - 1, W).astype(int)
        idx = (pick[:, None] * 32 + np.arange(32)[None, :]).ravel()
        tr, off = E.trace(q[idx], rho=0.3)
        Q = [parse(tr[off[i]:off[i + 1]].tolist()) for i in range(len(idx))]
        cnt = dict(round=0, sup=0, brick=0, probe=0, found=0, cell=0, pt=0, ins=0)
        lanecnt = dict(round=0, sup=0, brick=0, probe=0, found=0, cell=0, pt=0, ins=0)
        for x in Q:
            for rnd in x:
                lanecnt["round"] += 1
                for sup in rnd:
                    lanecnt["sup"] += 1
                    for st, cells in sup:
                        lanecnt["brick"] += 1; lanecnt["probe"] += st >= 1; lanecnt["found"] += st == 2
                        for npts, bits in cells:
                            lanecnt["cell"] += 1; lanecnt["pt"] += npts; lanecnt["ins"] += bin(bits).count("1")
        for w in range(W):
            lanes = Q[w * 32:(w + 1) * 32]
            for r in range(max(len(x) for x in lanes)):
                act = [x[r] for x in lanes if r < len(x)]
                cnt["round"] += 1
                for s in range(max(len(a) for a in act)):
                    sup = [a[s] for a in act if s < len(a)]
                    cnt["sup"] += 1
                    for b in range(max(len(y) for y in sup)):
                        br = [y[b] for y in sup if b < len(y)]
                        cnt["brick"] += 1
                        cnt["probe"] += any(z[0] >= 1 for z in br)
                        f = [z for z in br if z[0] == 2]
                        if f:
                            cnt["found"] += 1
                            for k in range(max(len(z[1]) for z in f)):
                                cs = [z[1][k] for z in f if k < len(z[1])]
                                cnt["cell"] += 1
                                mx = max(c[0] for c in cs)
                                cnt["pt"] += mx
                                cnt["ins"] += sum(any((c[1] >> j) & 1 for c in cs if j < c[0]) for j in range(mx))
        print(name, "lockstep iterations per WARP:", {k: round(v / W, 1) for k, v in cnt.items()})
        print(name, "lane iterations per QUERY   :", {k: round(v / (W * 32), 1) for k, v in lanecnt.items()})
        print(name, "warp-instr per query by level:", {k: round(cnt[k] / W * C[{'round':'round','sup':'sup','brick':'brick_pre','probe':'probe','found':'found','cell':'cell','pt':'pt','ins':'ins'}[k]] / 32) for k in cnt})


if __name__ == "__main__" and os.environ.get("BREAKDOWN"):
    breakdown()


def stream_cost(lanes, QC=32, K=dict(round=130, helper=45, brick=75, probe=30, found=30, cell=35, vote=9, drain0=8, it=35, ins=25)):
    """enumerate + stream (knn5_stream): warp instructions of one warp (list of parsed queries)"""
    tot = 150
    nr_rounds = max(len(q) for q in lanes)
    for r in range(nr_rounds):
        act = [q[r] for q in lanes if r < len(q)]
        tot += K["round"]
        # per lane: flat list of bricks of the round (super boundaries only add helper iterations)
        bl = []
        for a in act:
            bricks = []
            for sup in a:
                for j, b in enumerate(sup):
                    bricks.append((b, j == 0))   # first brick of a super: the helper loop ran
            bl.append(bricks)
        queue = [[] for _ in act]   # queued (npts, bits) per lane

        def drain():
            nonlocal tot
            tot += K["drain0"]
            seqs = []
            for qq in queue:
                s = []
                for npts, bits in qq:
                    s += [(bits >> j) & 1 for j in range(npts)]
                seqs.append(s)
                qq.clear()
            n = max((len(s) for s in seqs), default=0)
            for j in range(n):
                tot += K["it"] + (K["ins"] if any(s[j] for s in seqs if j < len(s)) else 0)

        for step in range(max(len(b) for b in bl) + 1):   # +1: the step in which the last lane finds no more bricks
            cur = [(i, b[step]) for i, b in enumerate(bl) if step < len(b)]
            tot += K["helper"] + K["brick"]
            if any(x[1][0][0] >= 1 for x in cur): tot += K["probe"]
            f = [(i, x[0]) for i, x in cur if x[0][0] == 2]
            if f:
                tot += K["found"]
                for k in range(max(len(b[1]) for _, b in f)):
                    tot += K["cell"] + K["vote"]
                    for i, b in f:
                        if k < len(b[1]): queue[i].append(b[1][k])
                    if any(len(qq) >= QC for qq in queue): drain()
            else:
                tot += K["vote"]
        drain()
    return tot


def stream_main():
    import cells_emul as ce
    mp = _c2_cached()
    E = ce.CellsEmul(mp, 0.15, hash_log2=22)
    W = int(os.environ.get("WARPS", "1500"))
    for name in ("pose_init", "pose_gt"):
        q = np.load(f'/tmp/c2_q_{name}.npy')
        nw = len(q) // 32
        pick = np.linspace(0, nw - 1, W).astype(int)
        idx = (pick[:, None] * 32 + np.arange(32)[None, :]).ravel()
        tr, off = E.trace(q[idx], rho=0.3)
        Q = [parse(tr[off[i]:off[i + 1]].tolist()) for i in range(len(idx))]
        for QC in (16, 32, 64, 10**6):
            tot = sum(stream_cost(Q[w * 32:(w + 1) * 32], QC) for w in range(W))
            print(f"{name}: enumerate+stream QC={QC}: {tot/(W*32):.0f} warp-instr/query -> {tot/(W*32)*240000/1e6:.0f} M per pass", flush=True)


if __name__ == "__main__" and os.environ.get("STREAM"):
    stream_main()
