"""A query in a hollow of a densely filled volume, placed next to a brick corner: every shell up to the last one ([1.44, 5] m^2) finds
fewer than five points, and the last one then finds 44 non-empty bricks among the 64 it enumerates -- more than one probing round of a
32-lane group may list if the list is sized for the other group widths (the defect this scene was built for: with 32 lanes per point
-- the choice for frames up to 14k points -- the group's brick list in shared memory held 32 entries and a round probes 64 bricks)."""
import numpy as np

DS = 0.15
BRICK = 8 * DS


def hollow_map_and_queries(n_queries=2):
    q0 = np.array([0.05, 0.07, -0.03])
    ax = np.arange(-3.6 + 0.05, 3.6, 0.25)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    p = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)
    p = (p + np.random.default_rng(2).uniform(-0.02, 0.02, p.shape)).astype(np.float32)
    lo = np.floor(p / np.float32(BRICK)) * BRICK
    e = np.maximum(0, np.maximum(lo - q0, q0 - (lo + BRICK)))
    mp = p[(e * e).sum(1) >= 1.44]                       # nothing in any brick closer than 1.2 m to the query
    rng = np.random.default_rng(3)
    qs = (q0 + rng.uniform(-0.03, 0.03, (n_queries, 3))).astype(np.float32)
    qs[0] = q0
    return mp, qs


def brute_force_sets(live, qs):
    """per query: the set of (up to five) nearest live points within d^2 <= 5, float32 arithmetic as the reference's"""
    out = []
    for q in qs:
        d = ((live - q) ** 2).astype(np.float32)
        dd = (d[:, 0] + d[:, 1]) + d[:, 2]
        o = np.argsort(dd, kind="stable")[:5]
        out.append(set(map(bytes, live[o][dd[o] <= 5])))
    return out
