"""Downsample boxes that share an existing point: the float boxes [fl(k ds), fl(fl(k ds) + ds)) of neighbouring cells overlap by one ulp
for some k, a point inside the overlap belongs to two boxes, and when BOTH receive new points in one Add_Points(downsample) batch the
reference's sequential walk (ikd_Tree.cpp:388-426) lets the second box see what the first one deleted. Scene: for twenty such cell pairs
(k-1, k) along x, 5 km from the origin, an existing point e in the overlap, a new point a at the centre of box k-1 (beats e there: e is
deleted) and a new point b in a corner of box k (loses against e -- if e is still there). Batch order decides what happens to b."""
import numpy as np

DS = 0.2


def overlapping_cells(ds, k0, n=200):
    """cells k (from k0 on) whose float box is overlapped by the box below by one ulp -> [(k, lower end of box k)]"""
    f = np.float32
    out = []
    for k in range(k0, k0 + n):
        mn = f(f(k) * f(ds))
        mx_prev = f(f(f(k - 1) * f(ds)) + f(ds))
        if mx_prev > mn:
            out.append((k, mn))
    return out


def scene():
    """-> base map (the shared points + filler), and three batches: all a then all b, all b then all a, interleaved"""
    f = np.float32
    cells = overlapping_cells(DS, -26050)
    assert len(cells) > 40
    y0, z0 = f(3640.1), f(-520.1)                      # the centre of a cell on the other two axes
    shared, news_a, news_b = [], [], []
    for k, mn in cells[:40:2]:                         # every other overlapping cell: the pairs stay apart from each other
        shared.append(np.array([mn, y0, z0], f))       # in box k (x >= mn) and in box k-1 (x < its upper end); 0.1 m from either centre
        news_a.append(np.array([mn - f(0.5 * DS), y0, z0], f))
        news_b.append(np.array([mn + f(0.19), y0 + f(0.09), z0 + f(0.09)], f))
    shared, news_a, news_b = np.array(shared), np.array(news_a), np.array(news_b)
    rng = np.random.default_rng(4)
    filler = (rng.uniform(-3, 3, (500, 3)) + np.array([-5200.0, 3640.0, -520.0])).astype(f)
    base = np.concatenate([shared, filler])
    batches = [np.concatenate([news_a, news_b]), np.concatenate([news_b, news_a]), np.stack([news_a, news_b], 1).reshape(-1, 3)]
    return base, batches
