"""CPU model of the slot-aligned shell's step count: for 300 warp tasks of the C2 scene (8 scan points each,
4 lanes per point) the number of candidate steps a warp executes -- per list slot the longest slab of the eight groups, rounded up to
batches of LI_KNN_U = 3 -- for the brick lists as enumerated, sorted by slab size, sorted by distance, against the ideal (all lanes busy);
and the number of lane-private top-5 inserts for the three orders. numpy only, no GPU. usage: python tools/sim_slot_steps.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_imu_init_b200 import scenes
import bench
t=time.time()
c = bench.make_workload("C2", 1, 240000, 5000000)
print('workload', time.time()-t)
ds = c["ds"]; B = 8*ds
mp = c["map_xyz"].astype(np.float32)
p = c["pose_init"]
body = c["body_xyz"]
world = (p.rot_end @ (p.R_LI @ body.T.astype(np.float64) + p.T_LI[:, None]) + p.pos_end[:, None]).T.astype(np.float32)
bk = np.floor(mp / np.float32(B)).astype(np.int64)
off = 1<<20
key = ((bk[:,0]+off)<<42) | ((bk[:,1]+off)<<21) | (bk[:,2]+off)
order = np.argsort(key, kind='stable')
ks = key[order]; mps = mp[order]
uk, first, cnt = np.unique(ks, return_index=True, return_counts=True)
print('bricks', len(uk), 'mean cnt', cnt.mean(), 'median', np.median(cnt), 'p90', np.percentile(cnt,90))
lut = dict(zip(uk.tolist(), range(len(uk))))
rng = np.random.default_rng(0)
G=4; Q=8; U=3
def steps(c):  # steps for a slab of c points at G lanes: batches of U
    per = -(-c//G)
    return U * (-(-per//U)) if c>0 else 0
def shell_bricks(q, lo2, hi2, thr):
    r = np.sqrt(min(hi2,5.0))
    lo = np.floor((q - r)/B).astype(int); hi = np.floor((q + r)/B).astype(int)
    out=[]
    for kz in range(lo[2],hi[2]+1):
      for ky in range(lo[1],hi[1]+1):
        for kx in range(lo[0],hi[0]+1):
            bl = np.array([kx,ky,kz])*B; bh = bl+B
            e = np.maximum(0, np.maximum(bl-q, q-bh)); d = float((e*e).sum())
            if d>=lo2 and d<hi2 and d<thr:
                k = ((kx+off)<<42)|((ky+off)<<21)|(kz+off)
                i = lut.get(k)
                if i is not None: out.append((i,d))
    return out
tot={'unsorted':0,'count_desc':0,'near_first':0,'ideal':0}
ins={'unsorted':0,'count_desc':0,'near_first':0}
nw=300
starts = rng.integers(0, len(world)//Q - 1, nw)*Q
def count_inserts(q, bricks_in_order, ld):
    # lane-private lists: G lanes; returns number of lane-inserts and updated lists
    n=0
    for (bi,_) in bricks_in_order:
        pts = mps[first[bi]:first[bi]+cnt[bi]]
        d = ((pts-q)**2).sum(1)
        for j,dj in enumerate(d):
            l = j % G
            if dj < ld[l][4] and dj<=5:
                ld[l][4]=dj; ld[l].sort(); n+=1
    return n
for s in starts:
    qs = world[s:s+Q].astype(np.float64)
    state=[{'lo2':0.0,'hi2':0.09,'done':False,'g5':np.inf,'n':0} for _ in range(Q)]
    lds={m:[[[np.inf]*5 for _ in range(G)] for _ in range(Q)] for m in ins}
    for shell in range(4):
        lists=[]
        for g in range(Q):
            st=state[g]
            if st['done']: lists.append([]); continue
            thr = st['g5'] if st['n']>=5 else np.inf
            lists.append(shell_bricks(qs[g], st['lo2'], st['hi2'], thr))
        if all(st['done'] for st in state): break
        for mode in ins:
            if mode=='unsorted': L=[l for l in lists]
            elif mode=='count_desc': L=[sorted(l, key=lambda x:-cnt[x[0]]) for l in lists]
            else: L=[sorted(l, key=lambda x:x[1]) for l in lists]
            ns = max(len(l) for l in L)
            for sl in range(ns):
                tot[mode]+= max(steps(cnt[l[sl][0]]) if sl<len(l) else 0 for l in L)
            for g in range(Q):
                ins[mode]+=count_inserts(qs[g], L[g], lds['unsorted' if False else mode][g])
        tot['ideal'] += sum(sum(cnt[b[0]] for b in l) for l in lists)/32.0
        # update state using exact merged result from 'unsorted' lists
        for g in range(Q):
            st=state[g]
            if st['done']: continue
            allv = sorted(v for lane in lds['unsorted'][g] for v in lane)[:5]
            st['n']=sum(1 for v in allv if v<np.inf); st['g5']=allv[4]
            last = st['hi2']>=5.0
            if last or (st['n']>=5 and st['g5']<=st['hi2']): st['done']=True
            else:
                st['lo2']=st['hi2']; st['hi2']=min(st['g5'],5.0) if st['n']>=5 else min(4*st['hi2'],5.0)
print('steps per warp-task:', {k:v/nw for k,v in tot.items()})
print('lane-inserts per point:', {k:v/(nw*Q) for k,v in ins.items()})
