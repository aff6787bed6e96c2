// map_kernels.cuh -- device-resident map: build / insert / downsample-insert.
//
// Replaces KD_TREE::Build (ikd_Tree.cpp:336-347), KD_TREE::Add_Points (ikd_Tree.cpp:381-456) and the
// classification of map_incremental (laserMapping.cpp:516-559). All batch semantics are derived in
// after Add_Points(batch, downsample_on=true) every voxel hit by the batch holds exactly the
// point closest to the voxel centre among {its live points} U {the batch's points in it} (new beats
// existing on ties, later beats earlier among new), which is what the reference's sequential loop
// produces, independent of order.
#pragma once
#include "common.cuh"

// ---- staging: strided host layout -> float4 -------------------------------------------------------
__global__ void k_repack(const float* __restrict__ src, int stride, int n, float4* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* s = src + (size_t)i * stride;
    dst[i] = make_float4(s[0], s[1], s[2], 0.f);
}

__global__ void k_map_clear(uint4* ent, uint4* aux, unsigned slots) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= slots) return;
    ent[i] = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u);
    aux[i] = make_uint4(0u, 0u, 0u, 0u);
}

__device__ __forceinline__ bool li_point_cells(const MapDev& M, float4 p, int& cx, int& cy, int& cz) {
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) return false;
    float lim = (float)(LI_CELL_LIMIT - 2) * M.ds;
    if (fabsf(p.x) >= lim || fabsf(p.y) >= lim || fabsf(p.z) >= lim) return false;
    cx = li_cell(p.x, M.ds);
    cy = li_cell(p.y, M.ds);
    cz = li_cell(p.z, M.ds);
    return true;
}

__device__ __forceinline__ unsigned li_voxel_in_brick(const MapDev& M, int cx, int cy, int cz) {
    int m = (1 << M.bshift) - 1;
    return (unsigned)(((cx & m) << (2 * M.bshift)) | ((cy & m) << M.bshift) | (cz & m));
}


// Index of the downsample box that CONTAINS x as the reference tests membership (Search_by_range / Delete_by_range,
// ikd_Tree.cpp:633,980): box k = [fl(k*ds), fl(fl(k*ds)+ds)). It equals the division index c = floor(fl(x/ds)) except
// within one ulp of a box edge, where x can sit in box c-1 / c+1 or in NO box (the float boxes leave one-ulp gaps).
// Grid-aligned synthetic scenes (walls at k*ds) hit this systematically, so it is reproduced exactly.
#define LI_NO_BOX 0xfffffffeu
__device__ __forceinline__ bool li_box_index(float x, float ds, int c, int& b) {
    float cf = (float)c;
    float mn = __fmul_rn(cf, ds), mx = __fadd_rn(mn, ds);
    if (x >= mn && x < mx) { b = c; return true; }
    if (x < mn) {
        float m1 = __fmul_rn(cf - 1.0f, ds), x1 = __fadd_rn(m1, ds);
        if (x >= m1 && x < x1) { b = c - 1; return true; }
    } else {
        float m1 = __fmul_rn(cf + 1.0f, ds), x1 = __fadd_rn(m1, ds);
        if (x >= m1 && x < x1) { b = c + 1; return true; }
    }
    b = c;
    return false;
}

// Where a map point is STORED: the brick of its box index (so that the box test of a later Add_Points finds it in
// the brick of that box) and its voxel-in-brick id, or LI_NO_BOX when it lies in no downsample box.
__device__ __forceinline__ void li_storage(const MapDev& M, float4 p, int cx, int cy, int cz, unsigned long long& brick_key, unsigned& vib) {
    int bx, by, bz;
    bool ok = li_box_index(p.x, M.ds, cx, bx);
    ok = li_box_index(p.y, M.ds, cy, by) && ok;
    ok = li_box_index(p.z, M.ds, cz, bz) && ok;
    if (ok) {
        brick_key = li_pack_key(bx >> M.bshift, by >> M.bshift, bz >> M.bshift);
        vib = li_voxel_in_brick(M, bx, by, bz);
    } else {
        brick_key = li_pack_key(cx >> M.bshift, cy >> M.bshift, cz >> M.bshift);
        vib = LI_NO_BOX;
    }
}

// add a hash slot to the batch's touched list exactly once (aux.w is the membership flag)
__device__ __forceinline__ void li_touch(const MapDev& M, int s) {
    if (atomicExch(&M.aux[s].w, 1u) == 0u) M.touched_list[atomicAdd(&M.counters[CNT_TOUCHED], 1)] = s;
}

// ---- plain insert (Build, Add_Points(.., false)) -------------------------------------------------
// pass 1: find-or-create the brick of every point, count pending points per brick.
__global__ void k_ins_count(MapDev M, const float4* __restrict__ pts, int n, const int* __restrict__ sel, int want,
                            int* __restrict__ slot_of) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    slot_of[i] = -1;
    if (sel && sel[i] != want) return;
    int cx, cy, cz;
    float4 p = pts[i];
    if (!li_point_cells(M, p, cx, cy, cz)) {
        atomicAdd(&M.counters[CNT_DROPPED], 1);
        return;
    }
    unsigned long long key;
    unsigned vib;
    li_storage(M, p, cx, cy, cz, key, vib);
    bool created = false;
    int s = li_brick_find_or_insert(M.ent, M.mask, key, &created);
    if (s < 0) {
        atomicOr(&M.counters[CNT_ERR], ERR_HASH_FULL);
        return;
    }
    slot_of[i] = s;
    if (created) atomicAdd(&M.counters[CNT_BRICKS], 1);
    atomicAdd(&M.aux[s].y, 1u);
    li_touch(M, s);
}

// pass 2: one warp per touched brick -- grow its slab when count + pending exceeds the capacity.
__global__ void k_ins_reserve(MapDev M) {
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= M.counters[CNT_TOUCHED]) return;
    int s = M.touched_list[w];
    uint4 e = M.ent[s];
    uint4 a = M.aux[s];
    unsigned need = e.w + a.y;
    if (need > a.x) {
        unsigned ncap = need + (need >> 1);
        if (ncap < 16u) ncap = 16u;
        ncap = (ncap + 7u) & ~7u;
        unsigned long long off = 0;
        if (lane == 0) off = atomicAdd(M.pool_top, (unsigned long long)ncap);
        off = __shfl_sync(LI_FULL, off, 0);
        if (off + ncap > M.pool_cap) {
            if (lane == 0) {
                atomicOr(&M.counters[CNT_ERR], ERR_POOL_FULL);
                M.aux[s].y = 0u;   // nothing will be appended to this brick
            }
            return;
        }
        for (unsigned j = lane; j < e.w; j += 32) M.pool[off + j] = M.pool[e.z + j];
        if (lane == 0) {
            M.ent[s].z = (unsigned)off;
            M.aux[s].x = ncap;
        }
    }
    if (lane == 0) M.aux[s].z = 0u;
}

// pass 3: append. ent.count is not modified until the commit, so first+count is the append base.
__global__ void k_ins_append(MapDev M, const float4* __restrict__ pts, int n, const int* __restrict__ slot_of) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = slot_of[i];
    if (s < 0) return;
    uint4 e = M.ent[s];
    uint4 a = M.aux[s];
    if (a.y == 0u) return;   // reservation failed
    unsigned j = atomicAdd(&M.aux[s].z, 1u);
    float4 p = pts[i];
    int cx = li_cell(p.x, M.ds), cy = li_cell(p.y, M.ds), cz = li_cell(p.z, M.ds);
    unsigned long long key;
    unsigned vib;
    li_storage(M, p, cx, cy, cz, key, vib);
    p.w = __uint_as_float(vib);
    M.pool[(size_t)e.z + e.w + j] = p;
}

// pass 4: commit counts.
__global__ void k_ins_commit(MapDev M) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M.counters[CNT_TOUCHED]) return;
    int s = M.touched_list[t];
    unsigned f = M.aux[s].z;
    M.ent[s].w += f;
    M.aux[s].y = 0u;
    M.aux[s].z = 0u;
    M.aux[s].w = 0u;
    atomicAdd(&M.counters[CNT_LIVE], (int)f);
}

// ---- downsample insert (Add_Points(.., true)) --------------------------------------------------
// Temporary voxel hash for the batch: vkeys (u64) + vbest (u64 = d_centre bits << 32 | ~index).
struct VoxTmp {
    unsigned long long* keys;
    unsigned long long* best;
    unsigned mask;
};

__global__ void k_vox_clear(VoxTmp V) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > V.mask) return;
    V.keys[i] = LI_EMPTY_KEY;
    V.best[i] = 0xffffffffffffffffull;
}

// Distance of p to the centre of its downsample box, float arithmetic of ikd_Tree.cpp:389-401.
__device__ __forceinline__ float li_center_dist(float4 p, float ds) {
    float mnx = __fmul_rn(floorf(__fdiv_rn(p.x, ds)), ds), mxx = __fadd_rn(mnx, ds);
    float mny = __fmul_rn(floorf(__fdiv_rn(p.y, ds)), ds), mxy = __fadd_rn(mny, ds);
    float mnz = __fmul_rn(floorf(__fdiv_rn(p.z, ds)), ds), mxz = __fadd_rn(mnz, ds);
    float mx = (float)__dadd_rn((double)mnx, __ddiv_rn((double)__fsub_rn(mxx, mnx), 2.0));
    float my = (float)__dadd_rn((double)mny, __ddiv_rn((double)__fsub_rn(mxy, mny), 2.0));
    float mz = (float)__dadd_rn((double)mnz, __ddiv_rn((double)__fsub_rn(mxz, mnz), 2.0));
    return li_dist2(p.x, p.y, p.z, mx, my, mz);
}
// Same for an EXISTING point q tested against the box of voxel (cx,cy,cz) (its own voxel).
__device__ __forceinline__ float li_center_dist_cell(float qx, float qy, float qz, int cx, int cy, int cz, float ds) {
    float mnx = __fmul_rn((float)cx, ds), mxx = __fadd_rn(mnx, ds);
    float mny = __fmul_rn((float)cy, ds), mxy = __fadd_rn(mny, ds);
    float mnz = __fmul_rn((float)cz, ds), mxz = __fadd_rn(mnz, ds);
    float mx = (float)__dadd_rn((double)mnx, __ddiv_rn((double)__fsub_rn(mxx, mnx), 2.0));
    float my = (float)__dadd_rn((double)mny, __ddiv_rn((double)__fsub_rn(mxy, mny), 2.0));
    float mz = (float)__dadd_rn((double)mnz, __ddiv_rn((double)__fsub_rn(mxz, mnz), 2.0));
    return li_dist2(qx, qy, qz, mx, my, mz);
}

// pass D1: per new point, vote for the voxel's best new point (min centre distance, later index wins ties).
__global__ void k_ds_vote(MapDev M, VoxTmp V, const float4* __restrict__ pts, int n, const int* __restrict__ sel /*optional: only sel[i]==want*/,
                          int want, int* __restrict__ vslot_of) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    vslot_of[i] = -1;
    if (sel && sel[i] != want) return;
    int cx, cy, cz;
    float4 p = pts[i];
    if (!li_point_cells(M, p, cx, cy, cz)) {
        atomicAdd(&M.counters[CNT_DROPPED], 1);
        return;
    }
    unsigned long long key = li_pack_key(cx, cy, cz);
    unsigned h = li_hash(key) & V.mask;
    int slot = -1;
    for (unsigned t = 0; t <= V.mask; t++) {
        unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(&V.keys[h]);
        if (k == key) { slot = (int)h; break; }
        if (k == LI_EMPTY_KEY) {
            unsigned long long old = atomicCAS(&V.keys[h], LI_EMPTY_KEY, key);
            if (old == LI_EMPTY_KEY || old == key) { slot = (int)h; break; }
        }
        h = (h + 1) & V.mask;
    }
    if (slot < 0) {
        atomicOr(&M.counters[CNT_ERR], ERR_HASH_FULL);
        return;
    }
    vslot_of[i] = slot;
    float d = li_center_dist(p, M.ds);
    unsigned long long v = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
    atomicMin(&V.best[slot], v);
}

// pass D2: voxel winners reserve room in the brick they would be STORED in and mark the brick of their BOX
// (where the existing points they compete with live) as touched. The two differ only for ulp-edge points.
__global__ void k_ds_reserve_votes(MapDev M, VoxTmp V, const float4* __restrict__ pts, int n, int* __restrict__ vslot_of /*in: voxel slot, out: box-brick slot*/,
                                   int* __restrict__ slot_of /*out: storage-brick slot*/) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    slot_of[i] = -1;
    int vs = vslot_of[i];
    vslot_of[i] = -1;
    if (vs < 0) return;
    unsigned long long b = V.best[vs];
    if ((unsigned)(b & 0xffffffffull) != 0xffffffffu - (unsigned)i) return;   // not this voxel's best new point
    float4 p = pts[i];
    int cx = li_cell(p.x, M.ds), cy = li_cell(p.y, M.ds), cz = li_cell(p.z, M.ds);
    unsigned long long skey, bkey = li_pack_key(cx >> M.bshift, cy >> M.bshift, cz >> M.bshift);
    unsigned vib;
    li_storage(M, p, cx, cy, cz, skey, vib);
    bool created = false;
    int s = li_brick_find_or_insert(M.ent, M.mask, skey, &created);
    if (s < 0) {
        atomicOr(&M.counters[CNT_ERR], ERR_HASH_FULL);
        return;
    }
    if (created) atomicAdd(&M.counters[CNT_BRICKS], 1);
    slot_of[i] = s;
    atomicAdd(&M.aux[s].y, 1u);
    li_touch(M, s);
    int sb = s;
    if (bkey != skey) {
        sb = -1;
        unsigned h = li_hash(bkey) & M.mask;
        for (unsigned t = 0; t <= M.mask; t++) {
            unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(&M.ent[h]);
            if (k == bkey) { sb = (int)h; break; }
            if (k == LI_EMPTY_KEY) break;
            h = (h + 1) & M.mask;
        }
        if (sb >= 0) li_touch(M, sb);
    }
    vslot_of[i] = sb;
}

// pass D3: one warp per voxel winner. Compare with the voxel's live points, tombstone the losers
// (w = 0xffffffff), append the new point if it wins. Different voxels of one brick touch disjoint
// slab entries, appends go through aux.fill, so warps of the same brick do not race.
__global__ void k_ds_apply(MapDev M, const float4* __restrict__ pts, int n, const int* __restrict__ slot_of,
                           const int* __restrict__ box_slot_of) {
    int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n) return;
    int s = slot_of[i];
    if (s < 0) return;
    if (M.aux[s].y == 0u) return;   // reservation failed (pool full)
    const int sb = box_slot_of[i];
    float4 p = pts[i];
    int cx = li_cell(p.x, M.ds), cy = li_cell(p.y, M.ds), cz = li_cell(p.z, M.ds);
    const unsigned vib = li_voxel_in_brick(M, cx, cy, cz);   // id of p's BOX inside the box brick
    float dp = li_center_dist(p, M.ds);
    // scan the live slab of the box brick for the points of this box
    float best_d = INFINITY;
    unsigned best_j = 0xffffffffu;
    unsigned n_exist = 0;
    uint4 e = make_uint4(0u, 0u, 0u, 0u);
    if (sb >= 0) e = M.ent[sb];
    for (unsigned base = 0; base < e.w; base += 32) {
        unsigned j = base + lane;
        bool mine = false;
        float d = INFINITY;
        if (j < e.w) {
            float4 q = M.pool[(size_t)e.z + j];
            if (__float_as_uint(q.w) == vib) {
                mine = true;
                d = li_center_dist_cell(q.x, q.y, q.z, cx, cy, cz, M.ds);
            }
        }
        unsigned mm = __ballot_sync(LI_FULL, mine);
        n_exist += __popc(mm);
        // lowest slot wins ties among existing points
        unsigned bits = mine ? __float_as_uint(d) : 0xffffffffu;
        unsigned mn = __reduce_min_sync(LI_FULL, bits);
        if (mm && mn < __float_as_uint(best_d)) {
            unsigned who = __ballot_sync(LI_FULL, mine && bits == mn);
            best_d = __uint_as_float(mn);
            best_j = base + (__ffs(who) - 1);
        }
    }
    bool existing_wins = (n_exist > 0) && (best_d < dp);   // strict: new point wins ties (ikd_Tree.cpp:405)
    // tombstone every existing point of the box except a winning existing one
    if (n_exist > 0) {
        for (unsigned base = 0; base < e.w; base += 32) {
            unsigned j = base + lane;
            if (j < e.w) {
                float4* qp = &M.pool[(size_t)e.z + j];
                if (__float_as_uint(qp->w) == vib && !(existing_wins && j == best_j)) qp->w = __uint_as_float(0xffffffffu);
            }
        }
    }
    if (lane == 0) {
        if (!existing_wins) {
            uint4 es = M.ent[s];
            unsigned j = atomicAdd(&M.aux[s].z, 1u);
            unsigned long long skey;
            unsigned svib;
            li_storage(M, p, cx, cy, cz, skey, svib);
            p.w = __uint_as_float(svib);
            M.pool[(size_t)es.z + es.w + j] = p;
        }
        if (!existing_wins || n_exist > 1) atomicAdd(&M.counters[CNT_CHANGED], 1);
    }
}

// pass D4: one warp per touched brick -- squeeze out tombstones over [0, count+fill), fix counts.
__global__ void k_ds_compact(MapDev M) {
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= M.counters[CNT_TOUCHED]) return;
    int s = M.touched_list[w];
    uint4 e = M.ent[s];
    uint4 a = M.aux[s];
    unsigned total = e.w + a.z;
    unsigned wr = 0;
    for (unsigned base = 0; base < total; base += 32) {
        unsigned j = base + lane;
        float4 q = make_float4(0, 0, 0, 0);
        bool live = false;
        if (j < total) {
            q = M.pool[(size_t)e.z + j];
            live = __float_as_uint(q.w) != 0xffffffffu;
        }
        unsigned mm = __ballot_sync(LI_FULL, live);
        if (live) M.pool[(size_t)e.z + wr + __popc(mm & ((1u << lane) - 1u))] = q;
        wr += __popc(mm);
        __syncwarp();
    }
    if (lane == 0) {
        M.ent[s].w = wr;
        M.aux[s].y = 0u;
        M.aux[s].z = 0u;
        M.aux[s].w = 0u;
        atomicAdd(&M.counters[CNT_LIVE], (int)wr - (int)e.w);
    }
}

// ---- flatten -------------------------------------------------------------------------------------
__global__ void k_map_flatten(MapDev M, unsigned slots, float* __restrict__ out, int cap, int* __restrict__ out_n) {
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if ((unsigned)w >= slots) return;
    uint4 e = M.ent[w];
    unsigned long long k = (unsigned long long)e.x | ((unsigned long long)e.y << 32);
    if (k == LI_EMPTY_KEY || e.w == 0u) return;
    int base = 0;
    if (lane == 0) base = atomicAdd(out_n, (int)e.w);
    base = __shfl_sync(LI_FULL, base, 0);
    for (unsigned j = lane; j < e.w; j += 32) {
        int o = base + (int)j;
        if (o < cap) {
            float4 q = M.pool[(size_t)e.z + j];
            out[3 * (size_t)o] = q.x;
            out[3 * (size_t)o + 1] = q.y;
            out[3 * (size_t)o + 2] = q.z;
        }
    }
}

// ---- map_incremental classification (laserMapping.cpp:516-559) ------------------------------------
// flag[i]: 0 = skip, 1 = PointToAdd (downsample insert), 2 = PointNoNeedDownsample (plain insert).
// world[i] = pointBodyToWorld(body[i]) with the FINAL state (:523). near_ids: retained from the last search.
__global__ void k_incr_classify(MapDev M, PoseD P, const float4* __restrict__ body, int n, const int* __restrict__ near_ids,
                                double ds, int flg_EKF_inited, float4* __restrict__ world, int* __restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 b = body[i];
    float wx, wy, wz;
    li_body_to_world(P, b.x, b.y, b.z, wx, wy, wz);
    world[i] = make_float4(wx, wy, wz, 0.f);
    int ids[5];
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        ids[j] = near_ids[(size_t)i * 5 + j];
        if (ids[j] >= 0) cnt++;
    }
    int f = 1;
    if (cnt > 0 && flg_EKF_inited) {
        // mid_point: double expression stored to float (laserMapping.cpp:530-535)
        float mx = (float)__dadd_rn(__dmul_rn(floor(__ddiv_rn((double)wx, ds)), ds), __dmul_rn(0.5, ds));
        float my = (float)__dadd_rn(__dmul_rn(floor(__ddiv_rn((double)wy, ds)), ds), __dmul_rn(0.5, ds));
        float mz = (float)__dadd_rn(__dmul_rn(floor(__ddiv_rn((double)wz, ds)), ds), __dmul_rn(0.5, ds));
        float dist = li_dist2(wx, wy, wz, mx, my, mz);
        float4 n0 = M.pool[ids[0]];
        double half = 0.5 * ds;
        if ((double)fabsf(__fsub_rn(n0.x, mx)) > half && (double)fabsf(__fsub_rn(n0.y, my)) > half &&
            (double)fabsf(__fsub_rn(n0.z, mz)) > half) {
            f = 2;
        } else {
            if (cnt >= 5) {
#pragma unroll
                for (int j = 0; j < 5; j++) {
                    float4 q = M.pool[ids[j]];
                    if (li_dist2(q.x, q.y, q.z, mx, my, mz) < dist) {
                        f = 0;
                        break;
                    }
                }
            }
        }
    }
    flag[i] = f;
    if (f == 1) atomicAdd(&M.counters[CNT_NADD], 1);
    if (f == 2) atomicAdd(&M.counters[CNT_NNOD], 1);
}

