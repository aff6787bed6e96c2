// knn_kernels.cuh -- exact bounded 5-NN on the brick hash (replaces KD_TREE::Nearest_Search,
// ikd_Tree.cpp:349-379, Search :825-968).
//
// G lanes (G = 8, 16 or 32) cooperate on one query, so a warp works on Q = 32/G queries AT ONCE and IN LOCKSTEP:
// every loop is warp-uniform (its trip count is the maximum over the warp's groups, idle groups are predicated
// off) and every shuffle / ballot uses the full mask. Sub-mask *_sync intrinsics make the hardware run the groups
// one after the other (measured in round 1: 11 of 32 lanes active); lockstep keeps all 32 lanes issuing together.
//
//   phase A  seed: one round probes the 2x2x2 bricks nearest to the query; the query's own brick is scanned first,
//            a cheap bound (5th smallest of the lanes' best candidates) prunes the other seven; merge -> (n, g5).
//            Done if g5 lies within the explored block.
//   phase A' (n < 5, sparse neighbourhood): ring expansion R = 1, 2, .. with the radius^2 = 5 bound until 5 are known.
//   phase B  closure: every unexplored brick whose box intersects the open ball of radius sqrt(g5) is probed and
//            scanned (bounding-box enumeration). After it the merged top-5 is exact: a point closer than g5 can
//            only live in a brick that intersects that ball.
//
// A brick's slab is read by G consecutive lanes -> contiguous 16*G-byte segments; every lane keeps a private sorted
// top-5 of the candidates IT saw; the group's top-5 is merged at phase boundaries.
// Semantics matched: candidates with d2 <= 5 only (ikd_Tree.cpp:842, sic), fp32
// (dx*dx+dy*dy)+dz*dz without FMA, ascending output; exact-tie handling is traversal dependent in the
// reference and therefore excluded from the parity claim.
#pragma once
#include "common.cuh"

template <int G>
struct Grp {
    static constexpr int Q = 32 / G;
    static constexpr unsigned GM = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
};

// ---- full-mask group primitives (all 32 lanes must call) -------------------------------------------------
template <int G>
__device__ __forceinline__ unsigned grp_ballot(bool p, int gbase) {
    return (__ballot_sync(LI_FULL, p) >> gbase) & Grp<G>::GM;
}
template <int G, class T>
__device__ __forceinline__ T grp_shfl(T v, int src_rel, int gbase) {
    return __shfl_sync(LI_FULL, v, gbase + src_rel);
}
template <int G>
__device__ __forceinline__ unsigned grp_min(unsigned v) {
    if constexpr (G == 32) {
        return __reduce_min_sync(LI_FULL, v);
    } else {
#pragma unroll
        for (int o = G / 2; o >= 1; o >>= 1) v = min(v, __shfl_xor_sync(LI_FULL, v, o));
        return v;
    }
}

// sorted insert into a lane-private top-5 (precondition: d < ld[4]); stable w.r.t. equal distances
__device__ __forceinline__ void local_insert(float (&ld)[5], int (&li)[5], float d, int id) {
    ld[4] = d;
    li[4] = id;
#pragma unroll
    for (int i = 4; i > 0; --i) {
        bool sw = ld[i] < ld[i - 1];
        float a = ld[i - 1], b = ld[i];
        int ia = li[i - 1], ib = li[i];
        ld[i - 1] = sw ? b : a;
        ld[i] = sw ? a : b;
        li[i - 1] = sw ? ib : ia;
        li[i] = sw ? ia : ib;
    }
}

// Group-wide top-5 of the G private lists -> gd/gi (uniform in the group).
template <int G>
__device__ __forceinline__ void group_merge(const float (&ld)[5], const int (&li)[5], float (&gd)[5], int (&gi)[5], int gl, int gbase) {
    float cd[5];
    int ci[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {
        cd[i] = ld[i];
        ci[i] = li[i];
    }
#pragma unroll
    for (int k = 0; k < 5; k++) {
        unsigned hb = __float_as_uint(cd[0]);
        unsigned mn = grp_min<G>(hb);
        unsigned who = grp_ballot<G>(hb == mn, gbase);
        int src = __ffs(who) - 1;
        gd[k] = __uint_as_float(mn);
        gi[k] = grp_shfl<G>(ci[0], src, gbase);
        if (gl == src) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                cd[i] = cd[i + 1];
                ci[i] = ci[i + 1];
            }
            cd[4] = INFINITY;
            ci[4] = -1;
        }
    }
}

// Cheap upper bound on the group's 5th-smallest candidate distance: the 5th smallest among the lanes' two best
// values (ld[0], ld[1]); +inf when fewer than 5 such candidates exist. Any 5 distinct candidates bound the true 5th.
template <int G>
__device__ __forceinline__ float group_bound5(const float (&ld)[5], int gl, int gbase) {
    unsigned a = __float_as_uint(ld[0]), b = __float_as_uint(ld[1]);   // a <= b (sorted list)
#pragma unroll
    for (int k = 0; k < 4; k++) {
        unsigned mn = grp_min<G>(a);
        unsigned who = grp_ballot<G>(a == mn, gbase);
        if (gl == __ffs(who) - 1) {   // pop this lane's head
            a = b;
            b = 0x7f800000u;
        }
    }
    return __uint_as_float(grp_min<G>(a));
}

struct KnnGeom {
    int bx, by, bz, dirx, diry, dirz, bs;
    float ds, margin;
};

// probe one brick (offset ox,oy,oz from the query's brick) if it can matter
__device__ __forceinline__ bool probe_brick(const MapDev& M, const KnnGeom& g, int ox, int oy, int oz, float qx, float qy, float qz,
                                            float bound2, bool full, unsigned& first, unsigned& count, float& dbox) {
    const int kx = g.bx + ox, ky = g.by + oy, kz = g.bz + oz;
    const int bs = g.bs;
    float lox = (float)(kx << bs) * g.ds - g.margin, hix = (float)((kx + 1) << bs) * g.ds + g.margin;
    float loy = (float)(ky << bs) * g.ds - g.margin, hiy = (float)((ky + 1) << bs) * g.ds + g.margin;
    float loz = (float)(kz << bs) * g.ds - g.margin, hiz = (float)((kz + 1) << bs) * g.ds + g.margin;
    float ex = fmaxf(0.f, fmaxf(lox - qx, qx - hix));
    float ey = fmaxf(0.f, fmaxf(loy - qy, qy - hiy));
    float ez = fmaxf(0.f, fmaxf(loz - qz, qz - hiz));
    dbox = (ex * ex + ey * ey + ez * ez) * (1.0f - 1e-6f);
    // only points with d2 <= 5 count, and only bricks whose box can beat the current bound on the 5th distance
    bool useful = (dbox <= 5.0f) && (dbox < bound2);
    (void)full;
    if (!useful) return false;
    bool found = li_brick_find(M.ent, M.mask, li_pack_key(kx, ky, kz), first, count);
    return found && count > 0u;
}

// Lockstep scan of the bricks found by the lanes of each group in the current probe round. thr: only candidates with
// d < thr (and d <= 5) matter. prune: skip bricks whose box distance is not below thr.
template <int G>
__device__ __forceinline__ void lockstep_scan_found(const float4* __restrict__ pool, bool found, unsigned first, unsigned count, float dbox,
                                                    float qx, float qy, float qz, float thr, float (&ld)[5], int (&li)[5], int gl, int gbase) {
    unsigned fm = grp_ballot<G>(found, gbase);
    while (__any_sync(LI_FULL, fm != 0u)) {
        const bool has = fm != 0u;
        const int src = has ? (__ffs(fm) - 1) : 0;
        fm &= fm - 1u;
        const float db = grp_shfl<G>(dbox, src, gbase);
        const unsigned f = grp_shfl<G>(first, src, gbase);
        const unsigned c = grp_shfl<G>(count, src, gbase);
        const unsigned cnt = (has && db < thr) ? c : 0u;
        for (unsigned j = gl; __any_sync(LI_FULL, j < cnt); j += G) {
            if (j < cnt) {
                float4 p = __ldg(&pool[(size_t)f + j]);
                float d = li_dist2(qx, qy, qz, p.x, p.y, p.z);
                if (d <= 5.0f && d < thr && d < ld[4]) local_insert(ld, li, d, (int)(f + j));
            }
        }
    }
}

// explored radius^2 after stage R (R = 0: the 2x2x2 half block)
__device__ __forceinline__ float explored_r2(const KnnGeom& g, int R, float qx, float qy, float qz) {
    int lo_x, hi_x, lo_y, hi_y, lo_z, hi_z;
    if (R == 0) {
        lo_x = min(g.bx, g.bx + g.dirx); hi_x = max(g.bx, g.bx + g.dirx);
        lo_y = min(g.by, g.by + g.diry); hi_y = max(g.by, g.by + g.diry);
        lo_z = min(g.bz, g.bz + g.dirz); hi_z = max(g.bz, g.bz + g.dirz);
    } else {
        lo_x = g.bx - R; hi_x = g.bx + R;
        lo_y = g.by - R; hi_y = g.by + R;
        lo_z = g.bz - R; hi_z = g.bz + R;
    }
    const int bs = g.bs;
    float rx = fminf(qx - (float)(lo_x << bs) * g.ds, (float)((hi_x + 1) << bs) * g.ds - qx);
    float ry = fminf(qy - (float)(lo_y << bs) * g.ds, (float)((hi_y + 1) << bs) * g.ds - qy);
    float rz = fminf(qz - (float)(lo_z << bs) * g.ds, (float)((hi_z + 1) << bs) * g.ds - qz);
    float r = fminf(rx, fminf(ry, rz)) - g.margin;
    if (r <= 0.f) return 0.f;
    return r * r * (1.0f - 1e-6f);
}

// Exact 5-NN of Q = 32/G queries by one warp in lockstep. ALL 32 lanes must call; `valid` is group-uniform.
// gd/gi: ascending distances / pool offsets (-1 = missing), uniform within each group.
template <int G>
__device__ __forceinline__ void knn5_lockstep(const MapDev& M, bool valid, float qx, float qy, float qz, float (&gd)[5], int (&gi)[5],
                                              int gl, int gbase) {
    float ld[5];
    int li[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {
        ld[i] = INFINITY;
        li[i] = -1;
        gd[i] = INFINITY;
        gi[i] = -1;
    }
    KnnGeom g;
    g.bs = M.bshift;
    g.ds = M.ds;
    const int bc = 1 << g.bs;
    const float B = (float)bc * g.ds;
    const float lim = (float)(LI_CELL_LIMIT - 16 * bc) * g.ds;
    bool act = valid && isfinite(qx) && isfinite(qy) && isfinite(qz) && fabsf(qx) < lim && fabsf(qy) < lim && fabsf(qz) < lim;
    if (!act) {
        qx = 0.f; qy = 0.f; qz = 0.f;
    }
    const int cx = li_cell(qx, g.ds), cy = li_cell(qy, g.ds), cz = li_cell(qz, g.ds);
    g.bx = cx >> g.bs; g.by = cy >> g.bs; g.bz = cz >> g.bs;
    const int half = bc >> 1;
    g.dirx = ((cx & (bc - 1)) < half) ? -1 : 1;
    g.diry = ((cy & (bc - 1)) < half) ? -1 : 1;
    g.dirz = ((cz & (bc - 1)) < half) ? -1 : 1;
    // slack for float cell assignment / edge products: relative 2^-23 effects, bounded generously
    g.margin = 1e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + 16.0f * B);

    // ---- phase A: seed ---------------------------------------------------------------------------------
    {
        float bound = INFINITY;
#pragma unroll
        for (int base = 0; base < 8; base += G) {
            unsigned first = 0, count = 0;
            float dbox = INFINITY;
            bool found = false;
            const int idx = base + gl;
            if (act && idx < 8) {
                int ox = (idx & 1) ? g.dirx : 0, oy = (idx & 2) ? g.diry : 0, oz = (idx & 4) ? g.dirz : 0;
                found = probe_brick(M, g, ox, oy, oz, qx, qy, qz, bound, false, first, count, dbox);
            }
            unsigned fm = grp_ballot<G>(found, gbase);   // bit 0 of round 0 = the query's own brick: taken first
            while (__any_sync(LI_FULL, fm != 0u)) {
                const bool has = fm != 0u;
                const int src = has ? (__ffs(fm) - 1) : 0;
                fm &= fm - 1u;
                const float db = grp_shfl<G>(dbox, src, gbase);
                const unsigned f = grp_shfl<G>(first, src, gbase);
                const unsigned c = grp_shfl<G>(count, src, gbase);
                const unsigned cnt = (has && db < bound) ? c : 0u;   // a brick whose box is beyond the bound holds none of the 5
                for (unsigned j = gl; __any_sync(LI_FULL, j < cnt); j += G) {
                    if (j < cnt) {
                        float4 p = __ldg(&M.pool[(size_t)f + j]);
                        float d = li_dist2(qx, qy, qz, p.x, p.y, p.z);
                        if (d <= 5.0f && d < bound && d < ld[4]) local_insert(ld, li, d, (int)(f + j));
                    }
                }
                bound = fminf(bound, group_bound5<G>(ld, gl, gbase));
            }
        }
    }
    group_merge<G>(ld, li, gd, gi, gl, gbase);
    float r2 = explored_r2(g, 0, qx, qy, qz);
    bool done = !act || r2 > 5.0f || (gi[4] >= 0 && gd[4] <= r2);
    int Rdone = 0;

    // ---- phase A': fewer than 5 known -> rings with the radius bound (sparse neighbourhoods, open air) ------
    const int Rmax = (int)ceilf(2.2360680f / B) + 1;
    for (int R = 1; R <= Rmax; R++) {
        const bool need = !done && gi[4] < 0;
        if (!__any_sync(LI_FULL, need)) break;
        const int S = 2 * R + 1, total = S * S * S;
        const float inv_s = 1.0f / (float)S, inv_ss = 1.0f / (float)(S * S);
        for (int base = 0; base < total; base += G) {
            const int idx = base + gl;
            bool want = need && idx < total;
            const int oz_ = (int)(((float)idx + 0.5f) * inv_ss);
            const int rem = idx - oz_ * S * S;
            const int oy_ = (int)(((float)rem + 0.5f) * inv_s);
            const int ox = rem - oy_ * S - R, oy = oy_ - R, oz = oz_ - R;
            if (R == 1) {
                if ((ox == 0 || ox == g.dirx) && (oy == 0 || oy == g.diry) && (oz == 0 || oz == g.dirz)) want = false;   // stage 0 block
            } else {
                if (max(abs(ox), max(abs(oy), abs(oz))) < R) want = false;   // inner cube already visited
            }
            unsigned first = 0, count = 0;
            float dbox = INFINITY;
            bool found = false;
            if (want) found = probe_brick(M, g, ox, oy, oz, qx, qy, qz, INFINITY, false, first, count, dbox);
            lockstep_scan_found<G>(M.pool, found, first, count, dbox, qx, qy, qz, INFINITY, ld, li, gl, gbase);
        }
        group_merge<G>(ld, li, gd, gi, gl, gbase);
        if (need) {
            Rdone = R;
            r2 = explored_r2(g, R, qx, qy, qz);
            done = r2 > 5.0f || (gi[4] >= 0 && gd[4] <= r2);
        }
    }
    if (gi[4] < 0) done = true;   // fewer than 5 points within the radius: everything within sqrt(5) was explored

    // ---- phase B: closure over the ball of radius sqrt(g5) -----------------------------------------------
    if (__any_sync(LI_FULL, !done)) {
        const bool need = !done;
        const float g5 = need ? gd[4] : 0.f;
        const float r = sqrtf(g5) * (1.0f + 1e-6f) + g.margin;
        // explored block (brick coordinates)
        int ex0, ex1, ey0, ey1, ez0, ez1;
        if (Rdone == 0) {
            ex0 = min(g.bx, g.bx + g.dirx); ex1 = max(g.bx, g.bx + g.dirx);
            ey0 = min(g.by, g.by + g.diry); ey1 = max(g.by, g.by + g.diry);
            ez0 = min(g.bz, g.bz + g.dirz); ez1 = max(g.bz, g.bz + g.dirz);
        } else {
            ex0 = g.bx - Rdone; ex1 = g.bx + Rdone;
            ey0 = g.by - Rdone; ey1 = g.by + Rdone;
            ez0 = g.bz - Rdone; ez1 = g.bz + Rdone;
        }
        const int lx = li_cell(qx - r, g.ds) >> g.bs, hx = li_cell(qx + r, g.ds) >> g.bs;
        const int ly = li_cell(qy - r, g.ds) >> g.bs, hy = li_cell(qy + r, g.ds) >> g.bs;
        const int lz = li_cell(qz - r, g.ds) >> g.bs, hz = li_cell(qz + r, g.ds) >> g.bs;
        const int nx = hx - lx + 1, ny = hy - ly + 1, nz = hz - lz + 1;
        const int nxy = nx * ny;
        const int total = need ? nxy * nz : 0;
        const float inv_nxy = 1.0f / (float)nxy, inv_nx = 1.0f / (float)nx;
        for (int base = 0; __any_sync(LI_FULL, base < total); base += G) {
            const int idx = base + gl;
            bool want = idx < total;
            const int iz = (int)(((float)idx + 0.5f) * inv_nxy);
            const int rem = idx - iz * nxy;
            const int iy = (int)(((float)rem + 0.5f) * inv_nx);
            const int ix = rem - iy * nx;
            const int kx = lx + ix, ky = ly + iy, kz = lz + iz;
            if (kx >= ex0 && kx <= ex1 && ky >= ey0 && ky <= ey1 && kz >= ez0 && kz <= ez1) want = false;   // explored
            unsigned first = 0, count = 0;
            float dbox = INFINITY;
            bool found = false;
            if (want) found = probe_brick(M, g, kx - g.bx, ky - g.by, kz - g.bz, qx, qy, qz, g5, true, first, count, dbox);
            lockstep_scan_found<G>(M.pool, found, first, count, dbox, qx, qy, qz, g5, ld, li, gl, gbase);
        }
        group_merge<G>(ld, li, gd, gi, gl, gbase);
    }
}

// ---- search kernel of an ICP pass: world transform + 5-NN for every scan point -----------------------
template <int G>
__global__ void __launch_bounds__(256) k_knn_scan(MapDev M, ScanDev S, PoseD P) {
    constexpr int Q = Grp<G>::Q;
    const int lane = threadIdx.x & 31;
    const int gl = lane % G, gid = lane / G, gbase = gid * G;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int qb = warp_global * Q; qb < S.n; qb += nwarps * Q) {   // warp-uniform
        const int q = qb + gid;
        const bool valid = q < S.n;
        float wx = 0.f, wy = 0.f, wz = 0.f;
        if (valid) {
            float4 b = __ldg(&S.body[q]);
            li_body_to_world(P, b.x, b.y, b.z, wx, wy, wz);
        }
        float gd[5];
        int gi[5];
        knn5_lockstep<G>(M, valid, wx, wy, wz, gd, gi, gl, gbase);
        if (valid && gl == 0) {
            S.world[q] = make_float4(wx, wy, wz, 0.f);
#pragma unroll
            for (int k = 0; k < 5; k++) S.near_ids[(size_t)q * 5 + k] = gi[k];
        }
    }
}

// ---- stand-alone Nearest_Search for arbitrary world-frame queries --------------------------------------
template <int G>
__global__ void __launch_bounds__(256) k_knn_queries(MapDev M, const float4* __restrict__ qpts, int n, int* __restrict__ ids,
                                                     float* __restrict__ d2) {
    constexpr int Q = Grp<G>::Q;
    const int lane = threadIdx.x & 31;
    const int gl = lane % G, gid = lane / G, gbase = gid * G;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int qb = warp_global * Q; qb < n; qb += nwarps * Q) {
        const int q = qb + gid;
        const bool valid = q < n;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) p = __ldg(&qpts[q]);
        float gd[5];
        int gi[5];
        knn5_lockstep<G>(M, valid, p.x, p.y, p.z, gd, gi, gl, gbase);
        if (valid && gl == 0) {
#pragma unroll
            for (int k = 0; k < 5; k++) {
                ids[(size_t)q * 5 + k] = gi[k];
                d2[(size_t)q * 5 + k] = (gi[k] >= 0) ? gd[k] : -1.f;
            }
        }
    }
}
