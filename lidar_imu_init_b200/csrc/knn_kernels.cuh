// knn_kernels.cuh -- exact bounded 5-NN on the brick hash (replaces KD_TREE::Nearest_Search,
// ikd_Tree.cpp:349-379, Search :825-968).
//
// G lanes (G = 4, 8, 16 or 32) cooperate on one query: a warp works on 32/G queries at once.
//   * stage 0 probes the 2x2x2 block of bricks nearest to the query, stage R >= 1 the shell of the
//     (2R+1)^3 block; a brick is probed only if its box can still hold a point closer than the current
//     5th neighbour (or than the radius^2 = 5 bound while fewer than 5 are known);
//   * a brick's slab is read with G consecutive lanes -> contiguous 16*G-byte segments;
//   * every lane keeps a private sorted top-5 of the candidates IT saw; the group's top-5 is merged with
//     redux.sync / shuffles at the end of each stage (the merge gives the pruning bound and the
//     termination test: stop when the 5th distance is within the explored radius);
//   * search is exact: ring expansion continues until no unexplored brick can intersect the ball.
// Semantics matched: candidates with d2 <= 5 only (ikd_Tree.cpp:842, sic), fp32
// (dx*dx+dy*dy)+dz*dz without FMA, ascending output; exact-tie handling is traversal dependent in the
// reference and therefore excluded from the parity claim.
#pragma once
#include "common.cuh"

template <int G>
struct Grp {
    static constexpr int QPW = 32 / G;
    __device__ static __forceinline__ unsigned mask(int lane) { return (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G)); }
};

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
__device__ __forceinline__ void group_merge(const float (&ld)[5], const int (&li)[5], float (&gd)[5], int (&gi)[5], unsigned gmask,
                                            int lane) {
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
        unsigned mn = __reduce_min_sync(gmask, hb);
        unsigned who = __ballot_sync(gmask, hb == mn) & gmask;
        int src = __ffs(who) - 1;
        gd[k] = __uint_as_float(mn);
        gi[k] = __shfl_sync(gmask, ci[0], src);
        if (lane == src) {
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

// scan the slabs of the bricks found by the lanes of this group in the current round
template <int G>
__device__ __forceinline__ void group_scan_found(const float4* __restrict__ pool, bool found, unsigned first, unsigned count,
                                                 float qx, float qy, float qz, float g5, float (&ld)[5], int (&li)[5],
                                                 unsigned gmask, int lane, int gl) {
    unsigned fm = __ballot_sync(gmask, found) & gmask;
    while (fm) {
        int src = __ffs(fm) - 1;
        fm &= fm - 1;
        unsigned f = __shfl_sync(gmask, first, src);
        unsigned c = __shfl_sync(gmask, count, src);
        for (unsigned j = gl; j < c; j += G) {
            float4 p = __ldg(&pool[(size_t)f + j]);
            float d = li_dist2(qx, qy, qz, p.x, p.y, p.z);
            if (d <= 5.0f && d < g5 && d < ld[4]) local_insert(ld, li, d, (int)(f + j));
        }
    }
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
    // while fewer than 5 are known every point with d2 <= 5 counts; afterwards only d2 < current 5th
    bool useful = full ? (dbox < bound2) : (dbox <= 5.0f);
    if (!useful) return false;
    bool found = li_brick_find(M.ent, M.mask, li_pack_key(kx, ky, kz), first, count);
    return found && count > 0u;
}

template <int G, int R>
__device__ __forceinline__ void knn_ring(const MapDev& M, const KnnGeom& g, float qx, float qy, float qz, float g5, bool full,
                                         float (&ld)[5], int (&li)[5], unsigned gmask, int lane, int gl) {
    constexpr int S = 2 * R + 1;
    constexpr int total = S * S * S;
    for (int base = 0; base < total; base += G) {
        int idx = base + gl;
        bool want = idx < total;
        int ox = idx % S - R, oy = (idx / S) % S - R, oz = idx / (S * S) - R;
        if (R == 1) {
            if ((ox == 0 || ox == g.dirx) && (oy == 0 || oy == g.diry) && (oz == 0 || oz == g.dirz)) want = false;   // stage 0 block
        } else {
            if (max(abs(ox), max(abs(oy), abs(oz))) < R) want = false;   // inner cube already visited
        }
        unsigned first = 0, count = 0;
        float dbox;
        bool found = false;
        if (want) found = probe_brick(M, g, ox, oy, oz, qx, qy, qz, g5, full, first, count, dbox);
        group_scan_found<G>(M.pool, found, first, count, qx, qy, qz, g5, ld, li, gmask, lane, gl);
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

// Cheap upper bound on the group's 5th-smallest candidate distance: the 5th smallest of the lanes' BEST values
// (each lane's ld[0]); +inf when fewer than 5 lanes hold a candidate. ~25 instructions instead of a full merge.
__device__ __forceinline__ float group_bound5(float best, unsigned gmask, int lane) {
    unsigned v = __float_as_uint(best);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        unsigned mn = __reduce_min_sync(gmask, v);
        unsigned who = __ballot_sync(gmask, v == mn) & gmask;
        if (lane == __ffs(who) - 1) v = 0x7f800000u;   // +inf
    }
    return __uint_as_float(__reduce_min_sync(gmask, v));
}

// scan one slab with all lanes of the group (thr: only candidates with d < thr and d <= 5 matter)
template <int G>
__device__ __forceinline__ void group_scan_slab(const float4* __restrict__ pool, unsigned f, unsigned c, float qx, float qy, float qz,
                                                float thr, float (&ld)[5], int (&li)[5], int gl) {
    for (unsigned j = gl; j < c; j += G) {
        float4 p = __ldg(&pool[(size_t)f + j]);
        float d = li_dist2(qx, qy, qz, p.x, p.y, p.z);
        if (d <= 5.0f && d < thr && d < ld[4]) local_insert(ld, li, d, (int)(f + j));
    }
}

// Exact 5-NN of one query by a group of G >= 8 lanes. gd/gi: ascending distances / pool offsets (-1 = missing).
//
//   phase A  seed: probe the 2x2x2 bricks nearest to the query (one round), scan the query's own brick first,
//            derive a cheap bound, scan the other seven only if their box can beat it; merge -> (n, g5).
//            Done if g5 lies within the explored block.
//   phase A' (n < 5, sparse neighbourhood): ring expansion R = 1, 2, .. with the radius^2 = 5 bound until 5 are known.
//   phase B  closure: every unexplored brick whose box intersects the open ball of radius sqrt(g5) is probed and
//            scanned (bounding-box enumeration, no further rings). After it the merged top-5 is exact: a point
//            closer than g5 can only live in a brick that intersects that ball.
template <int G>
__device__ __forceinline__ void knn5_group(const MapDev& M, float qx, float qy, float qz, float (&gd)[5], int (&gi)[5], unsigned gmask,
                                           int lane, int gl) {
    static_assert(G >= 8, "stage 0 needs 8 lanes");
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
    if (!(isfinite(qx) && isfinite(qy) && isfinite(qz))) return;
    const float lim = (float)(LI_CELL_LIMIT - 16 * bc) * g.ds;
    if (fabsf(qx) >= lim || fabsf(qy) >= lim || fabsf(qz) >= lim) return;
    const int cx = li_cell(qx, g.ds), cy = li_cell(qy, g.ds), cz = li_cell(qz, g.ds);
    g.bx = cx >> g.bs; g.by = cy >> g.bs; g.bz = cz >> g.bs;
    const int half = bc >> 1;
    g.dirx = ((cx & (bc - 1)) < half) ? -1 : 1;
    g.diry = ((cy & (bc - 1)) < half) ? -1 : 1;
    g.dirz = ((cz & (bc - 1)) < half) ? -1 : 1;
    const float B = (float)bc * g.ds;
    // slack for float cell assignment / edge products: relative 2^-23 effects, bounded generously
    g.margin = 1e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + 16.0f * B);

    // ---- phase A: seed -------------------------------------------------------------------------------
    {
        unsigned first = 0, count = 0;
        float dbox = INFINITY;
        bool found = false;
        if (gl < 8) {
            int ox = (gl & 1) ? g.dirx : 0, oy = (gl & 2) ? g.diry : 0, oz = (gl & 4) ? g.dirz : 0;
            found = probe_brick(M, g, ox, oy, oz, qx, qy, qz, INFINITY, false, first, count, dbox);
        }
        unsigned fm = __ballot_sync(gmask, found) & gmask;
        const int base_lane = lane - gl;
        float bound = INFINITY;
        if (fm & (1u << base_lane)) {   // the query's own brick
            unsigned f = __shfl_sync(gmask, first, base_lane), c = __shfl_sync(gmask, count, base_lane);
            group_scan_slab<G>(M.pool, f, c, qx, qy, qz, INFINITY, ld, li, gl);
            bound = group_bound5(ld[0], gmask, lane);
            fm &= ~(1u << base_lane);
        }
        while (fm) {
            int src = __ffs(fm) - 1;
            fm &= fm - 1;
            float db = __shfl_sync(gmask, dbox, src);
            if (!(db < bound)) continue;   // this brick cannot hold one of the 5 nearest
            unsigned f = __shfl_sync(gmask, first, src), c = __shfl_sync(gmask, count, src);
            group_scan_slab<G>(M.pool, f, c, qx, qy, qz, bound, ld, li, gl);
            bound = fminf(bound, group_bound5(ld[0], gmask, lane));
        }
    }
    group_merge<G>(ld, li, gd, gi, gmask, lane);
    float r2 = explored_r2(g, 0, qx, qy, qz);
    if (r2 > 5.0f || (gi[4] >= 0 && gd[4] <= r2)) return;
    int Rdone = 0;
    // ---- phase A': fewer than 5 known -> rings with the radius bound ----------------------------------
    if (gi[4] < 0) {
#define LI_RING(RR)                                                                                   \
    if (gi[4] < 0) {                                                                                  \
        knn_ring<G, RR>(M, g, qx, qy, qz, INFINITY, false, ld, li, gmask, lane, gl);                  \
        group_merge<G>(ld, li, gd, gi, gmask, lane);                                                  \
        Rdone = RR;                                                                                   \
        r2 = explored_r2(g, RR, qx, qy, qz);                                                          \
        if (r2 > 5.0f || (gi[4] >= 0 && gd[4] <= r2)) return;                                         \
    }
        LI_RING(1)
        LI_RING(2)
        LI_RING(3)
        LI_RING(4)
#undef LI_RING
        const int Rmax = (int)ceilf(2.2360680f / B) + 1;
        for (int R = 5; R <= Rmax && gi[4] < 0; R++) {
            const int S = 2 * R + 1, total = S * S * S;
            for (int base = 0; base < total; base += G) {
                int idx = base + gl;
                bool want = idx < total;
                int ox = idx % S - R, oy = (idx / S) % S - R, oz = idx / (S * S) - R;
                if (max(abs(ox), max(abs(oy), abs(oz))) < R) want = false;
                unsigned first = 0, count = 0;
                float dbox;
                bool found = false;
                if (want) found = probe_brick(M, g, ox, oy, oz, qx, qy, qz, INFINITY, false, first, count, dbox);
                group_scan_found<G>(M.pool, found, first, count, qx, qy, qz, INFINITY, ld, li, gmask, lane, gl);
            }
            group_merge<G>(ld, li, gd, gi, gmask, lane);
            Rdone = R;
            r2 = explored_r2(g, R, qx, qy, qz);
            if (r2 > 5.0f || (gi[4] >= 0 && gd[4] <= r2)) return;
        }
        if (gi[4] < 0) return;   // fewer than 5 points within the radius: everything within sqrt(5) was explored
    }
    // ---- phase B: closure over the ball of radius sqrt(g5) --------------------------------------------
    {
        const float g5 = gd[4];
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
        const int nxy = nx * ny, total = nxy * nz;
        const float inv_nxy = 1.0f / (float)nxy, inv_nx = 1.0f / (float)nx;
        for (int base = 0; base < total; base += G) {
            int idx = base + gl;
            bool want = idx < total;
            int iz = (int)(((float)idx + 0.5f) * inv_nxy);
            int rem = idx - iz * nxy;
            int iy = (int)(((float)rem + 0.5f) * inv_nx);
            int ix = rem - iy * nx;
            int kx = lx + ix, ky = ly + iy, kz = lz + iz;
            if (kx >= ex0 && kx <= ex1 && ky >= ey0 && ky <= ey1 && kz >= ez0 && kz <= ez1) want = false;   // explored
            unsigned first = 0, count = 0;
            float dbox;
            bool found = false;
            if (want) found = probe_brick(M, g, kx - g.bx, ky - g.by, kz - g.bz, qx, qy, qz, g5, true, first, count, dbox);
            group_scan_found<G>(M.pool, found, first, count, qx, qy, qz, g5, ld, li, gmask, lane, gl);
        }
        group_merge<G>(ld, li, gd, gi, gmask, lane);
    }
}

// ---- search kernel of an ICP pass: world transform + 5-NN for every scan point -----------------------
template <int G>
__global__ void __launch_bounds__(256) k_knn_scan(MapDev M, ScanDev S, PoseD P) {
    const int lane = threadIdx.x & 31;
    const int gl = lane % G;
    const unsigned gmask = Grp<G>::mask(lane);
    const int group_global = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int ngroups = (gridDim.x * blockDim.x) / G;
    for (int q = group_global; q < S.n; q += ngroups) {
        float4 b = __ldg(&S.body[q]);
        float wx, wy, wz;
        li_body_to_world(P, b.x, b.y, b.z, wx, wy, wz);
        float gd[5];
        int gi[5];
        knn5_group<G>(M, wx, wy, wz, gd, gi, gmask, lane, gl);
        if (gl == 0) {
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
    const int lane = threadIdx.x & 31;
    const int gl = lane % G;
    const unsigned gmask = Grp<G>::mask(lane);
    const int group_global = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int ngroups = (gridDim.x * blockDim.x) / G;
    for (int q = group_global; q < n; q += ngroups) {
        float4 p = __ldg(&qpts[q]);
        float gd[5];
        int gi[5];
        knn5_group<G>(M, p.x, p.y, p.z, gd, gi, gmask, lane, gl);
        if (gl == 0) {
#pragma unroll
            for (int k = 0; k < 5; k++) {
                ids[(size_t)q * 5 + k] = gi[k];
                d2[(size_t)q * 5 + k] = (gi[k] >= 0) ? gd[k] : -1.f;
            }
        }
    }
}
