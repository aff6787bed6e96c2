// knn_wq.cuh -- exact bounded 5-NN on the brick hash, ONE WARP PER SCAN POINT (replaces KD_TREE::Nearest_Search,
// ikd_Tree.cpp:349-379, Search :825-968). DESIGN.md section 3c.
//
// Why another shape. The lockstep search of knn_kernels.cuh marches 32/G scan points through every loop together: a warp
// pays for its slowest group in every shell, the lanes of a group insert into lane-private sorted lists (27 % of the
// instructions at 9 active lanes), and a shell scans every brick of its annulus although the first brick usually settles
// the 5th-neighbour bound. Here the 32 lanes work on ONE scan point, so every loop is as long as that point needs:
//
//   * bricks are enumerated as rings of the query's home brick (ring 1 = the 3x3x3 block, one brick per lane; ring r =
//     the (2r+1)^3 block without the (2r-1)^3 core), box distance dbox per lane exactly as in knn5_lockstep;
//   * probing is lazy: first only the bricks with dbox < rho^2 (the seed radius), then those below the current bound;
//   * found bricks are scanned NEAREST FIRST (warp min-reduction over dbox) and the walk stops at the first brick whose
//     dbox is not below the current 5th best -- only bricks that can hold a neighbour are ever read;
//   * a slab is read by the whole warp, 512 contiguous bytes per load instruction; a candidate below the bound is
//     appended to a per-warp list in shared memory through a ballot (no sorted insert in the candidate loop);
//   * the list is cut back to its five smallest once per brick: five rounds of {lane min, REDUX.MIN, ballot}.
//
// The search is exact under the argument of knn5_lockstep: a brick is skipped only when its (conservative) box distance
// is not below the 5th best known at that time, and it ends only when five neighbours are known whose 5th distance does
// not exceed the distance to the nearest unenumerated brick (cover2), or when the whole ball d^2 <= 5 has been covered.
// Semantics matched: candidates with d2 <= 5 only (ikd_Tree.cpp:842, sic), fp32
// (dx*dx+dy*dy)+dz*dz without FMA, ascending output; exact-distance ties at the 5th place are traversal dependent in the
// reference and excluded from the parity claim (here: the first candidate in walk order wins).
//
// Scheduling: tiles of LI_WQ_TILE consecutive scan points handed out by a monotone ticket (the next ticket is requested
// before the current tile is processed); lanes < TILE load and transform the tile's points (fp64, one point per lane), the
// coordinates reach the search through shuffles. HOST = true reads the caller's page-locked scan in place (one contiguous
// read of TILE*stride floats per tile over PCIe) and leaves the packed copy in S.body (liinit_scan_attach_host).
#pragma once
#include "common.cuh"

#ifndef LI_WQ_THREADS
#define LI_WQ_THREADS 256
#endif
#ifndef LI_WQ_MIN_BLOCKS
#define LI_WQ_MIN_BLOCKS 4
#endif
#ifndef LI_WQ_TILE
#define LI_WQ_TILE 8
#endif
#define LI_WQ_LIST 64            // list entries per warp: <= 32 carried into a load round + <= 32 appended by it
#define LI_WQ_INF 0x7f800000u

// The warp's candidate list lives in shared memory as float4 {x, y, z, d2}: the COORDINATES travel with the distance, so the five
// survivors can be written out as Nearest_Points copies (ScanDev::near_xyz) without going back to the map.

// Cut the list back to its five smallest entries (ascending, ties: lane order). On return the list holds them in [0, n),
// gd mirrors their distances (+inf for missing ranks), nfound = n. All 32 lanes must call.
__device__ __forceinline__ void wq_select(float4* __restrict__ list, int& n, float (&gd)[5], int lane) {
    __syncwarp();
    float4 e0 = make_float4(0.f, 0.f, 0.f, INFINITY), e1 = e0;
    if (lane < n) e0 = list[lane];
    if (lane + 32 < n) e1 = list[lane + 32];
    __syncwarp();   // every entry is in a register before the front of the list is rewritten
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const unsigned d0 = __float_as_uint(e0.w), d1 = __float_as_uint(e1.w);   // d2 >= 0: the bit pattern orders like the value
        const unsigned lm = min(d0, d1);
        const unsigned m = __reduce_min_sync(LI_FULL, lm);
        const unsigned who = __ballot_sync(LI_FULL, lm == m);
        const int src = __ffs(who) - 1;
        gd[k] = __uint_as_float(m);
        if (lane == src && m != LI_WQ_INF) {
            if (d0 == m) {
                list[k] = e0;
                e0.w = INFINITY;
            } else {
                list[k] = e1;
                e1.w = INFINITY;
            }
        }
    }
    n = min(n, 5);
    __syncwarp();
}

// Scan one slab [f, f+c) with the whole warp; candidates with d < tau are appended to the list. tau follows the 5th best
// whenever the list has to be cut inside the slab (more than 32 entries carried). Returns with the list possibly longer
// than 5 (the caller cuts it once per brick).
__device__ __forceinline__ void wq_scan_slab(const float4* __restrict__ pool, unsigned f, unsigned c, unsigned long long qxy, float qz,
                                             float& tau, float4* __restrict__ list, int& n, float (&gd)[5], int lane) {
    const float cap5 = __uint_as_float(0x40a00001u);   // smallest float above 5: d <= 5 <=> d < cap5
    const unsigned lt = (1u << lane) - 1u;
    const float4* __restrict__ ptr = pool + f + lane;
    for (unsigned base = 0; base < c; base += 64) {
        const bool ok0 = base + lane < c, ok1 = base + 32 + lane < c;
        float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
        if (ok0) p0 = __ldg(ptr + base);
        if (ok1) p1 = __ldg(ptr + base + 32);
        {
            const float d = li_dist2_packed(qxy, qz, p0);
            const bool pass = ok0 && d < tau;
            const unsigned bal = __ballot_sync(LI_FULL, pass);
            if (bal) {
                if (pass) list[n + __popc(bal & lt)] = make_float4(p0.x, p0.y, p0.z, d);
                n += __popc(bal);
                if (n > 32) {
                    wq_select(list, n, gd, lane);
                    tau = fminf(gd[4], cap5);
                }
            }
        }
        if (base + 32 < c) {   // warp-uniform
            const float d = li_dist2_packed(qxy, qz, p1);
            const bool pass = ok1 && d < tau;
            const unsigned bal = __ballot_sync(LI_FULL, pass);
            if (bal) {
                if (pass) list[n + __popc(bal & lt)] = make_float4(p1.x, p1.y, p1.z, d);
                n += __popc(bal);
                if (n > 32) {
                    wq_select(list, n, gd, lane);
                    tau = fminf(gd[4], cap5);
                }
            }
        }
    }
}

// Walk the bricks the lanes found (at most one per lane), nearest first, while their box distance is below the 5th best
// (gd[4] = +inf while fewer than five are known). A brick that is not walked is dropped: the bound only falls.
__device__ __forceinline__ void wq_walk_found(const float4* __restrict__ pool, bool found, unsigned first, unsigned count, float dbox,
                                              unsigned long long qxy, float qz, float& tau, float4* __restrict__ list, int& n, int& ntop,
                                              float (&gd)[5], int lane) {
    const float cap5 = __uint_as_float(0x40a00001u);
    while (true) {
        const unsigned key = found ? __float_as_uint(dbox) : 0xffffffffu;   // dbox >= 0: its bit pattern orders like the value
        const unsigned m = __reduce_min_sync(LI_FULL, key);
        if (m == 0xffffffffu) break;
        if (!(__uint_as_float(m) < gd[4])) break;   // every remaining brick is at least this far
        const unsigned who = __ballot_sync(LI_FULL, key == m);
        const int src = __ffs(who) - 1;
        const unsigned f = __shfl_sync(LI_FULL, first, src);
        const unsigned c = __shfl_sync(LI_FULL, count, src);
        if (lane == src) found = false;
        wq_scan_slab(pool, f, c, qxy, qz, tau, list, n, gd, lane);
        if (n > ntop) {
            wq_select(list, n, gd, lane);
            ntop = n;
            tau = fminf(gd[4], cap5);
        }
    }
}

// Exact 5-NN of ONE query by the whole warp. q* are warp-uniform; ALL 32 lanes must call. Returns the number of neighbours
// found (0..5); they are list[0 .. found) = {x, y, z, d2}, ascending.
__device__ __forceinline__ int knn5_warp(const MapDev& M, float rho2, float qx, float qy, float qz, float4* __restrict__ list, int lane) {
    float gd[5];   // distances of the current best five (+inf = rank not filled), uniform over the warp
#pragma unroll
    for (int i = 0; i < 5; i++) gd[i] = INFINITY;
    const int bs = M.bshift;
    const float ds = M.ds;
    const int bc = 1 << bs;
    const float B = (float)bc * ds;
    const float lim = (float)(LI_CELL_LIMIT - 16 * bc) * ds;
    const bool act = isfinite(qx) && isfinite(qy) && isfinite(qz) && fabsf(qx) < lim && fabsf(qy) < lim && fabsf(qz) < lim;
    if (!act) return 0;   // warp-uniform
    // slack for float cell assignment / edge products: relative 2^-23 effects, bounded generously (as knn5_lockstep)
    const float margin = 1e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + 16.0f * B);
    const float inv_ds = 1.0f / ds;
    // home brick: the reciprocal multiply may land one cell off near a face; only the enumeration centre depends on it,
    // the coverage radius below is computed from the block's real faces
    const int hx = (int)floorf(qx * inv_ds) >> bs, hy = (int)floorf(qy * inv_ds) >> bs, hz = (int)floorf(qz * inv_ds) >> bs;
    const unsigned long long qxy = li_pack_f32x2(qx, qy);
    const float cap5 = __uint_as_float(0x40a00001u);
    float tau = cap5;
    int n = 0, ntop = 0;

    for (int ring = 1;; ring++) {
        // distance from the query to the nearest face of the (2*ring+1)^3 block: every brick outside has dbox >= cover2
        float rc;
        {
            const float lox = (float)((hx - ring) << bs) * ds, hix = (float)((hx + ring + 1) << bs) * ds;
            const float loy = (float)((hy - ring) << bs) * ds, hiy = (float)((hy + ring + 1) << bs) * ds;
            const float loz = (float)((hz - ring) << bs) * ds, hiz = (float)((hz + ring + 1) << bs) * ds;
            rc = fminf(fminf(fminf(qx - lox, hix - qx), fminf(qy - loy, hiy - qy)), fminf(qz - loz, hiz - qz)) - margin;
            rc = fmaxf(rc, 0.f);
        }
        const float cover2 = rc * rc * (1.0f - 1e-6f);
        const bool last_ring = cover2 > 5.0f * (1.0f + 1e-6f);   // no brick outside the block can hold a point with d2 <= 5
        const int side = 2 * ring + 1, total = side * side * side;
        const float inv_side = 1.0f / (float)side, inv_side2 = inv_side * inv_side;
        for (int base = 0; base < total; base += 32) {
            const int idx = base + lane;
            bool want = idx < total;
            const int iz = (int)(((float)idx + 0.5f) * inv_side2);
            const int rem = idx - iz * side * side;
            const int iy = (int)(((float)rem + 0.5f) * inv_side);
            const int ix = rem - iy * side;
            const int ox = ix - ring, oy = iy - ring, oz = iz - ring;
            // the core was enumerated by the previous ring
            if (ring > 1 && max(max(abs(ox), abs(oy)), abs(oz)) < ring) want = false;
            const int kx = hx + ox, ky = hy + oy, kz = hz + oz;
            float dbox = INFINITY;
            if (want) {
                const float lox = (float)(kx << bs) * ds - margin, hix = (float)((kx + 1) << bs) * ds + margin;
                const float loy = (float)(ky << bs) * ds - margin, hiy = (float)((ky + 1) << bs) * ds + margin;
                const float loz = (float)(kz << bs) * ds - margin, hiz = (float)((kz + 1) << bs) * ds + margin;
                const float ex = fmaxf(0.f, fmaxf(lox - qx, qx - hix));
                const float ey = fmaxf(0.f, fmaxf(loy - qy, qy - hiy));
                const float ez = fmaxf(0.f, fmaxf(loz - qz, qz - hiz));
                dbox = (ex * ex + ey * ey + ez * ez) * (1.0f - 1e-6f);
                want = dbox <= 5.0f;   // the radius bound d2 <= 5 is inclusive (ikd_Tree.cpp:842)
            }
            bool pending = want;
            // ring 1: two probe rounds (seed radius, then whatever the bound still admits); outer rings: one
            for (int round = (ring == 1 ? 0 : 1); round < 2; round++) {
                const float limit = (round == 0) ? rho2 : INFINITY;
                const bool probe = pending && dbox < limit && dbox < gd[4];
                if (!__any_sync(LI_FULL, probe)) continue;
                unsigned first = 0, count = 0;
                bool found = false;
                if (probe) {
                    pending = false;
                    found = li_brick_find(M.ent, M.mask, li_pack_key(kx, ky, kz), first, count);
                    found = found && count > 0u;
                }
                wq_walk_found(M.pool, found, first, count, dbox, qxy, qz, tau, list, n, ntop, gd, lane);
                // five known within the probed radius of this block: nothing closer can be left anywhere
                if (round == 0 && gd[4] <= fminf(rho2, cover2)) return n;
            }
        }
        if (gd[4] <= cover2) return n;
        if (last_ring) return n;
    }
}

// ---- search kernel of an ICP pass: world transform + 5-NN for every scan point -----------------------
// Output per scan point: S.world and the five Nearest_Points copies S.near_xyz (w = 1 found / 0 missing rank), ascending.
template <bool HOST>
__global__ void __launch_bounds__(LI_WQ_THREADS, LI_WQ_MIN_BLOCKS)
k_knn_wq(MapDev M, ScanDev S, PoseD P, float rho2, const float* __restrict__ raw, int stride, unsigned* __restrict__ ticket,
         unsigned ticket_base) {
    constexpr int T = LI_WQ_TILE;
    __shared__ float4 s_list[LI_WQ_THREADS / 32][LI_WQ_LIST];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* list = s_list[warp];
    const unsigned ntiles = (unsigned)((S.n + T - 1) / T);
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(ticket, 1u) - ticket_base;
    t = __shfl_sync(LI_FULL, t, 0);
    while (t < ntiles) {   // warp-uniform
        unsigned tn = 0;
        if (lane == 0) tn = atomicAdd(ticket, 1u) - ticket_base;   // the next tile's ticket travels while this tile is searched
        const int q0 = (int)t * T;
        const int ql = q0 + lane;
        const bool mine = lane < T && ql < S.n;
        float bx = 0.f, by = 0.f, bz = 0.f, wx = 0.f, wy = 0.f, wz = 0.f;
        if (HOST) {
            if (stride == 3) {   // TILE * 3 consecutive floats: one contiguous request
                float v = 0.f;
                const long long fi = (long long)q0 * 3 + lane;
                if (lane < 3 * T && fi < (long long)S.n * 3) v = raw[fi];
                const int s0 = (lane < T) ? 3 * lane : 0;
                bx = __shfl_sync(LI_FULL, v, s0);
                by = __shfl_sync(LI_FULL, v, s0 + 1);
                bz = __shfl_sync(LI_FULL, v, s0 + 2);
            } else if (mine) {
                const float* s = raw + (size_t)ql * stride;
                bx = s[0]; by = s[1]; bz = s[2];
            }
            if (mine) S.body[ql] = make_float4(bx, by, bz, 0.f);
        } else if (mine) {
            const float4 b = __ldg(&S.body[ql]);
            bx = b.x; by = b.y; bz = b.z;
        }
        if (mine) {
            li_body_to_world(P, bx, by, bz, wx, wy, wz);
            S.world[ql] = make_float4(wx, wy, wz, 0.f);
        }
#pragma unroll 1
        for (int i = 0; i < T; i++) {
            const int q = q0 + i;
            if (q >= S.n) break;   // warp-uniform
            const float qx = __shfl_sync(LI_FULL, wx, i), qy = __shfl_sync(LI_FULL, wy, i), qz = __shfl_sync(LI_FULL, wz, i);
            const int nf = knn5_warp(M, rho2, qx, qy, qz, list, lane);
            if (lane < 5) {
                float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lane < nf) {
                    e = list[lane];
                    e.w = 1.0f;
                }
                S.near_xyz[(size_t)q * 5 + lane] = e;
            }
            __syncwarp();   // the list is rewritten by the next query
        }
        t = __shfl_sync(LI_FULL, tn, 0);
    }
}

// ---- stand-alone Nearest_Search for arbitrary world-frame queries (liinit_map_nearest_search) ---------
// out_xyz [nq*15], d2 [nq*5] (-1 = missing rank), cnt_ids [nq*5]: 0 for a found rank, -1 for a missing one
__global__ void __launch_bounds__(LI_WQ_THREADS, LI_WQ_MIN_BLOCKS)
k_knn_wq_queries(MapDev M, const float4* __restrict__ qpts, int nq, int* __restrict__ cnt_ids, float* __restrict__ out_xyz,
                 float* __restrict__ d2, float rho2) {
    __shared__ float4 s_list[LI_WQ_THREADS / 32][LI_WQ_LIST];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* list = s_list[warp];
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int q = warp_global; q < nq; q += nwarps) {
        const float4 p = __ldg(&qpts[q]);
        const int nf = knn5_warp(M, rho2, p.x, p.y, p.z, list, lane);
        if (lane < 5) {
            float4 e = make_float4(0.f, 0.f, 0.f, -1.f);
            if (lane < nf) e = list[lane];
            const size_t o = (size_t)q * 5 + lane;
            cnt_ids[o] = (lane < nf) ? 0 : -1;
            out_xyz[3 * o] = e.x;
            out_xyz[3 * o + 1] = e.y;
            out_xyz[3 * o + 2] = e.z;
            d2[o] = e.w;
        }
        __syncwarp();
    }
}
