// fused_kernel.cuh -- ONE kernel per search pass: world transform + exact 5-NN on the brick hash with TMA-staged neighbour
// tiles + plane fit + residual + Jacobian row + HtH/Htr reduction (replaces laserMapping.cpp:964-1071 + :1080 and
// KD_TREE::Nearest_Search, ikd_Tree.cpp:349-379,825-968). DESIGN.md section 3c.
//
// Work unit = a WARP TILE of 32 consecutive scan points, handed out by a ticket counter (dynamic, warp granular; no block-wide
// synchronisation anywhere in the kernel):
//   phase 0  one scan point per lane: body point (HBM, or the caller's page-locked buffer over PCIe), pointBodyToWorld in fp64
//   phase A  5-NN, G lanes per scan point, Q = 32/G points at once in lockstep (as knn_kernels.cuh), but
//            * the slabs are not read with per-lane loads: the lane that found a brick issues ONE bulk copy
//              (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes, the 1-D TMA path) of a <= C-point chunk into
//              the group's shared-memory slot, two slots per group (the copy of chunk s+1 is in flight while chunk s is scanned),
//              one mbarrier per warp and slot, armed with the warp's total byte count;
//            * the brick that holds the query is scanned first and twice from shared memory: pass 1 keeps the two smallest
//              distances per lane (three min/max instructions per candidate, no branch) -> the 5th smallest of the group's 2G
//              values is an upper bound tau0 of the 5th-neighbour distance; pass 2 and every other brick only KEEP candidates
//              with d <= tau0, and the first shell is the ball of radius sqrt(tau0): one shell, no guess, no closing step;
//            * a kept candidate is pushed onto a lane-private queue in shared memory (two predicated instructions); the sorted
//              insert into the lane-private top-5 runs once per shell over the queues, max-over-lanes(queue length) times,
//              instead of inside the candidate loop where any one of 32 lanes inserting made the whole warp walk through it;
//   phase B  one scan point per lane again: the five neighbours are gathered by id (L2 hits: the copies just came through),
//            tie order, esti_plane / residual / gate / Jacobian row (icp_kernels.cuh), warp reduce-scatter of the accumulators;
//   per tile the warp's V partial sums go to partials[tile]; a counter tree (fan-out 32) lets the warp that completes a node
//   sum its children IN INDEX ORDER: the result does not depend on which warp ran which tile -> bit-reproducible HtH/Htr.
// The search is exact under the same argument as knn5_lockstep; results are identical to k_knn_scan + k_icp_plane (same
// neighbour sets in the same order, same f32 normals and residuals), exact-distance ties at the 5th place excluded.
#pragma once
#include "common.cuh"
#include "icp_kernels.cuh"
#include "knn_kernels.cuh"

#ifndef LI_FUSED_THREADS
#define LI_FUSED_THREADS 128
#endif
#ifndef LI_FUSED_MIN_BLOCKS
#define LI_FUSED_MIN_BLOCKS 4
#endif
#ifndef LI_FUSED_CHUNK
#define LI_FUSED_CHUNK 32      // points per staged chunk
#endif
#ifndef LI_FUSED_QCAP
#define LI_FUSED_QCAP 12       // queue entries per lane (a chunk can add C/G)
#endif
#define LI_FUSED_FAN 32         // fan-out of the reduction tree

template <int G>
struct FusedCfg {
    static constexpr int Q = 32 / G;
    static constexpr int C = LI_FUSED_CHUNK;
    static constexpr int J = C / G;                                   // candidate steps per chunk
    static constexpr int SLOT_BYTES = C * 16;
    // two slots per group; groups of fewer than 8 lanes share a quarter warp: a 64-byte skew between neighbouring groups
    // makes the quarter's 16-byte reads cover all 32 banks once
    static constexpr int GROUP_BYTES = 2 * SLOT_BYTES + (G < 8 ? 64 : 0);
    static constexpr int WARP_STAGE_BYTES = Q * GROUP_BYTES;
    static constexpr int NW = LI_FUSED_THREADS / 32;
    static_assert(G == 4 || G == 8, "fused kernel: G = 4 or 8");
    static_assert(C % G == 0 && C * 16 % 128 == 0, "chunk size");
    static_assert(LI_FUSED_QCAP > C / G, "queue must hold one chunk's worth of pushes");
};

template <int G>
struct FusedSmem {
    alignas(128) unsigned char stage[FusedCfg<G>::NW][FusedCfg<G>::WARP_STAGE_BYTES];
    alignas(8) unsigned long long bar[FusedCfg<G>::NW][2];
    unsigned long long queue[LI_FUSED_QCAP][LI_FUSED_THREADS];     // [entry][thread]: conflict-free whatever entry a lane is at
    float4 sw[FusedCfg<G>::NW][32];                                 // world points of the warp tile
    float4 sb[FusedCfg<G>::NW][32];                                 // body points
    int sid[FusedCfg<G>::NW][5][32];                                // neighbour ids (pool offsets)
    double tot[FusedCfg<G>::NW][96];                                // root of the reduction tree (one warp uses it)
};

// ---- mbarrier / bulk-copy primitives (the CPU checker of tests/emul runs the copy as a memcpy and the wait as a rendezvous) ----
#ifdef LI_SIMT_EMUL
struct LiStage {
    unsigned char* base;
    int bar;
};
inline unsigned li_redux_add(unsigned v) { return __reduce_add_sync(LI_FULL, v); }
inline unsigned li_redux_max(unsigned v) { return __reduce_max_sync(LI_FULL, v); }
inline void li_mbar_init(unsigned long long*, unsigned) {}
inline void li_mbar_fence_init() {}
inline void li_mbar_expect_tx(unsigned long long*, unsigned) {}
inline void li_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long*) { memcpy(dst, src, bytes); }
inline void li_mbar_wait(unsigned long long*, unsigned) { __syncwarp(); }
#else
__device__ __forceinline__ unsigned li_redux_add(unsigned v) { return __reduce_add_sync(LI_FULL, v); }
__device__ __forceinline__ unsigned li_redux_max(unsigned v) { return __reduce_max_sync(LI_FULL, v); }
__device__ __forceinline__ unsigned li_saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void li_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(li_saddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void li_mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void li_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(li_saddr(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared (the TMA unit moves the bytes; no registers, no per-element instructions), completion
// signalled on the mbarrier as transaction bytes. dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void li_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(li_saddr(dst)), "l"(src),
                 "r"(bytes), "r"(li_saddr(bar))
                 : "memory");
}
__device__ __forceinline__ void li_mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LI_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra LI_DONE;\n"
        "bra LI_WAIT;\n"
        "LI_DONE:\n"
        "}\n" ::"r"(li_saddr(bar)),
        "r"(parity)
        : "memory");
}
#endif

__device__ __forceinline__ unsigned long long li_pack_cand(float d, int id) {
    return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(unsigned)id;
}

// Per-thread view of the search state that lives across chunks / rounds / shells of ONE query batch.
template <int G>
struct FusedLane {
    float ld[5];                    // lane-private sorted top-5 (distances)
    int li[5];                      //   and pool offsets
    int qn;                         // entries in this lane's queue
    unsigned long long* qcol;       // queue column of this thread: entry e at qcol[e * LI_FUSED_THREADS]
    unsigned char* gstage;          // this lane's group staging area (two slots)
    unsigned long long* bar;        // the warp's two barriers
    unsigned parity;                // bit s: parity the next wait on slot s expects (warp-uniform)
    int lane, gl, gbase;
};

// Queue -> lane-private sorted lists. Runs the sorted insert max-over-lanes(queue length) times for the whole warp.
template <int G>
__device__ __forceinline__ void fused_drain(FusedLane<G>& L) {
    const int nmax = (int)li_redux_max((unsigned)L.qn);
    for (int e = 0; e < nmax; e++) {
        if (e < L.qn) {
            const unsigned long long v = L.qcol[(size_t)e * LI_FUSED_THREADS];
            const float d = __uint_as_float((unsigned)(v >> 32));
            if (d < L.ld[4]) local_insert(L.ld, L.li, d, (int)(unsigned)v);
        }
    }
    L.qn = 0;
}

// Arm the slot's barrier with the warp's byte total and let every owner lane issue its copy. Warp-uniform return.
template <int G>
__device__ __forceinline__ bool fused_issue(FusedLane<G>& L, int slot, bool owner, const float4* src, int take) {
    const unsigned bytes = owner ? (unsigned)take * 16u : 0u;
    const unsigned tot = li_redux_add(bytes);
    if (tot == 0u) return false;
    if (L.lane == 0) li_mbar_expect_tx(L.bar + slot, tot);
    __syncwarp();
    if (bytes) li_bulk_g2s(L.gstage + slot * FusedCfg<G>::SLOT_BYTES, src, bytes, L.bar + slot);
    return true;
}
template <int G>
__device__ __forceinline__ void fused_wait(FusedLane<G>& L, int slot) {
    li_mbar_wait(L.bar + slot, (L.parity >> slot) & 1u);
    L.parity ^= 1u << slot;
}

// pass 1 over one staged chunk: the two smallest distances this lane sees (no ids, no branch)
template <int G, int JN>
__device__ __forceinline__ void fused_scan_top2(const float4* __restrict__ sm, int cnt, unsigned long long qxy, float qz, float& a0, float& a1, int gl) {
#pragma unroll
    for (int j = 0; j < JN; j++) {
        const float4 p = sm[j * G + gl];
        float d = li_dist2_packed(qxy, qz, p);
        d = (j * G + gl < cnt) ? d : INFINITY;
        const float t = fmaxf(a0, d);
        a0 = fminf(a0, d);
        a1 = fminf(a1, t);
    }
}
// pass 2 / every other chunk: candidates below tau go onto the lane's queue
template <int G, int JN>
__device__ __forceinline__ void fused_scan_push(const float4* __restrict__ sm, int cnt, int idbase, unsigned long long qxy, float qz, float tau,
                                                unsigned long long* __restrict__ qcol, int& qn, int gl) {
#pragma unroll
    for (int j = 0; j < JN; j++) {
        const float4 p = sm[j * G + gl];
        const float d = li_dist2_packed(qxy, qz, p);
        if (d < tau && j * G + gl < cnt) {
            qcol[(size_t)qn * LI_FUSED_THREADS] = li_pack_cand(d, idbase + j * G + gl);
            qn++;
        }
    }
}
template <int G>
__device__ __forceinline__ void fused_chunk_push(FusedLane<G>& L, int slot, int cnt, int idbase, unsigned long long qxy, float qz, float tau) {
    typedef FusedCfg<G> Cfg;
    // a chunk can add J entries to a queue: make room first (one vote per chunk)
    if (__any_sync(LI_FULL, L.qn > LI_FUSED_QCAP - Cfg::J)) {
        fused_drain<G>(L);
        tau = fminf(tau, L.ld[4]);
    }
    const float4* sm = reinterpret_cast<const float4*>(L.gstage + slot * Cfg::SLOT_BYTES);
    const int cmax = (int)li_redux_max((unsigned)cnt);
    if (cmax > Cfg::C / 2) fused_scan_push<G, Cfg::J>(sm, cnt, idbase, qxy, qz, tau, L.qcol, L.qn, L.gl);
    else fused_scan_push<G, Cfg::J / 2>(sm, cnt, idbase, qxy, qz, tau, L.qcol, L.qn, L.gl);
}

// 5th smallest of the group's 2G values (a0 <= a1 per lane); INFINITY when fewer than five are finite.
template <int G>
__device__ __forceinline__ float fused_fifth(float a0, float a1, int gl, int gbase) {
    float f = INFINITY;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const unsigned hb = __float_as_uint(a0);
        const unsigned mn = grp_min<G>(hb);
        const unsigned who = grp_ballot<G>(hb == mn, gbase);
        const int src = __ffs(who) - 1;
        f = __uint_as_float(mn);
        if (gl == src) {
            a0 = a1;
            a1 = INFINITY;
        }
    }
    return f;
}

// One round of found bricks (one brick per lane at most): seed phase for groups that have no bound yet, then the pipelined
// chunk loop over everything that is left. ALL 32 lanes call. thr5: min(strict bound from the previous merge, smallest float
// above 5); tau0p: the group's seed bound in exclusive form (d < tau0p <=> d <= tau0), INFINITY while unknown.
template <int G>
__device__ __forceinline__ void fused_round(FusedLane<G>& L, const float4* __restrict__ pool, bool found, unsigned first, unsigned count, float dbox,
                                            unsigned long long qxy, float qz, float thr5, float& tau0p) {
    typedef FusedCfg<G> Cfg;
    const int gl = L.gl, gbase = L.gbase;
    unsigned my_first = first;
    unsigned my_rem = found ? count : 0u;

    // ---- seed: the found brick nearest to the query, first 2C points, scanned twice from shared memory
    {
        const bool cand = found && !(tau0p < INFINITY);
        if (__any_sync(LI_FULL, cand)) {
            const unsigned key = cand ? __float_as_uint(dbox) : 0xffffffffu;   // dbox >= 0: its bit pattern is monotone
            const unsigned mn = grp_min<G>(key);
            const unsigned who = grp_ballot<G>(cand && key == mn, gbase);
            const bool has = who != 0u;
            const int src = has ? (__ffs(who) - 1) : 0;
            const unsigned sf = grp_shfl<G>(my_first, src, gbase);
            const unsigned sc = grp_shfl<G>(my_rem, src, gbase);
            const int n0 = has ? (int)min(sc, (unsigned)(2 * Cfg::C)) : 0;
            const int c0 = min(n0, Cfg::C), c1 = n0 - c0;
            const bool owner = has && gl == src;
            const bool i0 = fused_issue<G>(L, 0, owner, pool + sf, c0);
            const bool i1 = fused_issue<G>(L, 1, owner && c1 > 0, pool + sf + Cfg::C, c1);
            if (owner) {
                my_first += (unsigned)n0;
                my_rem -= (unsigned)n0;
            }
            const float4* s0 = reinterpret_cast<const float4*>(L.gstage);
            const float4* s1 = reinterpret_cast<const float4*>(L.gstage + Cfg::SLOT_BYTES);
            float a0 = INFINITY, a1 = INFINITY;
            if (i0) {
                fused_wait<G>(L, 0);
                fused_scan_top2<G, Cfg::J>(s0, c0, qxy, qz, a0, a1, gl);
            }
            if (i1) {
                fused_wait<G>(L, 1);
                fused_scan_top2<G, Cfg::J>(s1, c1, qxy, qz, a0, a1, gl);
            }
            const float t0 = fused_fifth<G>(a0, a1, gl, gbase);
            if (has && t0 < INFINITY) tau0p = __uint_as_float(__float_as_uint(t0) + 1u);   // smallest float above t0
            const float tau = fminf(fminf(thr5, tau0p), L.ld[4]);
            if (i0) fused_chunk_push<G>(L, 0, c0, (int)sf, qxy, qz, tau);
            if (i1) fused_chunk_push<G>(L, 1, c1, (int)sf + Cfg::C, qxy, qz, fminf(tau, L.ld[4]));
            __syncwarp();   // both slots are free again
        }
    }

    // ---- everything else, chunk by chunk; the copy of the next chunk is in flight while the current one is scanned
    unsigned cf = 0, nf = 0;
    int cc = 0, nc = 0;
    bool cown = false, nown = false;
    auto select = [&](unsigned& f, int& take, bool& own) {
        const float thr_g = fminf(thr5, tau0p);
        const bool live = my_rem > 0u && dbox < thr_g;
        const unsigned fm = grp_ballot<G>(live, gbase);
        const bool has = fm != 0u;
        const int src = has ? (__ffs(fm) - 1) : 0;
        f = grp_shfl<G>(my_first, src, gbase);
        const unsigned r = grp_shfl<G>(my_rem, src, gbase);
        take = has ? (int)min(r, (unsigned)Cfg::C) : 0;
        own = has && gl == src;
        if (own) {
            my_first += (unsigned)take;
            my_rem -= (unsigned)take;
        }
    };
    int slot = 0;
    select(cf, cc, cown);
    bool cur_issued = fused_issue<G>(L, slot, cown, pool + cf, cc);
    while (cur_issued) {
        select(nf, nc, nown);
        const bool nxt_issued = fused_issue<G>(L, slot ^ 1, nown, pool + nf, nc);
        fused_wait<G>(L, slot);
        const float tau = fminf(fminf(thr5, tau0p), L.ld[4]);
        fused_chunk_push<G>(L, slot, cc, (int)cf, qxy, qz, tau);
        __syncwarp();   // every lane is done with this slot before it is filled again
        cf = nf;
        cc = nc;
        cur_issued = nxt_issued;
        slot ^= 1;
    }
}

// Exact 5-NN of Q queries by one warp in lockstep (the staged counterpart of knn5_lockstep; same shell invariant).
template <int G>
__device__ __forceinline__ void fused_knn5(const MapDev& M, FusedLane<G>& L, float rho2, bool valid, float qx, float qy, float qz, float (&gd)[5],
                                           int (&gi)[5]) {
    const int gl = L.gl, gbase = L.gbase;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        L.ld[i] = INFINITY;
        L.li[i] = -1;
        gd[i] = INFINITY;
        gi[i] = -1;
    }
    L.qn = 0;
    const int bs = M.bshift;
    const float ds = M.ds;
    const int bc = 1 << bs;
    const float B = (float)bc * ds;
    const float lim = (float)(LI_CELL_LIMIT - 16 * bc) * ds;
    const bool act = valid && isfinite(qx) && isfinite(qy) && isfinite(qz) && fabsf(qx) < lim && fabsf(qy) < lim && fabsf(qz) < lim;
    if (!act) {
        qx = 0.f; qy = 0.f; qz = 0.f;
    }
    const float margin = 1e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + 16.0f * B);
    const float inv_ds = 1.0f / ds;
    const float slk = 0.02f + 4e-7f * (fabsf(qx) + fabsf(qy) + fabsf(qz)) * inv_ds;
    const unsigned long long qxy = li_pack_f32x2(qx, qy);
    const float cap5 = __uint_as_float(0x40a00001u);   // smallest float above 5: d <= 5 <=> d < cap5
    float tau0p = INFINITY;

    // ---- the brick that holds the query: seed + complete scan (it is skipped by the shells below)
    const int ox = li_cell(qx, ds) >> bs, oy = li_cell(qy, ds) >> bs, oz = li_cell(qz, ds) >> bs;
    {
        unsigned first = 0, count = 0;
        bool found = false;
        if (act && gl == 0) {
            found = li_brick_find(M.ent, M.mask, li_pack_key(ox, oy, oz), first, count);
            found = found && count > 0u;
        }
        fused_round<G>(L, M.pool, found, first, count, 0.f, qxy, qz, cap5, tau0p);
    }

    bool done = !act;
    float lo2 = 0.f;
    // with a seed bound the first shell is the ball that must hold the answer; without one, the guessed radius
    float hi2 = (tau0p < INFINITY) ? fminf(tau0p, 5.0f) : rho2;
    while (__any_sync(LI_FULL, !done)) {
        const bool need = !done;
        const bool last = hi2 >= 5.0f;   // the radius bound d2 <= 5 is inclusive (ikd_Tree.cpp:842)
        const float r = sqrtf(fminf(hi2, 5.0f)) * (1.0f + 1e-6f) + margin;
        const int lx = (int)floorf((qx - r) * inv_ds - slk) >> bs, hx = (int)floorf((qx + r) * inv_ds + slk) >> bs;
        const int ly = (int)floorf((qy - r) * inv_ds - slk) >> bs, hy = (int)floorf((qy + r) * inv_ds + slk) >> bs;
        const int lz = (int)floorf((qz - r) * inv_ds - slk) >> bs, hz = (int)floorf((qz + r) * inv_ds + slk) >> bs;
        const int nx = hx - lx + 1, ny = hy - ly + 1, nz = hz - lz + 1;
        const int nxy = nx * ny;
        const int total = need ? nxy * nz : 0;
        const float inv_nxy = 1.0f / (float)nxy, inv_nx = 1.0f / (float)nx;
        const float thr = (gi[4] >= 0) ? gd[4] : INFINITY;
        const float thr5 = fminf(thr, cap5);
        for (int base = 0; __any_sync(LI_FULL, base < total); base += G) {
            const int idx = base + gl;
            const bool want = idx < total;
            const int iz = (int)(((float)idx + 0.5f) * inv_nxy);
            const int rem = idx - iz * nxy;
            const int iy = (int)(((float)rem + 0.5f) * inv_nx);
            const int ix = rem - iy * nx;
            unsigned first = 0, count = 0;
            float dbox = INFINITY;
            bool found = false;
            const int kx = lx + ix, ky = ly + iy, kz = lz + iz;
            if (want && !(kx == ox && ky == oy && kz == oz)) {
                float lox = (float)(kx << bs) * ds - margin, hix = (float)((kx + 1) << bs) * ds + margin;
                float loy = (float)(ky << bs) * ds - margin, hiy = (float)((ky + 1) << bs) * ds + margin;
                float loz = (float)(kz << bs) * ds - margin, hiz = (float)((kz + 1) << bs) * ds + margin;
                float ex = fmaxf(0.f, fmaxf(lox - qx, qx - hix));
                float ey = fmaxf(0.f, fmaxf(loy - qy, qy - hiy));
                float ez = fmaxf(0.f, fmaxf(loz - qz, qz - hiz));
                dbox = (ex * ex + ey * ey + ez * ez) * (1.0f - 1e-6f);
                const bool in_shell = dbox >= lo2 && (last ? dbox <= 5.0f : dbox < hi2);
                if (in_shell && dbox < fminf(thr5, tau0p)) {
                    found = li_brick_find(M.ent, M.mask, li_pack_key(kx, ky, kz), first, count);
                    found = found && count > 0u;
                }
            }
            fused_round<G>(L, M.pool, found, first, count, dbox, qxy, qz, thr5, tau0p);
        }
        fused_drain<G>(L);
        group_merge<G>(L.ld, L.li, gd, gi, gl, gbase);
        if (need) {
            const bool full = gi[4] >= 0;
            if (last || (full && gd[4] <= hi2)) {
                done = true;
            } else {
                lo2 = hi2;
                hi2 = full ? fminf(gd[4] * (1.0f + 1e-6f), 5.0f) : fminf(4.0f * hi2, 5.0f);
                if (full && hi2 >= 5.0f) hi2 = 5.0f;
            }
        }
    }
}

// ---- deterministic reduction tree over warp tiles -------------------------------------------------------------------
// level 0 = one row of V doubles per warp tile; a node of level l+1 is the IN-ORDER sum of its <= 32 children. The warp that
// delivers the last child of a node sums the node and moves up; the warp that completes the root writes the 160-double result.
// buf: rows of all levels back to back; cnt: one counter per inner node, all zero between launches (the summing warp resets it).
template <bool IMU>
__device__ __forceinline__ void fused_tree_submit(double (&acc)[AccLayout<IMU>::K], int tile, int ntiles, double* __restrict__ buf,
                                                  unsigned* __restrict__ cnt, double* __restrict__ out160, double* __restrict__ s_tot, int lane) {
    typedef AccLayout<IMU> AL;
    int idx = tile, nlev = ntiles;
    size_t row_off = 0;
    int cnt_off = 0;
    while (nlev > 1) {
        double* row = buf + (row_off + (size_t)idx) * AL::V;
#pragma unroll
        for (int k = 0; k < AL::K; k++) row[32 * k + lane] = acc[k];
        __threadfence();
        __syncwarp();
        const int node = idx / LI_FUSED_FAN;
        const int nnodes = (nlev + LI_FUSED_FAN - 1) / LI_FUSED_FAN;
        const int size = min(LI_FUSED_FAN, nlev - node * LI_FUSED_FAN);
        unsigned old = 0;
        if (lane == 0) old = atomicAdd(&cnt[cnt_off + node], 1u);
        old = __shfl_sync(LI_FULL, old, 0);
        if ((int)old != size - 1) return;
        __threadfence();
        if (lane == 0) cnt[cnt_off + node] = 0u;
        const double* ch = buf + (row_off + (size_t)node * LI_FUSED_FAN) * AL::V;
#pragma unroll
        for (int k = 0; k < AL::K; k++) acc[k] = 0.0;
        for (int c = 0; c < size; c++) {
#pragma unroll
            for (int k = 0; k < AL::K; k++) acc[k] += __ldcg(ch + (size_t)c * AL::V + 32 * k + lane);
        }
        row_off += (size_t)nlev;
        cnt_off += nnodes;
        idx = node;
        nlev = nnodes;
    }
    // root: expand the packed accumulators to [HtH 144 | Htr 12 | res_sq | m | 0 0]
#pragma unroll
    for (int k = 0; k < AL::K; k++) s_tot[32 * k + lane] = acc[k];
    __syncwarp();
    for (int o = lane; o < 160; o += 32) {
        int src = -1;
        if (o < 144) {
            int a = o / 12, b = o % 12;
            if (a > b) { int tt = a; a = b; b = tt; }
            if (b < AL::NC) src = a * AL::NC - a * (a - 1) / 2 + (b - a);
        } else if (o < 156) {
            int a = o - 144;
            if (a < AL::NC) src = AL::NT + a;
        } else if (o == 156) {
            src = AL::NT + AL::NC;
        } else if (o == 157) {
            src = AL::NT + AL::NC + 1;
        }
        out160[o] = (src >= 0) ? s_tot[src] : 0.0;
    }
}

// rows / counters the tree needs for n tiles
inline void fused_tree_sizes(long long ntiles, long long& rows, long long& counters) {
    rows = 0;
    counters = 0;
    long long n = ntiles > 0 ? ntiles : 1;
    while (n > 1) {
        rows += n;
        n = (n + LI_FUSED_FAN - 1) / LI_FUSED_FAN;
        counters += n;
    }
    rows += 1;
    counters += 1;
}

// ---- the kernel ------------------------------------------------------------------------------------------------------
// ticket: monotone counter, never reset; ticket_base = its value before this launch (the host adds ntiles + warps per launch).
template <int G, bool IMU, bool HOST>
__global__ void __launch_bounds__(LI_FUSED_THREADS, LI_FUSED_MIN_BLOCKS)
k_icp_fused(MapDev M, ScanDev S, PoseD P, float rho2, const float* __restrict__ raw, int stride, unsigned* __restrict__ ticket, unsigned ticket_base,
            double* __restrict__ tree_buf, unsigned* __restrict__ tree_cnt, double* __restrict__ out160) {
    typedef FusedCfg<G> Cfg;
    typedef AccLayout<IMU> AL;
#ifdef LI_SIMT_EMUL
    static FusedSmem<G> sm_storage;
    FusedSmem<G>& sm = sm_storage;
#else
    extern __shared__ __align__(128) unsigned char li_fused_smem_raw[];
    FusedSmem<G>& sm = *reinterpret_cast<FusedSmem<G>*>(li_fused_smem_raw);
#endif
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    FusedLane<G> L;
    L.lane = lane;
    L.gl = lane % G;
    L.gbase = (lane / G) * G;
    L.qcol = &sm.queue[0][threadIdx.x];
    L.gstage = sm.stage[warp] + (lane / G) * Cfg::GROUP_BYTES;
    L.bar = sm.bar[warp];
    L.parity = 0u;
    L.qn = 0;
    if (lane == 0) {
        li_mbar_init(L.bar + 0, 1u);
        li_mbar_init(L.bar + 1, 1u);
        li_mbar_fence_init();
    }
    __syncwarp();

    const int ntiles = (S.n + 31) >> 5;
    for (;;) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1u) - ticket_base;
        t = __shfl_sync(LI_FULL, t, 0);
        if (t >= (unsigned)ntiles) break;
        const int tile = (int)t;
        const int q = tile * 32 + lane;
        const bool valid = q < S.n;

        // ---- phase 0: body point, pointBodyToWorld (laserMapping.cpp:209-220)
        {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (HOST) {
                if (stride == 3 || stride == 4) {   // contiguous read of the tile's floats, redistributed through shared memory
                    float* fs = reinterpret_cast<float*>(sm.sb[warp]);
                    const long long f0 = (long long)tile * 32 * stride, fe = (long long)S.n * stride;
                    for (int i = lane; i < 32 * stride; i += 32)
                        if (f0 + i < fe) fs[i] = raw[f0 + i];
                    __syncwarp();
                    float bx = fs[lane * stride], by = fs[lane * stride + 1], bz = fs[lane * stride + 2];
                    __syncwarp();
                    if (valid) b = make_float4(bx, by, bz, 0.f);
                } else if (valid) {
                    const float* s = raw + (size_t)q * stride;
                    b = make_float4(s[0], s[1], s[2], 0.f);
                }
                if (valid) S.body[q] = b;
            } else if (valid) {
                b = __ldg(&S.body[q]);
            }
            float wx = 0.f, wy = 0.f, wz = 0.f;
            if (valid) {
                li_body_to_world(P, b.x, b.y, b.z, wx, wy, wz);
                S.world[q] = make_float4(wx, wy, wz, 0.f);
            }
            sm.sb[warp][lane] = b;
            sm.sw[warp][lane] = make_float4(wx, wy, wz, 0.f);
            __syncwarp();
        }

        // ---- phase A: 5-NN, Q scan points at a time
#pragma unroll 1
        for (int bq = 0; bq < 32; bq += Cfg::Q) {
            const int lq = bq + lane / G;
            const float4 w = sm.sw[warp][lq];
            float gd[5];
            int gi[5];
            fused_knn5<G>(M, L, rho2, tile * 32 + lq < S.n, w.x, w.y, w.z, gd, gi);
            if (L.gl == 0) {
#pragma unroll
                for (int k = 0; k < 5; k++) sm.sid[warp][k][lq] = gi[k];
            }
        }
        __syncwarp();

        // ---- phase B: plane, residual, Jacobian row, accumulation (laserMapping.cpp:989-1071)
        double acc[AL::K];
#pragma unroll
        for (int k = 0; k < AL::K; k++) acc[k] = 0.0;
        {
            double row[AL::NC];
            double r = 0.0;
            bool sel = false;
#pragma unroll
            for (int i = 0; i < AL::NC; i++) row[i] = 0.0;
            if (valid) {
                int id[5];
#pragma unroll
                for (int k = 0; k < 5; k++) id[k] = sm.sid[warp][k][lane];
                if (id[4] >= 0) {
                    const float4 b = sm.sb[warp][lane];
                    const float4 w = sm.sw[warp][lane];
                    float4 nb[5];
#pragma unroll
                    for (int k = 0; k < 5; k++) nb[k] = __ldg(&M.pool[id[k]]);
                    float d[5];
#pragma unroll
                    for (int k = 0; k < 5; k++) d[k] = li_dist2(w.x, w.y, w.z, nb[k].x, nb[k].y, nb[k].z);
                    tie_order(nb, id, d);
                    float4 nvec;
                    sel = plane_and_row<IMU>(P, b.x, b.y, b.z, w.x, w.y, w.z, nb, nvec, row, r);
                    if (sel) S.normvec[q] = nvec;
                }
#pragma unroll
                for (int k = 0; k < 5; k++) S.near_ids[(size_t)q * 5 + k] = id[k];
                S.selected[q] = sel ? 1 : 0;
            }
            warp_accumulate<IMU>(row, r, sel, lane, acc);
        }
        __syncwarp();
        fused_tree_submit<IMU>(acc, tile, ntiles, tree_buf, tree_cnt, out160, sm.tot[warp], lane);
        __syncwarp();
    }
}
