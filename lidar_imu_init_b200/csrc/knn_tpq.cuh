// knn_tpq.cuh -- exact bounded 5-NN, one scan point per THREAD with warp-cooperative shared-memory staging.
//
// Why this shape: the group-cooperative kernel (knn_kernels.cuh) spends its issue slots on
// shuffles / merges and on lanes idling behind the slowest group; a thread-per-query search has neither, but its
// loads are scattered (every lane walks a different brick slab: 32 L1 wavefronts per LDG). Here the warp copies the
// 32 slabs (one per lane) into shared memory with coalesced 16-byte cp.async (16*CH bytes per slab, contiguous),
// and every lane then scans ITS tile with conflict-free LDS.128 (row stride CH+1 float4). All arithmetic is
// lane-private: a sorted top-5 in registers, no shuffles in the inner loop, no merge.
//
// Search = the same shell iteration on the brick-box distance as knn5_lockstep (exact):
//     invariant: every brick with dbox < lo2 has been scanned
//     step:      scan bricks with lo2 <= dbox < hi2 (bounding box of the ball of radius sqrt(hi2));
//                stop if 5 are known and d5 <= hi2, or hi2 >= 5; else lo2 = hi2, hi2 = d5 (closing) or 4*hi2 (sparse).
#pragma once
#include "common.cuh"
#include "knn_kernels.cuh"

__device__ __forceinline__ void li_cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void li_cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// CH: points per staged chunk and lane (tile row = CH+1 float4 to spread the rows over the 16-byte bank groups).
// NB: bricks probed per lane before their slabs are staged (the probes are independent loads).
template <int CH, int NB>
struct TpqCfg {
    static constexpr int ROW = CH + 1;
    static constexpr int WARP_TILE_F4 = 32 * ROW;
};

template <int CH, int NB>
__device__ __forceinline__ void knn5_tpq(const MapDev& M, float rho2, bool valid, float qx, float qy, float qz, float (&ld)[5],
                                         int (&li)[5], float4* __restrict__ tile /* this warp's [32][CH+1] */, int lane) {
    typedef TpqCfg<CH, NB> C;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        ld[i] = INFINITY;
        li[i] = -1;
    }
    const int bs = M.bshift;
    const float ds = M.ds;
    const int bc = 1 << bs;
    const float B = (float)bc * ds;
    const float lim = (float)(LI_CELL_LIMIT - 16 * bc) * ds;
    const bool act = valid && isfinite(qx) && isfinite(qy) && isfinite(qz) && fabsf(qx) < lim && fabsf(qy) < lim && fabsf(qz) < lim;
    if (!act) {
        qx = 0.f; qy = 0.f; qz = 0.f;
    }
    // slack for float cell assignment / edge products: relative 2^-23 effects, bounded generously
    const float margin = 1e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + 16.0f * B);
    float4* my_row = tile + lane * C::ROW;

    bool done = !act;
    float lo2 = 0.f, hi2 = rho2;
    while (__any_sync(LI_FULL, !done)) {
        const bool need = !done;
        const bool last = hi2 >= 5.0f;   // the radius bound d2 <= 5 is inclusive (ikd_Tree.cpp:842)
        const float r = sqrtf(fminf(hi2, 5.0f)) * (1.0f + 1e-6f) + margin;
        const int lx = li_cell(qx - r, ds) >> bs, hx = li_cell(qx + r, ds) >> bs;
        const int ly = li_cell(qy - r, ds) >> bs, hy = li_cell(qy + r, ds) >> bs;
        const int lz = li_cell(qz - r, ds) >> bs, hz = li_cell(qz + r, ds) >> bs;
        const int nx = hx - lx + 1, ny = hy - ly + 1, nz = hz - lz + 1;
        const int nxy = nx * ny;
        const int total = need ? nxy * nz : 0;
        const float inv_nxy = 1.0f / (float)nxy, inv_nx = 1.0f / (float)nx;

        for (int base = 0; __any_sync(LI_FULL, base < total); base += NB) {
            // ---- probe up to NB bricks of THIS lane's enumeration (independent hash lookups) -------------
            unsigned bf[NB], bcnt[NB];
#pragma unroll
            for (int u = 0; u < NB; u++) {
                bf[u] = 0u;
                bcnt[u] = 0u;
                const int idx = base + u;
                if (idx < total) {
                    const int iz = (int)(((float)idx + 0.5f) * inv_nxy);
                    const int rem = idx - iz * nxy;
                    const int iy = (int)(((float)rem + 0.5f) * inv_nx);
                    const int ix = rem - iy * nx;
                    const int kx = lx + ix, ky = ly + iy, kz = lz + iz;
                    float lox = (float)(kx << bs) * ds - margin, hix = (float)((kx + 1) << bs) * ds + margin;
                    float loy = (float)(ky << bs) * ds - margin, hiy = (float)((ky + 1) << bs) * ds + margin;
                    float loz = (float)(kz << bs) * ds - margin, hiz = (float)((kz + 1) << bs) * ds + margin;
                    float ex = fmaxf(0.f, fmaxf(lox - qx, qx - hix));
                    float ey = fmaxf(0.f, fmaxf(loy - qy, qy - hiy));
                    float ez = fmaxf(0.f, fmaxf(loz - qz, qz - hiz));
                    float dbox = (ex * ex + ey * ey + ez * ez) * (1.0f - 1e-6f);
                    const bool in_shell = dbox >= lo2 && (last ? dbox <= 5.0f : dbox < hi2);
                    if (in_shell && dbox < ld[4]) {   // ld[4] = +inf while fewer than 5 are known
                        unsigned f = 0u, c = 0u;
                        if (li_brick_find(M.ent, M.mask, li_pack_key(kx, ky, kz), f, c)) {
                            bf[u] = f;
                            bcnt[u] = c;
                        }
                    }
                }
            }
            // ---- stage + scan, one brick slot per round, CH points per chunk ------------------------------
#pragma unroll
            for (int u = 0; u < NB; u++) {
                const unsigned maxc = __reduce_max_sync(LI_FULL, bcnt[u]);
                for (unsigned off = 0; off < maxc; off += CH) {   // warp-uniform
                    // stage: slab chunk of lane l -> tile row l (lanes copy 16 B each, coalesced per slab)
                    unsigned have = __ballot_sync(LI_FULL, bcnt[u] > off);
                    while (have) {
                        const int l = __ffs(have) - 1;
                        have &= have - 1u;
                        const unsigned fl = __shfl_sync(LI_FULL, bf[u], l);
                        const unsigned cl = __shfl_sync(LI_FULL, bcnt[u], l);
                        const unsigned n = min(cl - off, (unsigned)CH);
#pragma unroll
                        for (int t = 0; t < (CH + 31) / 32; t++) {
                            const unsigned j = lane + 32u * t;
                            if (j < n) li_cp_async16(tile + l * C::ROW + j, M.pool + (size_t)fl + off + j);
                        }
                    }
                    li_cp_async_wait_all();
                    __syncwarp();
                    // scan my own row
                    const unsigned n_me = (bcnt[u] > off) ? min(bcnt[u] - off, (unsigned)CH) : 0u;
                    const unsigned n_max = min(maxc - off, (unsigned)CH);
                    const int id0 = (int)(bf[u] + off);
#pragma unroll 4
                    for (unsigned j = 0; j < n_max; j++) {   // warp-uniform trip count
                        if (j < n_me) {
                            float4 p = my_row[j];
                            float d = li_dist2(qx, qy, qz, p.x, p.y, p.z);
                            if (d <= 5.0f && d < ld[4]) local_insert(ld, li, d, id0 + (int)j);
                        }
                    }
                    __syncwarp();   // rows are rewritten by the next staging round
                }
            }
        }
        if (need) {
            const bool full = li[4] >= 0;
            if (last || (full && ld[4] <= hi2)) {
                done = true;
            } else {
                lo2 = hi2;
                hi2 = full ? fminf(ld[4] * (1.0f + 1e-6f), 5.0f) : fminf(4.0f * hi2, 5.0f);
            }
        }
    }
}

// ---- search kernel of an ICP pass: world transform + 5-NN for every scan point ---------------------------
template <int CH, int NB>
__global__ void __launch_bounds__(128) k_knn_scan_tpq(MapDev M, ScanDev S, PoseD P, float rho2) {
    extern __shared__ float4 li_tiles[];
    typedef TpqCfg<CH, NB> C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* tile = li_tiles + warp * C::WARP_TILE_F4;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int qb = warp_global * 32; qb < S.n; qb += nwarps * 32) {   // warp-uniform
        const int q = qb + lane;
        const bool valid = q < S.n;
        float wx = 0.f, wy = 0.f, wz = 0.f;
        if (valid) {
            float4 b = __ldg(&S.body[q]);
            li_body_to_world(P, b.x, b.y, b.z, wx, wy, wz);
        }
        float ld[5];
        int li[5];
        knn5_tpq<CH, NB>(M, rho2, valid, wx, wy, wz, ld, li, tile, lane);
        if (valid) {
            S.world[q] = make_float4(wx, wy, wz, 0.f);
#pragma unroll
            for (int k = 0; k < 5; k++) S.near_ids[(size_t)q * 5 + k] = li[k];
        }
    }
}

// ---- stand-alone Nearest_Search for arbitrary world-frame queries ------------------------------------------
template <int CH, int NB>
__global__ void __launch_bounds__(128) k_knn_queries_tpq(MapDev M, const float4* __restrict__ qpts, int n, int* __restrict__ ids,
                                                         float* __restrict__ d2, float rho2) {
    extern __shared__ float4 li_tiles[];
    typedef TpqCfg<CH, NB> C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* tile = li_tiles + warp * C::WARP_TILE_F4;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int qb = warp_global * 32; qb < n; qb += nwarps * 32) {
        const int q = qb + lane;
        const bool valid = q < n;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) p = __ldg(&qpts[q]);
        float ld[5];
        int li[5];
        knn5_tpq<CH, NB>(M, rho2, valid, p.x, p.y, p.z, ld, li, tile, lane);
        if (valid) {
#pragma unroll
            for (int k = 0; k < 5; k++) {
                ids[(size_t)q * 5 + k] = li[k];
                d2[(size_t)q * 5 + k] = (li[k] >= 0) ? ld[k] : -1.f;
            }
        }
    }
}
