// common.cuh -- device-side types and helpers shared by the map and ICP kernels.
//
// Map layout in HBM:
//   * brick hash: open-addressing table of 16-byte entries {key(8) | first(4) | count(4)};
//     a brick is a cube of (1<<bshift)^3 map voxels of edge ds = filter_size_map, i.e. the
//     voxel grid of KD_TREE::Add_Points (ikd_Tree.cpp:389-394) grouped 4x4x4 by default.
//   * aux[slot] = {cap, pending, fill, spare}: update-time bookkeeping, never read by searches.
//   * pool: float4 points; every brick owns one contiguous, 128-byte aligned slab
//     [first, first+cap) of which [first, first+count) are live. w carries the voxel-in-brick id.
#pragma once

// Invariants that only the CPU build of the library checks (tests/emul: every kernel compiled for the host): no instruction on the device.
#ifdef LI_SIMT_EMUL
#include <cstdio>
#include <cstdlib>
#define LI_EMUL_ASSERT(cond)                                                              \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            fprintf(stderr, "LI_EMUL_ASSERT failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
            abort();                                                                      \
        }                                                                                 \
    } while (0)
#else
#define LI_EMUL_ASSERT(cond) ((void)0)
#endif
#ifdef __CUDACC__
#include <cuda_runtime.h>
#elif defined(LI_SIMT_EMUL)   // tests/emul/simt_shim.h brings the vector types and constructors (and no CUDA API header)
#include <vector_types.h>
#else   // host-only parse (tests/emul: the cell-directory search is compiled for the CPU as its own checker)
#include <vector_types.h>
#include <vector_functions.h>
#endif
#include <stdint.h>

#define LI_FULL 0xffffffffu
#define LI_EMPTY_KEY 0xffffffffffffffffull
#define LI_CELL_LIMIT (1 << 20)   // |cell index| must stay below this (21-bit packing)

struct MapDev {
    uint4* ent;                    // hash entries
    uint4* aux;                    // {cap, pending, fill, spare}
    unsigned mask;                 // slots - 1
    float4* pool;
    unsigned long long pool_cap;   // points
    unsigned long long* pool_top;  // bump allocator (points)
    float ds;                      // filter_size_map as float (KD_TREE::downsample_size)
    int bshift;                    // log2(voxels per brick edge)
    int* touched_list;             // hash slots touched by the current update batch
    int* counters;                 // [0]=touched_n [1]=err flags [2]=n_live [3]=n_bricks [4]=changed voxels [5]=dropped pts
    int* brick_slots;              // hash slot of every brick ever created, in creation order [counters[CNT_BRICKS]]: whole-map kernels
                                   // (box delete, flatten, directory refresh) walk the bricks, not the (mostly empty) table
    // cell directory (cells.cuh; only with knn_index = cells, else null): per hash slot the occupancy mask of the brick's
    // 4x4x4 cells (cell = 2x2x2 voxels) and the start offset of every cell inside the slab, which is kept sorted by cell
    unsigned long long* cocc;
    unsigned short* cdir;          // [slots * 64]; cdir[slot*64] == 0xffff: brick not indexed (scan the whole slab)
    // super-brick table (cells.cuh): open addressing over 4x4x4-brick blocks; sb_occ bit b = brick b of the block was
    // created at some time (bits are never cleared: a stale bit only costs one brick probe)
    unsigned long long* sb_keys;
    unsigned long long* sb_occ;
    unsigned sb_mask;              // slots - 1
};

enum { CNT_TOUCHED = 0, CNT_ERR = 1, CNT_LIVE = 2, CNT_BRICKS = 3, CNT_CHANGED = 4, CNT_DROPPED = 5, CNT_NADD = 6, CNT_NNOD = 7, CNT_COUPLED = 8, CNT_COUNT = 16 };
enum { ERR_HASH_FULL = 1, ERR_POOL_FULL = 2, ERR_RANGE = 4 };

// Pose part of StatesGroup (common_lib.h:160-163), row-major doubles.
struct PoseD {
    double R[9];     // rot_end
    double p[3];     // pos_end
    double RLI[9];   // offset_R_L_I
    double TLI[3];   // offset_T_L_I
};

// Per-scan device state: the globals of laserMapping.cpp:102-125 the path touches.
struct ScanDev {
    float4* body;           // feats_down_body (xyz, w unused)
    float4* world;          // feats_down_world
    int* near_ids;          // [N*5] pool offsets, -1 = missing: hand-over from the id-writing search kernels to the plane pass of the same search pass
    float4* near_xyz;       // [N*5] Nearest_Points (laserMapping.cpp:107) as COPIES of the map points (xyz; w = 1 found, 0 missing rank)
    unsigned char* selected;  // point_selected_surf
    float4* normvec;        // (nx,ny,nz,pd2) f32
    int n;
};

// LI_HD: helpers that are also compiled for the host by the CPU checker of the cell-directory search (tests/emul)
#ifdef __CUDACC__
#define LI_HD __host__ __device__ __forceinline__
#else
#define LI_HD inline
#endif

LI_HD unsigned long long li_pack_key(int x, int y, int z) {
    return ((unsigned long long)(unsigned)(x + LI_CELL_LIMIT) << 42) | ((unsigned long long)(unsigned)(y + LI_CELL_LIMIT) << 21) |
           (unsigned long long)(unsigned)(z + LI_CELL_LIMIT);
}

// Spatial hash of a packed key: three 32-bit multiplies (Teschner et al. primes) + a final avalanche step.
LI_HD unsigned li_hash(unsigned long long k) {
    unsigned x = (unsigned)(k >> 42), y = (unsigned)(k >> 21) & 0x1fffffu, z = (unsigned)k & 0x1fffffu;
    unsigned h = (x * 73856093u) ^ (y * 19349663u) ^ (z * 83492791u);
    h ^= h >> 15;
    return h;
}

#if defined(__CUDACC__) || defined(LI_SIMT_EMUL)   // LI_SIMT_EMUL: tests/emul/simt_shim.h runs the warp-cooperative kernels on the CPU, 32 fibers per warp

// Voxel index exactly as the reference computes it: floor(x / downsample_size) in float
// (ikd_Tree.cpp:389). IEEE division, no reciprocal shortcut.
__device__ __forceinline__ int li_cell(float x, float ds) { return (int)floorf(__fdiv_rn(x, ds)); }

// fp32 squared distance with the reference's association and no FMA contraction
// (KD_TREE::calc_dist, ikd_Tree.cpp:1273-1277): (dx*dx + dy*dy) + dz*dz.
__device__ __forceinline__ float li_dist2(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Same value, two issue slots fewer: the x and y lanes go through Blackwell's packed fp32 pipe (sub.f32x2 / mul.f32x2 ->
// FADD2 / FMUL2 on sm_100a; each half is an IEEE round-to-nearest operation, so the result is bit-identical to li_dist2).
// qxy = {qx, qy} packed by li_pack_f32x2; p.x, p.y arrive adjacent from the 16-byte slab load, so the packing is free.
#ifdef LI_SIMT_EMUL   // CPU emulation of the kernels: the same values without PTX
inline unsigned long long li_pack_f32x2(float lo, float hi) {
    unsigned a, b;
    memcpy(&a, &lo, 4);
    memcpy(&b, &hi, 4);
    return (unsigned long long)a | ((unsigned long long)b << 32);
}
inline float li_dist2_packed(unsigned long long qxy, float qz, const float4& p) {
    const unsigned a = (unsigned)qxy, b = (unsigned)(qxy >> 32);
    float qx, qy;
    memcpy(&qx, &a, 4);
    memcpy(&qy, &b, 4);
    return li_dist2(qx, qy, qz, p.x, p.y, p.z);
}
#else
__device__ __forceinline__ unsigned long long li_pack_f32x2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float li_dist2_packed(unsigned long long qxy, float qz, const float4& p) {
    unsigned long long pxy = li_pack_f32x2(p.x, p.y), dxy, sq;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dxy) : "l"(qxy), "l"(pxy));
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(sq) : "l"(dxy));
    float sx, sy;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(sx), "=f"(sy) : "l"(sq));
    float dz = __fsub_rn(qz, p.z);
    return __fadd_rn(__fadd_rn(sx, sy), __fmul_rn(dz, dz));
}
#endif

// Lookup only (searches). Returns true and (first,count) when the brick exists.
__device__ __forceinline__ bool li_brick_find(const uint4* __restrict__ ent, unsigned mask, unsigned long long key,
                                              unsigned& first, unsigned& count) {
    unsigned h = li_hash(key) & mask;
    for (unsigned i = 0; i <= mask; i++) {
        uint4 e = __ldg(&ent[h]);
        unsigned long long k = (unsigned long long)e.x | ((unsigned long long)e.y << 32);
        if (k == key) {
            first = e.z;
            count = e.w;
            return true;
        }
        if (k == LI_EMPTY_KEY) return false;
        h = (h + 1) & mask;
    }
    return false;
}

// Find-or-create (updates). Returns the slot or -1 when the table is full.
__device__ __forceinline__ int li_brick_find_or_insert(uint4* ent, unsigned mask, unsigned long long key, bool* created) {
    unsigned h = li_hash(key) & mask;
    for (unsigned i = 0; i <= mask; i++) {
        unsigned long long* kp = reinterpret_cast<unsigned long long*>(&ent[h]);
        unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(kp);
        if (k == key) return (int)h;
        if (k == LI_EMPTY_KEY) {
            unsigned long long old = atomicCAS(kp, LI_EMPTY_KEY, key);
            if (old == LI_EMPTY_KEY) {
                if (created) *created = true;
                return (int)h;
            }
            if (old == key) return (int)h;
        }
        h = (h + 1) & mask;
    }
    return -1;
}

// Sort a singly linked list of point indices (head, next[] = -1 terminated) ASCENDING by index, in place: bottom-up merge sort on the
// links, O(k log k) steps and no extra storage. The per-leaf / per-box lists of the voxel grid and of Add_Points(downsample) are pushed
// by atomicExch in arrival order; their points must be visited in INPUT order (sequential float sums, sequential box replay). The first
// versions found "the next smallest index" by walking the whole list for every element: k^2 steps in one thread, seconds for the
// thousands of points a near-sensor leaf of a 2M-point raw scan can hold (advisor finding, round 1).
// true when the list holds more than `limit` nodes (walks at most limit + 1 of them)
__device__ __forceinline__ bool li_list_longer_than(int head, const int* __restrict__ next, int limit) {
    int k = 0;
    for (int t = head; t >= 0; t = next[t])
        if (++k > limit) return true;
    return false;
}
#define LI_LIST_SELECT_MAX 8   // up to this many nodes "next smallest index" by repeated walks is cheaper than sorting the links (no stores)

__device__ __forceinline__ int li_list_sort_ascending(int head, int* __restrict__ next) {
    if (head < 0 || next[head] < 0) return head;
    for (int width = 1;; width <<= 1) {
        int cur = head, tail = -1, new_head = -1, merges = 0;
        while (cur >= 0) {
            merges++;
            int a = cur, asz = 0, b = cur;
            for (int i = 0; i < width && b >= 0; i++) {   // b = first node of the second run
                b = next[b];
                asz++;
            }
            int bsz = width;
            while (asz > 0 || (bsz > 0 && b >= 0)) {
                int pick;
                if (asz == 0) { pick = b; b = next[b]; bsz--; }
                else if (bsz == 0 || b < 0 || a < b) { pick = a; a = next[a]; asz--; }
                else { pick = b; b = next[b]; bsz--; }
                if (tail >= 0) next[tail] = pick; else new_head = pick;
                tail = pick;
            }
            cur = b;
        }
        next[tail] = -1;
        head = new_head;
        if (merges <= 1) return head;
    }
}

// pointBodyToWorld (laserMapping.cpp:209-220): double math without contraction, float store.
__device__ __forceinline__ void li_body_to_world(const PoseD& P, float bx, float by, float bz, float& wx, float& wy, float& wz) {
    double x = (double)bx, y = (double)by, z = (double)bz;
    double t0 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.RLI[0], x), __dmul_rn(P.RLI[1], y)), __dmul_rn(P.RLI[2], z)), P.TLI[0]);
    double t1 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.RLI[3], x), __dmul_rn(P.RLI[4], y)), __dmul_rn(P.RLI[5], z)), P.TLI[1]);
    double t2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.RLI[6], x), __dmul_rn(P.RLI[7], y)), __dmul_rn(P.RLI[8], z)), P.TLI[2]);
    double g0 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.R[0], t0), __dmul_rn(P.R[1], t1)), __dmul_rn(P.R[2], t2)), P.p[0]);
    double g1 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.R[3], t0), __dmul_rn(P.R[4], t1)), __dmul_rn(P.R[5], t2)), P.p[1]);
    double g2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.R[6], t0), __dmul_rn(P.R[7], t1)), __dmul_rn(P.R[8], t2)), P.p[2]);
    wx = (float)g0;
    wy = (float)g1;
    wz = (float)g2;
}
#endif  // __CUDACC__
