// cells.cuh -- exact bounded 5-NN on a CELL DIRECTORY inside the brick hash, one scan point per thread
// (second implementation of KD_TREE::Nearest_Search, ikd_Tree.cpp:349-379 / Search :825-968; selected with
// liinit_config.knn_index = LIINIT_KNN_CELLS).
//
// Why: the lockstep brick search (knn_kernels.cuh) is instruction-issue bound because a brick (8x8x8 map voxels, 1.2 m at
// ds = 0.15) is the smallest unit it can prune: a query whose five neighbours lie within 0.2 m still evaluates every
// point of every brick its ball touches (~180 candidates per query at C2). Here every brick carries a directory of its
// 4x4x4 CELLS (cell = 2x2x2 voxels, 0.3 m):
//     cocc[slot]        64-bit occupancy mask, bit c = cell c holds at least one point, c = (cx << 4) | (cy << 2) | cz
//     cdir[slot*64 + c] offset of cell c's first point inside the slab (u16); cell c ends where cell c+1 starts
// and the slab is kept SORTED BY CELL (li_cells_refresh_brick, run on the touched bricks after every map update).
// A query enumerates the bricks of its ball's bounding box as before, but inside a brick it visits only the occupied
// cells of that box whose own box distance passes the shell / 5th-best tests, and scans just their points.
//
// The search is the same exact shell iteration as knn5_lockstep, stated on cells:
//     invariant  every cell with dcell < lo2 has been scanned (or was pruned against a 5th-best that only shrinks)
//     step       scan the cells with lo2 <= dcell < hi2 (<= 5 on the final radius) and dcell < current 5th best
//     stop       5 known and d5 <= hi2, or hi2 >= 5 (the reference's radius: squared distance <= 5, ikd_Tree.cpp:842)
//     next       lo2 = hi2;  hi2 = d5 if 5 are known (one closing step)  else 4 * hi2
// One thread owns one query: a sorted top-5 in registers, no shuffles, no merges, no lockstep.
//
// Everything marked LI_HD is plain C++ (no intrinsics; the library is built with --fmad=false, IEEE division and square
// root, so `a * b + c` is the same unfused arithmetic on both sides) and is ALSO compiled for the host by tests/emul,
// where the identical source is checked against brute force without a GPU.
#pragma once
#include <math.h>

#include "common.cuh"

#define LI_CELLS_BSHIFT 3            // the directory is defined for 8x8x8-voxel bricks only
#define LI_CDIR_UNINDEXED 0xffffu    // cdir[slot*64] of a brick too large for u16 offsets: scanned as one slab
#define LI_NO_BOX_W 0xfffffffeu      // == LI_NO_BOX (map_kernels.cuh): the point lies in no downsample box

#ifdef __CUDA_ARCH__
#define LC_LDG(p) __ldg(p)
#else
#define LC_LDG(p) (*(p))
#endif

LI_HD unsigned lc_f2u(float f) {
    union { float f; unsigned u; } c;
    c.f = f;
    return c.u;
}
LI_HD float lc_u2f(unsigned u) {
    union { float f; unsigned u; } c;
    c.u = u;
    return c.f;
}
LI_HD int lc_imax(int a, int b) { return a > b ? a : b; }
LI_HD int lc_imin(int a, int b) { return a < b ? a : b; }
LI_HD int lc_ctz64(unsigned long long m) {
#ifdef __CUDA_ARCH__
    return __ffsll((long long)m) - 1;
#else
    return __builtin_ctzll(m);
#endif
}

// Cell of a stored map point. Bricks and voxel-in-brick ids come from the point's BOX index (li_storage), so the cell
// does too; a point in no box (ulp gap between two float boxes) was filed under its division index.
LI_HD unsigned lc_cell_of(float4 p, float ds) {
    const unsigned w = lc_f2u(p.w);
    unsigned vx, vy, vz;
    if (w == LI_NO_BOX_W) {
        vx = (unsigned)((int)floorf(p.x / ds) & 7);
        vy = (unsigned)((int)floorf(p.y / ds) & 7);
        vz = (unsigned)((int)floorf(p.z / ds) & 7);
    } else {
        vx = (w >> 6) & 7u;
        vy = (w >> 3) & 7u;
        vz = w & 7u;
    }
    return ((vx >> 1) << 4) | ((vy >> 1) << 2) | (vz >> 1);
}

// (Re)build the directory of one brick: count per cell, prefix sums, in-place permutation of the slab (American-flag
// pass: every swap puts one point into its final cell range, so at most `count` swaps). One thread per brick: the
// slabs of a surface map hold tens of points, and the refresh runs once per map update, not per ICP pass.
LI_HD void li_cells_refresh_brick(const MapDev& M, unsigned slot) {
    const uint4 e = M.ent[slot];
    const unsigned first = e.z, n = e.w;
    unsigned short* dir = M.cdir + (size_t)slot * 64;
    if (n > 0xfff0u) {
        M.cocc[slot] = ~0ull;
        dir[0] = LI_CDIR_UNINDEXED;
        return;
    }
    unsigned short nxt[64], end[64];
    for (int c = 0; c < 64; c++) nxt[c] = 0;
    float4* slab = M.pool + first;
    for (unsigned j = 0; j < n; j++) nxt[lc_cell_of(slab[j], M.ds)]++;
    unsigned long long occ = 0ull;
    unsigned acc = 0;
    for (int c = 0; c < 64; c++) {
        const unsigned k = nxt[c];
        if (k) occ |= 1ull << c;
        dir[c] = (unsigned short)acc;
        nxt[c] = (unsigned short)acc;
        acc += k;
        end[c] = (unsigned short)acc;
    }
    M.cocc[slot] = occ;
    for (int c = 0; c < 64; c++) {
        while (nxt[c] < end[c]) {
            const float4 p = slab[nxt[c]];
            const unsigned k = lc_cell_of(p, M.ds);
            if (k == (unsigned)c) {
                nxt[c]++;
            } else {
                const unsigned dst = nxt[k]++;
                const float4 q = slab[dst];
                slab[dst] = p;
                slab[nxt[c]] = q;
            }
        }
    }
}

// sorted insert into the thread's top-5 (precondition d < ld[4]); equal distances keep their arrival order
LI_HD void lc_insert(float (&ld)[5], int (&li)[5], float d, int id) {
    ld[4] = d;
    li[4] = id;
#pragma unroll
    for (int i = 4; i > 0; --i) {
        const bool sw = ld[i] < ld[i - 1];
        const float a = ld[i - 1], b = ld[i];
        const int ia = li[i - 1], ib = li[i];
        ld[i - 1] = sw ? b : a;
        ld[i] = sw ? a : b;
        li[i - 1] = sw ? ib : ia;
        li[i] = sw ? ia : ib;
    }
}

// lookup returning the hash slot (-1 = no such brick)
LI_HD int lc_brick_find(const uint4* ent, unsigned mask, unsigned long long key, unsigned& first, unsigned& count) {
    unsigned h = li_hash(key) & mask;
    for (unsigned i = 0; i <= mask; i++) {
        const uint4 e = LC_LDG(&ent[h]);
        const unsigned long long k = (unsigned long long)e.x | ((unsigned long long)e.y << 32);
        if (k == key) {
            first = e.z;
            count = e.w;
            return (int)h;
        }
        if (k == LI_EMPTY_KEY) return -1;
        h = (h + 1) & mask;
    }
    return -1;
}

// occupancy-mask bits of the cells with coordinate in [a, b] (0 <= a <= b <= 3) on one axis
LI_HD unsigned long long lc_mask_x(int a, int b) { return (~0ull << (16 * a)) & (~0ull >> (48 - 16 * b)); }
LI_HD unsigned long long lc_mask_y(int a, int b) {
    const unsigned long long g = ((0xffffull << (4 * a)) & (0xffffull >> (12 - 4 * b))) & 0xffffull;
    return g * 0x0001000100010001ull;
}
LI_HD unsigned long long lc_mask_z(int a, int b) {
    const unsigned long long g = ((0xfull << a) & (0xfull >> (3 - b))) & 0xfull;
    return g * 0x1111111111111111ull;
}

// occupancy mask of a super-brick (block of 4x4x4 bricks; bit b = (bx&3) << 4 | (by&3) << 2 | (bz&3)); 0 = no such block
LI_HD unsigned long long lc_sb_find(const MapDev& M, unsigned long long key) {
    unsigned h = li_hash(key) & M.sb_mask;
    for (unsigned i = 0; i <= M.sb_mask; i++) {
        const unsigned long long k = LC_LDG(&M.sb_keys[h]);
        if (k == key) return LC_LDG(&M.sb_occ[h]);
        if (k == LI_EMPTY_KEY) return 0ull;
        h = (h + 1) & M.sb_mask;
    }
    return 0ull;
}

// Work counters of one query (host checker / cost model only; the device instantiation passes nullptr and COUNT = false).
struct LcStats {
    int rounds, supers, probes, found, cells, cells_scanned, points, inserts;
    int* tr;        // optional structure trace (knn5_boxes): -1 round, -2 super-brick, -3 brick + status, -4 cell + points + insert bits
    int ntr, cap;
};
#define LC_TR(v)                                                    \
    do {                                                            \
        if (COUNT && st->tr && st->ntr < st->cap) st->tr[st->ntr++] = (v); \
    } while (0)

LI_HD float lc_box_d2(float qx, float qy, float qz, float lox, float hix, float loy, float hiy, float loz, float hiz) {
    const float ex = fmaxf(0.f, fmaxf(lox - qx, qx - hix));
    const float ey = fmaxf(0.f, fmaxf(loy - qy, qy - hiy));
    const float ez = fmaxf(0.f, fmaxf(loz - qz, qz - hiz));
    return (ex * ex + ey * ey + ez * ez) * (1.0f - 1e-6f);
}

// candidate test of the inner loops. KD_TREE::calc_dist (ikd_Tree.cpp:1273-1277): (dx*dx + dy*dy) + dz*dz in f32, no FMA
#define LC_CANDIDATE(J)                                                  \
    do {                                                                 \
        const float4 p_ = LC_LDG(&pool[first + (J)]);                    \
        const float dx_ = qx - p_.x, dy_ = qy - p_.y, dz_ = qz - p_.z;   \
        const float d_ = (dx_ * dx_ + dy_ * dy_) + dz_ * dz_;            \
        if (COUNT) st->points++;                                         \
        if (d_ < tau) {                                                  \
            lc_insert(ld, li, d_, (int)(first + (J)));                   \
            tau = fminf(ld[4], cap5);                                    \
            if (COUNT) st->inserts++;                                    \
        }                                                                \
    } while (0)

// Exact 5-NN of one query. ld / li: ascending squared distances / pool offsets (-1 = missing).
// Three levels of 4x4x4 occupancy masks: super-brick (4.8 m at ds = 0.15) -> brick (1.2 m) -> cell (0.3 m); only
// bricks that exist are probed and only occupied cells inside the ball's bounding box are tested.
template <bool COUNT>
LI_HD void knn5_cells(const MapDev& M, float rho2, float qx, float qy, float qz, float (&ld)[5], int (&li)[5], LcStats* st) {
#pragma unroll
    for (int i = 0; i < 5; i++) {
        ld[i] = INFINITY;
        li[i] = -1;
    }
    const float ds = M.ds;
    const float cs = 2.0f * ds;   // cell edge (exact)
    const float B = 8.0f * ds;    // brick edge (exact)
    const float lim = (float)(LI_CELL_LIMIT - 16 * 8) * ds;
    if (!(fabsf(qx) < lim && fabsf(qy) < lim && fabsf(qz) < lim)) return;   // also rejects NaN / inf
    // Rounding slack, as in knn5_lockstep: box indices differ from real geometry by <~ |x| 2^-23; every pruning box is
    // inflated by `margin`, squared bounds carry a (1 - 1e-6) factor, the enumeration range a slack in voxels.
    const float l1 = fabsf(qx) + fabsf(qy) + fabsf(qz);
    const float margin = 1e-6f * (l1 + 16.0f * B);
    const float inv_ds = 1.0f / ds;
    const float slk = 0.02f + 4e-7f * l1 * inv_ds;
    const float cap5 = lc_u2f(0x40a00001u);   // smallest float above 5: d <= 5 <=> d < cap5
    const float4* __restrict__ pool = M.pool;
    float lo2 = 0.f, hi2 = rho2;
    for (;;) {
        if (COUNT) st->rounds++;
        const bool last = hi2 >= 5.0f;   // the radius bound d2 <= 5 is inclusive
        const float r = sqrtf(fminf(hi2, 5.0f)) * (1.0f + 1e-6f) + margin;
        // voxel range of the ball's bounding box (conservative); cells = voxels >> 1, bricks = voxels >> 3, super-bricks = voxels >> 5
        const int lvx = (int)floorf((qx - r) * inv_ds - slk), hvx = (int)floorf((qx + r) * inv_ds + slk);
        const int lvy = (int)floorf((qy - r) * inv_ds - slk), hvy = (int)floorf((qy + r) * inv_ds + slk);
        const int lvz = (int)floorf((qz - r) * inv_ds - slk), hvz = (int)floorf((qz + r) * inv_ds + slk);
        float tau = fminf(ld[4], cap5);
        for (int sz = lvz >> 5; sz <= (hvz >> 5); sz++) {
            const unsigned long long mz = lc_mask_z(lc_imax((lvz >> 3) - 4 * sz, 0), lc_imin((hvz >> 3) - 4 * sz, 3));
            for (int sy = lvy >> 5; sy <= (hvy >> 5); sy++) {
                const unsigned long long myz = mz & lc_mask_y(lc_imax((lvy >> 3) - 4 * sy, 0), lc_imin((hvy >> 3) - 4 * sy, 3));
                for (int sx = lvx >> 5; sx <= (hvx >> 5); sx++) {
                    if (COUNT) st->supers++;
                    unsigned long long bm = lc_sb_find(M, li_pack_key(sx, sy, sz)) & myz &
                                            lc_mask_x(lc_imax((lvx >> 3) - 4 * sx, 0), lc_imin((hvx >> 3) - 4 * sx, 3));
                    while (bm) {
                        const int b = lc_ctz64(bm);
                        bm &= bm - 1ull;
                        const int kx = 4 * sx + (b >> 4), ky = 4 * sy + ((b >> 2) & 3), kz = 4 * sz + (b & 3);
                        const float db = lc_box_d2(qx, qy, qz, (float)kx * B - margin, (float)(kx + 1) * B + margin, (float)ky * B - margin,
                                                   (float)(ky + 1) * B + margin, (float)kz * B - margin, (float)(kz + 1) * B + margin);
                        // a brick's box distance bounds its cells' from below: necessary conditions only (no lo2 test here)
                        if (!((last ? db <= 5.0f : db < hi2) && db < ld[4])) continue;
                        if (COUNT) st->probes++;
                        unsigned first = 0, count = 0;
                        const int slot = lc_brick_find(M.ent, M.mask, li_pack_key(kx, ky, kz), first, count);
                        if (slot < 0 || count == 0u) continue;
                        if (COUNT) st->found++;
                        const unsigned short* __restrict__ dir = M.cdir + (size_t)slot * 64;
                        if (LC_LDG(&dir[0]) == LI_CDIR_UNINDEXED) {
                            // oversized brick: one unit under the brick-level shell rule
                            if (db >= lo2)
                                for (unsigned j = 0; j < count; j++) LC_CANDIDATE(j);
                            continue;
                        }
                        unsigned long long m = LC_LDG(&M.cocc[slot]) &
                                               lc_mask_x(lc_imax((lvx >> 1) - 4 * kx, 0), lc_imin((hvx >> 1) - 4 * kx, 3)) &
                                               lc_mask_y(lc_imax((lvy >> 1) - 4 * ky, 0), lc_imin((hvy >> 1) - 4 * ky, 3)) &
                                               lc_mask_z(lc_imax((lvz >> 1) - 4 * kz, 0), lc_imin((hvz >> 1) - 4 * kz, 3));
                        while (m) {
                            const int c = lc_ctz64(m);
                            m &= m - 1ull;
                            if (COUNT) st->cells++;
                            const int gx = 4 * kx + (c >> 4), gy = 4 * ky + ((c >> 2) & 3), gz = 4 * kz + (c & 3);
                            const float dc = lc_box_d2(qx, qy, qz, (float)gx * cs - margin, (float)(gx + 1) * cs + margin, (float)gy * cs - margin,
                                                       (float)(gy + 1) * cs + margin, (float)gz * cs - margin, (float)(gz + 1) * cs + margin);
                            if (!(dc >= lo2 && (last ? dc <= 5.0f : dc < hi2) && dc < ld[4])) continue;
                            if (COUNT) st->cells_scanned++;
                            const unsigned s0 = LC_LDG(&dir[c]);
                            const unsigned e0 = (c == 63) ? count : (unsigned)LC_LDG(&dir[c + 1]);
                            for (unsigned j = s0; j < e0; j++) LC_CANDIDATE(j);
                        }
                    }
                }
            }
        }
        const bool full = li[4] >= 0;
        if (last || (full && ld[4] <= hi2)) break;
        lo2 = hi2;
        hi2 = full ? fminf(ld[4] * (1.0f + 1e-6f), 5.0f) : fminf(4.0f * hi2, 5.0f);
    }
}

// ---- variant 2: growing BOXES instead of shells -----------------------------------------------------------------
// The per-cell box distance of knn5_cells costs about as many instructions as scanning the ~3 points a cell of a
// downsampled surface map holds, so here a cell is never tested: round k scans EVERY occupied cell of the voxel box
// around the ball of radius r_k that was not already inside the box of round k-1 (the previous box is masked out, so no
// point is seen twice), and only the always-safe bound "unit farther than the current 5th best" prunes bricks and
// large cells. After round k every map point within r_k of the query has been examined, hence
//     stop   5 known and d5 <= r_k^2, or r_k^2 >= 5;     next   r^2 = d5 if 5 are known (one closing round) else 4 r^2.
// bits of the cells (or bricks) of block k whose coordinate on one axis lies in [lo, hi] (global index); 0 if none
LI_HD unsigned long long lc_range_x(int lo, int hi, int k) {
    const int a = lc_imax(lo - 4 * k, 0), b = lc_imin(hi - 4 * k, 3);
    return a <= b ? lc_mask_x(a, b) : 0ull;
}
LI_HD unsigned long long lc_range_y(int lo, int hi, int k) {
    const int a = lc_imax(lo - 4 * k, 0), b = lc_imin(hi - 4 * k, 3);
    return a <= b ? lc_mask_y(a, b) : 0ull;
}
LI_HD unsigned long long lc_range_z(int lo, int hi, int k) {
    const int a = lc_imax(lo - 4 * k, 0), b = lc_imin(hi - 4 * k, 3);
    return a <= b ? lc_mask_z(a, b) : 0ull;
}

#ifndef LI_CELLS_BIG
#define LI_CELLS_BIG 16   // cells with more points than this are tested against the 5th best before they are scanned
#endif

template <bool COUNT>
LI_HD void knn5_boxes(const MapDev& M, float rho2, float qx, float qy, float qz, float (&ld)[5], int (&li)[5], LcStats* st) {
#pragma unroll
    for (int i = 0; i < 5; i++) {
        ld[i] = INFINITY;
        li[i] = -1;
    }
    const float ds = M.ds;
    const float cs = 2.0f * ds;   // cell edge (exact)
    const float B = 8.0f * ds;    // brick edge (exact)
    const float lim = (float)(LI_CELL_LIMIT - 16 * 8) * ds;
    if (!(fabsf(qx) < lim && fabsf(qy) < lim && fabsf(qz) < lim)) return;   // also rejects NaN / inf
    const float l1 = fabsf(qx) + fabsf(qy) + fabsf(qz);
    const float margin = 1e-6f * (l1 + 16.0f * B);
    const float inv_ds = 1.0f / ds;
    const float slk = 0.02f + 4e-7f * l1 * inv_ds;
    const float cap5 = lc_u2f(0x40a00001u);   // smallest float above 5: d <= 5 <=> d < cap5
    const float4* __restrict__ pool = M.pool;
    float hi2 = rho2;
    // cell box of the previous round (empty before the first)
    int pcx0 = 1, pcx1 = 0, pcy0 = 1, pcy1 = 0, pcz0 = 1, pcz1 = 0;
    for (;;) {
        if (COUNT) st->rounds++;
        LC_TR(-1);
        const bool last = hi2 >= 5.0f;
        const float r = sqrtf(fminf(hi2, 5.0f)) * (1.0f + 1e-6f) + margin;
        // voxel box around the ball (conservative: every map point within sqrt(hi2) of the query is stored under a voxel
        // index inside it); cells = voxels >> 1, bricks = voxels >> 3, super-bricks = voxels >> 5
        const int lvx = (int)floorf((qx - r) * inv_ds - slk), hvx = (int)floorf((qx + r) * inv_ds + slk);
        const int lvy = (int)floorf((qy - r) * inv_ds - slk), hvy = (int)floorf((qy + r) * inv_ds + slk);
        const int lvz = (int)floorf((qz - r) * inv_ds - slk), hvz = (int)floorf((qz + r) * inv_ds + slk);
        // the box never shrinks: keep it a superset of the previous one, so "previous box" = "everything scanned so far"
        const int cx0 = lc_imin(lvx >> 1, pcx0 <= pcx1 ? pcx0 : (lvx >> 1)), cx1 = lc_imax(hvx >> 1, pcx0 <= pcx1 ? pcx1 : (hvx >> 1));
        const int cy0 = lc_imin(lvy >> 1, pcy0 <= pcy1 ? pcy0 : (lvy >> 1)), cy1 = lc_imax(hvy >> 1, pcy0 <= pcy1 ? pcy1 : (hvy >> 1));
        const int cz0 = lc_imin(lvz >> 1, pcz0 <= pcz1 ? pcz0 : (lvz >> 1)), cz1 = lc_imax(hvz >> 1, pcz0 <= pcz1 ? pcz1 : (hvz >> 1));
        float tau = fminf(ld[4], cap5);
        for (int sz = cz0 >> 4; sz <= (cz1 >> 4); sz++) {
            const unsigned long long mz = lc_range_z(cz0 >> 2, cz1 >> 2, sz);
            for (int sy = cy0 >> 4; sy <= (cy1 >> 4); sy++) {
                const unsigned long long myz = mz & lc_range_y(cy0 >> 2, cy1 >> 2, sy);
                for (int sx = cx0 >> 4; sx <= (cx1 >> 4); sx++) {
                    if (COUNT) st->supers++;
                    LC_TR(-2);
                    unsigned long long bm = lc_sb_find(M, li_pack_key(sx, sy, sz)) & myz & lc_range_x(cx0 >> 2, cx1 >> 2, sx);
                    while (bm) {
                        const int b = lc_ctz64(bm);
                        bm &= bm - 1ull;
                        const int kx = 4 * sx + (b >> 4), ky = 4 * sy + ((b >> 2) & 3), kz = 4 * sz + (b & 3);
                        // cells of this brick inside the current box and not inside the previous one
                        const unsigned long long want = lc_range_x(cx0, cx1, kx) & lc_range_y(cy0, cy1, ky) & lc_range_z(cz0, cz1, kz);
                        const unsigned long long seen = lc_range_x(pcx0, pcx1, kx) & lc_range_y(pcy0, pcy1, ky) & lc_range_z(pcz0, pcz1, kz);
                        LC_TR(-3);
                        int* trs = (COUNT && st->tr && st->ntr < st->cap) ? &st->tr[st->ntr++] : nullptr;   // brick status
                        if (trs) *trs = 0;
                        if ((want & ~seen) == 0ull) continue;
                        const float db = lc_box_d2(qx, qy, qz, (float)kx * B - margin, (float)(kx + 1) * B + margin, (float)ky * B - margin,
                                                   (float)(ky + 1) * B + margin, (float)kz * B - margin, (float)(kz + 1) * B + margin);
                        if (!(db < tau)) continue;   // nothing in this brick can enter the top 5 (safe at any time: tau only shrinks)
                        if (COUNT) st->probes++;
                        if (trs) *trs = 1;
                        unsigned first = 0, count = 0;
                        const int slot = lc_brick_find(M.ent, M.mask, li_pack_key(kx, ky, kz), first, count);
                        if (slot < 0 || count == 0u) continue;
                        if (COUNT) st->found++;
                        if (trs) *trs = 2;
                        const unsigned short* __restrict__ dir = M.cdir + (size_t)slot * 64;
                        if (LC_LDG(&dir[0]) == LI_CDIR_UNINDEXED) {
                            // oversized brick (no directory): one unit; scanned in the first round whose box touches it
                            if (seen == 0ull)
                                for (unsigned j = 0; j < count; j++) LC_CANDIDATE(j);
                            continue;
                        }
                        unsigned long long m = LC_LDG(&M.cocc[slot]) & want & ~seen;
                        while (m) {
                            const int c = lc_ctz64(m);
                            m &= m - 1ull;
                            if (COUNT) st->cells++;
                            const unsigned s0 = LC_LDG(&dir[c]);
                            const unsigned e0 = (c == 63) ? count : (unsigned)LC_LDG(&dir[c + 1]);
                            if (e0 - s0 > (unsigned)LI_CELLS_BIG) {
                                const int gx = 4 * kx + (c >> 4), gy = 4 * ky + ((c >> 2) & 3), gz = 4 * kz + (c & 3);
                                const float dc = lc_box_d2(qx, qy, qz, (float)gx * cs - margin, (float)(gx + 1) * cs + margin, (float)gy * cs - margin,
                                                           (float)(gy + 1) * cs + margin, (float)gz * cs - margin, (float)(gz + 1) * cs + margin);
                                if (!(dc < tau)) continue;
                            }
                            if (COUNT) st->cells_scanned++;
                            LC_TR(-4);
                            LC_TR((int)(e0 - s0));
                            const int ins0 = COUNT ? st->inserts : 0;
                            int* trb = (COUNT && st->tr && st->ntr < st->cap) ? &st->tr[st->ntr++] : nullptr;   // insert bits
                            if (trb) *trb = 0;
                            for (unsigned j = s0; j < e0; j++) {
                                const int before = COUNT ? st->inserts : 0;
                                LC_CANDIDATE(j);
                                if (trb && j - s0 < 31 && st->inserts != before) *trb |= 1 << (j - s0);
                            }
                            (void)ins0;
                        }
                    }
                }
            }
        }
        const bool full = li[4] >= 0;
        if (last || (full && ld[4] <= hi2)) break;
        pcx0 = cx0; pcx1 = cx1; pcy0 = cy0; pcy1 = cy1; pcz0 = cz0; pcz1 = cz1;
        hi2 = full ? fminf(ld[4] * (1.0f + 1e-6f), 5.0f) : fminf(4.0f * hi2, 5.0f);
    }
}

// ---- variant 3: growing boxes, ENUMERATE then STREAM ----------------------------------------------------------------
// Measured on B200 (profiles/r01_cells): the free-running nested loops of knn5_cells / knn5_boxes keep 5 of 32 lanes
// busy -- a warp pays, at every nesting level, for the longest trip count among its lanes (tools/cells_cost_model.py
// reproduces the instruction count from the CPU traces). Here the work of a round is split so that the long loop is FLAT:
//   A  enumerate: per lane a small state machine walks super-bricks -> bricks (probe, masks) and pushes the point range
//      (start, count) of every cell to scan into the lane's queue (shared memory on the device);
//   B  stream: all lanes walk their queues, ONE candidate per lane and iteration -- the trip count is the maximum over
//      the lanes of their TOTAL candidate count, not a sum of per-cell maxima.
// Every loop that contains a vote is warp-uniform (LC_ANY = __any_sync over the full warp on the device, the lane's own
// predicate on the host, where one lane runs alone); a full queue anywhere makes the whole warp drain. The set of
// candidates a lane evaluates, and their order, does not depend on the other lanes, so results are identical on both
// sides. Search logic = knn5_boxes.
#if defined(__CUDA_ARCH__) || defined(LI_SIMT_EMUL)   // (the SIMT shim of tests/emul votes over its 32 fibers)
#define LC_ANY(p) __any_sync(0xffffffffu, (p))
#else
#define LC_ANY(p) (p)
#endif

#ifndef LI_CELLS_QC
#define LI_CELLS_QC 32   // point ranges a lane can queue before the warp drains
#endif

struct LcQ {            // this lane's queue: entry i at [i * stride]
    unsigned* rstart;          // absolute pool offset of the range
    unsigned short* rcount;    // <= 0xfff0 points (cells of indexed bricks only)
    int stride;
};

#define LC_CANDIDATE_ABS(J)                                              \
    do {                                                                 \
        const float4 p_ = LC_LDG(&pool[(J)]);                            \
        const float dx_ = qx - p_.x, dy_ = qy - p_.y, dz_ = qz - p_.z;   \
        const float d_ = (dx_ * dx_ + dy_ * dy_) + dz_ * dz_;            \
        if (COUNT) st->points++;                                         \
        if (d_ < tau) {                                                  \
            lc_insert(ld, li, d_, (int)(J));                             \
            tau = fminf(ld[4], cap5);                                    \
            if (COUNT) st->inserts++;                                    \
        }                                                                \
    } while (0)

// phase B: evaluate the nr queued ranges of every lane, LI_CELLS_U candidates per lane and iteration: the loads of a batch are
// issued back to back before any distance is evaluated (measured with one load per iteration: issue slots 33 % busy, every
// iteration waits for its own L2 round trip). The order in which a lane sees its candidates is unchanged.
#ifndef LI_CELLS_U
#define LI_CELLS_U 4
#endif
template <bool COUNT>
LI_HD void lc_drain(const float4* __restrict__ pool, const LcQ& Q, int nr, float qx, float qy, float qz, float cap5, float& tau, float (&ld)[5],
                    int (&li)[5], LcStats* st) {
    int ri = 0;
    unsigned j = 0, e = 0;
    while (LC_ANY(ri < nr || j < e)) {
        float4 p[LI_CELLS_U];
        unsigned idx[LI_CELLS_U];
        bool ok[LI_CELLS_U];
#pragma unroll
        for (int u = 0; u < LI_CELLS_U; u++) {
            if (j >= e && ri < nr) {
                j = Q.rstart[ri * Q.stride];
                e = j + Q.rcount[ri * Q.stride];
                ri++;
            }
            ok[u] = j < e;
            idx[u] = j;
            if (ok[u]) {
                p[u] = LC_LDG(&pool[j]);
                j++;
            }
        }
#pragma unroll
        for (int u = 0; u < LI_CELLS_U; u++) {
            if (ok[u]) {
                const float dx_ = qx - p[u].x, dy_ = qy - p[u].y, dz_ = qz - p[u].z;
                const float d_ = (dx_ * dx_ + dy_ * dy_) + dz_ * dz_;   // KD_TREE::calc_dist association, no FMA
                if (COUNT) st->points++;
                if (d_ < tau) {
                    lc_insert(ld, li, d_, (int)idx[u]);
                    tau = fminf(ld[4], cap5);
                    if (COUNT) st->inserts++;
                }
            }
        }
    }
}

// ONE_ROUND: stop after the first box whatever it found (first stage of the hybrid search): returns true when the result is
// final (five neighbours within the box radius, or an invalid query), false when a wider search has to follow.
template <bool COUNT, bool ONE_ROUND = false>
LI_HD bool knn5_stream(const MapDev& M, float rho2, bool valid, float qx, float qy, float qz, float (&ld)[5], int (&li)[5], LcStats* st, const LcQ& Q) {
#pragma unroll
    for (int i = 0; i < 5; i++) {
        ld[i] = INFINITY;
        li[i] = -1;
    }
    const float ds = M.ds;
    const float cs = 2.0f * ds;   // cell edge (exact)
    const float B = 8.0f * ds;    // brick edge (exact)
    const float lim = (float)(LI_CELL_LIMIT - 16 * 8) * ds;
    bool done = !(valid && fabsf(qx) < lim && fabsf(qy) < lim && fabsf(qz) < lim);   // also rejects NaN / inf
    if (done) {
        qx = 0.f; qy = 0.f; qz = 0.f;
    }
    const float l1 = fabsf(qx) + fabsf(qy) + fabsf(qz);
    const float margin = 1e-6f * (l1 + 16.0f * B);
    const float inv_ds = 1.0f / ds;
    const float slk = 0.02f + 4e-7f * l1 * inv_ds;
    const float cap5 = lc_u2f(0x40a00001u);   // smallest float above 5: d <= 5 <=> d < cap5
    const float4* __restrict__ pool = M.pool;
    float hi2 = rho2;
    float tau = cap5;
    int pcx0 = 1, pcx1 = 0, pcy0 = 1, pcy1 = 0, pcz0 = 1, pcz1 = 0;   // cell box of the previous round (empty before the first)
    while (LC_ANY(!done)) {
        const bool need = !done;
        if (COUNT && need) st->rounds++;
        const bool last = hi2 >= 5.0f;
        const float r = sqrtf(fminf(hi2, 5.0f)) * (1.0f + 1e-6f) + margin;
        const int lvx = (int)floorf((qx - r) * inv_ds - slk), hvx = (int)floorf((qx + r) * inv_ds + slk);
        const int lvy = (int)floorf((qy - r) * inv_ds - slk), hvy = (int)floorf((qy + r) * inv_ds + slk);
        const int lvz = (int)floorf((qz - r) * inv_ds - slk), hvz = (int)floorf((qz + r) * inv_ds + slk);
        const bool hadp = pcx0 <= pcx1;
        const int cx0 = hadp ? lc_imin(lvx >> 1, pcx0) : (lvx >> 1), cx1 = hadp ? lc_imax(hvx >> 1, pcx1) : (hvx >> 1);
        const int cy0 = hadp ? lc_imin(lvy >> 1, pcy0) : (lvy >> 1), cy1 = hadp ? lc_imax(hvy >> 1, pcy1) : (hvy >> 1);
        const int cz0 = hadp ? lc_imin(lvz >> 1, pcz0) : (lvz >> 1), cz1 = hadp ? lc_imax(hvz >> 1, pcz1) : (hvz >> 1);
        // super-brick lattice of the box, walked with three counters (no division)
        const int sx0 = cx0 >> 4, sx1 = cx1 >> 4, sy0 = cy0 >> 4, sy1 = cy1 >> 4, sz1 = cz1 >> 4;
        int sx = sx0, sy = sy0, sz = cz0 >> 4;
        bool more = need;              // supers left to visit or bricks left in bm
        unsigned long long bm = 0ull;  // bricks of the current super-brick still to visit
        int bsx = 0, bsy = 0, bsz = 0; // ... and its coordinates
        int nr = 0;                    // ranges in this lane's queue
        // current brick of this lane: wanted cells still to queue (m), slab, directory
        unsigned long long m = 0ull;
        unsigned first = 0, count = 0;
        int kx = 0, ky = 0, kz = 0;
        const unsigned short* __restrict__ dir = M.cdir;
        for (;;) {
            const bool anymore = LC_ANY(more || m != 0ull);
            // ---- phase A1: a lane whose brick is used up takes its next one
            if (m == 0ull && more) {
                // short divergent helper loop over super-bricks without a wanted brick
                while (bm == 0ull && sz <= sz1) {
                    if (COUNT) st->supers++;
                    bm = lc_sb_find(M, li_pack_key(sx, sy, sz)) & lc_range_x(cx0 >> 2, cx1 >> 2, sx) & lc_range_y(cy0 >> 2, cy1 >> 2, sy) &
                         lc_range_z(cz0 >> 2, cz1 >> 2, sz);
                    bsx = sx; bsy = sy; bsz = sz;
                    if (++sx > sx1) {
                        sx = sx0;
                        if (++sy > sy1) {
                            sy = sy0;
                            ++sz;
                        }
                    }
                }
                if (bm == 0ull) {
                    more = false;
                } else {
                    const int b = lc_ctz64(bm);
                    bm &= bm - 1ull;
                    kx = 4 * bsx + (b >> 4); ky = 4 * bsy + ((b >> 2) & 3); kz = 4 * bsz + (b & 3);
                    // cells of this brick inside the current box and not inside the previous one
                    const unsigned long long want = lc_range_x(cx0, cx1, kx) & lc_range_y(cy0, cy1, ky) & lc_range_z(cz0, cz1, kz);
                    const unsigned long long seen = hadp ? (lc_range_x(pcx0, pcx1, kx) & lc_range_y(pcy0, pcy1, ky) & lc_range_z(pcz0, pcz1, kz)) : 0ull;
                    if ((want & ~seen) != 0ull) {
                        const float db = lc_box_d2(qx, qy, qz, (float)kx * B - margin, (float)(kx + 1) * B + margin, (float)ky * B - margin,
                                                   (float)(ky + 1) * B + margin, (float)kz * B - margin, (float)(kz + 1) * B + margin);
                        if (db < tau) {   // else nothing in this brick can enter the top 5 (safe at any time: tau only shrinks)
                            if (COUNT) st->probes++;
                            const int slot = lc_brick_find(M.ent, M.mask, li_pack_key(kx, ky, kz), first, count);
                            if (slot >= 0 && count > 0u) {
                                if (COUNT) st->found++;
                                dir = M.cdir + (size_t)slot * 64;
                                if (LC_LDG(&dir[0]) == LI_CDIR_UNINDEXED) {
                                    // no directory (oversized brick): one unit, scanned when the box first touches it
                                    if (seen == 0ull)
                                        for (unsigned j = 0; j < count; j++) LC_CANDIDATE_ABS(first + j);
                                } else {
                                    m = LC_LDG(&M.cocc[slot]) & want & ~seen;
                                }
                            }
                        }
                    }
                }
            }
            // ---- phase A2: queue the point ranges of the wanted cells until some lane's queue is full
            while (LC_ANY(m != 0ull) && !LC_ANY(nr >= LI_CELLS_QC)) {
                if (m != 0ull) {
                    const int c = lc_ctz64(m);
                    m &= m - 1ull;
                    if (COUNT) st->cells++;
                    const unsigned s0 = LC_LDG(&dir[c]);
                    const unsigned e0 = (c == 63) ? count : (unsigned)LC_LDG(&dir[c + 1]);
                    bool take = true;
                    if (e0 - s0 > (unsigned)LI_CELLS_BIG) {   // large cell: worth a test against the 5th best
                        const int gx = 4 * kx + (c >> 4), gy = 4 * ky + ((c >> 2) & 3), gz = 4 * kz + (c & 3);
                        const float dc = lc_box_d2(qx, qy, qz, (float)gx * cs - margin, (float)(gx + 1) * cs + margin, (float)gy * cs - margin,
                                                   (float)(gy + 1) * cs + margin, (float)gz * cs - margin, (float)(gz + 1) * cs + margin);
                        take = dc < tau;
                    }
                    if (take) {
                        if (COUNT) st->cells_scanned++;
                        Q.rstart[nr * Q.stride] = first + s0;
                        Q.rcount[nr * Q.stride] = (unsigned short)(e0 - s0);
                        nr++;
                    }
                }
            }
            // ---- phase B (the only call site): a full queue somewhere, or nothing left to enumerate
            if (!anymore || LC_ANY(nr >= LI_CELLS_QC)) {
                lc_drain<COUNT>(pool, Q, nr, qx, qy, qz, cap5, tau, ld, li, st);
                nr = 0;
            }
            if (!anymore) break;
        }
        if (need) {
            const bool full = li[4] >= 0;
            if (last || (full && ld[4] <= hi2)) {
                done = true;
            } else {
                pcx0 = cx0; pcx1 = cx1; pcy0 = cy0; pcy1 = cy1; pcz0 = cz0; pcz1 = cz1;
                hi2 = full ? fminf(ld[4] * (1.0f + 1e-6f), 5.0f) : fminf(4.0f * hi2, 5.0f);
            }
        }
        if (ONE_ROUND) break;   // (the loop condition above was a vote: every lane leaves together)
    }
    return done;
}

// the three searches: 1 = shells on cells (knn5_cells), 2 = growing boxes (knn5_boxes), 3 = growing boxes, enumerate + stream
#ifndef LI_CELLS_SEARCH_DEFAULT
#define LI_CELLS_SEARCH_DEFAULT 3
#endif

#if defined(__CUDACC__) || defined(LI_SIMT_EMUL)
// ---- super-brick table: called by the insert kernels when they CREATE a brick (map_kernels.cuh) -------------
__device__ __forceinline__ void li_sb_mark(const MapDev& M, unsigned long long brick_key) {
    if (!M.sb_keys) return;
    const int kx = (int)(unsigned)(brick_key >> 42) - LI_CELL_LIMIT, ky = (int)((unsigned)(brick_key >> 21) & 0x1fffffu) - LI_CELL_LIMIT,
              kz = (int)((unsigned)brick_key & 0x1fffffu) - LI_CELL_LIMIT;
    const unsigned long long key = li_pack_key(kx >> 2, ky >> 2, kz >> 2);
    const unsigned long long bit = 1ull << (((kx & 3) << 4) | ((ky & 3) << 2) | (kz & 3));
    unsigned h = li_hash(key) & M.sb_mask;
    for (unsigned i = 0; i <= M.sb_mask; i++) {
        unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(&M.sb_keys[h]);
        if (k == LI_EMPTY_KEY) {
            k = atomicCAS(&M.sb_keys[h], LI_EMPTY_KEY, key);
            if (k == LI_EMPTY_KEY) k = key;
        }
        if (k == key) {
            atomicOr(&M.sb_occ[h], bit);
            return;
        }
        h = (h + 1) & M.sb_mask;
    }
    atomicOr(&M.counters[CNT_ERR], ERR_HASH_FULL);
}
__global__ void k_sb_clear(MapDev M) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > M.sb_mask) return;
    M.sb_keys[i] = LI_EMPTY_KEY;
    M.sb_occ[i] = 0ull;
}

// ---- directory maintenance ---------------------------------------------------------------------------
// after k_ins_commit / k_ds_compact: the bricks of the batch's touched list (count on the device: fixed grid)
__global__ void k_cells_refresh_touched(MapDev M) {
    const int nt = gridDim.x * blockDim.x, ntouched = M.counters[CNT_TOUCHED];
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ntouched; t += nt) li_cells_refresh_brick(M, (unsigned)M.touched_list[t]);
}
// Same job, one WARP per brick (the thread-per-brick version above costs ~0.6 ms per map update at C2: three dependent passes over
// a slab per thread). Slabs of up to 32 * LI_REFRESH_K points are held in registers: coalesced loads, a shared-memory
// histogram over the 64 cells, prefix sums by shuffle, then every lane drops its points at cursor[cell]++ -- all loads
// happen before the first store (__syncwarp), so the permutation is in place. Larger slabs fall back to lane 0 running
// li_cells_refresh_brick. The order of the points INSIDE a cell is whatever the atomics give (as for appends).
#ifndef LI_REFRESH_K
#define LI_REFRESH_K 8
#endif
__device__ __forceinline__ void li_cells_refresh_brick_warp(const MapDev& M, unsigned slot, int* cnt /* [64] */, int* cur /* [64] */, int lane) {
    const uint4 e = M.ent[slot];
    const unsigned first = e.z, n = e.w;
    if (n > 32u * LI_REFRESH_K) {
        if (lane == 0) li_cells_refresh_brick(M, slot);
        __syncwarp();
        return;
    }
    float4* slab = M.pool + first;
    float4 p[LI_REFRESH_K];
    unsigned key[LI_REFRESH_K];
    cnt[lane] = 0;
    cnt[lane + 32] = 0;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < LI_REFRESH_K; k++) {
        const unsigned j = k * 32 + lane;
        key[k] = 0xffffffffu;
        if (j < n) {
            p[k] = slab[j];
            key[k] = lc_cell_of(p[k], M.ds);
            atomicAdd(&cnt[key[k]], 1);
        }
    }
    __syncwarp();
    // lane l owns cells 2l and 2l+1
    const int c0 = cnt[2 * lane], c1 = cnt[2 * lane + 1];
    int incl = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int s0 = incl - (c0 + c1), s1 = s0 + c0;
    reinterpret_cast<ushort2*>(M.cdir + (size_t)slot * 64)[lane] = make_ushort2((unsigned short)s0, (unsigned short)s1);
    const unsigned b0 = __ballot_sync(0xffffffffu, c0 > 0), b1 = __ballot_sync(0xffffffffu, c1 > 0);
    if (lane == 0) {
        unsigned long long occ = 0ull;   // bit 2l from b0, bit 2l+1 from b1
#pragma unroll
        for (int l = 0; l < 32; l++) occ |= ((unsigned long long)((b0 >> l) & 1u) << (2 * l)) | ((unsigned long long)((b1 >> l) & 1u) << (2 * l + 1));
        M.cocc[slot] = occ;
    }
    cur[2 * lane] = s0;
    cur[2 * lane + 1] = s1;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < LI_REFRESH_K; k++)
        if (key[k] != 0xffffffffu) slab[atomicAdd(&cur[key[k]], 1)] = p[k];
    __syncwarp();
}
__global__ void __launch_bounds__(128) k_cells_refresh_touched_warp(MapDev M) {
    __shared__ int s_cnt[4][64], s_cur[4][64];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5, ntouched = M.counters[CNT_TOUCHED];
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < ntouched; t += nw)
        li_cells_refresh_brick_warp(M, (unsigned)M.touched_list[t], s_cnt[wib], s_cur[wib], lane);
}
__global__ void __launch_bounds__(128) k_cells_refresh_all_warp(MapDev M, unsigned slots) {
    __shared__ int s_cnt[4][64], s_cur[4][64];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned nw = (gridDim.x * blockDim.x) >> 5;
    for (unsigned i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < slots; i += nw) {
        const uint4 e = M.ent[i];
        if (((unsigned long long)e.x | ((unsigned long long)e.y << 32)) == LI_EMPTY_KEY) continue;   // warp-uniform
        li_cells_refresh_brick_warp(M, i, s_cnt[wib], s_cur[wib], lane);
    }
}

// after a box delete (any slab may have been squeezed)
__global__ void k_cells_refresh_all(MapDev M, unsigned slots) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= slots) return;
    const uint4 e = M.ent[i];
    if (((unsigned long long)e.x | ((unsigned long long)e.y << 32)) == LI_EMPTY_KEY) return;
    li_cells_refresh_brick(M, i);
}

// ---- search kernel of an ICP pass: world transform + 5-NN, one scan point per thread --------------------
// HOST = true: the scan is still the caller's page-locked host buffer (liinit_scan_attach_host); consecutive lanes read
// consecutive points, so for packed xyz a warp pulls one contiguous 384-byte run over PCIe.
// SEARCH: 1 / 2 / 3 as above. MINB: resident blocks per SM the kernel is compiled for (register budget).
#ifndef LI_CELLS_THREADS
#define LI_CELLS_THREADS 128
#endif
template <int SEARCH>
__device__ __forceinline__ void li_cells_search(const MapDev& M, float rho2, bool valid, float x, float y, float z, float (&ld)[5], int (&li)[5]) {
    if constexpr (SEARCH == 3) {
        __shared__ unsigned s_start[LI_CELLS_QC * LI_CELLS_THREADS];
        __shared__ unsigned short s_count[LI_CELLS_QC * LI_CELLS_THREADS];
        LcQ Q;
        Q.rstart = s_start + threadIdx.x;   // entry i of this lane at [i * blockDim.x]: conflict-free whatever i each lane is at
        Q.rcount = s_count + threadIdx.x;
        Q.stride = LI_CELLS_THREADS;
        knn5_stream<false>(M, rho2, valid, x, y, z, ld, li, nullptr, Q);
    } else {
        if (!valid) {
#pragma unroll
            for (int i = 0; i < 5; i++) {
                ld[i] = INFINITY;
                li[i] = -1;
            }
            return;
        }
        if constexpr (SEARCH == 2) knn5_boxes<false>(M, rho2, x, y, z, ld, li, nullptr);
        else knn5_cells<false>(M, rho2, x, y, z, ld, li, nullptr);
    }
}

template <bool HOST, int MINB, int SEARCH>
__global__ void __launch_bounds__(LI_CELLS_THREADS, MINB) k_knn_cells_scan(MapDev M, ScanDev S, PoseD P, float rho2, const float* __restrict__ raw, int stride) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = q < S.n;   // no early exit: the stream search votes over the full warp
    float bx = 0.f, by = 0.f, bz = 0.f, wx = 0.f, wy = 0.f, wz = 0.f;
    if (valid) {
        if (HOST) {
            const float* s = raw + (size_t)q * stride;
            bx = s[0]; by = s[1]; bz = s[2];
            S.body[q] = make_float4(bx, by, bz, 0.f);
        } else {
            const float4 b = __ldg(&S.body[q]);
            bx = b.x; by = b.y; bz = b.z;
        }
        li_body_to_world(P, bx, by, bz, wx, wy, wz);
    }
    float ld[5];
    int li[5];
    li_cells_search<SEARCH>(M, rho2, valid, wx, wy, wz, ld, li);
    if (valid) {
        S.world[q] = make_float4(wx, wy, wz, 0.f);
#pragma unroll
        for (int k = 0; k < 5; k++) S.near_ids[(size_t)q * 5 + k] = li[k];
    }
}

// Same pass with DYNAMIC scheduling: a persistent grid (MINB blocks per SM), every warp pulls batches of 32 scan points from a ticket
// counter until the scan is used up. Why: with one block per 128 points a block lives ~300 k cycles (dependent L2 round trips, 24
// warps per SM), so the last wave of blocks runs on a nearly empty GPU -- SMs were active 61 % of the kernel in the ncu capture of the
// static version (profiles/r01_cells/ncu_full_stream_final_metrics.txt). With warp-granular batches the tail shrinks to one batch.
// Checked on the CPU (tests/test_liinit_emul.py); NOT yet measured on a GPU (round-1 GPU budget was spent): opt-in via
// LIINIT_CELLS_SCHED=dynamic until it has been.
template <bool HOST, int MINB, int SEARCH>
__global__ void __launch_bounds__(LI_CELLS_THREADS, MINB) k_knn_cells_scan_dyn(MapDev M, ScanDev S, PoseD P, float rho2, const float* __restrict__ raw, int stride,
                                                                                unsigned* __restrict__ ticket) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(ticket, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= (unsigned)S.n) break;   // warp-uniform
        const int q = (int)base + lane;
        const bool valid = q < S.n;
        float bx = 0.f, by = 0.f, bz = 0.f, wx = 0.f, wy = 0.f, wz = 0.f;
        if (valid) {
            if (HOST) {
                const float* s = raw + (size_t)q * stride;
                bx = s[0]; by = s[1]; bz = s[2];
                S.body[q] = make_float4(bx, by, bz, 0.f);
            } else {
                const float4 b = __ldg(&S.body[q]);
                bx = b.x; by = b.y; bz = b.z;
            }
            li_body_to_world(P, bx, by, bz, wx, wy, wz);
        }
        float ld[5];
        int li[5];
        li_cells_search<SEARCH>(M, rho2, valid, wx, wy, wz, ld, li);
        if (valid) {
            S.world[q] = make_float4(wx, wy, wz, 0.f);
#pragma unroll
            for (int k = 0; k < 5; k++) S.near_ids[(size_t)q * 5 + k] = li[k];
        }
    }
}

// ---- stand-alone Nearest_Search for arbitrary world-frame queries -----------------------------------------
template <int SEARCH>
__global__ void __launch_bounds__(LI_CELLS_THREADS) k_knn_cells_queries(MapDev M, const float4* __restrict__ qpts, int n, int* __restrict__ ids,
                                                                        float* __restrict__ d2, float rho2) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = q < n;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) p = __ldg(&qpts[q]);
    float ld[5];
    int li[5];
    li_cells_search<SEARCH>(M, rho2, valid, p.x, p.y, p.z, ld, li);
    if (valid) {
#pragma unroll
        for (int k = 0; k < 5; k++) {
            ids[(size_t)q * 5 + k] = li[k];
            d2[(size_t)q * 5 + k] = (li[k] >= 0) ? ld[k] : -1.f;
        }
    }
}
#endif  // __CUDACC__
