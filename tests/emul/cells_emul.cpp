// cells_emul.cpp -- CPU checker of the cell-directory 5-NN (lidar_imu_init_b200/csrc/cells.cuh).
//
// TEST INFRASTRUCTURE, not product: it compiles the SAME LI_HD source the sm_100a kernels are built from
// (li_cells_refresh_brick, knn5_cells) for the host, over a brick hash laid out in host memory by the storage rule of
// map_kernels.cuh (li_storage: brick and voxel id from the float BOX index, ikd_Tree.cpp:633,980), so that the search
// logic (shell iteration on cells, range masks, directory offsets, rounding margins) can be checked against brute force
// without a GPU. Nothing under lidar_imu_init_b200/ loads this file; the product has no CPU path.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -I/usr/local/cuda/include cells_emul.cpp -o libcells_emul.so
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "../../lidar_imu_init_b200/csrc/cells.cuh"

#include "emul_map.h"

extern "C" {

void* emul_create(const float* xyz, int n, float ds, int hash_log2) {
    Emul* E = new Emul();
    E->ds = ds;
    E->hash_log2 = hash_log2;
    for (int i = 0; i < n; i++) {
        float4 p = make_float4(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], 0.f);
        unsigned long long key;
        storage(ds, p, key);
        E->bricks[key].push_back(p);
    }
    layout(E, true);
    return E;
}

// plain append (slab order: sorted old content, then the new points in batch order, as k_ins_append leaves it) + refresh
void emul_add(void* h, const float* xyz, int n) {
    Emul* E = (Emul*)h;
    for (int i = 0; i < n; i++) {
        float4 p = make_float4(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], 0.f);
        unsigned long long key;
        storage(E->ds, p, key);
        E->bricks[key].push_back(p);
    }
    layout(E, true);
}

void emul_destroy(void* h) { delete (Emul*)h; }

int emul_num_bricks(void* h) { return (int)((Emul*)h)->bricks.size(); }

// directory invariants of every brick: offsets monotone, cell c = [dir[c], dir[c+1]), every point in the cell the
// directory says, occupancy mask == non-empty cells. Returns the number of violations.
int emul_check_directory(void* h) {
    Emul* E = (Emul*)h;
    int bad = 0;
    for (size_t s = 0; s <= E->M.mask; s++) {
        uint4 e = E->ent[s];
        if (e.x == 0xffffffffu && e.y == 0xffffffffu) continue;
        const unsigned short* dir = &E->cdir[s * 64];
        if (dir[0] == LI_CDIR_UNINDEXED) {
            if (e.w <= 0xfff0u) bad++;
            continue;
        }
        if (dir[0] != 0) bad++;
        unsigned long long occ = 0;
        for (int c = 0; c < 64; c++) {
            unsigned s0 = dir[c], e0 = (c == 63) ? e.w : dir[c + 1];
            if (e0 < s0 || e0 > e.w) { bad++; continue; }
            if (e0 > s0) occ |= 1ull << c;
            for (unsigned j = s0; j < e0; j++)
                if (lc_cell_of(E->pool[e.z + j], E->ds) != (unsigned)c) bad++;
        }
        if (occ != E->cocc[s]) bad++;
    }
    return bad;
}

// 5-NN of n queries (packed xyz). variant: 0 = the kernels' default, 1 = shells on cells, 2 = growing boxes, 3 = growing boxes (enumerate + stream).
// out_xyz [n*15], out_d2 [n*5] (-1 = missing), out_cnt [n], stats [n*8] or NULL.
void emul_knn(void* h, const float* q, int n, float rho2, int variant, float* out_xyz, float* out_d2, int* out_cnt, int* stats) {
    Emul* E = (Emul*)h;
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; i++) {
        float ld[5];
        int li[5];
        LcStats st;
        std::memset(&st, 0, sizeof(st));
        const int v = variant ? variant : LI_CELLS_SEARCH_DEFAULT;   // 0 = what the kernels run by default
        if (v == 3) {
            unsigned rs[LI_CELLS_QC];
            unsigned short rc[LI_CELLS_QC];   // one lane alone: its queue is a local array
            LcQ Q;
            Q.rstart = rs; Q.rcount = rc; Q.stride = 1;
            knn5_stream<true>(E->M, rho2, true, q[3 * (size_t)i], q[3 * (size_t)i + 1], q[3 * (size_t)i + 2], ld, li, &st, Q);
        } else if (v == 2) {
            knn5_boxes<true>(E->M, rho2, q[3 * (size_t)i], q[3 * (size_t)i + 1], q[3 * (size_t)i + 2], ld, li, &st);
        } else {
            knn5_cells<true>(E->M, rho2, q[3 * (size_t)i], q[3 * (size_t)i + 1], q[3 * (size_t)i + 2], ld, li, &st);
        }
        int cnt = 0;
        for (int k = 0; k < 5; k++) {
            if (li[k] >= 0) {
                cnt++;
                const float4 p = E->pool[li[k]];
                out_xyz[15 * (size_t)i + 3 * k] = p.x;
                out_xyz[15 * (size_t)i + 3 * k + 1] = p.y;
                out_xyz[15 * (size_t)i + 3 * k + 2] = p.z;
                out_d2[5 * (size_t)i + k] = ld[k];
            } else {
                out_xyz[15 * (size_t)i + 3 * k] = out_xyz[15 * (size_t)i + 3 * k + 1] = out_xyz[15 * (size_t)i + 3 * k + 2] = 0.f;
                out_d2[5 * (size_t)i + k] = -1.f;
            }
        }
        out_cnt[i] = cnt;
        if (stats) std::memcpy(stats + 8 * (size_t)i, &st, 8 * sizeof(int));
    }
}

// structure traces of knn5_boxes for the lockstep cost model (tools/cells_cost_model.py): per query a slice of `tr`
// ([off[i], off[i+1])), tokens as documented at LcStats.
long long emul_trace(void* h, const float* q, int n, float rho2, int* tr, long long cap, long long* off) {
    Emul* E = (Emul*)h;
    long long pos = 0;
    for (int i = 0; i < n; i++) {
        float ld[5];
        int li[5];
        LcStats st;
        std::memset(&st, 0, sizeof(st));
        st.tr = tr + pos;
        st.cap = (int)std::min<long long>(cap - pos, 1 << 20);
        knn5_boxes<true>(E->M, rho2, q[3 * (size_t)i], q[3 * (size_t)i + 1], q[3 * (size_t)i + 2], ld, li, &st);
        off[i] = pos;
        pos += st.ntr;
    }
    off[n] = pos;
    return pos;
}

}  // extern "C"
