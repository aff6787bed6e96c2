// oracle/ikd_ref_wrap.cpp -- TEST INFRASTRUCTURE ONLY (never on the product path).
//
// Thin C-ABI around the reference's *verbatim* ikd-Tree, which is compiled
// from where it lies under /root/reference/include/ikd-Tree/ikd_Tree.{h,cpp}
// by oracle/Makefile (target _ref) against the type shim in oracle/shim/.
// Nothing from the reference is copied into this repository: the Makefile
// passes -I/root/reference/include/ikd-Tree and compiles ikd_Tree.cpp in place.
//
// Reference API wrapped (include/ikd-Tree/ikd_Tree.h:165-187):
//   KD_TREE::set_downsample_param, Build, Nearest_Search, Add_Points, size,
//   validnum, flatten.
#include "ikd_Tree.h"

#include <omp.h>
#include <cstdint>
#include <vector>

namespace {
inline PointType mk(const float* p) {
    PointType q;
    q.x = p[0];
    q.y = p[1];
    q.z = p[2];
    return q;
}
}  // namespace

extern "C" {

// KD_TREE embeds a ~90 MB operation-log array (ikd_Tree.h:17,82) -> heap only.
void* ikdref_create(float ds) {
    KD_TREE* t = new KD_TREE();            // defaults as the node's global (laserMapping.cpp:125)
    t->set_downsample_param(ds);           // laserMapping.cpp:923
    return t;
}

void ikdref_destroy(void* h) { delete static_cast<KD_TREE*>(h); }

void ikdref_build(void* h, const float* xyz, int n) {
    PointVector v(n);
    for (int i = 0; i < n; i++) v[i] = mk(xyz + 3 * (size_t)i);
    static_cast<KD_TREE*>(h)->Build(v);    // laserMapping.cpp:928
}

int ikdref_size(void* h) { return static_cast<KD_TREE*>(h)->size(); }
int ikdref_validnum(void* h) { return static_cast<KD_TREE*>(h)->validnum(); }

// Nearest_Search for nq queries (laserMapping.cpp:980). out_xyz: nq*k*3 floats,
// out_d2: nq*k floats, out_cnt: nq ints (number actually found, <= k).
void ikdref_nearest(void* h, const float* q, int nq, int k, double max_dist,
                    float* out_xyz, float* out_d2, int* out_cnt, int nthreads) {
    KD_TREE* t = static_cast<KD_TREE*>(h);
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int i = 0; i < nq; i++) {
        PointVector near;
        std::vector<float> d2;
        t->Nearest_Search(mk(q + 3 * (size_t)i), k, near, d2, max_dist);
        int c = (int)near.size();
        out_cnt[i] = c;
        for (int j = 0; j < k; j++) {
            size_t o = (size_t)i * k + j;
            if (j < c) {
                out_xyz[3 * o + 0] = near[j].x;
                out_xyz[3 * o + 1] = near[j].y;
                out_xyz[3 * o + 2] = near[j].z;
                out_d2[o] = d2[j];
            } else {
                out_xyz[3 * o + 0] = out_xyz[3 * o + 1] = out_xyz[3 * o + 2] = 0.f;
                out_d2[o] = -1.f;
            }
        }
    }
}

int ikdref_add_points(void* h, const float* xyz, int n, int downsample_on) {
    PointVector v(n);
    for (int i = 0; i < n; i++) v[i] = mk(xyz + 3 * (size_t)i);
    return static_cast<KD_TREE*>(h)->Add_Points(v, downsample_on != 0);   // laserMapping.cpp:556-557
}

// All live (non-deleted) points. Returns the count; writes min(count, cap).
int ikdref_flatten(void* h, float* out_xyz, int cap) {
    KD_TREE* t = static_cast<KD_TREE*>(h);
    PointVector v;
    if (t->Root_Node != nullptr) t->flatten(t->Root_Node, v, NOT_RECORD);
    int n = (int)v.size();
    for (int i = 0; i < n && i < cap; i++) {
        out_xyz[3 * (size_t)i + 0] = v[i].x;
        out_xyz[3 * (size_t)i + 1] = v[i].y;
        out_xyz[3 * (size_t)i + 2] = v[i].z;
    }
    return n;
}

}  // extern "C"
