// oracle/liinit_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of LI-Init's per-scan point-to-plane ICP measurement model
// (the block inlined in main() at /root/reference/src/laserMapping.cpp:957-1134)
// plus the map update that follows it (laserMapping.cpp:516-559 and the
// semantics of KD_TREE::Build / Nearest_Search / Add_Points,
// include/ikd-Tree/ikd_Tree.cpp:336-456,825-968).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library, and only as the checker / the timed
// CPU baseline. The product (lidar_imu_init_b200/) never links or loads it.
//
// PARITY PIN: the reference ships no tests or golden vectors for this path
// (SURVEY.md section 4). The pin is (i) the reference's ikd-Tree compiled
// verbatim (oracle/_ref, see Makefile) which this file's restated kd-tree /
// Add_Points are checked against in tests/test_oracle.py, (ii) analytic planar
// scenes with closed-form residuals/Jacobians, (iii) numpy lstsq vs the QR
// restated here. Eigen/PCL/ROS are not installed, so laserMapping.cpp itself
// cannot be compiled: "parity unpinned" by reference-owned vectors.
//
// Two map back-ends implement the same interface:
//   backend 0: restated static kd-tree + sequential Add_Points restatement
//              (always available, self-contained).
//   backend 1: the reference's verbatim KD_TREE (only when compiled with
//              -DORACLE_WITH_IKD, i.e. oracle/_ref/liboracle_ref.so).
//
// Deliberate deviation (SURVEY.md fact box): dynamic sizes instead of the
// reference's fixed 100000-point arrays (laserMapping.cpp:108-109,117-119).
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <unordered_map>
#include <vector>

#ifdef ORACLE_WITH_IKD
#include "ikd_Tree.h"
#endif

namespace orc {

struct P3 {
    float x, y, z;
};

// ---------------------------------------------------------------------------
// fp32 squared distance exactly as KD_TREE::calc_dist (ikd_Tree.cpp:1273-1277):
// (dx*dx + dy*dy) + dz*dz, evaluated in float without contraction (the
// reference is built -O3 for baseline x86-64: no FMA, CMakeLists.txt:8).
// This file is compiled with -ffp-contract=off.
static inline float dist2f(const P3& a, const P3& b) {
    float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    float d = dx * dx + dy * dy + dz * dz;
    return d;
}

// Ordering of PointType_CMP (ikd_Tree.h:50-61): distances closer than 1e-10
// are ties broken by smaller x.
struct Cand {
    P3 p;
    float d;
    int id;
};
static inline bool cand_less(const Cand& a, const Cand& b) {
    if (std::fabs(a.d - b.d) < 1e-10) return a.p.x < b.p.x;
    return a.d < b.d;
}

// Bounded "max-priority queue" with the semantics Search() relies on
// (ikd_Tree.cpp:840-847): size(), top() = greatest under cand_less, pop, push.
// Kept as an ascending sorted array (k is 5).
struct TopK {
    Cand a[16];
    int n = 0;
    inline int size() const { return n; }
    inline const Cand& top() const { return a[n - 1]; }
    inline void pop() {
        if (n > 0) n--;
    }
    inline void push(const Cand& c) {
        int i = n++;
        while (i > 0 && cand_less(c, a[i - 1])) {
            a[i] = a[i - 1];
            i--;
        }
        a[i] = c;
    }
};

// ---------------------------------------------------------------------------
// Restated static kd-tree: same construction rule as KD_TREE::BuildTree
// (ikd_Tree.cpp:536-584: split at the median of the longest axis, first axis
// wins ties) and the same pruning rule as KD_TREE::Search (ikd_Tree.cpp:825-968).
struct KdNode {
    P3 p;
    int id;
    int axis;
    int left, right;
    float lo[3], hi[3];
};

struct KdTree {
    std::vector<KdNode> nodes;
    int root = -1;

    int build_rec(std::vector<std::pair<P3, int>>& s, int l, int r) {
        if (l > r) return -1;
        int mid = (l + r) >> 1;
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int i = l; i <= r; i++) {
            const P3& p = s[i].first;
            mn[0] = std::min(mn[0], p.x); mx[0] = std::max(mx[0], p.x);
            mn[1] = std::min(mn[1], p.y); mx[1] = std::max(mx[1], p.y);
            mn[2] = std::min(mn[2], p.z); mx[2] = std::max(mx[2], p.z);
        }
        int ax = 0;
        float rg[3] = {mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]};
        for (int i = 1; i < 3; i++)
            if (rg[i] > rg[ax]) ax = i;
        auto cmp = [ax](const std::pair<P3, int>& a, const std::pair<P3, int>& b) {
            const float* pa = &a.first.x;
            const float* pb = &b.first.x;
            return pa[ax] < pb[ax];
        };
        std::nth_element(s.begin() + l, s.begin() + mid, s.begin() + r + 1, cmp);
        int me = (int)nodes.size();
        nodes.push_back(KdNode());
        nodes[me].p = s[mid].first;
        nodes[me].id = s[mid].second;
        nodes[me].axis = ax;
        for (int i = 0; i < 3; i++) {
            nodes[me].lo[i] = mn[i];
            nodes[me].hi[i] = mx[i];
        }
        int L = build_rec(s, l, mid - 1);
        int R = build_rec(s, mid + 1, r);
        nodes[me].left = L;
        nodes[me].right = R;
        return me;
    }

    void build(const std::vector<P3>& pts, const std::vector<int>& ids) {
        nodes.clear();
        nodes.reserve(pts.size());
        std::vector<std::pair<P3, int>> s(pts.size());
        for (size_t i = 0; i < pts.size(); i++) s[i] = {pts[i], ids[i]};
        root = build_rec(s, 0, (int)s.size() - 1);
    }

    // calc_box_dist (ikd_Tree.cpp:1279-1289), float accumulation.
    inline float box_dist(int n, const P3& q) const {
        if (n < 0) return INFINITY;
        const KdNode& nd = nodes[n];
        float m = 0.0f;
        if (q.x < nd.lo[0]) m += (q.x - nd.lo[0]) * (q.x - nd.lo[0]);
        if (q.x > nd.hi[0]) m += (q.x - nd.hi[0]) * (q.x - nd.hi[0]);
        if (q.y < nd.lo[1]) m += (q.y - nd.lo[1]) * (q.y - nd.lo[1]);
        if (q.y > nd.hi[1]) m += (q.y - nd.hi[1]) * (q.y - nd.hi[1]);
        if (q.z < nd.lo[2]) m += (q.z - nd.lo[2]) * (q.z - nd.lo[2]);
        if (q.z > nd.hi[2]) m += (q.z - nd.hi[2]) * (q.z - nd.hi[2]);
        return m;
    }

    void search(int n, int k, const P3& q, TopK& h, double max_dist) const {
        if (n < 0) return;
        double cur = box_dist(n, q);
        if (cur > max_dist * max_dist) return;          // ikd_Tree.cpp:827-828 (sic: box uses max_dist^2)
        const KdNode& nd = nodes[n];
        float d = dist2f(q, nd.p);
        if (d <= max_dist && (h.size() < k || d < h.top().d)) {   // :842 (sic: d^2 vs un-squared max_dist)
            if (h.size() >= k) h.pop();
            h.push(Cand{nd.p, d, nd.id});
        }
        float dl = box_dist(nd.left, q), dr = box_dist(nd.right, q);
        if (h.size() < k || (dl < h.top().d && dr < h.top().d)) {
            if (dl <= dr) {
                search(nd.left, k, q, h, max_dist);
                if (h.size() < k || dr < h.top().d) search(nd.right, k, q, h, max_dist);
            } else {
                search(nd.right, k, q, h, max_dist);
                if (h.size() < k || dl < h.top().d) search(nd.left, k, q, h, max_dist);
            }
        } else {
            if (dl < h.top().d) search(nd.left, k, q, h, max_dist);
            if (dr < h.top().d) search(nd.right, k, q, h, max_dist);
        }
    }
};

// ---------------------------------------------------------------------------
// Map interface
struct MapBase {
    virtual ~MapBase() {}
    virtual void build(const float* xyz, int n) = 0;
    virtual int add_points(const float* xyz, int n, bool downsample_on) = 0;
    virtual int delete_boxes(const float* boxes, int nbox) = 0;   // KD_TREE::Delete_Point_Boxes
    virtual int size() = 0;       // KD_TREE::size(): nodes incl. lazily deleted
    virtual int validnum() = 0;   // live points
    virtual int flatten(float* out, int cap) = 0;
    // returns count found (<=k); out ascending
    virtual int knn(const P3& q, int k, double max_dist, P3* out, float* d2, int* ids) = 0;
    virtual void prepare() {}
    virtual bool empty() = 0;
};

// Backend 0: restatement. Live set + spatial hash for the downsample boxes +
// lazily rebuilt static kd-tree for searches.
struct MapRestated : MapBase {
    float ds;                       // KD_TREE::downsample_size (float, ikd_Tree.h:134)
    std::vector<P3> pts;
    std::vector<uint8_t> alive;
    int n_alive = 0;
    std::unordered_map<uint64_t, std::vector<int>> vox;   // voxel -> indices (may hold dead ones)
    KdTree tree;
    bool dirty = true;

    explicit MapRestated(float ds_) : ds(ds_) {}

    static inline uint64_t key(int ix, int iy, int iz) {
        return ((uint64_t)(uint32_t)(ix + (1 << 20)) << 42) | ((uint64_t)(uint32_t)(iy + (1 << 20)) << 21) |
               (uint64_t)(uint32_t)(iz + (1 << 20));
    }
    inline void vidx(const P3& p, int& ix, int& iy, int& iz) const {
        ix = (int)std::floor(p.x / ds);
        iy = (int)std::floor(p.y / ds);
        iz = (int)std::floor(p.z / ds);
    }
    void push_live(const P3& p) {
        int id = (int)pts.size();
        pts.push_back(p);
        alive.push_back(1);
        n_alive++;
        int ix, iy, iz;
        vidx(p, ix, iy, iz);
        vox[key(ix, iy, iz)].push_back(id);
        dirty = true;
    }
    void build(const float* xyz, int n) override {   // KD_TREE::Build: no downsampling (ikd_Tree.cpp:336-347)
        pts.clear(); alive.clear(); vox.clear(); n_alive = 0;
        for (int i = 0; i < n; i++) push_live(P3{xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2]});
        dirty = true;
    }
    // Live points inside the float box [mn, mx) exactly as Search_by_range /
    // Delete_by_range test them (ikd_Tree.cpp:633,980): mn <= c && mx > c.
    void in_box(const float mn[3], const float mx[3], int ix, int iy, int iz, std::vector<int>& out) {
        out.clear();
        for (int dx = -1; dx <= 1; dx++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dz = -1; dz <= 1; dz++) {
                    auto it = vox.find(key(ix + dx, iy + dy, iz + dz));
                    if (it == vox.end()) continue;
                    for (int id : it->second) {
                        if (!alive[id]) continue;
                        const P3& c = pts[id];
                        if (mn[0] <= c.x && mx[0] > c.x && mn[1] <= c.y && mx[1] > c.y && mn[2] <= c.z && mx[2] > c.z)
                            out.push_back(id);
                    }
                }
    }
    int add_points(const float* xyz, int n, bool downsample_on) override {   // ikd_Tree.cpp:381-456
        int counter = 0;
        std::vector<int> box;
        for (int i = 0; i < n; i++) {
            P3 p{xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2]};
            if (!downsample_on) {
                push_live(p);
                continue;
            }
            float mn[3], mx[3];
            mn[0] = std::floor(p.x / ds) * ds; mx[0] = mn[0] + ds;
            mn[1] = std::floor(p.y / ds) * ds; mx[1] = mn[1] + ds;
            mn[2] = std::floor(p.z / ds) * ds; mx[2] = mn[2] + ds;
            P3 mid;
            // float + (float-float)/2.0 -> double expr stored to float (ikd_Tree.cpp:397-399)
            mid.x = (float)(mn[0] + (mx[0] - mn[0]) / 2.0);
            mid.y = (float)(mn[1] + (mx[1] - mn[1]) / 2.0);
            mid.z = (float)(mn[2] + (mx[2] - mn[2]) / 2.0);
            int ix, iy, iz;
            vidx(p, ix, iy, iz);
            in_box(mn, mx, ix, iy, iz, box);
            float min_dist = dist2f(p, mid);
            P3 result = p;
            for (int id : box) {
                float t = dist2f(pts[id], mid);
                if (t < min_dist) {
                    min_dist = t;
                    result = pts[id];
                }
            }
            bool same = std::fabs(p.x - result.x) < 1e-6 && std::fabs(p.y - result.y) < 1e-6 &&
                        std::fabs(p.z - result.z) < 1e-6;    // same_point, EPSS (ikd_Tree.cpp:1269-1271)
            if (box.size() > 1 || same) {
                for (int id : box) {
                    alive[id] = 0;
                    n_alive--;
                }
                push_live(result);
                counter++;
            }
        }
        dirty = true;
        return counter;
    }
    int delete_boxes(const float* boxes, int nbox) override {   // Delete_by_range point test (ikd_Tree.cpp:633)
        int c = 0;
        for (size_t i = 0; i < pts.size(); i++) {
            if (!alive[i]) continue;
            const P3& p = pts[i];
            for (int b = 0; b < nbox; b++) {
                const float* x = boxes + 6 * b;
                if (x[0] <= p.x && x[3] > p.x && x[1] <= p.y && x[4] > p.y && x[2] <= p.z && x[5] > p.z) {
                    alive[i] = 0;
                    n_alive--;
                    c++;
                    break;
                }
            }
        }
        dirty = true;
        return c;
    }
    int size() override { return (int)pts.size(); }
    int validnum() override { return n_alive; }
    bool empty() override { return pts.empty(); }
    int flatten(float* out, int cap) override {
        int c = 0;
        for (size_t i = 0; i < pts.size(); i++) {
            if (!alive[i]) continue;
            if (c < cap) {
                out[3 * (size_t)c] = pts[i].x; out[3 * (size_t)c + 1] = pts[i].y; out[3 * (size_t)c + 2] = pts[i].z;
            }
            c++;
        }
        return c;
    }
    void prepare() override {
        if (!dirty) return;
        std::vector<P3> lp;
        std::vector<int> ids;
        lp.reserve(n_alive); ids.reserve(n_alive);
        for (size_t i = 0; i < pts.size(); i++)
            if (alive[i]) {
                lp.push_back(pts[i]);
                ids.push_back((int)i);
            }
        tree.build(lp, ids);
        dirty = false;
    }
    int knn(const P3& q, int k, double max_dist, P3* out, float* d2, int* ids) override {
        TopK h;
        tree.search(tree.root, k, q, h, max_dist);
        int c = std::min(k, h.size());
        for (int j = 0; j < c; j++) {
            out[j] = h.a[j].p;
            d2[j] = h.a[j].d;
            if (ids) ids[j] = h.a[j].id;
        }
        return c;
    }
};

#ifdef ORACLE_WITH_IKD
// Backend 1: the reference's own KD_TREE.
struct MapIkd : MapBase {
    KD_TREE* t;
    explicit MapIkd(float ds) {
        t = new KD_TREE();
        t->set_downsample_param(ds);
    }
    ~MapIkd() override { delete t; }
    static PointType mk(const float* p) {
        PointType q;
        q.x = p[0]; q.y = p[1]; q.z = p[2];
        return q;
    }
    void build(const float* xyz, int n) override {
        PointVector v(n);
        for (int i = 0; i < n; i++) v[i] = mk(xyz + 3 * (size_t)i);
        t->Build(v);
    }
    int add_points(const float* xyz, int n, bool downsample_on) override {
        PointVector v(n);
        for (int i = 0; i < n; i++) v[i] = mk(xyz + 3 * (size_t)i);
        return t->Add_Points(v, downsample_on);
    }
    int delete_boxes(const float* boxes, int nbox) override {
        std::vector<BoxPointType> v(nbox);
        for (int b = 0; b < nbox; b++)
            for (int a = 0; a < 3; a++) {
                v[b].vertex_min[a] = boxes[6 * b + a];
                v[b].vertex_max[a] = boxes[6 * b + 3 + a];
            }
        return t->Delete_Point_Boxes(v);
    }
    int size() override { return t->size(); }
    int validnum() override { return t->validnum(); }
    bool empty() override { return t->Root_Node == nullptr; }
    int flatten(float* out, int cap) override {
        PointVector v;
        if (t->Root_Node) t->flatten(t->Root_Node, v, NOT_RECORD);
        for (int i = 0; i < (int)v.size() && i < cap; i++) {
            out[3 * (size_t)i] = v[i].x; out[3 * (size_t)i + 1] = v[i].y; out[3 * (size_t)i + 2] = v[i].z;
        }
        return (int)v.size();
    }
    int knn(const P3& q, int k, double max_dist, P3* out, float* d2, int* ids) override {
        PointType p;
        p.x = q.x; p.y = q.y; p.z = q.z;
        PointVector near;
        std::vector<float> dd;
        t->Nearest_Search(p, k, near, dd, max_dist);
        int c = (int)near.size();
        for (int j = 0; j < c && j < k; j++) {
            out[j] = P3{near[j].x, near[j].y, near[j].z};
            d2[j] = dd[j];
            if (ids) ids[j] = -1;
        }
        return c;
    }
};
#endif

// ---------------------------------------------------------------------------
// Small dense math (Eigen stand-ins), all double.
static inline void mat3_mul_vec(const double* R, const double* v, double* o) {   // row-major 3x3
    for (int i = 0; i < 3; i++) o[i] = R[3 * i] * v[0] + R[3 * i + 1] * v[1] + R[3 * i + 2] * v[2];
}
static inline void mat3T_mul_vec(const double* R, const double* v, double* o) {
    for (int i = 0; i < 3; i++) o[i] = R[i] * v[0] + R[3 + i] * v[1] + R[6 + i] * v[2];
}
static inline void mat3_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static inline void mat3_T(const double* A, double* T) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) T[3 * i + j] = A[3 * j + i];
}
static inline void skew(const double* v, double* K) {   // so3_math.h:8
    K[0] = 0; K[1] = -v[2]; K[2] = v[1];
    K[3] = v[2]; K[4] = 0; K[5] = -v[0];
    K[6] = -v[1]; K[7] = v[0]; K[8] = 0;
}
// Exp(v1,v2,v3) (so3_math.h:61-80): identity below 1e-5.
static void so3_exp(double v1, double v2, double v3, double* R) {
    double n = std::sqrt(v1 * v1 + v2 * v2 + v3 * v3);
    double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (n > 0.00001) {
        double r[3] = {v1 / n, v2 / n, v3 / n};
        double K[9], KK[9];
        skew(r, K);
        mat3_mul(K, K, KK);
        double s = std::sin(n), c = 1.0 - std::cos(n);
        for (int i = 0; i < 9; i++) R[i] = I[i] + s * K[i] + c * KK[i];
    } else {
        for (int i = 0; i < 9; i++) R[i] = I[i];
    }
}
// Log(R) (so3_math.h:100-107)
static void so3_log(const double* R, double* o) {
    double tr = R[0] + R[4] + R[8];
    double theta = (tr > 3.0 - 1e-6) ? 0.0 : std::acos(0.5 * (tr - 1));
    double K[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
    double f = (std::fabs(theta) < 0.001) ? 0.5 : (0.5 * theta / std::sin(theta));
    for (int i = 0; i < 3; i++) o[i] = f * K[i];
}

// n x n inverse by partial-pivot LU (what Eigen's .inverse() does for n > 4).
static bool mat_inverse(const double* A, double* Ainv, int n) {
    std::vector<double> lu(A, A + (size_t)n * n);
    std::vector<int> piv(n);
    for (int i = 0; i < n; i++) piv[i] = i;
    for (int k = 0; k < n; k++) {
        int p = k;
        double best = std::fabs(lu[(size_t)k * n + k]);
        for (int i = k + 1; i < n; i++) {
            double v = std::fabs(lu[(size_t)i * n + k]);
            if (v > best) { best = v; p = i; }
        }
        if (best == 0.0) return false;
        if (p != k) {
            for (int j = 0; j < n; j++) std::swap(lu[(size_t)k * n + j], lu[(size_t)p * n + j]);
            std::swap(piv[k], piv[p]);
        }
        for (int i = k + 1; i < n; i++) {
            lu[(size_t)i * n + k] /= lu[(size_t)k * n + k];
            double f = lu[(size_t)i * n + k];
            for (int j = k + 1; j < n; j++) lu[(size_t)i * n + j] -= f * lu[(size_t)k * n + j];
        }
    }
    for (int c = 0; c < n; c++) {
        std::vector<double> y(n);
        for (int i = 0; i < n; i++) {
            double s = (piv[i] == c) ? 1.0 : 0.0;
            for (int j = 0; j < i; j++) s -= lu[(size_t)i * n + j] * y[j];
            y[i] = s;
        }
        for (int i = n - 1; i >= 0; i--) {
            double s = y[i];
            for (int j = i + 1; j < n; j++) s -= lu[(size_t)i * n + j] * Ainv[(size_t)j * n + c];
            Ainv[(size_t)i * n + c] = s / lu[(size_t)i * n + i];
        }
    }
    return true;
}

// Least squares min ||A x - b||, A 5x3, by column-pivoted Householder QR: the
// algorithm behind Eigen's colPivHouseholderQr().solve() used by esti_plane
// (common_lib.h:252). Any backward-stable LSQ agrees to ~1e-12 on full-rank A.
static void lsq_5x3_colpiv(const double Ain[5][3], const double bin[5], double x[3]) {
    double A[5][3], b[5];
    for (int i = 0; i < 5; i++) {
        b[i] = bin[i];
        for (int j = 0; j < 3; j++) A[i][j] = Ain[i][j];
    }
    int perm[3] = {0, 1, 2};
    double cn[3];
    double maxn = 0;
    for (int j = 0; j < 3; j++) {
        cn[j] = 0;
        for (int i = 0; i < 5; i++) cn[j] += A[i][j] * A[i][j];
        maxn = std::max(maxn, cn[j]);
    }
    const double eps = 2.220446049250313e-16;
    double thresh = (std::sqrt(maxn) * eps) * (std::sqrt(maxn) * eps) / 5.0;
    int rank = 3;
    for (int k = 0; k < 3; k++) {
        int p = k;
        double best = -1;
        for (int j = k; j < 3; j++) {
            double s = 0;
            for (int i = k; i < 5; i++) s += A[i][j] * A[i][j];
            cn[j] = s;
            if (s > best) { best = s; p = j; }
        }
        if (best < thresh * (double)(5 - k)) { rank = k; break; }   // Eigen: biggest_col_sq_norm < threshold_helper * (rows - k)
        if (p != k) {
            for (int i = 0; i < 5; i++) std::swap(A[i][k], A[i][p]);
            std::swap(perm[k], perm[p]);
        }
        double alpha = A[k][k];
        double tail = 0;
        for (int i = k + 1; i < 5; i++) tail += A[i][k] * A[i][k];
        if (tail == 0.0) continue;   // already upper-triangular in this column
        double nrm = std::sqrt(alpha * alpha + tail);
        double beta = (alpha >= 0) ? -nrm : nrm;
        double tau = (beta - alpha) / beta;
        double scale = 1.0 / (alpha - beta);
        double v[5];
        v[k] = 1.0;
        for (int i = k + 1; i < 5; i++) v[i] = A[i][k] * scale;
        A[k][k] = beta;
        for (int i = k + 1; i < 5; i++) A[i][k] = 0.0;
        for (int j = k + 1; j < 3; j++) {
            double s = 0;
            for (int i = k; i < 5; i++) s += v[i] * A[i][j];
            s *= tau;
            for (int i = k; i < 5; i++) A[i][j] -= s * v[i];
        }
        double s = 0;
        for (int i = k; i < 5; i++) s += v[i] * b[i];
        s *= tau;
        for (int i = k; i < 5; i++) b[i] -= s * v[i];
    }
    double y[3] = {0, 0, 0};
    for (int i = rank - 1; i >= 0; i--) {
        double s = b[i];
        for (int j = i + 1; j < rank; j++) s -= A[i][j] * y[j];
        y[i] = s / A[i][i];
    }
    for (int j = 0; j < 3; j++) x[perm[j]] = y[j];
}

// esti_plane<double>(pabcd, points, threshold) (common_lib.h:236-269).
static bool esti_plane(double pabcd[4], const P3 nb[5], double threshold) {
    double A[5][3], b[5];
    for (int j = 0; j < 5; j++) {
        A[j][0] = nb[j].x; A[j][1] = nb[j].y; A[j][2] = nb[j].z;
        b[j] = -1.0;
    }
    double nv[3];
    lsq_5x3_colpiv(A, b, nv);
    double n = std::sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
    pabcd[0] = nv[0] / n; pabcd[1] = nv[1] / n; pabcd[2] = nv[2] / n;
    pabcd[3] = 1.0 / n;
    for (int j = 0; j < 5; j++) {
        if (std::fabs(pabcd[0] * nb[j].x + pabcd[1] * nb[j].y + pabcd[2] * nb[j].z + pabcd[3]) > threshold) return false;
    }
    return true;
}

// ---------------------------------------------------------------------------
// Per-scan state: the file-scope globals of laserMapping.cpp:102-125 that the
// path touches, with dynamic sizes.
struct Scan {
    int N = 0;
    std::vector<P3> body;               // feats_down_body (xyz)
    std::vector<P3> world;              // feats_down_world
    std::vector<int> near_cnt;          // Nearest_Points[i].size()
    std::vector<P3> near_pts;           // Nearest_Points[i][0..4]
    std::vector<float> near_d2;
    std::vector<int> near_ids;          // backend-0 ids (diagnostics)
    std::vector<uint8_t> selected;      // point_selected_surf
    std::vector<float> res_last;
    std::vector<float> normvec;         // (nx,ny,nz,pd2) as stored f32 (laserMapping.cpp:1004-1008)
    // outputs of the last iterate()
    int effect_feat_num = 0;
    std::vector<int> sel_index;         // compaction map k -> i
    std::vector<double> Hsub;           // m x 12
    std::vector<double> meas;           // m
};

struct Pose {
    double rot_end[9], pos_end[3], R_LI[9], T_LI[3];
};

// pointBodyToWorld (laserMapping.cpp:209-220): double math, float store.
static inline P3 body_to_world(const Pose& s, const P3& b) {
    double pb[3] = {b.x, b.y, b.z};
    double t[3], g[3];
    mat3_mul_vec(s.R_LI, pb, t);
    for (int i = 0; i < 3; i++) t[i] += s.T_LI[i];
    mat3_mul_vec(s.rot_end, t, g);
    for (int i = 0; i < 3; i++) g[i] += s.pos_end[i];
    return P3{(float)g[0], (float)g[1], (float)g[2]};
}

// One ICP iteration: laserMapping.cpp:959-1071 + the reduction of :1080.
// Outputs the UNWEIGHTED HtH (12x12 row-major) and Htr (12), r = meas = -pd2;
// the reference's R_inv = 1000 weight (:1050,1068) is applied by the caller.
static void icp_iterate(Scan& sc, MapBase* map, const Pose& st, bool imu_en, bool nearest_search_en, int nthreads,
                        double* HtH, double* Htr, int* m_out) {
    const int N = sc.N;
    if (nthreads < 1) nthreads = 1;
    if (nearest_search_en) map->prepare();
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 256)
    for (int i = 0; i < N; i++) {
        const P3& pb = sc.body[i];
        double p_body[3] = {pb.x, pb.y, pb.z};
        P3 pw = body_to_world(st, pb);
        sc.world[i] = pw;
        float d2[5] = {0, 0, 0, 0, 0};
        if (nearest_search_en) {
            int c = map->knn(pw, 5, 5.0, &sc.near_pts[5 * (size_t)i], d2, &sc.near_ids[5 * (size_t)i]);
            sc.near_cnt[i] = c;
            for (int j = 0; j < 5; j++) sc.near_d2[5 * (size_t)i + j] = (j < c) ? d2[j] : -1.f;
            if (c < 5)
                sc.selected[i] = 0;
            else
                sc.selected[i] = !(d2[4] > 5);
        }
        sc.res_last[i] = -1000.0f;
        if (!sc.selected[i] || sc.near_cnt[i] < 5) {
            sc.selected[i] = 0;
            continue;
        }
        sc.selected[i] = 0;
        double pabcd[4] = {0, 0, 0, 0};
        if (esti_plane(pabcd, &sc.near_pts[5 * (size_t)i], 0.1)) {
            float pd2 = (float)(pabcd[0] * pw.x + pabcd[1] * pw.y + pabcd[2] * pw.z + pabcd[3]);
            double pn = std::sqrt(p_body[0] * p_body[0] + p_body[1] * p_body[1] + p_body[2] * p_body[2]);
            float s = (float)(1 - 0.9 * std::fabs(pd2) / std::sqrt(pn));
            if (s > 0.9) {
                sc.selected[i] = 1;
                sc.normvec[4 * (size_t)i + 0] = (float)pabcd[0];
                sc.normvec[4 * (size_t)i + 1] = (float)pabcd[1];
                sc.normvec[4 * (size_t)i + 2] = (float)pabcd[2];
                sc.normvec[4 * (size_t)i + 3] = pd2;
                sc.res_last[i] = std::fabs(pd2);
            }
        }
    }
    // compaction (laserMapping.cpp:1013-1020)
    sc.sel_index.clear();
    for (int i = 0; i < N; i++)
        if (sc.selected[i]) sc.sel_index.push_back(i);
    const int m = (int)sc.sel_index.size();
    sc.effect_feat_num = m;
    sc.Hsub.assign((size_t)m * 12, 0.0);
    sc.meas.assign(m, 0.0);
    double RendT[9], RLIT[9];
    mat3_T(st.rot_end, RendT);
    mat3_T(st.R_LI, RLIT);
    // Jacobian rows (laserMapping.cpp:1035-1071)
    for (int k = 0; k < m; k++) {
        int i = sc.sel_index[k];
        double pL[3] = {sc.body[i].x, sc.body[i].y, sc.body[i].z};
        double pI[3];
        mat3_mul_vec(st.R_LI, pL, pI);
        for (int a = 0; a < 3; a++) pI[a] += st.T_LI[a];
        double cm[9];
        skew(pI, cm);
        double nv[3] = {sc.normvec[4 * (size_t)i], sc.normvec[4 * (size_t)i + 1], sc.normvec[4 * (size_t)i + 2]};
        double* row = &sc.Hsub[(size_t)k * 12];
        double M[9], A[3];
        mat3_mul(cm, RendT, M);              // (point_crossmat * rot_end^T) * n, left to right as Eigen
        mat3_mul_vec(M, nv, A);
        row[0] = A[0]; row[1] = A[1]; row[2] = A[2];
        row[3] = nv[0]; row[4] = nv[1]; row[5] = nv[2];
        if (imu_en) {
            double cL[9], M1[9], M2[9], B[3], C[3];
            skew(pL, cL);
            mat3_mul(cL, RLIT, M1);
            mat3_mul(M1, RendT, M2);
            mat3_mul_vec(M2, nv, B);
            mat3_mul_vec(RendT, nv, C);
            row[6] = B[0]; row[7] = B[1]; row[8] = B[2];
            row[9] = C[0]; row[10] = C[1]; row[11] = C[2];
        }
        sc.meas[k] = -(double)sc.normvec[4 * (size_t)i + 3];
    }
    for (int a = 0; a < 144; a++) HtH[a] = 0;
    for (int a = 0; a < 12; a++) Htr[a] = 0;
    for (int k = 0; k < m; k++) {
        const double* row = &sc.Hsub[(size_t)k * 12];
        for (int a = 0; a < 12; a++) {
            for (int b = 0; b < 12; b++) HtH[a * 12 + b] += row[a] * row[b];
            Htr[a] += row[a] * sc.meas[k];
        }
    }
    *m_out = m;
}

// StatesGroup (common_lib.h:68-169), flat.
struct State {
    double rot_end[9], pos_end[3], R_LI[9], T_LI[3], vel[3], bg[3], ba[3], grav[3];
    double cov[24 * 24];
};
static void state_boxplus(State& s, const double* d) {    // operator+= (common_lib.h:124-135)
    double E[9], R[9];
    so3_exp(d[0], d[1], d[2], E);
    mat3_mul(s.rot_end, E, R);
    std::memcpy(s.rot_end, R, sizeof(R));
    for (int i = 0; i < 3; i++) s.pos_end[i] += d[3 + i];
    so3_exp(d[6], d[7], d[8], E);
    mat3_mul(s.R_LI, E, R);
    std::memcpy(s.R_LI, R, sizeof(R));
    for (int i = 0; i < 3; i++) {
        s.T_LI[i] += d[9 + i];
        s.vel[i] += d[12 + i];
        s.bg[i] += d[15 + i];
        s.ba[i] += d[18 + i];
        s.grav[i] += d[21 + i];
    }
}
static void state_boxminus(const State& a, const State& b, double* o) {   // a - b (common_lib.h:137-151)
    double bT[9], rd[9];
    mat3_T(b.rot_end, bT);
    mat3_mul(bT, a.rot_end, rd);
    so3_log(rd, o);
    mat3_T(b.R_LI, bT);
    mat3_mul(bT, a.R_LI, rd);
    so3_log(rd, o + 6);
    for (int i = 0; i < 3; i++) {
        o[3 + i] = a.pos_end[i] - b.pos_end[i];
        o[9 + i] = a.T_LI[i] - b.T_LI[i];
        o[12 + i] = a.vel[i] - b.vel[i];
        o[15 + i] = a.bg[i] - b.bg[i];
        o[18 + i] = a.ba[i] - b.ba[i];
        o[21 + i] = a.grav[i] - b.grav[i];
    }
}

// IESKF update, literal form of laserMapping.cpp:1080-1087 (with the m-wide K).
// Returns solution (24) and K*Hsub (24x12) for the covariance update (:1113).
static void ieskf_update_literal(const Scan& sc, State& st, const State& prop, double* solution, double* KH /*24x12*/) {
    const int m = sc.effect_feat_num;
    const int D = 24;
    std::vector<double> HTH(D * D, 0.0), covinv(D * D), S(D * D), K1(D * D);
    for (int k = 0; k < m; k++) {
        const double* row = &sc.Hsub[(size_t)k * 12];
        for (int a = 0; a < 12; a++)
            for (int b = 0; b < 12; b++) HTH[a * D + b] += (row[a] * 1000) * row[b];   // Hsub_T_R_inv * Hsub
    }
    mat_inverse(st.cov, covinv.data(), D);
    for (int i = 0; i < D * D; i++) S[i] = HTH[i] + covinv[i];
    mat_inverse(S.data(), K1.data(), D);
    std::vector<double> K((size_t)D * std::max(m, 1), 0.0);   // K = K_1[:, 0:12] * Hsub_T_R_inv  (24 x m)
    for (int r = 0; r < D; r++)
        for (int k = 0; k < m; k++) {
            const double* row = &sc.Hsub[(size_t)k * 12];
            double s = 0;
            for (int a = 0; a < 12; a++) s += K1[r * D + a] * (row[a] * 1000);
            K[(size_t)r * m + k] = s;
        }
    double vec[24];
    state_boxminus(prop, st, vec);
    for (int r = 0; r < D; r++)
        for (int c = 0; c < 12; c++) {
            double s = 0;
            for (int k = 0; k < m; k++) s += K[(size_t)r * m + k] * sc.Hsub[(size_t)k * 12 + c];
            KH[r * 12 + c] = s;
        }
    for (int r = 0; r < D; r++) {
        double s = 0;
        for (int k = 0; k < m; k++) s += K[(size_t)r * m + k] * sc.meas[k];
        double t = 0;
        for (int c = 0; c < 12; c++) t += KH[r * 12 + c] * vec[c];
        solution[r] = s + vec[r] - t;
    }
    state_boxplus(st, solution);
}

// map_incremental (laserMapping.cpp:516-559). ds is the node's double
// filter_size_map_min; uses Nearest_Points left over from the last search.
static int map_incremental(Scan& sc, MapBase* map, const Pose& st, double ds, bool flg_EKF_inited, int* n_add, int* n_nod,
                           int* add_flag /*N, optional: 0 skip 1 add 2 no-downsample*/) {
    std::vector<float> to_add, no_ds;
    for (int i = 0; i < sc.N; i++) {
        P3 pw = body_to_world(st, sc.body[i]);
        sc.world[i] = pw;
        int flag = 0;
        if (sc.near_cnt[i] > 0 && flg_EKF_inited) {
            const P3* near = &sc.near_pts[5 * (size_t)i];
            bool need_add = true;
            P3 mid;
            mid.x = (float)(std::floor(pw.x / ds) * ds + 0.5 * ds);
            mid.y = (float)(std::floor(pw.y / ds) * ds + 0.5 * ds);
            mid.z = (float)(std::floor(pw.z / ds) * ds + 0.5 * ds);
            float dist = dist2f(pw, mid);
            if (std::fabs(near[0].x - mid.x) > 0.5 * ds && std::fabs(near[0].y - mid.y) > 0.5 * ds &&
                std::fabs(near[0].z - mid.z) > 0.5 * ds) {
                no_ds.push_back(pw.x); no_ds.push_back(pw.y); no_ds.push_back(pw.z);
                if (add_flag) add_flag[i] = 2;
                continue;
            }
            for (int j = 0; j < 5; j++) {
                if (sc.near_cnt[i] < 5) break;
                if (dist2f(near[j], mid) < dist) {
                    need_add = false;
                    break;
                }
            }
            if (need_add) {
                to_add.push_back(pw.x); to_add.push_back(pw.y); to_add.push_back(pw.z);
                flag = 1;
            }
        } else {
            to_add.push_back(pw.x); to_add.push_back(pw.y); to_add.push_back(pw.z);
            flag = 1;
        }
        if (add_flag) add_flag[i] = flag;
    }
    int c = map->add_points(to_add.data(), (int)to_add.size() / 3, true);
    map->add_points(no_ds.data(), (int)no_ds.size() / 3, false);
    if (n_add) *n_add = (int)to_add.size() / 3;
    if (n_nod) *n_nod = (int)no_ds.size() / 3;
    return c;
}

}  // namespace orc

// ===========================================================================
// C-ABI used by tests / bench via ctypes.
using namespace orc;

extern "C" {

int oracle_has_ikd() {
#ifdef ORACLE_WITH_IKD
    return 1;
#else
    return 0;
#endif
}

void* oracle_map_create(int backend, float ds) {
    if (backend == 0) return new MapRestated(ds);
#ifdef ORACLE_WITH_IKD
    if (backend == 1) return new MapIkd(ds);
#endif
    return nullptr;
}
void oracle_map_destroy(void* m) { delete static_cast<MapBase*>(m); }
void oracle_map_build(void* m, const float* xyz, int n) { static_cast<MapBase*>(m)->build(xyz, n); }
int oracle_map_add_points(void* m, const float* xyz, int n, int downsample_on) {
    return static_cast<MapBase*>(m)->add_points(xyz, n, downsample_on != 0);
}
int oracle_map_delete_boxes(void* m, const float* boxes, int nbox) { return static_cast<MapBase*>(m)->delete_boxes(boxes, nbox); }
int oracle_map_size(void* m) { return static_cast<MapBase*>(m)->size(); }
int oracle_map_validnum(void* m) { return static_cast<MapBase*>(m)->validnum(); }
int oracle_map_flatten(void* m, float* out, int cap) { return static_cast<MapBase*>(m)->flatten(out, cap); }
void oracle_map_knn(void* m, const float* q, int nq, int k, double max_dist, float* out_xyz, float* out_d2, int* out_cnt,
                    int* out_ids, int nthreads) {
    MapBase* mp = static_cast<MapBase*>(m);
    mp->prepare();
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 256)
    for (int i = 0; i < nq; i++) {
        P3 out[16];
        float d2[16];
        int ids[16];
        int c = mp->knn(P3{q[3 * (size_t)i], q[3 * (size_t)i + 1], q[3 * (size_t)i + 2]}, k, max_dist, out, d2, ids);
        out_cnt[i] = c;
        for (int j = 0; j < k; j++) {
            size_t o = (size_t)i * k + j;
            if (j < c) {
                out_xyz[3 * o] = out[j].x; out_xyz[3 * o + 1] = out[j].y; out_xyz[3 * o + 2] = out[j].z;
                out_d2[o] = d2[j];
                if (out_ids) out_ids[o] = ids[j];
            } else {
                out_xyz[3 * o] = out_xyz[3 * o + 1] = out_xyz[3 * o + 2] = 0.f;
                out_d2[o] = -1.f;
                if (out_ids) out_ids[o] = -1;
            }
        }
    }
}

// Brute-force exact kNN (O(nq*M)), independent of any tree: the arbiter for
// small cases. Same fp32 distance and (dist, x) ordering.
void oracle_knn_bruteforce(const float* map_xyz, int M, const float* q, int nq, int k, double max_dist, float* out_xyz,
                           float* out_d2, int* out_cnt, int* out_ids, int nthreads) {
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 64)
    for (int i = 0; i < nq; i++) {
        P3 qq{q[3 * (size_t)i], q[3 * (size_t)i + 1], q[3 * (size_t)i + 2]};
        TopK h;
        for (int j = 0; j < M; j++) {
            P3 p{map_xyz[3 * (size_t)j], map_xyz[3 * (size_t)j + 1], map_xyz[3 * (size_t)j + 2]};
            float d = dist2f(qq, p);
            if (d <= max_dist && (h.size() < k || d < h.top().d)) {
                if (h.size() >= k) h.pop();
                h.push(Cand{p, d, j});
            }
        }
        int c = std::min(k, h.size());
        out_cnt[i] = c;
        for (int j = 0; j < k; j++) {
            size_t o = (size_t)i * k + j;
            if (j < c) {
                out_xyz[3 * o] = h.a[j].p.x; out_xyz[3 * o + 1] = h.a[j].p.y; out_xyz[3 * o + 2] = h.a[j].p.z;
                out_d2[o] = h.a[j].d;
                if (out_ids) out_ids[o] = h.a[j].id;
            } else {
                out_xyz[3 * o] = out_xyz[3 * o + 1] = out_xyz[3 * o + 2] = 0.f;
                out_d2[o] = -1.f;
                if (out_ids) out_ids[o] = -1;
            }
        }
    }
}

void* oracle_scan_create(const float* body_xyz, int N) {
    Scan* s = new Scan();
    s->N = N;
    s->body.resize(N);
    for (int i = 0; i < N; i++) s->body[i] = P3{body_xyz[3 * (size_t)i], body_xyz[3 * (size_t)i + 1], body_xyz[3 * (size_t)i + 2]};
    s->world.assign(N, P3{0, 0, 0});
    s->near_cnt.assign(N, 0);
    s->near_pts.assign((size_t)5 * N, P3{0, 0, 0});
    s->near_d2.assign((size_t)5 * N, -1.f);
    s->near_ids.assign((size_t)5 * N, -1);
    s->selected.assign(N, 0);
    s->res_last.assign(N, 0.f);
    s->normvec.assign((size_t)4 * N, 0.f);
    return s;
}
void oracle_scan_destroy(void* s) { delete static_cast<Scan*>(s); }

static Pose mk_pose(const double* rot_end, const double* pos_end, const double* R_LI, const double* T_LI) {
    Pose p;
    std::memcpy(p.rot_end, rot_end, 72);
    std::memcpy(p.pos_end, pos_end, 24);
    std::memcpy(p.R_LI, R_LI, 72);
    std::memcpy(p.T_LI, T_LI, 24);
    return p;
}

void oracle_icp_iterate(void* scan, void* map, const double* rot_end, const double* pos_end, const double* R_LI,
                        const double* T_LI, int imu_en, int nearest_search_en, int nthreads, double* HtH, double* Htr,
                        int* m) {
    Pose p = mk_pose(rot_end, pos_end, R_LI, T_LI);
    icp_iterate(*static_cast<Scan*>(scan), static_cast<MapBase*>(map), p, imu_en != 0, nearest_search_en != 0, nthreads,
                HtH, Htr, m);
}

// Per-point outputs of the last iterate (any pointer may be null).
void oracle_scan_get(void* scan, float* world_xyz, int* near_cnt, float* near_xyz, float* near_d2, unsigned char* selected,
                     float* normvec, float* res_last) {
    Scan* s = static_cast<Scan*>(scan);
    int N = s->N;
    if (world_xyz) std::memcpy(world_xyz, s->world.data(), sizeof(P3) * N);
    if (near_cnt) std::memcpy(near_cnt, s->near_cnt.data(), sizeof(int) * N);
    if (near_xyz) std::memcpy(near_xyz, s->near_pts.data(), sizeof(P3) * 5 * (size_t)N);
    if (near_d2) std::memcpy(near_d2, s->near_d2.data(), sizeof(float) * 5 * (size_t)N);
    if (selected) std::memcpy(selected, s->selected.data(), N);
    if (normvec) std::memcpy(normvec, s->normvec.data(), sizeof(float) * 4 * (size_t)N);
    if (res_last) std::memcpy(res_last, s->res_last.data(), sizeof(float) * N);
}
int oracle_scan_get_H(void* scan, double* Hsub, double* meas, int* sel_index) {
    Scan* s = static_cast<Scan*>(scan);
    int m = s->effect_feat_num;
    if (Hsub) std::memcpy(Hsub, s->Hsub.data(), sizeof(double) * 12 * (size_t)m);
    if (meas) std::memcpy(meas, s->meas.data(), sizeof(double) * m);
    if (sel_index) std::memcpy(sel_index, s->sel_index.data(), sizeof(int) * m);
    return m;
}

// state layout for the C-ABI: 24 + 9 + 9 ... flat struct State (see above):
// rot_end[9] pos_end[3] R_LI[9] T_LI[3] vel[3] bg[3] ba[3] grav[3] cov[576] = 612 doubles.
int oracle_state_doubles() { return (int)(sizeof(State) / sizeof(double)); }

// Literal IESKF update on the rows of the last iterate. state is updated in place.
void oracle_ieskf_update(void* scan, double* state, const double* state_propagat, double* solution, double* KH) {
    State st, pr;
    std::memcpy(&st, state, sizeof(State));
    std::memcpy(&pr, state_propagat, sizeof(State));
    ieskf_update_literal(*static_cast<Scan*>(scan), st, pr, solution, KH);
    std::memcpy(state, &st, sizeof(State));
}

// Whole per-scan update (laserMapping.cpp:936-1134): iteration/rematch policy,
// convergence test and covariance update. state in/out. Returns iterations run.
// stats[0]=search passes, stats[1]=last effect_feat_num.
int oracle_scan_update(void* scan, void* map, double* state, int max_iter, int imu_en, int nthreads, int* stats) {
    Scan& sc = *static_cast<Scan*>(scan);
    MapBase* mp = static_cast<MapBase*>(map);
    State st, prop;
    std::memcpy(&st, state, sizeof(State));
    prop = st;                                   // state_propagat = state (laserMapping.cpp:910)
    int rematch_num = 0;
    bool nearest_search_en = true;
    int iters = 0, searches = 0;
    double HtH[144], Htr[12];
    int m = 0;
    for (int it = 0; it < max_iter; it++) {
        Pose p = mk_pose(st.rot_end, st.pos_end, st.R_LI, st.T_LI);
        if (nearest_search_en) searches++;
        icp_iterate(sc, mp, p, imu_en != 0, nearest_search_en, nthreads, HtH, Htr, &m);
        double sol[24], KH[24 * 12];
        ieskf_update_literal(sc, st, prop, sol, KH);
        iters++;
        double rn = std::sqrt(sol[0] * sol[0] + sol[1] * sol[1] + sol[2] * sol[2]);
        double tn = std::sqrt(sol[3] * sol[3] + sol[4] * sol[4] + sol[5] * sol[5]);
        bool converged = (rn * 57.3 < 0.01) && (tn * 100 < 0.015);
        nearest_search_en = false;
        if (converged || ((rematch_num == 0) && (it == (max_iter - 2)))) {
            nearest_search_en = true;
            rematch_num++;
        }
        if (rematch_num >= 2 || (it == max_iter - 1)) {
            // cov = (I - G) cov, G[:, 0:12] = K*Hsub (laserMapping.cpp:1112-1114)
            std::vector<double> nc(24 * 24, 0.0);
            for (int r = 0; r < 24; r++)
                for (int c = 0; c < 24; c++) {
                    double s = st.cov[r * 24 + c];
                    for (int a = 0; a < 12; a++) s -= KH[r * 12 + a] * st.cov[a * 24 + c];
                    nc[r * 24 + c] = s;
                }
            std::memcpy(st.cov, nc.data(), sizeof(st.cov));
            break;
        }
    }
    std::memcpy(state, &st, sizeof(State));
    if (stats) {
        stats[0] = searches;
        stats[1] = m;
    }
    return iters;
}

int oracle_map_incremental(void* scan, void* map, const double* rot_end, const double* pos_end, const double* R_LI,
                           const double* T_LI, double ds, int flg_EKF_inited, int* n_add, int* n_nod, int* add_flag) {
    Pose p = mk_pose(rot_end, pos_end, R_LI, T_LI);
    return map_incremental(*static_cast<Scan*>(scan), static_cast<MapBase*>(map), p, ds, flg_EKF_inited != 0, n_add, n_nod,
                           add_flag);
}

// Exp(ang_vel, dt) (so3_math.h:39-59)
static void so3_exp_dt(const double* w, double dt, double* R) {
    const double n = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (n > 0.0000001) {
        double r[3] = {w[0] / n, w[1] / n, w[2] / n};
        double K[9], KK[9];
        skew(r, K);
        mat3_mul(K, K, KK);
        const double a = n * dt, s = std::sin(a), c = 1.0 - std::cos(a);
        for (int i = 0; i < 9; i++) R[i] = I[i] + s * K[i] + c * KK[i];
    } else {
        for (int i = 0; i < 9; i++) R[i] = I[i];
    }
}

// The un-distortion loop of Forward_propagation_without_imu (IMU_Processing.hpp:208-212,246-266): time-sorted cloud walked
// backwards from the last point down to (not including) the first. xyz in/out, t_ms = PointType.curvature.
void oracle_undistort_cv(float* xyz, const float* t_ms, int n, const double* omega, const double* rot_end, const double* vel_end) {
    std::vector<int> ord(n);
    for (int i = 0; i < n; i++) ord[i] = i;
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return t_ms[a] < t_ms[b]; });   // time_list (:15)
    const double t_end = t_ms[ord[n - 1]] / double(1000);
    double RT[9], vb[3];
    mat3_T(rot_end, RT);
    mat3_mul_vec(RT, vel_end, vb);
    for (int k = n - 1; k != 0; k--) {
        float* p = xyz + 3 * (size_t)ord[k];
        const double dt_j = t_end - t_ms[ord[k]] / double(1000);
        double R[9];
        so3_exp_dt(omega, -dt_j, R);
        double P[3] = {p[0], p[1], p[2]}, o[3];
        mat3_mul_vec(R, P, o);
        for (int a = 0; a < 3; a++) p[a] = (float)(o[a] + (-vb[a] * dt_j));
    }
}

// The back-propagation loop of propagation_and_undist (IMU_Processing.hpp:390-415). poses: npose x 22 doubles
// {offset_time, acc[3], gyr[3], vel[3], pos[3], rot[9]}.
void oracle_undistort_imu(float* xyz, const float* t_ms, int n, const double* poses, int npose, const double* rot_end,
                          const double* pos_end, const double* R_LI, const double* T_LI) {
    std::vector<int> ord(n);
    for (int i = 0; i < n; i++) ord[i] = i;
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return t_ms[a] < t_ms[b]; });
    double RT[9], RLIT[9];
    mat3_T(rot_end, RT);
    mat3_T(R_LI, RLIT);
    int it = n - 1;
    bool stop = false;
    for (int kp = npose - 1; kp != 0 && !stop; kp--) {
        const double* h = poses + (size_t)(kp - 1) * 22;
        for (; t_ms[ord[it]] / double(1000) > h[0]; it--) {
            float* p = xyz + 3 * (size_t)ord[it];
            const double dt = t_ms[ord[it]] / double(1000) - h[0];
            double E[9], Ri[9], Pi[3], q[3], u[3], v[3], w[3];
            so3_exp_dt(h + 4, dt, E);
            mat3_mul(h + 13, E, Ri);
            for (int a = 0; a < 3; a++) Pi[a] = h[10 + a] + h[7 + a] * dt + 0.5 * h[1 + a] * dt * dt;
            double pin[3] = {p[0], p[1], p[2]};
            mat3_mul_vec(R_LI, pin, q);
            for (int a = 0; a < 3; a++) q[a] += T_LI[a];
            mat3_mul_vec(Ri, q, u);
            for (int a = 0; a < 3; a++) u[a] = u[a] + Pi[a] - pos_end[a];
            mat3_mul_vec(RT, u, v);
            for (int a = 0; a < 3; a++) v[a] -= T_LI[a];
            mat3_mul_vec(RLIT, v, w);
            for (int a = 0; a < 3; a++) p[a] = (float)w[a];
            if (it == 0) { stop = true; break; }
        }
    }
}

// PCL VoxelGrid<PointT>::applyFilter restated for xyz (PCL >= 1.8, pcl/filters/impl/voxel_grid.hpp; third-party, pinned
// only as ">= 1.8" by the reference README:57; call site laserMapping.cpp:122,823,917-918). Output in PCL's order (ascending
// leaf index); the summation order inside a leaf is the input order (PCL's std::sort leaves it unspecified).
// Returns the number of output points, or -1 on the "leaf size too small" overflow.
int oracle_voxel_grid(const float* xyz, int n, float leaf, float* out) {
    const float inv = 1.0f / leaf;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = 0; i < n; i++) {
        const float* p = xyz + 3 * (size_t)i;
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
        for (int a = 0; a < 3; a++) {
            mn[a] = std::min(mn[a], p[a]);
            mx[a] = std::max(mx[a], p[a]);
        }
    }
    if (!(mn[0] <= mx[0])) return -1;
    // "Leaf size is too small for the input dataset. Integer indices would overflow." (PCL checks this first)
    {
        double dx = std::floor((double)((mx[0] - mn[0]) * inv)) + 1, dy = std::floor((double)((mx[1] - mn[1]) * inv)) + 1,
               dz = std::floor((double)((mx[2] - mn[2]) * inv)) + 1;
        if (dx * dy * dz > 2147483647.0) return -1;
    }
    int min_b[3], div_b[3];
    for (int a = 0; a < 3; a++) {
        min_b[a] = (int)std::floor(mn[a] * inv);
        int max_b = (int)std::floor(mx[a] * inv);
        div_b[a] = max_b - min_b[a] + 1;
    }
    if ((int64_t)div_b[0] * (int64_t)div_b[1] * (int64_t)div_b[2] > (int64_t)2147483647) return -1;
    const int m1 = div_b[0], m2 = div_b[0] * div_b[1];
    std::vector<std::pair<unsigned, int>> iv;
    iv.reserve(n);
    for (int i = 0; i < n; i++) {
        const float* p = xyz + 3 * (size_t)i;
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
        int i0 = (int)(std::floor(p[0] * inv) - (float)min_b[0]);
        int i1 = (int)(std::floor(p[1] * inv) - (float)min_b[1]);
        int i2 = (int)(std::floor(p[2] * inv) - (float)min_b[2]);
        iv.emplace_back((unsigned)(i0 + i1 * m1 + i2 * m2), i);
    }
    std::stable_sort(iv.begin(), iv.end(), [](const std::pair<unsigned, int>& a, const std::pair<unsigned, int>& b) { return a.first < b.first; });
    int k = 0;
    size_t s = 0;
    while (s < iv.size()) {
        size_t e = s;
        float cx = 0.f, cy = 0.f, cz = 0.f;
        while (e < iv.size() && iv[e].first == iv[s].first) {
            const float* p = xyz + 3 * (size_t)iv[e].second;
            cx += p[0]; cy += p[1]; cz += p[2];
            e++;
        }
        float c = (float)(e - s);
        out[3 * (size_t)k] = cx / c; out[3 * (size_t)k + 1] = cy / c; out[3 * (size_t)k + 2] = cz / c;
        k++;
        s = e;
    }
    return k;
}

// esti_plane alone (for the numpy lstsq cross-check). nb: 15 floats. Returns valid flag.
int oracle_esti_plane(const float* nb, double* pabcd) {
    P3 p[5];
    for (int j = 0; j < 5; j++) p[j] = P3{nb[3 * j], nb[3 * j + 1], nb[3 * j + 2]};
    return esti_plane(pabcd, p, 0.1) ? 1 : 0;
}

void oracle_so3_exp(const double* v, double* R) { so3_exp(v[0], v[1], v[2], R); }
void oracle_so3_log(const double* R, double* v) { so3_log(R, v); }
int oracle_mat_inverse(const double* A, double* Ainv, int n) { return mat_inverse(A, Ainv, n) ? 1 : 0; }

}  // extern "C"
