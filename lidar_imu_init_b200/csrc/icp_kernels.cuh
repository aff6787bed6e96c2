// icp_kernels.cuh -- the hot path: fused 5-NN + plane fit + residual + Jacobian row + HtH/Htr reduction.
//
// Replaces laserMapping.cpp:959-1071 (+ the dense reduction of :1080), esti_plane (common_lib.h:236-269) and
// KD_TREE::Nearest_Search (ikd_Tree.cpp:349-379,825-968). DESIGN.md section 5 has the derivations.
//
// Phase 1 (warp-cooperative, one scan point at a time): exact bounded 5-NN on the brick hash by ring expansion,
//   32 lanes scanning a brick's slab with coalesced float4 loads, top-5 kept warp-uniform in registers.
// Phase 2 (lane-parallel, one scan point per lane): fp64 column-pivoted Householder LSQ of the 5x3 system,
//   residual, gating, Jacobian row; warp reduce-scatter of the 92 (imu) / 29 (lidar-only) accumulators.
#pragma once
#include "common.cuh"

// ----------------------------------------------------------------------------------------------
// warp-uniform sorted top-5
struct Top5 {
    float d[5];
    int id[5];
    int n;
};

__device__ __forceinline__ void top5_init(Top5& t) {
#pragma unroll
    for (int i = 0; i < 5; i++) {
        t.d[i] = INFINITY;
        t.id[i] = -1;
    }
    t.n = 0;
}

// uniform insert of (dm, idm): after every entry <= dm (first come first kept on ties, as the
// reference's strict '<' replacement test, ikd_Tree.cpp:842).
__device__ __forceinline__ void top5_insert(Top5& t, float dm, int idm) {
    int p = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) p += (t.d[i] <= dm) ? 1 : 0;
#pragma unroll
    for (int i = 4; i >= 1; i--) {
        if (i > p) {
            t.d[i] = t.d[i - 1];
            t.id[i] = t.id[i - 1];
        }
    }
#pragma unroll
    for (int i = 0; i < 5; i++) {
        if (i == p) {
            t.d[i] = dm;
            t.id[i] = idm;
        }
    }
    t.n = min(t.n + 1, 5);
}

// All lanes call. valid: this lane carries candidate (dc, idc). max_d2 = 5 (laserMapping.cpp:980).
__device__ __forceinline__ void top5_consider(Top5& t, float dc, int idc, bool valid, float max_d2, int lane) {
    bool pass = valid && (dc <= max_d2) && (dc < t.d[4]);   // d[4] = +inf while fewer than 5
    unsigned m = __ballot_sync(LI_FULL, pass);
    while (m) {
        unsigned bits = pass ? __float_as_uint(dc) : 0xffffffffu;
        unsigned mn = __reduce_min_sync(LI_FULL, bits);
        unsigned who = __ballot_sync(LI_FULL, pass && bits == mn);
        int src = __ffs(who) - 1;
        int idm = __shfl_sync(LI_FULL, idc, src);
        top5_insert(t, __uint_as_float(mn), idm);
        if (lane == src) pass = false;
        pass = pass && (dc < t.d[4]);
        m = __ballot_sync(LI_FULL, pass);
    }
}

// Scan one brick slab.
__device__ __forceinline__ void knn_scan_slab(const float4* __restrict__ pool, unsigned first, unsigned count, float qx, float qy,
                                              float qz, Top5& t, int lane) {
    for (unsigned base = 0; base < count; base += 32) {
        unsigned j = base + lane;
        bool valid = j < count;
        float dc = INFINITY;
        if (valid) {
            float4 p = __ldg(&pool[(size_t)first + j]);
            dc = li_dist2(qx, qy, qz, p.x, p.y, p.z);
        }
        top5_consider(t, dc, (int)(first + j), valid, 5.0f, lane);
    }
}

// Exact 5-NN of (qx,qy,qz) within squared distance 5, warp-cooperative. Result warp-uniform in t.
__device__ __forceinline__ void knn5_warp(const MapDev& M, float qx, float qy, float qz, Top5& t, int lane) {
    top5_init(t);
    const int bs = M.bshift;
    const float ds = M.ds;
    const int bc = 1 << bs;
    if (!(isfinite(qx) && isfinite(qy) && isfinite(qz))) return;
    const float lim = (float)(LI_CELL_LIMIT - 8 * bc) * ds;
    if (fabsf(qx) >= lim || fabsf(qy) >= lim || fabsf(qz) >= lim) return;
    const int cx = li_cell(qx, ds), cy = li_cell(qy, ds), cz = li_cell(qz, ds);
    const int bx = cx >> bs, by = cy >> bs, bz = cz >> bs;
    const int half = bc >> 1;
    const int dirx = ((cx & (bc - 1)) < half) ? -1 : 1;
    const int diry = ((cy & (bc - 1)) < half) ? -1 : 1;
    const int dirz = ((cz & (bc - 1)) < half) ? -1 : 1;
    const float B = (float)bc * ds;
    // slack for float cell assignment / edge products: relative 2^-23 effects, bounded generously
    const float margin = 1e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + 16.0f * B);
    const int Rmax = (int)ceilf(2.2360680f / B) + 1;

    for (int R = 0; R <= Rmax; R++) {
        // candidate bricks of this stage
        int S, total;
        if (R == 0) {
            S = 2;
            total = 8;
        } else {
            S = 2 * R + 1;
            total = S * S * S;
        }
        for (int base = 0; base < total; base += 32) {
            int idx = base + lane;
            bool want = idx < total;
            int ox = 0, oy = 0, oz = 0;
            if (want) {
                if (R == 0) {
                    ox = (idx & 1) ? dirx : 0;
                    oy = (idx & 2) ? diry : 0;
                    oz = (idx & 4) ? dirz : 0;
                } else {
                    ox = idx % S - R;
                    oy = (idx / S) % S - R;
                    oz = idx / (S * S) - R;
                    if (R == 1) {
                        // skip the 2x2x2 half-block already visited in stage 0
                        if ((ox == 0 || ox == dirx) && (oy == 0 || oy == diry) && (oz == 0 || oz == dirz)) want = false;
                    } else {
                        if (max(abs(ox), max(abs(oy), abs(oz))) < R) want = false;   // inner cube done
                    }
                }
            }
            unsigned first = 0, count = 0;
            float dbox = INFINITY;
            bool found = false;
            if (want) {
                const int kx = bx + ox, ky = by + oy, kz = bz + oz;
                // box distance (lower bound, widened by margin)
                float lox = (float)(kx << bs) * ds - margin, hix = (float)((kx + 1) << bs) * ds + margin;
                float loy = (float)(ky << bs) * ds - margin, hiy = (float)((ky + 1) << bs) * ds + margin;
                float loz = (float)(kz << bs) * ds - margin, hiz = (float)((kz + 1) << bs) * ds + margin;
                float ex = fmaxf(0.f, fmaxf(lox - qx, qx - hix));
                float ey = fmaxf(0.f, fmaxf(loy - qy, qy - hiy));
                float ez = fmaxf(0.f, fmaxf(loz - qz, qz - hiz));
                dbox = (ex * ex + ey * ey + ez * ez) * (1.0f - 1e-6f);
                if (dbox <= 5.0f) found = li_brick_find(M.ent, M.mask, li_pack_key(kx, ky, kz), first, count);
                found = found && count > 0u;
            }
            unsigned fm = __ballot_sync(LI_FULL, found);
            while (fm) {
                int src = __ffs(fm) - 1;
                fm &= fm - 1;
                float db = __shfl_sync(LI_FULL, dbox, src);
                unsigned f = __shfl_sync(LI_FULL, first, src);
                unsigned c = __shfl_sync(LI_FULL, count, src);
                if (t.n == 5 && db >= t.d[4]) continue;   // cannot improve (strict '<' rule)
                knn_scan_slab(M.pool, f, c, qx, qy, qz, t, lane);
            }
        }
        // explored region after this stage: bricks [bx-alo, bx+ahi] per axis
        int lo_x, hi_x, lo_y, hi_y, lo_z, hi_z;
        if (R == 0) {
            lo_x = min(bx, bx + dirx); hi_x = max(bx, bx + dirx);
            lo_y = min(by, by + diry); hi_y = max(by, by + diry);
            lo_z = min(bz, bz + dirz); hi_z = max(bz, bz + dirz);
        } else {
            lo_x = bx - R; hi_x = bx + R;
            lo_y = by - R; hi_y = by + R;
            lo_z = bz - R; hi_z = bz + R;
        }
        float rx = fminf(qx - (float)(lo_x << bs) * ds, (float)((hi_x + 1) << bs) * ds - qx);
        float ry = fminf(qy - (float)(lo_y << bs) * ds, (float)((hi_y + 1) << bs) * ds - qy);
        float rz = fminf(qz - (float)(lo_z << bs) * ds, (float)((hi_z + 1) << bs) * ds - qz);
        float r = fminf(rx, fminf(ry, rz)) - margin;
        if (r > 0.f) {
            float r2 = r * r * (1.0f - 1e-6f);
            if (r2 > 5.0f) break;                      // everything within the search radius was seen
            if (t.n == 5 && t.d[4] <= r2) break;       // no unseen point can be strictly closer
        }
    }
}

// ----------------------------------------------------------------------------------------------
// Phase 2 math (per scan point, fp64)

// min ||A x - b||, A 5x3 (rows = neighbours), b = -1: column-pivoted Householder QR (the method behind
// Eigen's colPivHouseholderQr().solve, common_lib.h:252). Operation order mirrors oracle/liinit_oracle.cpp.
__device__ __forceinline__ void lsq5x3(double (&A)[5][3], double (&x)[3]) {
    double b[5] = {-1.0, -1.0, -1.0, -1.0, -1.0};
    int p0 = 0, p1 = 1, p2 = 2;
    double maxn = 0.0;
    {
        double c0 = 0, c1 = 0, c2 = 0;
#pragma unroll
        for (int i = 0; i < 5; i++) {
            c0 += A[i][0] * A[i][0];
            c1 += A[i][1] * A[i][1];
            c2 += A[i][2] * A[i][2];
        }
        maxn = fmax(c0, fmax(c1, c2));
    }
    const double eps = 2.220446049250313e-16;
    const double sq = sqrt(maxn) * eps;
    const double thresh = sq * sq / 5.0;
    int rank = 3;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (rank < 3) continue;   // uniform per lane; remaining steps skipped once rank-deficient
        // column norms of the trailing block, pick the first maximum
        double cn[3] = {-1.0, -1.0, -1.0};
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (j >= k) {
                double s = 0;
#pragma unroll
                for (int i = 0; i < 5; i++)
                    if (i >= k) s += A[i][j] * A[i][j];
                cn[j] = s;
            }
        }
        int p = k;
        double best = cn[k];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            if (j > k && cn[j] > best) {
                best = cn[j];
                p = j;
            }
        }
        if (best < thresh) {
            rank = k;
            continue;
        }
        if (p != k) {
            // swap columns k and p (static indices via unrolled selects)
#pragma unroll
            for (int j = 0; j < 3; j++) {
                if (j > k && j == p) {
#pragma unroll
                    for (int i = 0; i < 5; i++) {
                        double tmp = A[i][k];
                        A[i][k] = A[i][j];
                        A[i][j] = tmp;
                    }
                }
            }
            // permutation bookkeeping
            int pk = (k == 0) ? p0 : ((k == 1) ? p1 : p2);
            int pp = (p == 0) ? p0 : ((p == 1) ? p1 : p2);
            if (k == 0) p0 = pp; else if (k == 1) p1 = pp; else p2 = pp;
            if (p == 0) p0 = pk; else if (p == 1) p1 = pk; else p2 = pk;
        }
        double alpha = A[k][k];
        double tail = 0;
#pragma unroll
        for (int i = 0; i < 5; i++)
            if (i > k) tail += A[i][k] * A[i][k];
        if (tail != 0.0) {
            double nrm = sqrt(alpha * alpha + tail);
            double beta = (alpha >= 0) ? -nrm : nrm;
            double tau = (beta - alpha) / beta;
            double scale = 1.0 / (alpha - beta);
            double v[5];
#pragma unroll
            for (int i = 0; i < 5; i++) v[i] = (i == k) ? 1.0 : ((i > k) ? A[i][k] * scale : 0.0);
            A[k][k] = beta;
#pragma unroll
            for (int i = 0; i < 5; i++)
                if (i > k) A[i][k] = 0.0;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                if (j > k) {
                    double s = 0;
#pragma unroll
                    for (int i = 0; i < 5; i++)
                        if (i >= k) s += v[i] * A[i][j];
                    s *= tau;
#pragma unroll
                    for (int i = 0; i < 5; i++)
                        if (i >= k) A[i][j] -= s * v[i];
                }
            }
            double s = 0;
#pragma unroll
            for (int i = 0; i < 5; i++)
                if (i >= k) s += v[i] * b[i];
            s *= tau;
#pragma unroll
            for (int i = 0; i < 5; i++)
                if (i >= k) b[i] -= s * v[i];
        }
    }
    double y0 = 0, y1 = 0, y2 = 0;
    if (rank >= 3) y2 = b[2] / A[2][2];
    if (rank >= 2) y1 = (b[1] - A[1][2] * y2) / A[1][1];
    if (rank >= 1) y0 = (b[0] - A[0][1] * y1 - A[0][2] * y2) / A[0][0];
    // x[perm[j]] = y[j]
    x[0] = (p0 == 0) ? y0 : ((p1 == 0) ? y1 : y2);
    x[1] = (p0 == 1) ? y0 : ((p1 == 1) ? y1 : y2);
    x[2] = (p0 == 2) ? y0 : ((p1 == 2) ? y1 : y2);
}

// Number of accumulators: upper triangle of HtH + Htr + sum r^2 + count, padded to a multiple of 32.
template <bool IMU>
struct AccLayout {
    static constexpr int NC = IMU ? 12 : 6;
    static constexpr int NT = NC * (NC + 1) / 2;        // 78 / 21
    static constexpr int NV = NT + NC + 2;              // 92 / 29
    static constexpr int V = ((NV + 31) / 32) * 32;     // 96 / 32
    static constexpr int K = V / 32;                    // 3 / 1
};

// esti_plane + residual + gate + Jacobian row for one scan point (laserMapping.cpp:995-1010,1035-1071).
// Returns selected; fills row[NC] and r = meas = -pd2 (zeros when not selected).
template <bool IMU>
__device__ __forceinline__ bool plane_and_row(const PoseD& P, float bxf, float byf, float bzf, float wx, float wy, float wz,
                                              const float4 (&nb)[5], float4& normvec_out,
                                              double (&row)[AccLayout<IMU>::NC], double& r) {
    typedef AccLayout<IMU> L;
#pragma unroll
    for (int i = 0; i < L::NC; i++) row[i] = 0.0;
    r = 0.0;
    double A[5][3];
#pragma unroll
    for (int j = 0; j < 5; j++) {
        A[j][0] = (double)nb[j].x;
        A[j][1] = (double)nb[j].y;
        A[j][2] = (double)nb[j].z;
    }
    double nv[3];
    lsq5x3(A, nv);
    double nn = sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
    double pa = nv[0] / nn, pb = nv[1] / nn, pc = nv[2] / nn, pd = 1.0 / nn;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        double e = fabs(pa * (double)nb[j].x + pb * (double)nb[j].y + pc * (double)nb[j].z + pd);
        if (!(e <= 0.1)) ok = false;    // reference: reject if > threshold (NaN also rejects via !(<=)... see note)
    }
    // note: the reference tests `fabs(..) > threshold -> return false`; a NaN plane passes that test there and is
    // rejected later by `s > 0.9` being false. Both forms reject NaN; finite cases are identical.
    if (!ok) return false;
    float pd2 = (float)(pa * (double)wx + pb * (double)wy + pc * (double)wz + pd);
    double bx = (double)bxf, by = (double)byf, bz = (double)bzf;
    double pn = sqrt(bx * bx + by * by + bz * bz);
    float s = (float)(1.0 - 0.9 * (double)fabsf(pd2) / sqrt(pn));
    if (!((double)s > 0.9)) return false;
    float nxf = (float)pa, nyf = (float)pb, nzf = (float)pc;
    normvec_out = make_float4(nxf, nyf, nzf, pd2);
    // Jacobian row from the stored f32 normal (laserMapping.cpp:1046-1047 reads corr_normvect)
    double n0 = (double)nxf, n1 = (double)nyf, n2 = (double)nzf;
    double pI0 = P.RLI[0] * bx + P.RLI[1] * by + P.RLI[2] * bz + P.TLI[0];
    double pI1 = P.RLI[3] * bx + P.RLI[4] * by + P.RLI[5] * bz + P.TLI[1];
    double pI2 = P.RLI[6] * bx + P.RLI[7] * by + P.RLI[8] * bz + P.TLI[2];
    // C = rot_end^T n
    double C0 = P.R[0] * n0 + P.R[3] * n1 + P.R[6] * n2;
    double C1 = P.R[1] * n0 + P.R[4] * n1 + P.R[7] * n2;
    double C2 = P.R[2] * n0 + P.R[5] * n1 + P.R[8] * n2;
    // A = [pI]x C
    row[0] = pI1 * C2 - pI2 * C1;
    row[1] = pI2 * C0 - pI0 * C2;
    row[2] = pI0 * C1 - pI1 * C0;
    row[3] = n0;
    row[4] = n1;
    row[5] = n2;
    if (IMU) {
        // B = [pL]x (R_LI^T C), C as above (laserMapping.cpp:1057-1059)
        double D0 = P.RLI[0] * C0 + P.RLI[3] * C1 + P.RLI[6] * C2;
        double D1 = P.RLI[1] * C0 + P.RLI[4] * C1 + P.RLI[7] * C2;
        double D2 = P.RLI[2] * C0 + P.RLI[5] * C1 + P.RLI[8] * C2;
        row[6 % L::NC] = by * D2 - bz * D1;
        row[7 % L::NC] = bz * D0 - bx * D2;
        row[8 % L::NC] = bx * D1 - by * D0;
        row[9 % L::NC] = C0;
        row[10 % L::NC] = C1;
        row[11 % L::NC] = C2;
    }
    r = -(double)pd2;   // meas_vec (laserMapping.cpp:1070)
    return true;
}

// Warp reduce-scatter: on entry every lane holds V values; on exit lane l holds the warp sums of
// indices [K*l, K*l+K) in v[0..K). V(1-1/32) shuffles instead of 5V.
template <int N, int S>
struct RS {
    template <int V>
    static __device__ __forceinline__ void run(double (&v)[V], int lane) {
        const bool upper = (lane & S) != 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            double keep = upper ? v[i + N] : v[i];
            double send = upper ? v[i] : v[i + N];
            v[i] = keep + __shfl_xor_sync(LI_FULL, send, S);
        }
        RS<N / 2, S / 2>::run(v, lane);
    }
};
template <int N>
struct RS<N, 0> {
    template <int V>
    static __device__ __forceinline__ void run(double (&)[V], int) {}
};

// Add one scan point per lane into the warp's accumulators. The V values (upper triangle of row^T row, row*r,
// r^2, 1) are formed and reduce-scattered 32 at a time so that at most 32 doubles are live: afterwards lane l
// holds in acc[g] the running warp sum of value index 32*g + l.
template <bool IMU>
__device__ __forceinline__ void warp_accumulate(const double (&row)[AccLayout<IMU>::NC], double r, bool sel, int lane,
                                                double (&acc)[AccLayout<IMU>::K]) {
    typedef AccLayout<IMU> L;
#pragma unroll
    for (int g = 0; g < L::K; g++) {
        double v[32];
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = 0.0;
        int idx = 0;
#pragma unroll
        for (int a = 0; a < L::NC; a++) {
#pragma unroll
            for (int b = a; b < L::NC; b++) {
                if (idx / 32 == g) v[idx % 32] = row[a] * row[b];
                idx++;
            }
        }
#pragma unroll
        for (int a = 0; a < L::NC; a++) {
            if (idx / 32 == g) v[idx % 32] = row[a] * r;
            idx++;
        }
        if (idx / 32 == g) v[idx % 32] = r * r;
        idx++;
        if (idx / 32 == g) v[idx % 32] = sel ? 1.0 : 0.0;
        RS<16, 16>::run(v, lane);
        acc[g] += v[0];
    }
}

// Block-level accumulation + deterministic grid reduction.
//   acc[K]: this lane's running sums (value index 32*k + lane). partials: [gridDim.x][V]. out160: final layout
//   [HtH 144 | Htr 12 | res_sq | m | 0 0].
template <bool IMU>
__device__ __forceinline__ void block_finish(const double (&acc)[AccLayout<IMU>::K], double* __restrict__ partials,
                                             unsigned* __restrict__ done_counter, double* __restrict__ out160) {
    typedef AccLayout<IMU> L;
    __shared__ double spart[8][L::V];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < L::K; k++) spart[warp][32 * k + lane] = acc[k];
    __syncthreads();
    if (threadIdx.x < L::V) {
        double s = 0;
        for (int w = 0; w < nwarps; w++) s += spart[w][threadIdx.x];
        partials[(size_t)blockIdx.x * L::V + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned ticket = atomicAdd(done_counter, 1u);
        s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last block: fixed-order sum over blocks -> bit-reproducible result for a given grid size
    for (int o = threadIdx.x; o < 160; o += blockDim.x) {
        int src = -1;
        if (o < 144) {
            int a = o / 12, b = o % 12;
            if (a > b) { int tt = a; a = b; b = tt; }
            if (b < L::NC) src = a * L::NC - a * (a - 1) / 2 + (b - a);
        } else if (o < 156) {
            int a = o - 144;
            if (a < L::NC) src = L::NT + a;
        } else if (o == 156) {
            src = L::NT + L::NC;
        } else if (o == 157) {
            src = L::NT + L::NC + 1;
        }
        double s = 0;
        if (src >= 0) {
            const volatile double* pp = partials;
            for (unsigned b = 0; b < gridDim.x; b++) s += pp[(size_t)b * L::V + src];
        }
        out160[o] = s;
    }
    if (threadIdx.x == 0) *done_counter = 0u;
}

struct ScanDev {
    const float4* body;     // feats_down_body (xyz, w unused)
    float4* world;          // feats_down_world
    int* near_ids;          // [N*5] pool offsets, -1 = missing (Nearest_Points)
    unsigned char* selected;  // point_selected_surf
    float4* normvec;        // (nx,ny,nz,pd2) f32
    int n;
};

// Order the 5 neighbours as PointType_CMP does (ikd_Tree.h:57-60): ascending distance, distances closer than
// 1e-10 are ties broken by smaller x. The search already delivers ascending distances, only ties can move.
__device__ __forceinline__ void tie_order(float4 (&nb)[5], int (&id)[5], float (&d)[5]) {
#pragma unroll
    for (int pass = 0; pass < 4; pass++) {
#pragma unroll
        for (int k = 0; k < 4 - pass; k++) {
            if (fabsf(d[k] - d[k + 1]) < 1e-10f && nb[k + 1].x < nb[k].x) {
                float4 tq = nb[k]; nb[k] = nb[k + 1]; nb[k + 1] = tq;
                int ti = id[k]; id[k] = id[k + 1]; id[k + 1] = ti;
                float td = d[k]; d[k] = d[k + 1]; d[k + 1] = td;
            }
        }
    }
}

// ---- fused search pass -------------------------------------------------------------------------
// One warp owns a tile of TILE consecutive scan points: transform (lane-parallel), 5-NN (warp-cooperative,
// point after point), then plane/residual/Jacobian (lane-parallel) and the reduction.
template <int TILE, bool IMU>
__global__ void __launch_bounds__(256) k_icp_search(MapDev M, ScanDev S, PoseD P, double* __restrict__ partials,
                                                    unsigned* __restrict__ done_counter, double* __restrict__ out160) {
    typedef AccLayout<IMU> L;
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps_total = (gridDim.x * blockDim.x) >> 5;
    const int ntiles = (S.n + TILE - 1) / TILE;
    double acc[L::K];
#pragma unroll
    for (int k = 0; k < L::K; k++) acc[k] = 0.0;

    for (int tile = warp_global; tile < ntiles; tile += nwarps_total) {
        const int q = tile * TILE + lane;
        const bool active = (lane < TILE) && (q < S.n);
        float bx = 0, by = 0, bz = 0, wx = 0, wy = 0, wz = 0;
        if (active) {
            float4 b = __ldg(&S.body[q]);
            bx = b.x; by = b.y; bz = b.z;
            li_body_to_world(P, bx, by, bz, wx, wy, wz);
            S.world[q] = make_float4(wx, wy, wz, 0.f);
        }
        int my_id[5] = {-1, -1, -1, -1, -1};
        float my_d[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        const int in_tile = min(TILE, S.n - tile * TILE);
        for (int j = 0; j < in_tile; j++) {
            float qx = __shfl_sync(LI_FULL, wx, j), qy = __shfl_sync(LI_FULL, wy, j), qz = __shfl_sync(LI_FULL, wz, j);
            Top5 t;
            knn5_warp(M, qx, qy, qz, t, lane);
            if (lane == j) {
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    my_id[k] = t.id[k];
                    my_d[k] = t.d[k];
                }
            }
        }
        double row[L::NC];
        double r = 0.0;
        bool sel = false;
#pragma unroll
        for (int i = 0; i < L::NC; i++) row[i] = 0.0;
        if (active) {
            // selection gate (laserMapping.cpp:981-984): 5 found; d2[4] <= 5 holds by construction
            if (my_id[4] >= 0) {
                float4 nb[5];
#pragma unroll
                for (int k = 0; k < 5; k++) nb[k] = __ldg(&M.pool[my_id[k]]);
                tie_order(nb, my_id, my_d);
                float4 nvec;
                sel = plane_and_row<IMU>(P, bx, by, bz, wx, wy, wz, nb, nvec, row, r);
                if (sel) S.normvec[q] = nvec;
            }
#pragma unroll
            for (int k = 0; k < 5; k++) S.near_ids[(size_t)q * 5 + k] = my_id[k];
            S.selected[q] = sel ? 1 : 0;
        }
        warp_accumulate<IMU>(row, r, sel, lane, acc);
    }
    block_finish<IMU>(acc, partials, done_counter, out160);
}

// ---- reuse pass (nearest_search_en == false, laserMapping.cpp:989-994) ------------------------------
// One scan point per lane: reuses the stored neighbour ids and the previous selection flag.
template <bool IMU>
__global__ void __launch_bounds__(256) k_icp_reuse(MapDev M, ScanDev S, PoseD P, double* __restrict__ partials,
                                                   unsigned* __restrict__ done_counter, double* __restrict__ out160) {
    typedef AccLayout<IMU> L;
    const int lane = threadIdx.x & 31;
    double acc[L::K];
#pragma unroll
    for (int k = 0; k < L::K; k++) acc[k] = 0.0;
    const int stride = gridDim.x * blockDim.x;
    const int nround = (S.n + stride - 1) / stride;
    for (int it = 0; it < nround; it++) {
        const int q = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
        double row[L::NC];
        double r = 0.0;
        bool sel = false;
#pragma unroll
        for (int i = 0; i < L::NC; i++) row[i] = 0.0;
        if (q < S.n) {
            float4 b = __ldg(&S.body[q]);
            float wx, wy, wz;
            li_body_to_world(P, b.x, b.y, b.z, wx, wy, wz);
            S.world[q] = make_float4(wx, wy, wz, 0.f);
            if (S.selected[q]) {
                int id4 = S.near_ids[(size_t)q * 5 + 4];
                if (id4 >= 0) {
                    float4 nb[5];
#pragma unroll
                    for (int k = 0; k < 5; k++) nb[k] = __ldg(&M.pool[S.near_ids[(size_t)q * 5 + k]]);
                    float4 nvec;
                    sel = plane_and_row<IMU>(P, b.x, b.y, b.z, wx, wy, wz, nb, nvec, row, r);
                    if (sel) S.normvec[q] = nvec;
                }
            }
            S.selected[q] = sel ? 1 : 0;
        }
        warp_accumulate<IMU>(row, r, sel, lane, acc);
    }
    block_finish<IMU>(acc, partials, done_counter, out160);
}

// ---- stand-alone Nearest_Search for arbitrary queries (tests / KD_TREE::Nearest_Search drop-in) ------
__global__ void __launch_bounds__(256) k_knn_queries(MapDev M, const float4* __restrict__ q, int n, int* __restrict__ ids,
                                                     float* __restrict__ d2) {
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp_global; i < n; i += nw) {
        float4 p = __ldg(&q[i]);
        Top5 t;
        knn5_warp(M, p.x, p.y, p.z, t, lane);
        if (lane < 5) {
            int id = (lane == 0) ? t.id[0] : (lane == 1) ? t.id[1] : (lane == 2) ? t.id[2] : (lane == 3) ? t.id[3] : t.id[4];
            float d = (lane == 0) ? t.d[0] : (lane == 1) ? t.d[1] : (lane == 2) ? t.d[2] : (lane == 3) ? t.d[3] : t.d[4];
            ids[(size_t)i * 5 + lane] = id;
            d2[(size_t)i * 5 + lane] = (id >= 0) ? d : -1.f;
        }
    }
}
