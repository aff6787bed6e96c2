// liinit_host.cpp -- host side of the drop-in (see liinit_host.h). Everything here is O(24^3) per iteration.
#include "liinit_host.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace {

constexpr int D = LIINIT_DIM_STATE;

inline void mul33(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
inline void mulT33(const double* A, const double* B, double* C) {   // A^T * B
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}

// Dense inverse by LU with partial pivoting (what Eigen's .inverse() does for a 24x24, laserMapping.cpp:1081).
// ncols < n: only the first ncols columns of the inverse are solved for (the others of X are left untouched) -- the gain K_1 enters
// the update through its first 12 columns only; a column of the inverse does not depend on which other columns are computed.
bool invert(const double* A, double* X, int n, int ncols = -1) {
    if (ncols < 0) ncols = n;
    double lu[D * D];
    int perm[D];
    std::memcpy(lu, A, sizeof(double) * n * n);
    for (int i = 0; i < n; i++) perm[i] = i;
    for (int k = 0; k < n; k++) {
        int piv = k;
        double best = std::fabs(lu[k * n + k]);
        for (int i = k + 1; i < n; i++)
            if (std::fabs(lu[i * n + k]) > best) {
                best = std::fabs(lu[i * n + k]);
                piv = i;
            }
        if (best == 0.0) return false;
        if (piv != k) {
            for (int j = 0; j < n; j++) {
                double t = lu[k * n + j];
                lu[k * n + j] = lu[piv * n + j];
                lu[piv * n + j] = t;
            }
            int t = perm[k];
            perm[k] = perm[piv];
            perm[piv] = t;
        }
        const double inv = 1.0 / lu[k * n + k];
        for (int i = k + 1; i < n; i++) {
            const double f = lu[i * n + k] * inv;
            lu[i * n + k] = f;
            for (int j = k + 1; j < n; j++) lu[i * n + j] -= f * lu[k * n + j];
        }
    }
    for (int c = 0; c < ncols; c++) {
        double y[D];
        for (int i = 0; i < n; i++) {
            double s = (perm[i] == c) ? 1.0 : 0.0;
            for (int j = 0; j < i; j++) s -= lu[i * n + j] * y[j];
            y[i] = s;
        }
        for (int i = n - 1; i >= 0; i--) {
            double s = y[i];
            for (int j = i + 1; j < n; j++) s -= lu[i * n + j] * X[j * n + c];
            X[i * n + c] = s / lu[i * n + i];
        }
    }
    return true;
}

}  // namespace

extern "C" {

void liinit_so3_exp(const double v[3], double R[9]) {
    const double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    static const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (n > 0.00001) {   // so3_math.h:66
        const double r[3] = {v[0] / n, v[1] / n, v[2] / n};
        const double K[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
        double KK[9];
        mul33(K, K, KK);
        const double s = std::sin(n), c = 1.0 - std::cos(n);
        for (int i = 0; i < 9; i++) R[i] = I[i] + s * K[i] + c * KK[i];
    } else {
        std::memcpy(R, I, sizeof(I));
    }
}

void liinit_so3_log(const double R[9], double v[3]) {
    const double tr = R[0] + R[4] + R[8];
    const double theta = (tr > 3.0 - 1e-6) ? 0.0 : std::acos(0.5 * (tr - 1));
    const double f = (std::fabs(theta) < 0.001) ? 0.5 : (0.5 * theta / std::sin(theta));
    v[0] = f * (R[7] - R[5]);
    v[1] = f * (R[2] - R[6]);
    v[2] = f * (R[3] - R[1]);
}

void liinit_state_init(liinit_state* s) {
    std::memset(s, 0, sizeof(*s));
    s->rot_end[0] = s->rot_end[4] = s->rot_end[8] = 1.0;
    s->offset_R_L_I[0] = s->offset_R_L_I[4] = s->offset_R_L_I[8] = 1.0;
    for (int i = 0; i < D; i++) s->cov[i * D + i] = (i >= 15) ? 0.00001 : 1.0;   // INIT_COV, common_lib.h:79-80
}

void liinit_state_boxplus(liinit_state* s, const double* d) {
    double E[9], R[9];
    liinit_so3_exp(d, E);
    mul33(s->rot_end, E, R);
    std::memcpy(s->rot_end, R, sizeof(R));
    liinit_so3_exp(d + 6, E);
    mul33(s->offset_R_L_I, E, R);
    std::memcpy(s->offset_R_L_I, R, sizeof(R));
    for (int i = 0; i < 3; i++) {
        s->pos_end[i] += d[3 + i];
        s->offset_T_L_I[i] += d[9 + i];
        s->vel_end[i] += d[12 + i];
        s->bias_g[i] += d[15 + i];
        s->bias_a[i] += d[18 + i];
        s->gravity[i] += d[21 + i];
    }
}

void liinit_state_boxminus(const liinit_state* a, const liinit_state* b, double* o) {
    double rd[9];
    mulT33(b->rot_end, a->rot_end, rd);
    liinit_so3_log(rd, o);
    mulT33(b->offset_R_L_I, a->offset_R_L_I, rd);
    liinit_so3_log(rd, o + 6);
    for (int i = 0; i < 3; i++) {
        o[3 + i] = a->pos_end[i] - b->pos_end[i];
        o[9 + i] = a->offset_T_L_I[i] - b->offset_T_L_I[i];
        o[12 + i] = a->vel_end[i] - b->vel_end[i];
        o[15 + i] = a->bias_g[i] - b->bias_g[i];
        o[18 + i] = a->bias_a[i] - b->bias_a[i];
        o[21 + i] = a->gravity[i] - b->gravity[i];
    }
}

// cov_inv: state.cov^-1 when the caller already has it (the covariance does not change between the iterations of a scan,
// laserMapping.cpp:1081 vs :1127: liinit_scan_update inverts it once), else nullptr.
static int ieskf_update(liinit_state* st, const liinit_state* prop, const double* HtH, const double* Htr, double* sol, double* KH,
                        const double* cov_inv) {
    double S[D * D], K1[D * D];
    if (cov_inv) std::memcpy(S, cov_inv, sizeof(S));
    else if (!invert(st->cov, S, D)) return LIINIT_ERR_INVALID;      // cov^-1
    for (int a = 0; a < 12; a++)
        for (int b = 0; b < 12; b++) S[a * D + b] += 1000.0 * HtH[a * 12 + b];   // R_inv = 1000 (laserMapping.cpp:1050)
    if (!invert(S, K1, D, 12)) return LIINIT_ERR_INVALID;            // (columns 0..11 of K_1 are all the update reads)
    double vec[D];
    liinit_state_boxminus(prop, st, vec);
    double G[D * 12];   // K*Hsub
    for (int r = 0; r < D; r++)
        for (int c = 0; c < 12; c++) {
            double s = 0;
            for (int a = 0; a < 12; a++) s += K1[r * D + a] * (1000.0 * HtH[a * 12 + c]);
            G[r * 12 + c] = s;
        }
    for (int r = 0; r < D; r++) {
        double s = 0, t = 0;
        for (int a = 0; a < 12; a++) {
            s += K1[r * D + a] * (1000.0 * Htr[a]);
            t += G[r * 12 + a] * vec[a];
        }
        sol[r] = s + vec[r] - t;
    }
    liinit_state_boxplus(st, sol);
    if (KH) std::memcpy(KH, G, sizeof(G));
    return LIINIT_OK;
}

int liinit_ieskf_update(liinit_state* st, const liinit_state* prop, const double* HtH, const double* Htr, double* sol, double* KH) {
    return ieskf_update(st, prop, HtH, Htr, sol, KH, nullptr);
}

int liinit_scan_update(liinit_ctx* h, liinit_state* state, int max_iter, int imu_en, liinit_scan_stats* stats) {
    if (!h || !state || max_iter < 1) return LIINIT_ERR_INVALID;
    liinit_state prop = *state;            // state_propagat = state (laserMapping.cpp:910)
    int rematch_num = 0, nearest_search_en = 1;
    liinit_scan_stats st{};
    double HtH[144], Htr[12], sol[D], KH[D * 12];
    double cov_inv[D * D];                 // state.cov is only replaced at the end of the scan (:1127): one inversion for all iterations
    if (!invert(state->cov, cov_inv, D)) return LIINIT_ERR_INVALID;
    for (int it = 0; it < max_iter; it++) {
        int m = 0;
        double rs = 0;
        st.search_passes += nearest_search_en;
        int rc = liinit_icp_iterate(h, state->rot_end, state->pos_end, state->offset_R_L_I, state->offset_T_L_I, imu_en,
                                    nearest_search_en, HtH, Htr, &m, &rs);
        if (rc != LIINIT_OK) return rc;
        rc = ieskf_update(state, &prop, HtH, Htr, sol, KH, cov_inv);
        if (rc != LIINIT_OK) return rc;
        st.iterations++;
        st.effect_feat_num = m;
        st.res_sq = rs;
        const double rn = std::sqrt(sol[0] * sol[0] + sol[1] * sol[1] + sol[2] * sol[2]);
        const double tn = std::sqrt(sol[3] * sol[3] + sol[4] * sol[4] + sol[5] * sol[5]);
        st.last_rot_deg = rn * 57.3;
        st.last_trans_cm = tn * 100;
        const bool converged = (rn * 57.3 < 0.01) && (tn * 100 < 0.015);       // :1093-1094
        st.converged = converged;
        nearest_search_en = 0;                                                 // :1102-1106
        if (converged || (rematch_num == 0 && it == max_iter - 2)) {
            nearest_search_en = 1;
            rematch_num++;
        }
        if (rematch_num >= 2 || it == max_iter - 1) {                          // :1109-1133
            double nc[D * D];
            for (int r = 0; r < D; r++)
                for (int c = 0; c < D; c++) {
                    double s = state->cov[r * D + c];
                    for (int a = 0; a < 12; a++) s -= KH[r * 12 + a] * state->cov[a * D + c];
                    nc[r * D + c] = s;
                }
            std::memcpy(state->cov, nc, sizeof(nc));
            break;
        }
    }
    if (stats) *stats = st;
    return LIINIT_OK;
}

// Exp(ang_vel, dt) (so3_math.h:39-59): gated on the angular VELOCITY (|w| > 1e-7), not on the angle like Exp(v1, v2, v3) (:61-80,
// |v| > 1e-5) -- for 1e-7 < |w| < 1e-5 / dt the two differ (identity vs a small rotation). The un-distortion kernel uses the same
// form (li_so3_exp_dt, undistort_kernels.cuh).
static void so3_exp_dt(const double w[3], double dt, double R[9]) {
    const double n = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (!(n > 0.0000001)) return;
    const double a[3] = {w[0] / n, w[1] / n, w[2] / n};
    const double K[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
    double KK[9];
    mul33(K, K, KK);
    const double ang = n * dt, sn = std::sin(ang), cs = 1.0 - std::cos(ang);
    for (int i = 0; i < 9; i++) R[i] += sn * K[i] + cs * KK[i];
}

void liinit_propagate_cv(liinit_state* s, double dt, const double* cov_gyr_scale, const double* cov_acc_scale) {
    // F is the identity except three 3x3 blocks, so F cov F^T touches rows / columns 0:6 only:
    //   rows 0:3 <- E rows 0:3 + dt rows 15:18 ; rows 3:6 <- rows 3:6 + dt rows 12:15 ; then the same on the columns.
    double E[9];
    so3_exp_dt(s->bias_g, -dt, E);   // Exp(state_inout.bias_g, -dt) (IMU_Processing.hpp:226)
    std::vector<double> T(D * D);
    double* P = s->cov;
    for (int c = 0; c < D; c++) {
        for (int r = 0; r < D; r++) T[r * D + c] = P[r * D + c];
        for (int r = 0; r < 3; r++)
            T[r * D + c] = E[3 * r] * P[0 * D + c] + E[3 * r + 1] * P[1 * D + c] + E[3 * r + 2] * P[2 * D + c] + dt * P[(15 + r) * D + c];
        for (int r = 0; r < 3; r++) T[(3 + r) * D + c] = P[(3 + r) * D + c] + dt * P[(12 + r) * D + c];
    }
    for (int r = 0; r < D; r++) {
        for (int c = 0; c < D; c++) P[r * D + c] = T[r * D + c];
        for (int c = 0; c < 3; c++)
            P[r * D + c] = T[r * D + 0] * E[3 * c] + T[r * D + 1] * E[3 * c + 1] + T[r * D + 2] * E[3 * c + 2] + dt * T[r * D + 15 + c];
        for (int c = 0; c < 3; c++) P[r * D + 3 + c] = T[r * D + 3 + c] + dt * T[r * D + 12 + c];
    }
    for (int i = 0; i < 3; i++) {
        P[(15 + i) * D + 15 + i] += cov_gyr_scale[i] * dt * dt;
        P[(12 + i) * D + 12 + i] += cov_acc_scale[i] * dt * dt;
    }
    double Ef[9], R[9];
    so3_exp_dt(s->bias_g, dt, Ef);   // Exp(state_inout.bias_g, dt) (IMU_Processing.hpp:239)
    mul33(s->rot_end, Ef, R);
    std::memcpy(s->rot_end, R, sizeof(R));
    for (int i = 0; i < 3; i++) s->pos_end[i] += s->vel_end[i] * dt;
}

int liinit_fov_segment(const double* pos, double cube_len, double det_range, float* box, int* initialized, float* out) {
    const double MOV_THRESHOLD = 1.5;   // laserMapping.cpp:58
    if (!*initialized) {                 // :266-272
        for (int i = 0; i < 3; i++) {
            box[i] = (float)(pos[i] - cube_len / 2.0);
            box[3 + i] = (float)(pos[i] + cube_len / 2.0);
        }
        *initialized = 1;
        return 0;
    }
    float dist[3][2];
    bool need_move = false;
    for (int i = 0; i < 3; i++) {        // :274-281
        dist[i][0] = (float)std::fabs(pos[i] - (double)box[i]);
        dist[i][1] = (float)std::fabs(pos[i] - (double)box[3 + i]);
        if (dist[i][0] <= MOV_THRESHOLD * det_range || dist[i][1] <= MOV_THRESHOLD * det_range) need_move = true;
    }
    if (!need_move) return 0;
    float nb[6];
    std::memcpy(nb, box, sizeof(nb));
    const double a = (cube_len - 2.0 * MOV_THRESHOLD * det_range) * 0.5 * 0.9, b = det_range * (MOV_THRESHOLD - 1);
    const float mov_dist = (float)(a > b ? a : b);   // :285-286
    int n = 0;
    for (int i = 0; i < 3; i++) {        // :287-301
        float tmp[6];
        std::memcpy(tmp, box, sizeof(tmp));
        if (dist[i][0] <= MOV_THRESHOLD * det_range) {
            nb[3 + i] -= mov_dist;
            nb[i] -= mov_dist;
            tmp[i] = box[3 + i] - mov_dist;
            std::memcpy(out + 6 * n++, tmp, sizeof(tmp));
        } else if (dist[i][1] <= MOV_THRESHOLD * det_range) {
            nb[3 + i] += mov_dist;
            nb[i] += mov_dist;
            tmp[3 + i] = box[i] + mov_dist;
            std::memcpy(out + 6 * n++, tmp, sizeof(tmp));
        }
    }
    std::memcpy(box, nb, sizeof(nb));
    return n;
}

}  // extern "C"
