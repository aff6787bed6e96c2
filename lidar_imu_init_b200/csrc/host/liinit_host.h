/* liinit_host.h -- host side of the drop-in: the part of laserMapping.cpp:936-1134 that the north star keeps on
 * the CPU (IESKF solve, state update, convergence / rematch policy, covariance update), re-expressed over the
 * C-ABI of include/liinit_gpu.h so that only HtH (12x12), Htr (12) and m cross the PCIe bus per iteration
 * (SURVEY.md 8a row A8). Plain C++17, no Eigen/PCL/ROS (none is installed here); see INTEGRATION.md for the
 * Eigen-typed call sites in the node.
 */
#ifndef LIINIT_HOST_H
#define LIINIT_HOST_H

#include "liinit_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define LIINIT_DIM_STATE 24 /* include/common_lib.h:24 */

/* StatesGroup (include/common_lib.h:68-169), flat, row-major. State vector order for boxplus / boxminus:
 * [rot 0:3 | pos 3:6 | R_LI 6:9 | T_LI 9:12 | vel 12:15 | bg 15:18 | ba 18:21 | g 21:24]. */
typedef struct liinit_state {
    double rot_end[9];
    double pos_end[3];
    double offset_R_L_I[9];
    double offset_T_L_I[3];
    double vel_end[3];
    double bias_g[3];
    double bias_a[3];
    double gravity[3];
    double cov[LIINIT_DIM_STATE * LIINIT_DIM_STATE];
} liinit_state;

typedef struct liinit_scan_stats {
    int iterations;        /* passes run (<= max_iteration) */
    int search_passes;     /* of which nearest_search_en was true */
    int effect_feat_num;   /* m of the last pass */
    int converged;         /* flg_EKF_converged of the last pass */
    double last_rot_deg;   /* deltaR, laserMapping.cpp:1096 */
    double last_trans_cm;  /* deltaT, :1097 */
    double res_sq;         /* sum of pd2^2 over the selected points of the last pass */
} liinit_scan_stats;

/* StatesGroup() defaults (common_lib.h:70-81): identity rotations, zero vectors, cov = I with [15:24] = 1e-5. */
void liinit_state_init(liinit_state* s);
/* operator+= (common_lib.h:124-135) and operator- (:137-151). */
void liinit_state_boxplus(liinit_state* s, const double delta[LIINIT_DIM_STATE]);
void liinit_state_boxminus(const liinit_state* a, const liinit_state* b, double out[LIINIT_DIM_STATE]);
/* Exp(v1,v2,v3) (so3_math.h:61-80) and Log(R) (:100-107). */
void liinit_so3_exp(const double v[3], double R[9]);
void liinit_so3_log(const double R[9], double v[3]);

/* One IESKF update from the device accumulators: the information-form equivalent of laserMapping.cpp:1080-1087.
 *   K_1 = (blkdiag(1000*HtH, 0) + cov^-1)^-1 ; solution = K_1[:,0:12]*(1000*Htr) + vec - K_1[:,0:12]*(1000*HtH)*vec[0:12],
 *   vec = state_propagat (-) state ; state (+)= solution.  KH (24x12, may be NULL) = K_1[:,0:12]*(1000*HtH) = K*Hsub (:1113). */
int liinit_ieskf_update(liinit_state* state, const liinit_state* state_propagat, const double HtH[144], const double Htr[12],
                        double solution[LIINIT_DIM_STATE], double* KH);

/* The whole per-scan update, laserMapping.cpp:936-1134, on an uploaded scan (liinit_scan_upload): iterates
 * liinit_icp_iterate + liinit_ieskf_update with the reference's rematch / convergence policy and covariance update.
 * state: in = propagated prior (state_propagat = state, :910), out = posterior. */
int liinit_scan_update(liinit_ctx* h, liinit_state* state, int max_iteration, int imu_en, liinit_scan_stats* stats);

/* State propagation of the LiDAR-only (constant-velocity) mode, ImuProcess::Forward_propagation_without_imu without its
 * point loop (src/IMU_Processing.hpp:212-243; the loop :246-266 is liinit_raw_undistort_cv): bias_g stands for the
 * angular velocity, vel_end for the linear velocity. cov <- F cov F^T + Q with F[0:3,0:3] = Exp(-bias_g dt),
 * F[0:3,15:18] = F[3:6,12:15] = I dt, Q[15:18] = cov_gyr_scale dt^2, Q[12:15] = cov_acc_scale dt^2; then
 * rot_end <- rot_end Exp(bias_g dt), pos_end += vel_end dt. dt: scan-to-scan time (0.1 for the first frame, :217-223). */
void liinit_propagate_cv(liinit_state* s, double dt, const double cov_gyr_scale[3], const double cov_acc_scale[3]);

/* lasermap_fov_segment (laserMapping.cpp:260-305): keeps a cube_len-sized local map box around the LiDAR and, when the
 * LiDAR comes within MOV_THRESHOLD*det_range of a face, shifts the box and returns the slabs that fell out of it
 * (cub_needrm) -- the boxes to hand to liinit_map_delete_boxes. local_box: {min xyz, max xyz}, state kept by the caller;
 * *initialized: Localmap_Initialized. boxes_out: room for 3 boxes (18 floats). Returns the number of boxes (0..3). */
int liinit_fov_segment(const double pos_LiD[3], double cube_len, double det_range, float local_box[6], int* initialized,
                       float boxes_out[18]);

#ifdef __cplusplus
}
#endif
#endif
