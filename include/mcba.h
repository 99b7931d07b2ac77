/* mcba.h — C-ABI of the B200 bundle-adjustment engine (libmcba.so).
 *
 * Drop-in boundary for ONE numeric seam of the reference (multical):
 *     scipy.optimize.least_squares(evaluate, self.param_vec, jac_sparsity=..., x_scale='jac',
 *                                  f_scale=..., ftol=..., max_nfev=..., method='trf', loss=...)
 *     -- multical/optimization/calibration.py:209-210 (inside Calibration.bundle_adjust, 199-212)
 * plus the residual closure it drives (calibration.py:204-206) and the reprojection-error /
 * outlier statistics that bracket every call (calibration.py:134-141, 240-252; tables.py:244-249).
 *
 * Plain C types only: opaque handle, host pointers + sizes, integer status (0 = ok; message via
 * mcba_last_error).  All host arrays are borrowed for the duration of the call and copied.
 * There is no CPU fallback: every entry point fails with MCBA_ERR_CUDA if no sm_100 device works.
 *
 * Index contract (bit-exact with the reference): packed corner k corresponds to row k of
 * np.argwhere(calib.inliers) = (camera, frame, board, point) in row-major order of the dense
 * [C,F,B,P] table (calibration.py:206 boolean-mask order); residual vector element 2k is u, 2k+1 is v.
 * Parameter vector layout = reference order (calibration.py:146-153, parameters.py:104-106):
 *   [camera_poses 6C][board_poses 6B][motion 6F | 12F rolling | 12 hand-eye][cameras (5+nd)C][boards 3*B*P]   (enabled blocks only; the boards
 *   block is the padded stack of tables.stack_boards -- the host strips the padding of boards with fewer points)
 *   pose = [rx ry rz tx ty tz] (transform/rtvec.py:16-27), camera = [fx fy cx cy skew dist...]
 *   (camera.py:144-155).
 */
#ifndef MCBA_H
#define MCBA_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcba_ctx mcba_ctx;

enum { MCBA_OK = 0, MCBA_ERR_ARG = 1, MCBA_ERR_CUDA = 2, MCBA_ERR_STATE = 3, MCBA_ERR_NCCL = 4,
       MCBA_ERR_NONFINITE = 5, MCBA_ERR_UNSUPPORTED = 6 };

/* camera models: camera.py:43-48 (cv2.projectPoints, 5/8/12/14 coefficients) and
 * camera_fisheye.py:113-117 (cv2.fisheye.projectPoints, 4 coefficients) */
enum { MCBA_MODEL_STANDARD = 0, MCBA_MODEL_RATIONAL = 1, MCBA_MODEL_THIN_PRISM = 2, MCBA_MODEL_FISHEYE = 3,
       MCBA_MODEL_TILTED = 4 /* 14 coefficients: thin prism + sensor tilt (tauX, tauY) */ };

/* loss names of scipy.optimize.least_squares (config/arguments.py:59) */
enum { MCBA_LOSS_LINEAR = 0, MCBA_LOSS_SOFT_L1 = 1, MCBA_LOSS_HUBER = 2, MCBA_LOSS_CAUCHY = 3, MCBA_LOSS_ARCTAN = 4 };

/* which parameter blocks are free: Calibration.optimize (calibration.py:28-35,155-161) */
enum { MCBA_OPT_CAMERA_POSES = 1, MCBA_OPT_BOARD_POSES = 2, MCBA_OPT_MOTION = 4, MCBA_OPT_CAMERAS = 8,
       MCBA_OPT_BOARDS = 16 /* board points as parameters (board/charuco.py:112-117); block layout = padded [B][P][3] */,
       MCBA_OPT_FIX_ASPECT = 256 /* camera.py:147-148,159-160 */ };
/* motion model = the `motion` argument of Calibration (calibration.py:44-46), or-ed into mcba_problem_desc.optimize:
 *   (none)                StaticFrames: one rig pose per frame (motion/static_frames.py:29-42)
 *   MCBA_MOTION_ROLLING   RollingFrames: a start and an end pose per frame; a corner is transformed by both and the two
 *                         camera-frame points are blended by its observed row / image height (motion/rolling_frames.py:15-41,
 *                         66-150).  motion block of the parameter vector = [start F x 6 | end F x 6] (135-140).
 *   MCBA_MOTION_HAND_EYE  HandEye: frame pose f = gripper_wrt_camera @ base_wrt_gripper[f] @ world_wrt_base with the arm
 *                         poses fixed (motion/hand_eye.py:14-90).  motion block = [world_wrt_base 6 | gripper_wrt_camera 6]
 *                         (76-81), shared by every residual (89-90).
 */
enum { MCBA_MOTION_ROLLING = 1 << 16, MCBA_MOTION_HAND_EYE = 1 << 17 };

typedef struct {
  int32_t C, F, B, P;       /* Calibration.size (calibration.py:64-67); F = frames held by THIS rank   */
  int32_t model;            /* MCBA_MODEL_*                                                            */
  int32_t optimize;         /* bitmask MCBA_OPT_*                                                      */
  int64_t N;                /* packed inlier corners held by this rank                                 */
} mcba_problem_desc;

typedef struct {
  double ftol, xtol, gtol;  /* scipy defaults 1e-8; the reference passes ftol=tolerance (1e-4)         */
  double f_scale;           /* least_squares f_scale                                                   */
  int32_t max_nfev;         /* max_iterations -> max_nfev (calibration.py:199,210)                     */
  int32_t loss;             /* MCBA_LOSS_*                                                             */
} mcba_solve_opts;

typedef struct {            /* one row of scipy's verbose=2 table (calibration.py:208)                 */
  int32_t iteration, nfev;
  double cost, cost_reduction, step_norm, optimality;
} mcba_log_row;

typedef struct {
  double cost, initial_cost, optimality;
  int32_t nfev, njev, status;   /* scipy termination status -1,0,1,2,3,4 (_lsq/common.py:705-717)      */
  int32_t n_log;                /* rows written to the log buffer                                      */
  double device_ms;             /* CUDA-event time of the solve on the context stream                  */
  int32_t kernel_launches;      /* kernels launched by this solve                                      */
  int32_t chol_retries;
} mcba_solve_result;

/* -- context ------------------------------------------------------------------------------------ */
int  mcba_create(int device_ordinal, mcba_ctx** out);
void mcba_destroy(mcba_ctx* ctx);
const char* mcba_last_error(const mcba_ctx* ctx);   /* ctx may be NULL: returns the creation error     */
int  mcba_set_stream(mcba_ctx* ctx, void* cuda_stream);   /* e.g. torch.cuda.current_stream().cuda_stream */
int  mcba_version(void);

/* -- multi-GPU: one context per process/GPU; frames are sharded across ranks ----
 * mcba_comm_init records (rank, world <= 16); mcba_comm_unique_id is kept for the usual bootstrap shape (rank 0 makes an id, the host
 * broadcasts it) and returns a constant tag.  The exchanges of a solve -- shared gradient / diagonal / cost, g_h^T A g_h, the reduced
 * normal equations S and rhs, the subspace sums: four per LM iteration -- run INSIDE the solver kernel over NVLink peer memory
 * (push to every rank's slot, sequence flag, reduce in rank order: bit-identical results on every rank; csrc/lm_kernel.cuh exchange()).
 * Every rank exports one buffer as an IPC handle (64 bytes), the host all-gathers the handles (torch.distributed / NCCL is the
 * bootstrap), every rank imports all of them.  cap_doubles per slot must hold n_s^2 + 2 n_s + 16; a rank that stops answering makes
 * its peers' solves fail with MCBA_ERR_NCCL after a bounded wait instead of hanging their GPUs. */
int  mcba_comm_unique_id(mcba_ctx* ctx, char out_id[128]);
int  mcba_comm_init(mcba_ctx* ctx, const char id[128], int rank, int world);
int  mcba_peer_export(mcba_ctx* ctx, int64_t cap_doubles, char out_handle[64]);
int  mcba_peer_import(mcba_ctx* ctx, const char* handles /* world x 64 bytes in rank order */);

/* -- problem upload: replaces what `evaluate` closes over (calibration.py:204-206) --------------
 * cam/frame/board/point: int32[N] rows of np.argwhere(inliers) (frame ids local to this rank);
 * obs: f64[N][2] = point_table.points[inliers] (calibration.py:206);
 * board_points: f64[B][P][3] = tables.stack_boards(boards).points (tables.py:385-394).            */
int  mcba_upload(mcba_ctx* ctx, const mcba_problem_desc* desc,
                 const int32_t* cam, const int32_t* frame, const int32_t* board, const int32_t* point,
                 const double* obs, const double* board_points);

/* Dense variant (the reference's own wire format): mask uint8[C][F][B][P] = calib.inliers (calibration.py:78-81),
 * points f64[C][F][B][P][2] = point_table.points (tables.make_point_table, tables.py:68-81).  Packing into the
 * frame-major corner arrays happens on the device; *n_corners receives the number of selected corners.  The packed
 * order seen through mcba_residuals is still np.argwhere(mask) order.                                              */
int  mcba_upload_dense(mcba_ctx* ctx, const mcba_problem_desc* desc, const uint8_t* mask, const double* points,
                       const double* board_points, int64_t* n_corners);
/* The same with the mask in the two parts the reference builds it from (calibration.py:73-81: `valid` = point_table.valid &
 * pose validity broadcast over the points): valid uint8[C][F][B][P] = point_table.valid as detected, view_valid uint8[C][F][B] =
 * camera_poses.valid x motion.valid x board_poses.valid (may be NULL: all views valid).  The conjunction is taken on the device, so
 * the caller hands over the detection table untouched (no 1-byte-per-point host pass before the copy).                          */
int  mcba_upload_dense_views(mcba_ctx* ctx, const mcba_problem_desc* desc, const uint8_t* valid, const uint8_t* view_valid,
                             const double* points, const double* board_points, int64_t* n_corners);
/* The same for a float32 table, points f32[C][F][B][P][2]: make_point_table keeps the dtype of the detector's corners (tables.py:15-17
 * fill_sparse: `dtype=values.dtype`; cv2's aruco / charuco detectors return float32), so this IS the reference's table for real
 * detections -- half the bytes over the link and no float64 copy of the table on the host.  The packed observations are f64 (exact). */
int  mcba_upload_dense_views_f32(mcba_ctx* ctx, const mcba_problem_desc* desc, const uint8_t* valid, const uint8_t* view_valid,
                                 const float* points, const double* board_points, int64_t* n_corners);

/* full parameter state (also the values of disabled/fixed blocks):
 * cam_rt f64[C][6], board_rt f64[B][6], frame_rt f64[F][6] (PoseSet.params, pose_set.py:51-53),
 * intrinsics f64[C][5+nd] (Camera.params, camera.py:144-155)                                      */
int  mcba_set_params(mcba_ctx* ctx, const double* cam_rt, const double* board_rt,
                     const double* frame_rt, const double* intrinsics);
int  mcba_get_params(mcba_ctx* ctx, double* cam_rt, double* board_rt, double* frame_rt, double* intrinsics);
/* the same state as 4x4 pose matrices, f64[C+B+F][4][4] in the order cameras, boards, frames (PoseSet.poses,
 * pose_set.py:40-42): matrix <-> rtvec (transform/rtvec.py:24-32) is done on the device */
int  mcba_set_state_matrices(mcba_ctx* ctx, const double* pose_matrices, const double* intrinsics);
int  mcba_get_state_matrices(mcba_ctx* ctx, double* pose_matrices, double* intrinsics);
/* motion-model state (after the upload, next to mcba_set_params / mcba_set_state_matrices):
 * rolling: the frames of the calls above are the START poses (mcba_set_params: frame_rt is f64[F][12], start | end of a
 *   frame adjacent); end poses f64[F][4][4] and image heights f64[C] (rolling_times, rolling_frames.py:15-19) come here.
 * hand-eye: the frames of the calls above are ignored on input (mcba_get_state_matrices returns the derived frame poses);
 *   base_wrt_gripper f64[F][4][4] (constants), world_wrt_base and gripper_wrt_camera f64[4][4] (the parameters). */
int  mcba_set_rolling(mcba_ctx* ctx, const double* end_pose_matrices, const double* image_heights);
int  mcba_get_rolling(mcba_ctx* ctx, double* end_pose_matrices);
int  mcba_set_hand_eye(mcba_ctx* ctx, const double* base_wrt_gripper, const double* world_wrt_base,
                       const double* gripper_wrt_camera);
int  mcba_get_hand_eye(mcba_ctx* ctx, double* world_wrt_base, double* gripper_wrt_camera);
int  mcba_num_params(mcba_ctx* ctx, int64_t* n);      /* length of param_vec for the enabled blocks    */
int  mcba_get_param_vec(mcba_ctx* ctx, double* x);    /* Parameters.param_vec (parameters.py:44-46)    */
int  mcba_set_param_vec(mcba_ctx* ctx, const double* x);   /* with_param_vec (parameters.py:48-50)     */

/* -- parity hooks -------------------------------------------------------------------------------
 * residuals: r f64[2N] = evaluate(x) (calibration.py:204-206); cost = 0.5*sum(rho) as scipy.       */
int  mcba_residuals(mcba_ctx* ctx, const double* x /*NULL = current params*/, double* r, double* cost);
/* linearize at x: dense J^T J (f64[n][n], row-major, parameter order of param_vec) and J^T r (f64[n])
 * of the (loss-scaled) residual -- what scipy's FD Jacobian + compute_grad produce (trf.py).       */
int  mcba_linearize(mcba_ctx* ctx, const double* x, double* JtJ, double* Jtr, double* cost);
/* per-corner reprojection error norm, f64[N] (tables.py:244-249 restricted to the packed corners)  */
int  mcba_reprojection_error(mcba_ctx* ctx, double* err);

/* -- the solve: replaces scipy.optimize.least_squares(...) at calibration.py:209-210 ------------ */
int  mcba_solve(mcba_ctx* ctx, const mcba_solve_opts* opts, mcba_solve_result* result,
                mcba_log_row* log, int32_t log_capacity);

/* -- resident point table: the outlier loop without host round trips --------------------------------
 * Calibration.adjust_outliers (calibration.py:250-266) alternates  report -> select threshold (np.quantile of the
 * per-corner error over `valid`, calibration.py:37-40) -> reject_outliers (calibration.py:240-252, errors from
 * tables.reprojection_error, tables.py:244-249) -> bundle_adjust.  With the table resident, the dense
 * [C,F,B,P] observations cross PCIe once; each round moves a few order statistics and one mask count.
 *
 * mcba_table_upload: valid uint8[C][F][B][P] = point_table.valid & pose validity (calibration.py:73-76),
 *   points f64[C][F][B][P][2]; the inlier mask starts equal to `valid`; the packed set is `valid`.
 * mcba_table_from_detections: the same table built on the device from the per-image detection lists the
 *   reference keeps before tables.make_point_table (tables.py:68-81; fill_sparse 15-21): list w = (c*F+f)*B+b
 *   holds det_ids[det_start[w]..det_start[w+1]) (point ids on board b) and det_xy (pixel corners).  Ids outside
 *   [0,P) fail with MCBA_ERR_ARG; a repeated id within one list keeps an unspecified one of its corners.
 * Both reset the parameter state to zero like mcba_upload; set it afterwards.                                    */
typedef struct mcba_table_stats {
  int64_t n_valid, n_inliers;          /* corners in `valid` / in the current inlier mask                        */
  double  sumsq_valid, sumsq_inliers;  /* sum of squared pixel errors over each set (mse = sumsq / n)            */
} mcba_table_stats;
enum { MCBA_TABLE_VALID = 0, MCBA_TABLE_INLIERS = 1 };

int  mcba_table_upload(mcba_ctx* ctx, const mcba_problem_desc* desc, const uint8_t* valid, const double* points,
                       const double* board_points, int64_t* n_valid);
int  mcba_table_from_detections(mcba_ctx* ctx, const mcba_problem_desc* desc, const int64_t* det_start,
                                const int32_t* det_ids, const double* det_xy, const double* board_points,
                                int64_t* n_valid);
/* read the resident table back (either pointer may be NULL): valid uint8[C][F][B][P], points f64[C][F][B][P][2] */
int  mcba_table_download(mcba_ctx* ctx, uint8_t* valid, double* points);
/* inlier mask uint8[C][F][B][P]; set: NULL restores inliers = valid, otherwise the mask is and-ed with `valid` */
int  mcba_table_set_inliers(mcba_ctx* ctx, const uint8_t* mask);
int  mcba_table_get_inliers(mcba_ctx* ctx, uint8_t* mask);
/* pack `valid` or the inlier mask as the problem the solver / residual entry points work on; the parameter state
 * is kept.  *n_corners (may be NULL) receives the packed count. */
int  mcba_table_select(mcba_ctx* ctx, int which, int64_t* n_corners);
/* per-corner error over `valid` at the current parameters, kept on the device sorted (all valid, and the inlier
 * subset) for mcba_table_error_ranks / mcba_table_reject; selects MCBA_TABLE_VALID as a side effect. */
int  mcba_table_errors(mcba_ctx* ctx, mcba_table_stats* stats);
/* out[i] = ranks[i]-th smallest error (0-based) of the chosen set -- the order statistics np.quantile interpolates */
int  mcba_table_error_ranks(mcba_ctx* ctx, int which, const int64_t* ranks, int32_t n, double* out);
/* counts[i] = number of errors of the chosen set below thresholds[i]: the other half of a quantile over SEVERAL ranks' tables
 * (frames sharded across GPUs: every rank sorts its own errors; ranks and counts are merged on the host) */
int  mcba_table_count_below(mcba_ctx* ctx, int which, const double* thresholds, int32_t n, int64_t* counts);
/* inliers = valid & (error < threshold) (calibration.py:243-244) with the errors of the last mcba_table_errors;
 * MCBA_ERR_STATE if parameters or selection changed since.  Follow with mcba_table_select(MCBA_TABLE_INLIERS). */
int  mcba_table_reject(mcba_ctx* ctx, double threshold, int64_t* n_valid, int64_t* n_keep);

/* -- batched board-pose initialisation (what feeds the path) -------------------------------
 * Replaces the C*F*B calls of board.estimate_pose_points (board/common.py:36-47: camera.undistort_points, then
 * cv2.solvePnPGeneric with the camera matrix and no distortion) that tables.make_pose_table makes (tables.py:44-66).
 * Detection lists as for mcba_table_from_detections; intrinsics f64[C][5+nd] (Camera.params); board_grid int32[B][5] =
 * {id-grid width, height, id divisor (1 ChArUco corner ids, 4 AprilGrid tag corners), min_points, min_rows} of
 * has_min_detections_grid (board/common.py:30-34; charuco.py:104-106, aprilgrid.py:197-199).
 * Out, per list w = (c*F+f)*B+b: pose f64[4][4] (board wrt camera), reprojection RMS over the 2n scalar residuals as
 * solvePnPGeneric reports it, number of corners, valid flag; a view without the minimum detections gets the reference's
 * invalid_pose (identity, 0, 0, false; tables.py:38).  Does not touch the uploaded problem.                       */
int  mcba_pnp_views(mcba_ctx* ctx, const mcba_problem_desc* desc, const int64_t* det_start, const int32_t* det_ids,
                    const double* det_xy, const double* board_points, const double* intrinsics, const int32_t* board_grid,
                    double* poses, double* errors, int32_t* num_points, uint8_t* valid);

/* -- measurement hooks (bench.py): launch one kernel family on the context stream -------------- */
enum { MCBA_BENCH_LINEARIZE = 0, MCBA_BENCH_RESIDUAL = 1, MCBA_BENCH_COST = 2,
       MCBA_BENCH_NO_PREPARE = 256 /* or-ed in: reuse the pose tables of the previous call (times the kernel alone) */ };
int  mcba_bench_launch(mcba_ctx* ctx, int which, int repeats);
int  mcba_bench_info(mcba_ctx* ctx, int which, int64_t* corners, int64_t* bytes_per_launch, int32_t* launches_per_call);

#ifdef __cplusplus
}
#endif
#endif
