/*
 * C handles of the C++ host layer of flvis_b200 (flvis_b200/host/), for callers that cannot include the C++
 * headers (the ctypes tests, other languages).  The C++ classes keep the reference's names and semantics:
 *   flv::PoseLMBag  <- src/backend/include/poselmbag.h:24-63
 *   flv::LocalMap   <- LocalMapNodeletClass::frame_callback, src/backend/vo_localmap.cpp:87-380
 */
#ifndef FLVIS_B200_HOST_H
#define FLVIS_B200_HOST_H
#include "flvis_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct flv_localmap flv_localmap;

flv_localmap* flv_localmap_create(flv_ctx* ctx, int window_size, double fx, double fy, double cx, double cy);
void flv_localmap_destroy(flv_localmap* lm);
void flv_localmap_reset(flv_localmap* lm);      /* KFMSG_CMD_RESET_LM, vo_localmap.cpp:89-98 */
/* One KeyFrame message (msg/KeyFrame.msg without the images): lm_2d[n][2] undistorted px, lm_3d[n][3] world,
 * T_c_w = [qx qy qz qw tx ty tz].  Returns 1 when a solve ran and the CorrectionInf outputs were written
 * (msg/CorrectionInf.msg), 0 when the window is still filling, <0 on error (buffers too small = FLV_ERR_OVERFLOW, a due solve
 * that failed = FLV_ERR_CUDA with flv_last_error(ctx) set).  window_size: 3..100 like the reference
 * (vo_localmap.cpp:441-447): up to 25 poses the reduced camera system lives in shared memory (csrc/ba.cu), larger windows run
 * on the global-memory solver (csrc/ba_big.cu). */
int flv_localmap_add_keyframe(flv_localmap* lm, int64_t frame_id, int n, const int64_t* lm_id, const double* lm_2d,
                              const double* lm_3d, const double* T_c_w, int64_t* out_frame_id, double* out_T_c_w,
                              int* out_lm_count, int64_t* out_lm_id, double* out_lm_3d, int lm_cap,
                              int* out_outlier_count, int64_t* out_outlier_id, int outlier_cap,
                              flv_ba_stats* out_stats);

/* ---- flv::VIMOTION <- src/processing/include/vi_motion.h, src/processing/vi_motion.cpp:3-464 -------------------
 * Poses are [qx qy qz qw tx ty tz]; orientation outputs are Eigen order [qw qx qy qz] like the reference's
 * Quaterniond.  flv_vimotion_imu_feed = F2FTracking::imu_feed (src/frontend/f2f_tracking.cpp:46-57). */
typedef struct flv_vimotion flv_vimotion;
flv_vimotion* flv_vimotion_create(const double* T_i_c, double magnitude_g, double para1, double para2, double para3,
                                  double para4, double para5, double para6);
void flv_vimotion_destroy(flv_vimotion* vm);
int flv_vimotion_imu_feed(flv_vimotion* vm, double t, const double* acc, const double* gyro, double* q_wxyz,
                          double* pos, double* vel);                      /* returns 1 once imu_initialized */
int flv_vimotion_vision_trigger(flv_vimotion* vm, double* q_wxyz);
int flv_vimotion_correction(flv_vimotion* vm, double t_curr, const double* Tcw_curr, double t_last, const double* Tcw_last);
int flv_vimotion_corr_frame_state(flv_vimotion* vm, double t, double* T_c_w);     /* 1 found, 0 not in queue */
int flv_vimotion_rp_compensation(flv_vimotion* vm, double t, double* T_c_w_inout);
int flv_vimotion_get_bias(flv_vimotion* vm, double* acc_bias, double* gyro_bias);
int flv_vimotion_queue_size(flv_vimotion* vm);

/* ---- flv::F2FTracking <- src/frontend/include/f2f_tracking.h:24-78, src/frontend/f2f_tracking.cpp:5-453 --------
 * One handle = one camera sequence (private single-stream context).  cam_type: 0 = DEPTH_D435 (img1 = u16 depth),
 * 1 = STEREO_RECT, 2 = STEREO_UNRECT (img1 = u8 right image).  Poses are [qx qy qz qw tx ty tz]. */
typedef struct flv_f2f flv_f2f;
typedef struct {
  int cam_type, img_w, img_h;
  double cam0[4];          /* fx fy cx cy */
  double cam1[4];          /* K1 (stereo) */
  double depth_scale;
  double P0[12], P1[12];
  double T_cam1_cam0[7];
  double T_i_c0[7];
  double feature_para[6], vi_para[6], dc_para[3];
  int skip_first_n_imgs;
} flv_f2f_config;
/* RANSAC hooks replacing the host stand-ins (tests inject OpenCV's results through these) */
typedef int (*flv_f2f_fmat_fn)(void* user, int n, const float* from_xy, const float* to_xy, uint8_t* mask_out);
typedef int (*flv_f2f_pnp_fn)(void* user, int n, const float* p3d, const float* p2d, const double* K4, int use_guess,
                              double* T_c_w_inout, int* inlier_idx_out, int* n_inliers_out);
flv_f2f* flv_f2f_create(const flv_f2f_config* cfg, int device);
void flv_f2f_destroy(flv_f2f* f);
/* STEREO_UNRECT (cam_type 2; EuRoC raw images): raw lens model of camera `cam` (0 | 1) = DepthCamera::K0/D0/R0 or K1/D1/R1
 * (src/processing/depth_camera.cpp:27-72): K4 = fx fy cx cy, D14 = OpenCV order k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 0 0,
 * R9 = rectification rotation (row-major).  The rectified projection P0 / P1 comes from flv_f2f_config.  Call before the
 * first image_feed.  need_equal_hist (f2f_tracking.cpp:125-145): cv::equalizeHist of every ingested image, on the GPU. */
int flv_f2f_set_lens(flv_f2f* f, int cam, const double* K4, const double* D14, const double* R9);
/* DepthCamera::setSteroCamInfo (src/processing/depth_camera.cpp:27-90) followed by F2FTracking::init, with the matrices the
 * tracking nodelet hands over (src/frontend/vo_tracking.cpp:198-214 D435 stereo, :245-262 EuRoC, :276-302 KITTI): row-major
 * K 3x3, D (OpenCV order, 14 slots, unused = 0), R 3x3 and P 3x4 of cv::stereoRectify, T_c0_c1 = [qx qy qz qw tx ty tz].
 * cam_type 1 = STEREO_RECT, 2 = STEREO_UNRECT.  need_equal_hist as in F2FTracking::init (EuRoC: 1). */
typedef struct {
  int cam_type, img_w, img_h;
  double K0[9], D0[14], R0[9], P0[12];
  double K1[9], D1[14], R1[9], P1[12];
  double T_c0_c1[7];
  double T_i_c0[7];
  double feature_para[6], vi_para[6], dc_para[3];
  int skip_first_n_imgs, need_equal_hist;
} flv_f2f_stereo_config;
flv_f2f* flv_f2f_create_stereo(const flv_f2f_stereo_config* cfg, int device);
/* The DepthCamera fields setSteroCamInfo derives (no GPU needed): cam0 / cam1 = fx fy cx cy taken from P0 / P1
 * (depth_camera.cpp:73-82), T_cam1_cam0 = T_c0_c1^-1 (:60-61). */
int flv_host_stereo_cam_info(const flv_f2f_stereo_config* cfg, double* cam0_4, double* cam1_4, double* T_cam1_cam0_7);
int flv_f2f_set_equalize_hist(flv_f2f* f, int enable);
/* the two per-point OpenCV calls of the UNRECT path, exported for the parity tests (flvis_b200/host/undistort.h) */
int flv_host_undistort_points(const double* K4, const double* D14, const double* R9, const double* P12, int n, const float* in_xy,
                              float* out_xy);
/* the host stand-ins of the two RANSAC calls (flvis_b200/host/ransac.h; selectable instead of the device kernels), exported
 * for the CPU tests.  Return value: fundamental 1 / 0 (model found), PnP = number of inliers (0 = failed). */
int flv_host_fundamental_ransac(int n, const float* from_xy, const float* to_xy, double thr_px, double conf, uint8_t* mask, double* F9);
int flv_host_pnp_ransac(int n, const float* p3d, const float* p2d, const double* K4, double* T_c_w_inout, int iterations, double thr_px,
                        double conf, uint8_t* mask);
int flv_host_project_points(const double* K4, const double* D14, const double* Rcw9, const double* t3, int n, const float* xyz,
                            float* out_xy);
const char* flv_f2f_last_error(flv_f2f* f);
void flv_f2f_set_ransac_hooks(flv_f2f* f, flv_f2f_fmat_fn fmat, flv_f2f_pnp_fn pnp, void* user);
int flv_f2f_imu_feed(flv_f2f* f, double t, const double* acc, const double* gyro);
int flv_f2f_image_feed(flv_f2f* f, double t, const uint8_t* img0, const void* img1, int* new_keyframe, int* reset_cmd);
/* F2FTracking::correction_feed (src/frontend/f2f_tracking.cpp:40-44): hand a CorrectionInf message (msg/CorrectionInf.msg: the
 * output of flv_localmap_add_keyframe) back to the tracker; it is applied at the start of the next tracked frame (:189-219:
 * the pose record of that keyframe and everything after it are re-based, the last frame's landmarks get the optimised world
 * points, the local map's outliers are flagged).  The reference's tracking nodelet drops the message (vo_tracking.cpp:373-385),
 * so nothing calls this unless the integrator wires the feedback. */
int flv_f2f_correction_feed(flv_f2f* f, double t, int64_t frame_id, const double* T_c_w, int lm_count, const int64_t* lm_id,
                            const double* lm_3d, int outlier_count, const int64_t* outlier_id);
int flv_f2f_state(flv_f2f* f);     /* 0 UnInit, 1 Tracking, 2 TrackingFail */
/* landmarks of the current frame (after image_feed): returns the count (<= cap) */
int flv_f2f_get_frame(flv_f2f* f, double* T_c_w, int64_t* lm_id, double* plane_xy, double* undist_xy, double* p3d_w,
                      uint8_t* has_3d, uint8_t* is_inlier, int cap);
/* the remaining per-landmark state of the current frame (LandMarkInFrame::lm_3d_c, lm_1st_obs_2d, lm_1st_obs_frame_pose,
 * landmark.h:8-36) and F2FTracking::T_c_w_last_keyframe; returns the count (<= cap).  With flv_f2f_get_frame and
 * flv_f2f_get_imu_states this is the tracker's complete continuous state (the tests re-seed the oracle from it). */
int flv_f2f_get_frame_ex(flv_f2f* f, double* p3d_c, double* first_obs_2d, double* first_obs_pose, double* T_c_w_last_keyframe, int cap);
/* VIMOTION::states (vi_motion.h:30): out[i] = {t, qw qx qy qz, px py pz, vx vy vz}; returns the queue length (<= cap) */
int flv_f2f_get_imu_states(flv_f2f* f, double* out11, int cap);
/* VIMOTION::acc_bias / gyro_bias of the tracker's IMU filter (vi_motion.cpp:322-330); returns has_imu */
int flv_f2f_get_imu_bias(flv_f2f* f, double* acc_bias, double* gyro_bias);
int flv_f2f_tracking_counts(flv_f2f* f, int* of_inliers, int* f_inliers, int* pnp_inliers);

/* ---- trajectory files <- src/independ_modules/vo_repub_rec.cpp:80-126 -------------------------------------------------
 * format 0: the recorder's pose line "stamp x y z qw qx qy qz" (qw first, precision 6, stamp as sec.nsec);
 * format 1: KITTI odometry, 12 numbers = row-major [R | t], precision 6.
 * Poses are given as the tracker's T_c_w = [qx qy qz qw tx ty tz]; the file holds the camera pose in the world (T_c_w^-1),
 * which is what FLVIS publishes and the recorder subscribes to. */
typedef struct flv_traj flv_traj;
flv_traj* flv_traj_open(const char* path, int format);
int flv_traj_write(flv_traj* t, double stamp, const double* T_c_w);
void flv_traj_close(flv_traj* t);

/* ---- flv_localmap_batch: the local-map thread for S sequences ------------------------------------------------------
 * One worker thread (FLVIS runs the local map in its own nodelet thread fed by a queue, vo_localmap.cpp:464-467) owns S
 * LocalMap state machines and solves all windows that are due after a submission with ONE flv_ba_optimize launch on its own
 * CUDA stream.  submit() returns immediately; wait() drains the queue; result() returns the latest CorrectionInf of a stream
 * (return value = number of solves so far for that stream, 0 = none yet).  Keyframes of one submission are concatenated:
 * lm_id / lm_2d / lm_3d hold lm_counts[0] entries of keyframe 0, then lm_counts[1] of keyframe 1, ...; T_c_w is [n_kf][7]. */
typedef struct flv_localmap_batch flv_localmap_batch;
flv_localmap_batch* flv_localmap_batch_create(int device, int n_streams, int window_size, double fx, double fy, double cx, double cy);
void flv_localmap_batch_destroy(flv_localmap_batch* b);
const char* flv_localmap_batch_last_error(flv_localmap_batch* b);
int flv_localmap_batch_submit(flv_localmap_batch* b, int n_kf, const int* streams, const int64_t* frame_ids, const int* lm_counts,
                              const int64_t* lm_id, const double* lm_2d, const double* lm_3d, const double* T_c_w);
int flv_localmap_batch_wait(flv_localmap_batch* b);
/* solve_ms2 = {wall ms inside the solver calls, wall ms of graph editing + array packing on the worker thread} */
int flv_localmap_batch_stats(flv_localmap_batch* b, long long* n_keyframes, long long* n_solves, long long* n_launches, double* solve_ms2);
/* totals over all windows solved so far: {edges, landmarks, poses, LM iterations} (for the BA roofline: SURVEY.md 8(d) bytes) */
int flv_localmap_batch_problem_totals(flv_localmap_batch* b, double* totals4);
int flv_localmap_batch_result(flv_localmap_batch* b, int stream, int64_t* out_frame_id, double* out_T_c_w, int* out_lm_count,
                              int64_t* out_lm_id, double* out_lm_3d, int lm_cap, int* out_outlier_count, int64_t* out_outlier_id,
                              int outlier_cap);

/* ---- flv_f2f_batch: S camera sequences of the same sensor advanced together ------------------------------------------
 * The batched form of flv::F2FTracking (src/frontend/f2f_tracking.cpp:5-453): per stream the same state machine, IMU filter,
 * landmark id counter and rand() stream as one flv_f2f handle, but every stage (pyramids, frame->frame LK, keep rule,
 * F / PnP RANSAC, pose-only BA, reprojection cull, FeatureDEM redetect, left->right LK, depth innovation) is ONE launch for
 * all streams and the landmark lists stay on the device between stages; the host synchronises once per frame.
 * All streams share `cfg` (one sensor model).  Results per stream are identical to S separate flv_f2f handles
 * (tests/test_batch_tracker_gpu.py).
 *   imu_feed:   F2FTracking::imu_feed for one stream (may be called from another thread than image_feed).
 *   image_feed: t[S]; img0 = S left images back to back (u8, tightly packed); img1 = S right images (u8) or S depth images
 *               (u16); mem says where the images live (pinned host memory gives asynchronous copies);
 *               new_keyframe[S] / reset_cmd[S] as F2FTracking::image_feed returns them per stream. */
typedef struct flv_f2f_batch flv_f2f_batch;
flv_f2f_batch* flv_f2f_batch_create(const flv_f2f_config* cfg, int n_streams, int device);
/* The same object split into `groups` equal groups of streams (n_streams % groups == 0, else one group): each group has its
 * own kernel context, CUDA stream and host thread, so the one-CTA-per-stream stages of one group overlap the other groups'
 * kernels and per-frame host work.  Per-stream results are unchanged (streams never interact). */
flv_f2f_batch* flv_f2f_batch_create_grouped(const flv_f2f_config* cfg, int n_streams, int device, int groups);
int flv_f2f_batch_groups(flv_f2f_batch* b);
void flv_f2f_batch_destroy(flv_f2f_batch* b);
const char* flv_f2f_batch_last_error(flv_f2f_batch* b);
flv_ctx* flv_f2f_batch_context(flv_f2f_batch* b);      /* the batch's kernel context (stream selection, launch counter) */
int flv_f2f_batch_set_lens(flv_f2f_batch* b, int cam, const double* K4, const double* D14, const double* R9);
int flv_f2f_batch_set_equalize_hist(flv_f2f_batch* b, int enable);
void flv_f2f_batch_set_ransac_hooks(flv_f2f_batch* b, flv_f2f_fmat_fn fmat, flv_f2f_pnp_fn pnp, void* user);
int flv_f2f_batch_imu_feed(flv_f2f_batch* b, int stream, double t, const double* acc, const double* gyro);
int flv_f2f_batch_image_feed(flv_f2f_batch* b, const double* t, const uint8_t* img0, const void* img1, flv_memspace mem,
                             int* new_keyframe, int* reset_cmd);
/* Pipelined form of imu_feed_many + image_feed (worth it for grouped batches): for every group in turn the previous frame is
 * finished, the group's IMU samples of this call are fed, and its new frame is enqueued, so the GPU works on the other
 * groups while the host serves one.  Per stream the order of operations is that of the synchronous calls; results
 * (accessors, keyframe hand-off to an attached local map) lag by one call.  flv_f2f_batch_sync finishes the frames in
 * flight and returns the flags of every stream's last frame; the synchronous image_feed syncs first. */
int flv_f2f_batch_frame_async(flv_f2f_batch* b, const double* t, const uint8_t* img0, const void* img1, flv_memspace mem, int n_imu,
                              const int* imu_streams, const double* imu_t, const double* imu_acc, const double* imu_gyro);
int flv_f2f_batch_sync(flv_f2f_batch* b, int* new_keyframe, int* reset_cmd);
int flv_f2f_batch_state(flv_f2f_batch* b, int stream);
int flv_f2f_batch_get_frame(flv_f2f_batch* b, int stream, double* T_c_w, int64_t* lm_id, double* plane_xy, double* undist_xy,
                            double* p3d_w, uint8_t* has_3d, uint8_t* is_inlier, int cap);
int flv_f2f_batch_get_frame_ex(flv_f2f_batch* b, int stream, double* p3d_c, double* first_obs_2d, double* first_obs_pose,
                               double* T_c_w_last_keyframe, int cap);
int flv_f2f_batch_get_imu_states(flv_f2f_batch* b, int stream, double* out11, int cap);
int flv_f2f_batch_get_imu_bias(flv_f2f_batch* b, int stream, double* acc_bias, double* gyro_bias);
int flv_f2f_batch_tracking_counts(flv_f2f_batch* b, int stream, int* of_inliers, int* f_inliers, int* pnp_inliers);
long long flv_f2f_batch_launch_count(flv_f2f_batch* b);
/* n IMU samples in one call (sample i belongs to streams[i]; acc / gyro are [n][3]) */
int flv_f2f_batch_imu_feed_many(flv_f2f_batch* b, int n, const int* streams, const double* t, const double* acc, const double* gyro);
/* Per-stage device time (CUDA events on the compute stream at the stage boundaries, accumulated over frames): stage_ms9 =
 * {ingest + pyramids, frame->frame LK, keep rule + F RANSAC, PnP RANSAC, pose-only BA, reprojection cull, FeatureDEM redetect,
 *  left->right LK, depth innovation + finish}.  set_profile resets the accumulators. */
int flv_f2f_batch_set_profile(flv_f2f_batch* b, int enable);
int flv_f2f_batch_get_profile(flv_f2f_batch* b, double* stage_ms9, long long* frames);
/* host wall time of image_feed over the same frames: {per-stream decisions, enqueue of the frame's work, waiting for the
 * device, per-stream post-frame work (IMU feedback, keyframe rule, keyframe hand-off)} */
int flv_f2f_batch_get_host_profile(flv_f2f_batch* b, double* host_ms4);
/* Hand every new keyframe (KeyFrame message content: ids / undistorted pixels / world points of the inlier landmarks with
 * depth + T_c_w, keyframe_msg.cpp:30-110) to `lm` from inside image_feed; NULL detaches.  The local map never blocks tracking. */
int flv_f2f_batch_attach_localmap(flv_f2f_batch* b, flv_localmap_batch* lm);
/* What image_feed copies back per frame besides the per-stream summaries: full = 1 (default) the complete landmark lists
 * (everything the get_frame / get_frame_ex accessors return), full = 0 only what a KeyFrame message and the pose consumers
 * need (ids, undistorted pixels, world points, flags, counts, poses: 50 instead of 162 bytes per landmark slot); with
 * full = 0 get_frame's plane_xy and get_frame_ex's outputs are not refreshed. */
int flv_f2f_batch_set_readback(flv_f2f_batch* b, int full);
/* Per-frame result records for a gather across GPUs: after every finished frame, buf[frame % block_frames][stream][8] =
 * {T_c_w as qx qy qz qw tx ty tz, landmark count} (host memory, e.g. the pinned staging block of the collective);
 * buf = NULL stops logging; the frame counter restarts at 0 with every call. */
int flv_f2f_batch_set_result_log(flv_f2f_batch* b, double* buf, int block_frames);

#ifdef __cplusplus
}
#endif
#endif
