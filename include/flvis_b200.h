/*
 * flvis_b200 -- C ABI of the Blackwell-native FLVIS hot path (libflvis_b200.so).
 *
 * The reference (HKPolyU-UAV/FLVIS) has no FFI around its hot path: the heavy calls are C++
 * member functions that call OpenCV / g2o in-process.  Each entry point below names the
 * reference call it replaces (paths relative to the reference tree).  The C++ host classes
 * in flvis_b200/host/ (same names and argument meaning as the reference's) forward to these.
 *
 * Conventions
 *   - every function returns FLV_OK (0) or a negative flv_status; nothing throws or aborts;
 *   - one flv_ctx per GPU; it owns `max_streams` independent stream slots (one camera
 *     sequence each) that are processed by ONE launch per stage (batched);
 *   - `mem` says where the caller's arrays live: FLV_MEM_HOST (the library stages them
 *     through pinned buffers and synchronises before returning) or FLV_MEM_DEVICE (device
 *     pointers on ctx's device, work is enqueued on the ctx stream, no synchronisation);
 *   - per-stream arrays are laid out [stream][max_pts] with the ctx's `max_pts` stride;
 *   - images are 8-bit single channel, row-major (flv_upload_color_images converts 3- / 4-channel frames);
 *   - a context is not thread-safe: one thread drives it at a time (several contexts, e.g. one per camera sequence or
 *     per GPU, are independent); FLV_MEM_HOST calls return results, FLV_MEM_DEVICE calls only enqueue work.
 */
#ifndef FLVIS_B200_H
#define FLVIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct flv_ctx flv_ctx;

typedef enum {
  FLV_OK = 0,
  FLV_ERR_INVALID = -1,   /* bad argument */
  FLV_ERR_CUDA = -2,      /* CUDA runtime error (see flv_last_error) */
  FLV_ERR_NOMEM = -3,
  FLV_ERR_UNSUPPORTED = -4, /* parameter combination the kernels do not implement */
  FLV_ERR_OVERFLOW = -5   /* a device-side capacity was exceeded (candidate list, window) */
} flv_status;

typedef enum { FLV_MEM_HOST = 0, FLV_MEM_DEVICE = 1 } flv_memspace;

#define FLV_MAX_LEVELS 4      /* 31x31 window: every FLVIS image size gives 4 pyramid levels */
#define FLV_NUM_SLOTS 4       /* image/pyramid slots per stream (prev0, cur0, cur1, spare) */
#define FLV_NUM_REGIONS 16    /* FeatureDEM's 4x4 grid, feature_dem.cpp:38-53 */

/* ---- context ---------------------------------------------------------------------------- */
int flv_create(flv_ctx** out, int device, int max_streams, int img_w, int img_h, int max_pts);
void flv_destroy(flv_ctx* ctx);
/* Enqueue all work on `cuda_stream` (a cudaStream_t, e.g. torch's current stream). NULL = own stream. */
int flv_set_stream(flv_ctx* ctx, void* cuda_stream);
int flv_sync(flv_ctx* ctx);
const char* flv_last_error(flv_ctx* ctx);
const char* flv_version(void);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
long long flv_launch_count(flv_ctx* ctx);
/* Geometry of pyramid level `level` (0 = full image): returns width/height/pitch/offset. */
int flv_level_info(flv_ctx* ctx, int level, int* w, int* h, int* pitch, size_t* offset);
int flv_num_levels(flv_ctx* ctx);

/* Colour frames (F2FTracking::image_feed converts 3- / 4-channel input with cvtColor BGR2GRAY / BGRA2GRAY, or the RGB
 * variants when mbRGB is set: src/frontend/f2f_tracking.cpp:78-110): interleaved u8 pixels, `channels` = 3 | 4,
 * is_rgb = 0 for BGR(A) order.  Bit-exact with cv2.cvtColor; equalizeHist (below) applies after the conversion. */
int flv_upload_color_images(flv_ctx* ctx, int slot, int n_streams, const uint8_t* imgs, size_t row_stride_bytes,
                            size_t img_stride_bytes, int channels, int is_rgb, flv_memspace mem);

/* need_equal_hist (src/frontend/f2f_tracking.cpp:125-145): when enabled, every image ingested by flv_upload_images goes
 * through cv::equalizeHist (per-image histogram, LUT) before level 0 is written.  Bit-exact with cv2.equalizeHist. */
int flv_set_equalize_hist(flv_ctx* ctx, int enable);
/* ---- images + pyramid (K1) ---------------------------------------------------------------
 * Replaces the image hand-over of F2FTracking::image_feed (src/frontend/f2f_tracking.cpp:59-145)
 * and cv::buildOpticalFlowPyramid inside cv::calcOpticalFlowPyrLK (called at
 * src/processing/lkorb_tracking.cpp:64-73, src/processing/camera_frame.cpp:124-128).
 * `imgs` holds n_streams images back to back: image s starts at imgs + s*img_stride_bytes,
 * rows are row_stride_bytes apart.  Level 0 of slot `slot` is overwritten. */
int flv_upload_images(flv_ctx* ctx, int slot, int n_streams, const uint8_t* imgs,
                      size_t row_stride_bytes, size_t img_stride_bytes, flv_memspace mem);
/* pyrDown chain (5x5 Gaussian, (s+128)>>8, REFLECT_101) for levels 1..L-1 of `slot`. */
int flv_build_pyramid(flv_ctx* ctx, int slot, int n_streams);
/* Copy pyramid level `level` of (slot, stream) out as a tight w*h image (tests / debugging). */
int flv_download_level(flv_ctx* ctx, int slot, int stream, int level, uint8_t* out, flv_memspace mem);

/* ---- pyramidal Lucas-Kanade (K2 frame->frame, K3 left->right) ------------------------------
 * Replaces cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err, Size(31,31),
 * maxLevel, TermCriteria(COUNT+EPS, 30, 1e-3), OPTFLOW_USE_INITIAL_FLOW) at
 * src/processing/lkorb_tracking.cpp:64-73 (maxLevel 10) and src/processing/camera_frame.cpp:124-128
 * (maxLevel 5).  Pyramids of both slots must have been built.  init_xy is the initial flow
 * (USE_INITIAL_FLOW); next_xy may alias init_xy.  win must be 31.  FLV_MEM_DEVICE callers that do not need OpenCV's `err`
 * output (the reference never reads it) may pass err = NULL: the kernel then skips the final error pass. */
typedef struct {
  int win;          /* 31 */
  int max_level;    /* levels used = min(max_level, built levels-1)+1 */
  int max_iter;     /* 30 */
  double eps;       /* 1e-3 (squared internally, like OpenCV) */
  double min_eig_threshold; /* 1e-4 */
} flv_lk_params;
int flv_lk_track(flv_ctx* ctx, int src_slot, int dst_slot, int n_streams, const int* n_pts,
                 const float* prev_xy, const float* init_xy, float* next_xy, uint8_t* status,
                 float* err, const flv_lk_params* prm, flv_memspace mem);

/* Keep rule of LKORBTracking::tracking (src/processing/lkorb_tracking.cpp:98-119): a tracked point survives
 * iff status==1 && 0<x<w-1 && 0<y<h-1.  keep[s][i] = 1/0; out_xy[s][i] = next if kept else prev (f32) and the
 * same as f64 in out_xy_f64 (the Vec2 list FeatureDEM::redetect takes); either output may be NULL.
 * Device pointers only (FLV_MEM_DEVICE); this is glue for device-resident pipelines. */
int flv_select_tracked(flv_ctx* ctx, int n_streams, const int* n_pts, const float* prev_xy,
                       const float* next_xy, const uint8_t* status, uint8_t* keep, float* out_xy,
                       double* out_xy_f64);

/* ---- Shi-Tomasi corners (K4) ---------------------------------------------------------------
 * Replaces cv::goodFeaturesToTrack(img, corners, maxCorners, quality, minDistance[, mask=255])
 * at src/processing/feature_dem.cpp:160 and :221 (blockSize 3, Sobel 3, no Harris).
 * Output: xy_out[s][i] integer-valued float coordinates in response-descending order,
 * n_out[s] corners (<= max_corners <= ctx max_corner capacity). */
int flv_gftt(flv_ctx* ctx, int slot, int n_streams, int max_corners, double quality,
             double min_distance, float* xy_out, int* n_out, int out_stride_pts, flv_memspace mem);
/* Debug: the fused kernel normally never writes the response map; enable=1 makes flv_gftt also store it so
 * flv_download_eig can return the f32 min-eigenvalue map of `stream` from the last flv_gftt (tests). */
int flv_gftt_keep_response(flv_ctx* ctx, int enable);
int flv_download_eig(flv_ctx* ctx, int stream, float* out, flv_memspace mem);
/* Capacity of the corner output (max value of max_corners). */
int flv_gftt_capacity(flv_ctx* ctx);
/* Device-mode callers: read (into flags_out[n_streams], may be NULL) and clear the per-stream capacity-overflow flags of the
 * Shi-Tomasi / FeatureDEM kernels (1 = candidates, 2 = accepted corners, 4 = region kept, 8 = sort depth, 16 = max_pts).  Host-mode
 * calls check them on their own.  Synchronises the context's stream; returns FLV_ERR_OVERFLOW if any flag was set. */
int flv_get_flags(flv_ctx* ctx, int n_streams, int* flags_out);

/* ---- FeatureDEM region selection (K5) ------------------------------------------------------
 * Replaces FeatureDEM::detect (src/processing/feature_dem.cpp:215-266) and
 * FeatureDEM::redetect (:124-213) including calHarrisR (:59-88) and fillIntoRegion (:92-121);
 * runs flv_gftt internally (2*gftt_num corners for detect, gftt_num for redetect).
 * existing_xy: [s][max_pts][2] f64 (redetect only; the reference passes vector<Vec2>).
 * new_xy: [s][max_pts][2] f32, n_new[s]; ordered region 0..15, within region by acceptance. */
typedef struct {
  int max_region_feature_num;  /* feature_para1 */
  int min_region_feature_num;  /* feature_para2 (parsed, unused by the reference) */
  int boundary_dis;            /* floor(feature_para3/2) */
  int gftt_num;                /* feature_para4 */
  double gftt_ql;              /* feature_para5 */
  int gftt_dis;                /* feature_para6 */
} flv_feature_params;
int flv_feature_detect(flv_ctx* ctx, int slot, int n_streams, const flv_feature_params* prm,
                       float* new_xy, int* n_new, flv_memspace mem);
int flv_feature_redetect(flv_ctx* ctx, int slot, int n_streams, const flv_feature_params* prm,
                         const double* existing_xy, const int* n_existing, float* new_xy,
                         int* n_new, flv_memspace mem);

/* Optional: start the Shi-Tomasi part (corner response + min-distance selection, feature_dem.cpp:160 / :221) of the next
 * flv_feature_detect (redetect = 0) / flv_feature_redetect (redetect = 1) call for `slot` NOW, on an internal auxiliary
 * stream, so it overlaps whatever the caller enqueues next (the reference runs goodFeaturesToTrack after tracking; it
 * only needs the image).  The next detect / redetect call with the same slot / n_streams / params consumes the result;
 * any other GFTT call waits for it and discards it.  Results are identical with or without this call. */
int flv_feature_prepare(flv_ctx* ctx, int slot, int n_streams, const flv_feature_params* prm, int redetect);

/* ---- local bundle adjustment (K7-K10) and pose-only BA ---------------------------------------
 * Replaces the g2o calls of LocalMapNodeletClass::frame_callback
 * (src/backend/vo_localmap.cpp:292-319: initializeOptimization(); optimize(12); chi2>3 edge
 * cull; initializeOptimization(); optimize(8)) -- BlockSolver_6_3 + LM + Huber(1) +
 * EdgeSE3ProjectXYZ -- and of OptimizeInFrame::optimize (src/processing/optimize_in_frame.cpp:10-90).
 * One problem per stream; the whole Levenberg-Marquardt loop runs on the device.
 * Poses are g2o SE3Quat order: [qx qy qz qw tx ty tz] of T_c_w. */
typedef struct {
  int n_poses;        /* P  (slots 0..P-1) */
  int n_landmarks;    /* L */
  int n_edges;        /* E */
  int fixed_pose;     /* index of the fixed pose vertex, -1 = none */
  int fix_landmarks;  /* 1 = pose-only BA (all points fixed, optimize_in_frame.cpp:41) */
  double fx, fy, cx, cy;
} flv_ba_problem;
typedef struct {
  int iters1;         /* 12 (local map) / 2 (pose only) */
  int iters2;         /* 8 / 2 */
  double huber_delta; /* 1.0 */
  double cull_chi2;   /* 3.0 */
  int min_edges_after_cull; /* 0 (local map) / 10 (pose only: fail if fewer) */
  int ws_slot0;       /* workspace slot of the first problem of this call: 0 unless several calls are in flight
                         concurrently on different streams (then give them disjoint [ws_slot0, ws_slot0+n) ranges) */
} flv_ba_params;
typedef struct {
  int iterations_run; /* total LM iterations executed (both phases) */
  int n_culled;       /* edges removed by the chi2 cull */
  int ok;             /* 0 if the solve failed / too few edges */
  int reserved;
  double chi2_initial, chi2_after1, chi2_final; /* robustified active chi2 */
  double lambda_final;
} flv_ba_stats;
/* Arrays are per stream with fixed strides given at flv_ba_reserve:
 *   poses [s][max_poses][7], landmarks [s][max_lms][3], edge_pose/edge_lm [s][max_edges] (indices
 *   into the stream's pose / landmark arrays), edge_uv [s][max_edges][2],
 *   edge_active [s][max_edges] (in: 0 = already culled; out: 0 for edges culled by this call).
 * Vertex ordering inside the solver follows g2o: poses by index, landmarks by index, edges in
 * array order (the host keeps them in g2o id / insertion order).
 * Window sizes: up to 100 poses (the reference's limit, vo_localmap.cpp:441-447).  Windows of <= 25 poses run on the
 * shared-memory solver (one cluster of 1/2/4 CTAs per window); a call that contains a larger window -- or, for FLV_MEM_DEVICE
 * calls, a context reserved for more than 25 poses -- runs on the global-memory solver (one cluster of 8 CTAs per window). */
int flv_ba_reserve(flv_ctx* ctx, int max_poses, int max_landmarks, int max_edges);
int flv_ba_optimize(flv_ctx* ctx, int n_streams, const flv_ba_problem* problems,
                    const flv_ba_params* prm, double* poses, double* landmarks,
                    const int* edge_pose, const int* edge_lm, const double* edge_uv,
                    uint8_t* edge_active, flv_ba_stats* stats, flv_memspace mem);

/* TEST AID, no GPU involved: the large-window solver's driver and phase code (csrc/ba_big.cu) executed sequentially on the host
 * over host arrays of ONE window, `emulated_threads` (32..2048, multiple of 32) playing the kernel's threads one after the
 * other.  Lets the CPU suite compare the arithmetic with the oracle; the product never calls it. */
int flv_ba_big_emulate_host(const flv_ba_problem* problem, const flv_ba_params* prm, double* poses, double* landmarks,
                            const int* edge_pose, const int* edge_lm, const double* edge_uv, uint8_t* edge_active,
                            flv_ba_stats* stats, int emulated_threads);

/* ---- per-landmark geometry (K6) -------------------------------------------------------------------
 * flv_depth_innovation replaces CameraFrame::depthInnovation (src/processing/camera_frame.cpp:271-330) and the
 * measurement sources it calls: recover3DPts_c_FromTriangulation (:236-270, Triangulation::triangulationPt,
 * src/processing/triangulation.cpp:9-39,80-97), the part of recover3DPts_c_FromStereo after its LK call (:133-179;
 * run flv_lk_track for the LK, then undistort on the host or pass rectified points) and recover3DPts_c_FromDepthImg
 * (:182-234; the caller samples the u16 depth image at round(lm_2d_plane) into depth_at_pts).
 * dummy_rand[s][i] holds the next d_rand values of the reference's rand() stream for that sequence, in draw order
 * (flvis_b200/host/glibc_rand.h reproduces glibc's sequence); n_rand_used[s] returns how many were consumed.
 * flv_reprojection_inliers replaces CameraFrame::calReprjInlierOutlier (:43-91).  All arrays [s][max_pts][...]. */
typedef struct {
  double fx, fy, cx, cy;   /* DepthCamera::cam0_fx .. cam0_cy */
  double P0[12], P1[12];   /* DepthCamera::P0_, P1_ (row-major 3x4), stereo only */
  int cam_type;            /* 0 = DEPTH_D435, 1 = STEREO_RECT / STEREO_UNRECT */
  double depth_scale;      /* cam_scale_factor (depth image units per metre) */
} flv_camera;
typedef struct {
  float iir_ratio;   /* dr_para1 */
  float range;       /* dr_para2 */
  int dummy_depth;   /* dr_para3 >= 0.5 */
} flv_depth_params;
int flv_depth_innovation(flv_ctx* ctx, int n_streams, const int* n_lms, const flv_camera* cam,
                         const flv_depth_params* prm, const double* T_c_w, const double* lm_2d_plane,
                         const double* lm_2d_undist, double* lm_3d_w, double* lm_3d_c, uint8_t* has_3d,
                         const double* first_obs_2d, const double* first_obs_pose, const double* stereo_pt1_undist,
                         const uint8_t* stereo_status, const uint16_t* depth_at_pts, const float* dummy_rand,
                         int* n_rand_used, flv_memspace mem);
int flv_reprojection_inliers(flv_ctx* ctx, int n_streams, const int* n_lms, const flv_camera* cam, const double* T_c_w,
                             const double* lm_2d_undist, const double* lm_3d_w, double sh_over_med,
                             uint8_t* is_inlier, double* mean_prjerr, flv_memspace mem);

/* ---- geometric verification (K11) -------------------------------------------------------------------
 * Batched RANSAC replacing cv::findFundamentalMat(from, to, FM_RANSAC, 5.0, 0.99) (src/processing/lkorb_tracking.cpp:134)
 * and cv::solvePnPRansac(p3d, p2d, K, D=0, rvec, tvec, guess, 100, 3.0, 0.99) (:170-177) for all streams in one launch.
 * Same models, thresholds and error measures as OpenCV; OpenCV's private RNG stream / LAPACK solvers are not
 * reproducible, so the inlier sets agree with cv2 statistically, not bit for bit (tests/test_ransac_gpu.py).  A fixed
 * budget of hypotheses (256 / 128) is evaluated in parallel: `confidence` and `max_iterations` are accepted for
 * signature compatibility only.  Arrays are [s][max_pts][..]; mask[s][i] = 1 for inliers; F[s][9] row-major maps
 * `from` to epipolar lines in `to` (x_to^T F x_from = 0); poses are [qx qy qz qw tx ty tz]; T_c_w_in is the pose prior
 * (IMU prediction or the previous frame's pose), K4 = fx fy cx cy per stream. */
typedef struct { double threshold_px; double confidence; int max_iterations; } flv_ransac_params;
int flv_fundamental_ransac(flv_ctx* ctx, int n_streams, const int* n_pts, const float* from_xy, const float* to_xy,
                           const flv_ransac_params* prm, uint8_t* mask, double* F, int* n_inliers, flv_memspace mem);
int flv_pnp_ransac(flv_ctx* ctx, int n_streams, const int* n_pts, const float* p3d, const float* p2d, const double* K4,
                   const double* T_c_w_in, const flv_ransac_params* prm, double* T_c_w_out, uint8_t* mask, int* n_inliers,
                   flv_memspace mem);

/* Run subsequent flv_ba_optimize calls on `cuda_stream` instead of the context stream (enable=1), the analogue of
 * FLVIS's separate local-map thread: the BA of keyframe k overlaps the tracking of the following frames.
 * enable=0 reverts to the context stream.  The caller orders the streams (events) as it needs. */
int flv_set_ba_stream(flv_ctx* ctx, void* cuda_stream, int enable);
/* Thread-block cluster size (CTAs per window: 1, 2 or 4) of flv_ba_optimize.  A window that optimises landmarks is split over
 * the cluster by landmark chunks; the partial reduced camera systems are summed through distributed shared memory.
 * host_mode_cluster: FLV_MEM_HOST calls (default 4, env FLV_BA_CLUSTER; pose-only problems always run on 1 CTA);
 * device_mode_cluster: FLV_MEM_DEVICE calls (default 1: the problems cannot be inspected on the host). */
int flv_set_ba_cluster(flv_ctx* ctx, int host_mode_cluster, int device_mode_cluster);

/* Debug: SM cycle counters of the last flv_ba_optimize for `stream`:
 * out16 = {chi2, build pose pass, schur products, cholesky, substitution, update, setup, -, schur init, schur staging,
 * build edge pass, build landmark pass, -...}. Synchronises. */
int flv_ba_profile(flv_ctx* ctx, int stream, long long* out16);
/* Debug: Levenberg-Marquardt trace of the last flv_ba_optimize for `stream`, one row per iteration (both phases, in order):
 * out[it][4] = {robustified chi2 at the end of the iteration, lambda after it, rho of its last trial, trials used}
 * (optimization_algorithm_levenberg.cpp:58-150).  `iterations_run` = flv_ba_stats.iterations_run of that solve.
 * Returns the number of rows written (<= cap_iters, <= 32).  Synchronises. */
int flv_ba_trace(flv_ctx* ctx, int stream, double* out, int cap_iters, int iterations_run);
/* Debug: residual + analytic Jacobians exactly as ba_kernel evaluates them (EdgeSE3ProjectXYZ::computeError /
 * linearizeOplus, types_six_dof_expmap.cpp:389-433) for n independent (pose[7], point[3], uv[2]) triples:
 * r[n][2], A[n][2][3] = d r / d point, B[n][2][6] = d r / d pose (rotation first, then translation).  Host arrays. */
int flv_ba_debug_edges(flv_ctx* ctx, int n, const double* poses7, const double* pts3, const double* uv2, const double* K4,
                       double* r, double* A, double* B);

#ifdef __cplusplus
}
#endif
#endif /* FLVIS_B200_H */
