// Host-side frame-to-frame tracker with the reference's class surface (one class per reference class, same method names and
// argument meaning; cv::Mat image arguments become a GPU context + pyramid slot):
//   LandMark / LandMarkInFrame   src/processing/include/landmark.h:8-36, src/processing/landmark.cpp:3-44
//   DepthCamera                  src/processing/include/depth_camera.h:6-77, depth_camera.cpp:92-150
//   CameraFrame                  src/processing/include/camera_frame.h:44-78, camera_frame.cpp:8-529
//   LKORBTracking                src/processing/include/lkorb_tracking.h:14-20, lkorb_tracking.cpp:9-202
//   OptimizeInFrame              src/processing/include/optimize_in_frame.h:31, optimize_in_frame.cpp:10-90
//   FeatureDEM                   src/processing/include/feature_dem.h:35-50
//   F2FTracking                  src/frontend/include/f2f_tracking.h:24-78, f2f_tracking.cpp:5-453
// cv::Mat arguments become raw image pointers; every OpenCV / g2o call goes to the GPU through the C ABI
// (include/flvis_b200.h), including the two RANSAC calls (K11: flv_fundamental_ransac / flv_pnp_ransac; caller callbacks
// or the host stand-ins of ransac.h can be selected instead).
// Sensor types: DEPTH_D435, STEREO_RECT, STEREO_UNRECT (EuRoC raw: LK runs on the distorted images, points go through
// cv::undistortPoints / cv::projectPoints restated in undistort.h); need_equal_hist = cv::equalizeHist on ingest (GPU).
#pragma once
#include <cstdint>
#include <deque>
#include <memory>
#include <vector>
#include "../../include/flvis_b200.h"
#include "glibc_rand.h"
#include "ransac.h"
#include "sophus_lite.h"
#include "undistort.h"
#include "vi_motion.h"
#include "local_map.h"

namespace flv {

enum TYPEOFCAMERA { DEPTH_D435 = 0, STEREO_RECT = 1, STEREO_UNRECT = 2 };

struct DepthCamera {
  int cam_type = DEPTH_D435;
  int img_w = 0, img_h = 0;
  double cam0_fx = 0, cam0_fy = 0, cam0_cx = 0, cam0_cy = 0;
  double cam1_fx = 0, cam1_fy = 0, cam1_cx = 0, cam1_cy = 0;   // K1 (stereo)
  double cam_scale_factor = 1000.0;
  double P0_[12] = {0}, P1_[12] = {0};                          // rectified projection matrices, row-major 3x4
  SE3 T_cam1_cam0;
  LensModel lens0, lens1;                                       // K0/D0/R0/P0, K1/D1/R1/P1 (STEREO_UNRECT: raw lens models)
  // DepthCamera::setDepthCamInfo (depth_camera.cpp:6-25): pinhole D435 colour camera + depth scale; K0_rect = K, D0_rect = 0
  void setDepthCamInfo(int w_in, int h_in, double fx, double fy, double cx, double cy, double scale_factor, int cam_type_in = DEPTH_D435);
  // DepthCamera::setSteroCamInfo (depth_camera.cpp:27-90), cv::Mat arguments as row-major arrays: K 3x3, D up to 14 OpenCV
  // coefficients (nD given), R 3x3, P 3x4.  K*_rect / D*_rect are accepted for signature parity (the reference stores them and
  // never reads them on the hot path: K_rect is the 3x3 part of P, D_rect is zero).  cam0/cam1 intrinsics come from P0 / P1.
  void setSteroCamInfo(int w_in, int h_in, const double* K0_in, const double* D0_in, int nD0, const double* K0_rect_in,
                       const double* D0_rect_in, const double* R0_in, const double* P0_in, const double* K1_in, const double* D1_in,
                       int nD1, const double* K1_rect_in, const double* D1_rect_in, const double* R1_in, const double* P1_in,
                       const SE3& T_c0_c1_in, int cam_type_in);
  SE3 T_cam0_cam1;
  Vec2 camera2pixel(const Vec3& p_c) const { return Vec2{cam0_fx * p_c[0] / p_c[2] + cam0_cx, cam0_fy * p_c[1] / p_c[2] + cam0_cy}; }
  static Vec3 world2cameraT_c_w(const Vec3& p_w, const SE3& T) { const Vec3 r = q_rot(T.q, p_w); return Vec3{r[0] + T.t[0], r[1] + T.t[1], r[2] + T.t[2]}; }
  static Vec3 camera2worldT_c_w(const Vec3& p_c, const SE3& T) { const SE3 Ti = T.inverse(); const Vec3 r = q_rot(Ti.q, p_c); return Vec3{r[0] + Ti.t[0], r[1] + Ti.t[1], r[2] + Ti.t[2]}; }
};

struct LandMarkInFrame {
  int64_t lm_id = 0;
  Vec3 lm_3d_w{0, 0, 0}, lm_3d_c{0, 0, 0};
  Vec2 lm_2d_plane{0, 0}, lm_2d_undistort{0, 0};
  bool has_3d = false, is_belong_to_kf = false, is_tracking_inlier = true;
  Vec2 lm_1st_obs_2d{0, 0};
  SE3 lm_1st_obs_frame_pose;
  bool hasDepthInf() const { return has_3d; }
};

class CameraFrame {
 public:
  int64_t frame_id = 0;
  double frame_time = 0;
  // img0 / img1 of the reference (cv::Mat members) live on the GPU: pyramid slots slot0 / slot1 of `ctx`
  flv_ctx* ctx = nullptr;
  int slot0 = 0, slot1 = 1;
  GlibcRand* rand_stream = nullptr;               // the sequence's rand() stream (dummy depths, camera_frame.cpp:153,168,198,222)
  std::vector<uint16_t> d_img;                    // depth image (DEPTH_D435), host copy for the nearest-pixel lookups
  DepthCamera d_camera;
  std::vector<LandMarkInFrame> landmarks;
  SE3 T_c_w;
  double reprojection_error = 0;
  void clear();
  // CameraFrame::calReprjInlierOutlier (camera_frame.cpp:43-91) on the GPU (flv_reprojection_inliers); `outlier` receives
  // the pixel positions of the landmarks flagged as outliers (the reference's debug list)
  int calReprjInlierOutlier(double& mean_prjerr, std::vector<Vec2>& outlier, double sh_over_med = 3.0);
  // CameraFrame::depthInnovation (camera_frame.cpp:271-330, incl. the left->right LK of recover3DPts_c_FromStereo :93-131)
  int depthInnovation(float iir_ratio, float range, bool dummy_depth);
  void eraseReprjOutlier();
  void eraseNoDepthPoint();
  int validLMCount() const;
  void updateLMState(const std::vector<uint8_t>& status);
  std::vector<Vec2> get2dPlaneVec() const;
  void getKeyFrameInf(std::vector<int64_t>& lm_id, std::vector<Vec2>& lm_2d, std::vector<Vec3>& lm_3d) const;
};

// ---- FeatureDEM <- src/processing/include/feature_dem.h:35-50, feature_dem.cpp:12-266 ------------------------------------
// cv::Mat arguments become (context, pyramid slot): the image is already on the GPU.
class FeatureDEM {
 public:
  FeatureDEM(flv_ctx* ctx, int image_width, int image_height, const double f_para[6]);
  int detect(int slot, std::vector<P2f>& newPts);
  int redetect(int slot, const std::vector<Vec2>& existedPts, std::vector<P2f>& newPts, int& newPtscount);
  const flv_feature_params& params() const { return prm_; }

 private:
  flv_ctx* ctx_;
  int width, height;
  flv_feature_params prm_{};
};

// RANSAC hooks (default = host stand-ins of ransac.h).  Return 0 on success.
typedef int (*flv_fmat_fn)(void* user, int n, const float* from_xy, const float* to_xy, uint8_t* mask_out);
typedef int (*flv_pnp_fn)(void* user, int n, const float* p3d, const float* p2d, const double* K4, int use_guess,
                          double* T_c_w_inout /*[qx qy qz qw tx ty tz]*/, int* inlier_idx_out, int* n_inliers_out);

// ---- LKORBTracking <- src/processing/include/lkorb_tracking.h:8-21, lkorb_tracking.cpp:9-202 ----------------------------
class LKORBTracking {
  int width, height;

 public:
  DepthCamera d_camera;
  LKORBTracking(int width_in, int height_in) : width(width_in), height(height_in) {}
  // lm2d_from / lm2d_to: pixel positions of the PnP inliers in both frames, outlier: positions rejected on the way
  // (the reference's debug lists, lkorb_tracking.cpp:190-200)
  bool tracking(CameraFrame& from, CameraFrame& to, SE3 T_c_w_guess, bool use_guess, std::vector<P2f>& lm2d_from,
                std::vector<P2f>& lm2d_to, std::vector<P2f>& outlier);
  // the two OpenCV RANSAC calls: device kernels (default), host stand-ins, or caller hooks
  flv_fmat_fn fmat_fn = nullptr; flv_pnp_fn pnp_fn = nullptr; void* hook_user = nullptr;
  bool host_ransac = false;
  int last_of_inliers = 0, last_f_inliers = 0, last_pnp_inliers = 0;   // "pnp|F|of" (:191)
};

// ---- OptimizeInFrame <- src/processing/include/optimize_in_frame.h:27-32, optimize_in_frame.cpp:10-90 --------------------
class OptimizeInFrame {
 public:
  static bool optimize(CameraFrame& frame);
};

class F2FTracking {
 public:
  enum TRACKINGSTATE { UnInit = 0, Tracking, TrackingFail };
  // one tracker = one sequence = one private flv_ctx with a single stream
  F2FTracking();
  ~F2FTracking();
  // dc, T_i_c0, feature_para(6), vi_para(6), dc_para(3) as in f2f_tracking.cpp:5-38
  int init(const DepthCamera& dc, const SE3& T_i_c0, const double feature_para[6], const double vi_para[6],
           const double dc_para[3], int skip_first_n_imgs, bool need_equal_hist, int device = 0);
  void imu_feed(double time, const Vec3& acc, const Vec3& gyro, Quat& q_w_i, Vec3& pos_w_i, Vec3& vel_w_i);
  // img1: u8 right image (stereo) or u16 depth image (DEPTH_D435); rows tightly packed
  int image_feed(double time, const uint8_t* img0, const void* img1, bool& new_keyframe, bool& reset_cmd);
  void set_ransac_hooks(flv_fmat_fn f, flv_pnp_fn p, void* user) { if (lkorb_tracker) { lkorb_tracker->fmat_fn = f; lkorb_tracker->pnp_fn = p; lkorb_tracker->hook_user = user; } }
  // raw lens model of camera `cam` (K, D[14], R rectification); P stays the rectified projection given at init
  int set_lens(int cam, const double* K4, const double* D14, const double* R9) {
    for (DepthCamera* dc : {&d_camera, lkorb_tracker ? &lkorb_tracker->d_camera : (DepthCamera*)nullptr}) {
      if (!dc) continue;
      LensModel& m = cam == 0 ? dc->lens0 : dc->lens1;
      m.fx = K4[0]; m.fy = K4[1]; m.cx = K4[2]; m.cy = K4[3];
      for (int i = 0; i < 14; ++i) m.k[i] = D14[i];
      for (int i = 0; i < 9; ++i) m.R[i] = R9[i];
    }
    return 0;
  }
  // F2FTracking::correction_feed (f2f_tracking.cpp:40-44): the local map's CorrectionInf for a past keyframe; applied at the
  // start of the next tracked frame (:189-219).  The reference's nodelet never calls it (vo_tracking.cpp:373-385 unpacks the
  // message and drops it), so it is off unless the integrator wires it.
  void correction_feed(double time, const CorrectionInfStruct& corr) { (void)time; correction_inf = corr; has_localmap_feedback = true; }
  void set_host_ransac(bool on) { if (lkorb_tracker) lkorb_tracker->host_ransac = on; }
  int set_equalize_hist(bool enable) { need_equal_hist = enable; return flv_set_equalize_hist(ctx_, enable ? 1 : 0); }

  std::shared_ptr<CameraFrame> curr_frame, last_frame;
  TRACKINGSTATE vo_tracking_state = UnInit;
  bool has_imu = false;
  int frameCount = 0;
  VIMOTION* vimotion = nullptr;
  const char* last_error() const { return err_; }
  const SE3& last_keyframe_pose() const { return T_c_w_last_keyframe; }
  FeatureDEM* feature_dem = nullptr;               // the modules the reference allocates in init (f2f_tracking.cpp:16-19)
  LKORBTracking* lkorb_tracker = nullptr;

 private:
  flv_ctx* ctx_ = nullptr;
  DepthCamera d_camera;
  float iir_ratio = 0.9f, range = 50.f;
  bool enable_dummy = false, need_equal_hist = false;
  int skip_n_imgs = 0, cam_type = DEPTH_D435;
  int64_t id_index = 100;                         // landmark.cpp:3 (process-global there, per sequence here)
  GlibcRand rand_;
  SE3 T_c_w_last_keyframe;
  struct ID_POSE { int64_t frame_id; SE3 T_c_w; };
  std::deque<ID_POSE> pose_records;
  int continus_tracking_fail_cnt = 0, fail_cnt = 0;
  bool has_localmap_feedback = false;
  CorrectionInfStruct correction_inf;
  void apply_localmap_feedback();
  char err_[256] = {0};
  int slot_toggle = 0;

  LandMarkInFrame make_landmark(const Vec2& pt2d, const Vec2& pt2d_undist, const SE3& T_c_w, bool is_inlier);
  bool init_frame();
};

}  // namespace flv
