#include "f2f_tracking.h"
#include "nvtx_range.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace flv {

namespace {
constexpr int MAX_PTS = 512;

SE3 se3_from_R(const double* R, const Vec3& t) { return SE3(R_to_q(R), t); }
Pose7 to7(const SE3& T) { return Pose7{T.q.x, T.q.y, T.q.z, T.q.w, T.t[0], T.t[1], T.t[2]}; }
SE3 from7(const double* p) { return SE3(Quat{p[3], p[0], p[1], p[2]}, Vec3{p[4], p[5], p[6]}); }

Vec3 so3_log(const Quat& q) {          // Sophus SO3::logAndTheta (so3.cpp:127-164), SMALL_EPS = 1e-10
  const double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z), w = q.w;
  double f;
  if (n < 1e-10) f = 2. / w - 2. * (n * n) / (w * w * w);
  else f = 2 * std::atan(n / w) / n;
  return Vec3{f * q.x, f * q.y, f * q.z};
}
}  // namespace

// ---- CameraFrame --------------------------------------------------------------------------------
void CameraFrame::clear() {                                     // camera_frame.cpp:8-15
  T_c_w = SE3();
  d_img.clear();
  landmarks.clear();
}
void CameraFrame::eraseReprjOutlier() {                         // :18-28
  for (int i = (int)landmarks.size() - 1; i >= 0; i--)
    if (!landmarks[i].is_tracking_inlier) landmarks.erase(landmarks.begin() + i);
}
void CameraFrame::eraseNoDepthPoint() {                         // :30-40
  for (int i = (int)landmarks.size() - 1; i >= 0; i--)
    if (!landmarks[i].has_3d) landmarks.erase(landmarks.begin() + i);
}
int CameraFrame::validLMCount() const {                         // :399-413
  int ret = 0;
  for (const LandMarkInFrame& lm : landmarks) ret += (lm.has_3d && lm.is_tracking_inlier);
  return ret;
}
void CameraFrame::updateLMState(const std::vector<uint8_t>& status) {   // :445-461
  int indexLM = 0;
  for (LandMarkInFrame& lm : landmarks)
    if (lm.has_3d && lm.is_tracking_inlier) {
      if (status[indexLM] == 0) lm.is_tracking_inlier = false;
      indexLM += 1;
    }
}
std::vector<Vec2> CameraFrame::get2dPlaneVec() const {
  std::vector<Vec2> r;
  for (const LandMarkInFrame& lm : landmarks) r.push_back(lm.lm_2d_plane);
  return r;
}
void CameraFrame::getKeyFrameInf(std::vector<int64_t>& lm_id, std::vector<Vec2>& lm_2d, std::vector<Vec3>& lm_3d) const {  // :515-529
  lm_id.clear(); lm_2d.clear(); lm_3d.clear();
  for (const LandMarkInFrame& lm : landmarks)
    if (lm.has_3d && lm.is_tracking_inlier) { lm_3d.push_back(lm.lm_3d_w); lm_2d.push_back(lm.lm_2d_undistort); lm_id.push_back(lm.lm_id); }
}

// ---- DepthCamera ---------------------------------------------------------------------------------
void DepthCamera::setDepthCamInfo(int w_in, int h_in, double fx, double fy, double cx, double cy, double scale_factor, int cam_type_in) {
  img_w = w_in; img_h = h_in;
  cam0_fx = fx; cam0_fy = fy; cam0_cx = cx; cam0_cy = cy;
  lens0 = LensModel();                                             // K0_rect = K, D0_rect = 0 (depth_camera.cpp:20-22)
  lens0.fx = fx; lens0.fy = fy; lens0.cx = cx; lens0.cy = cy;
  lens0.P[0] = fx; lens0.P[2] = cx; lens0.P[5] = fy; lens0.P[6] = cy;
  for (int i = 0; i < 12; ++i) P0_[i] = lens0.P[i];
  cam_scale_factor = scale_factor;
  cam_type = cam_type_in;
}
void DepthCamera::setSteroCamInfo(int w_in, int h_in, const double* K0_in, const double* D0_in, int nD0, const double*, const double*,
                                  const double* R0_in, const double* P0_in, const double* K1_in, const double* D1_in, int nD1,
                                  const double*, const double*, const double* R1_in, const double* P1_in, const SE3& T_c0_c1_in,
                                  int cam_type_in) {
  img_w = w_in; img_h = h_in;
  auto fill = [](LensModel& m, const double* K, const double* D, int nD, const double* R, const double* P) {
    m = LensModel();
    m.fx = K[0]; m.fy = K[4]; m.cx = K[2]; m.cy = K[5];
    for (int i = 0; i < nD && i < 14; ++i) m.k[i] = D[i];
    for (int i = 0; i < 9; ++i) m.R[i] = R[i];
    for (int i = 0; i < 12; ++i) m.P[i] = P[i];
  };
  fill(lens0, K0_in, D0_in, nD0, R0_in, P0_in);
  fill(lens1, K1_in, D1_in, nD1, R1_in, P1_in);
  T_cam0_cam1 = T_c0_c1_in;
  T_cam1_cam0 = T_cam0_cam1.inverse();
  for (int i = 0; i < 12; ++i) { P0_[i] = P0_in[i]; P1_[i] = P1_in[i]; }
  cam0_fx = P0_[0]; cam0_fy = P0_[5]; cam0_cx = P0_[2]; cam0_cy = P0_[6];     // depth_camera.cpp:73-82
  cam1_fx = P1_[0]; cam1_fy = P1_[5]; cam1_cx = P1_[2]; cam1_cy = P1_[6];
  cam_type = cam_type_in;
}

// ---- F2FTracking ---------------------------------------------------------------------------------
F2FTracking::F2FTracking() {}
F2FTracking::~F2FTracking() { delete vimotion; delete feature_dem; delete lkorb_tracker; if (ctx_) flv_destroy(ctx_); }

int F2FTracking::init(const DepthCamera& dc, const SE3& T_i_c0, const double feature_para[6], const double vi_para[6],
                      const double dc_para[3], int skip_first_n_imgs, bool need_equal_hist_in, int device) {   // f2f_tracking.cpp:5-38
  skip_n_imgs = skip_first_n_imgs; need_equal_hist = need_equal_hist_in;
  int rc = flv_create(&ctx_, device, 1, dc.img_w, dc.img_h, MAX_PTS);
  if (rc) { snprintf(err_, sizeof(err_), "flv_create: %s", flv_last_error(ctx_)); return rc; }
  if (need_equal_hist && (rc = flv_set_equalize_hist(ctx_, 1))) return rc;    // f2f_tracking.cpp:125-145
  if ((rc = flv_ba_reserve(ctx_, 1, MAX_PTS, MAX_PTS))) return rc;            // workspace of the per-frame pose-only BA, once
  feature_dem = new FeatureDEM(ctx_, dc.img_w, dc.img_h, feature_para);       // f2f_tracking.cpp:16-19
  lkorb_tracker = new LKORBTracking(dc.img_w, dc.img_h);
  lkorb_tracker->d_camera = dc;
  vimotion = new VIMOTION(T_i_c0, 9.81, vi_para[0], vi_para[1], vi_para[2], vi_para[3]);
  curr_frame = std::make_shared<CameraFrame>();
  last_frame = std::make_shared<CameraFrame>();
  cam_type = dc.cam_type;
  d_camera = curr_frame->d_camera = last_frame->d_camera = dc;
  curr_frame->slot0 = 0; curr_frame->slot1 = 1; last_frame->slot0 = 2; last_frame->slot1 = 3;
  curr_frame->ctx = last_frame->ctx = ctx_;
  curr_frame->rand_stream = last_frame->rand_stream = &rand_;
  iir_ratio = (float)dc_para[0];
  range = (float)dc_para[1];
  enable_dummy = !(dc_para[2] < 0.5);
  frameCount = 0;
  vo_tracking_state = UnInit;
  return FLV_OK;
}

void F2FTracking::imu_feed(double time, const Vec3& acc, const Vec3& gyro, Quat& q_w_i, Vec3& pos_w_i, Vec3& vel_w_i) {   // :46-57
  IMUSTATE s; s.timestamp = time; s.acc_raw = acc; s.gyro_raw = gyro;
  if (!vimotion->imu_initialized) { has_imu = true; vimotion->viIMUinitialization(s, q_w_i, pos_w_i, vel_w_i); }
  else vimotion->viIMUPropagation(s, q_w_i, pos_w_i, vel_w_i);
}

LandMarkInFrame F2FTracking::make_landmark(const Vec2& pt2d, const Vec2& pt2d_undist, const SE3& T_c_w, bool is_inlier) {   // landmark.cpp:18-39
  LandMarkInFrame lm;
  lm.lm_id = id_index++;
  lm.lm_1st_obs_2d = lm.lm_2d_undistort = pt2d_undist;
  lm.lm_2d_plane = pt2d;
  lm.lm_1st_obs_frame_pose = T_c_w;
  lm.is_tracking_inlier = is_inlier;
  return lm;
}

// ---- FeatureDEM ------------------------------------------------------------------------------------
FeatureDEM::FeatureDEM(flv_ctx* ctx, int image_width, int image_height, const double f_para[6])   // feature_dem.cpp:12-54
    : ctx_(ctx), width(image_width), height(image_height) {
  prm_.max_region_feature_num = (int)f_para[0];
  prm_.min_region_feature_num = (int)f_para[1];
  prm_.boundary_dis = (int)std::floor(f_para[2] / 2.0);
  prm_.gftt_num = (int)f_para[3];
  prm_.gftt_ql = f_para[4];
  prm_.gftt_dis = (int)f_para[5];
}
int FeatureDEM::detect(int slot, std::vector<P2f>& newPts) {                                       // :215-266
  std::vector<float> out(2 * MAX_PTS); int n = 0;
  int rc = flv_feature_detect(ctx_, slot, 1, &prm_, out.data(), &n, FLV_MEM_HOST);
  if (rc) return rc;
  newPts.resize(n);
  memcpy(newPts.data(), out.data(), (size_t)n * 8);
  return FLV_OK;
}
int FeatureDEM::redetect(int slot, const std::vector<Vec2>& existedPts, std::vector<P2f>& newPts, int& newPtscount) {   // :124-213
  std::vector<double> ex(2 * MAX_PTS, 0.0);
  int ne = (int)std::min<size_t>(existedPts.size(), MAX_PTS), n = 0;
  for (int i = 0; i < ne; ++i) { ex[2 * i] = existedPts[i][0]; ex[2 * i + 1] = existedPts[i][1]; }
  std::vector<float> out(2 * MAX_PTS);
  int rc = flv_feature_redetect(ctx_, slot, 1, &prm_, ex.data(), &ne, out.data(), &n, FLV_MEM_HOST);
  if (rc) return rc;
  newPts.resize(n);
  memcpy(newPts.data(), out.data(), (size_t)n * 8);
  newPtscount = n;
  return FLV_OK;
}

int F2FTracking::image_feed(double time, const uint8_t* img0, const void* img1, bool& new_keyframe, bool& reset_cmd) {   // :59-400
  new_keyframe = false; reset_cmd = false;
  NvtxStages nvtx;                                   // NVTX range per step of the reference's image_feed
  nvtx.next("flv: ingest + pyramids");
  frameCount++;
  last_frame.swap(curr_frame);
  curr_frame->clear();
  curr_frame->frame_id = frameCount;
  curr_frame->frame_time = time;
  const int w = d_camera.img_w, h = d_camera.img_h;
  int rc = flv_upload_images(ctx_, curr_frame->slot0, 1, img0, w, (size_t)w * h, FLV_MEM_HOST);
  if (rc) return rc;
  if (cam_type == DEPTH_D435) curr_frame->d_img.assign((const uint16_t*)img1, (const uint16_t*)img1 + (size_t)w * h);
  else if ((rc = flv_upload_images(ctx_, curr_frame->slot1, 1, (const uint8_t*)img1, w, (size_t)w * h, FLV_MEM_HOST))) return rc;
  if (skip_n_imgs > 0) { skip_n_imgs--; return FLV_OK; }
  if ((rc = flv_build_pyramid(ctx_, curr_frame->slot0, 1))) return rc;
  if (cam_type != DEPTH_D435 && (rc = flv_build_pyramid(ctx_, curr_frame->slot1, 1))) return rc;

  switch (vo_tracking_state) {
    case UnInit: {
      const double R_w_c[9] = {0, 0, 1, -1, 0, 0, 0, -1, 0};
      curr_frame->T_c_w = se3_from_R(R_w_c, Vec3{0, 0, 0}).inverse();
      if (has_imu) {
        if (vimotion->imu_initialized) {
          Quat q_init;
          vimotion->viVisiontrigger(q_init);
          double Ra[9], Rb[9], Rc[9];
          q_to_R(q_init, Ra); q_to_R(vimotion->T_i_c.q, Rb);
          for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rc[3 * i + j] = Ra[3 * i] * Rb[j] + Ra[3 * i + 1] * Rb[3 + j] + Ra[3 * i + 2] * Rb[6 + j];
          curr_frame->T_c_w = se3_from_R(Rc, Vec3{0, 0, 0}).inverse();
        } else break;
      }
      if (init_frame()) { new_keyframe = true; vo_tracking_state = Tracking; }
      break;
    }
    case Tracking: {
      // STEP1: local-map feedback (:189-219); only ever set through correction_feed, which the reference's nodelet never calls
      nvtx.next("flv: STEP1-3 LKORBTracking::tracking (LK, F RANSAC, PnP RANSAC)");
      if (has_localmap_feedback) apply_localmap_feedback();
      SE3 imu_guess; bool has_imu_guess = false;
      if (has_imu) has_imu_guess = vimotion->viGetCorrFrameState(time, imu_guess);
      std::vector<P2f> lm2d_from, lm2d_to, outlier_tracking;
      const bool tracking_success = lkorb_tracker->tracking(*last_frame, *curr_frame, imu_guess, has_imu_guess, lm2d_from, lm2d_to, outlier_tracking);
      if (!tracking_success) {
        continus_tracking_fail_cnt++;
        last_frame.swap(curr_frame);
        if (continus_tracking_fail_cnt >= 2) { vo_tracking_state = TrackingFail; continus_tracking_fail_cnt = 0; }
        break;
      }
      continus_tracking_fail_cnt = 0;
      nvtx.next("flv: STEP4 OptimizeInFrame (pose-only BA)");
      if (has_imu) vimotion->viVisionRPCompensation(curr_frame->frame_time, curr_frame->T_c_w);
      if (!OptimizeInFrame::optimize(*curr_frame)) {
        continus_tracking_fail_cnt++;
        last_frame.swap(curr_frame);
        if (continus_tracking_fail_cnt >= 2) { vo_tracking_state = TrackingFail; continus_tracking_fail_cnt = 0; }
        break;
      }
      nvtx.next("flv: STEP5 reprojection cull");
      std::vector<Vec2> outlier_reproject;
      double mean_reprojection_error = 0;
      if ((rc = curr_frame->calReprjInlierOutlier(mean_reprojection_error, outlier_reproject, 1.5))) return rc;
      curr_frame->reprojection_error = mean_reprojection_error;
      curr_frame->eraseReprjOutlier();
      if (has_imu)
        vimotion->viCorrectionFromVision(curr_frame->frame_time, curr_frame->T_c_w, last_frame->frame_time, last_frame->T_c_w,
                                         curr_frame->reprojection_error);
      nvtx.next("flv: STEP6 FeatureDEM redetect");
      std::vector<P2f> pts2d;
      const int orig_size = (int)curr_frame->landmarks.size();
      int newPtsCount = 0;
      if ((rc = feature_dem->redetect(curr_frame->slot0, curr_frame->get2dPlaneVec(), pts2d, newPtsCount))) return rc;
      const bool add_as_inliers = orig_size < 60;
      for (const P2f& p : pts2d) {      // DEPTH_D435 / STEREO_RECT: pts2d_undistort = pts2d (:294-299); UNRECT: undistortPoints (:300-303)
        P2f u = p;
        if (cam_type == STEREO_UNRECT) undistort_point(d_camera.lens0, p.x, p.y, u.x, u.y);
        curr_frame->landmarks.push_back(make_landmark(Vec2{p.x, p.y}, Vec2{u.x, u.y}, curr_frame->T_c_w, add_as_inliers));
      }
      nvtx.next("flv: STEP7 depth innovation (left->right LK, triangulation)");
      if ((rc = curr_frame->depthInnovation(iir_ratio, range, enable_dummy))) return rc;
      curr_frame->eraseNoDepthPoint();
      pose_records.push_back(ID_POSE{curr_frame->frame_id, curr_frame->T_c_w});
      if (pose_records.size() >= 1000) pose_records.pop_front();
      const SE3 T_diff = T_c_w_last_keyframe * curr_frame->T_c_w.inverse();
      const Vec3 r = so3_log(T_diff.q);
      const double t_norm = std::fabs(T_diff.t[0]) + std::fabs(T_diff.t[1]) + std::fabs(T_diff.t[2]);
      const double r_norm = std::fabs(r[0]) + std::fabs(r[1]) + std::fabs(r[2]);
      if (frameCount < 40 && (frameCount % 5) == 0) { new_keyframe = true; T_c_w_last_keyframe = curr_frame->T_c_w; }
      if (t_norm >= 0.05 || r_norm >= 0.2) { new_keyframe = true; T_c_w_last_keyframe = curr_frame->T_c_w; }
      break;
    }
    case TrackingFail: {
      fail_cnt++;
      if ((fail_cnt % 3) == 0) {
        if (vimotion->viGetCorrFrameState(curr_frame->frame_time, curr_frame->T_c_w)) {
          if (init_frame()) { new_keyframe = true; vo_tracking_state = Tracking; }
          else last_frame.swap(curr_frame);
        } else last_frame.swap(curr_frame);
        fail_cnt = 0;
      } else {
        last_frame.swap(curr_frame);
        if ((fail_cnt % 2) == 0) reset_cmd = true;
      }
      break;
    }
  }
  return FLV_OK;
}

void F2FTracking::apply_localmap_feedback() {                       // f2f_tracking.cpp:189-219
  const int corr_id = (int)correction_inf.frame_id;                  // `int corr_id = correction_inf.frame_id` (kept)
  int old_pose_idx = 0;
  for (int i = (int)pose_records.size() - 1; i >= 0; i--)
    if (pose_records[i].frame_id == corr_id) { old_pose_idx = i; break; }
  if (!pose_records.empty()) {
    const SE3 old_T_c_w_inv = pose_records[old_pose_idx].T_c_w.inverse();
    const Pose7& u = correction_inf.T_c_w;
    const SE3 update_T_c_w(Quat{u[3], u[0], u[1], u[2]}, Vec3{u[4], u[5], u[6]});
    for (size_t i = old_pose_idx; i < pose_records.size(); i++) {
      const SE3 T_diff = pose_records[i].T_c_w * old_T_c_w_inv;
      pose_records[i].T_c_w = T_diff * update_T_c_w;
    }
    const SE3 T_diff = last_frame->T_c_w * old_T_c_w_inv;
    last_frame->T_c_w = T_diff * update_T_c_w;
    // correctLMP3DWByLMP3DCandT (camera_frame.cpp:332-342) iterates its landmarks BY VALUE: no effect, nothing to do here
    // forceCorrectLM3DW (:344-359): ids are compared through an `int` (kept), first match only
    for (int i = 0; i < correction_inf.lm_count && i < (int)correction_inf.lm_id.size(); i++) {
      const int id = (int)correction_inf.lm_id[i];
      for (LandMarkInFrame& lm : last_frame->landmarks)
        if (lm.lm_id == id) { lm.lm_3d_w = correction_inf.lm_3d[i]; break; }
    }
    // forceMarkOutlier (:361-376): every match, no break
    for (int i = 0; i < correction_inf.lm_outlier_count && i < (int)correction_inf.lm_outlier_id.size(); i++) {
      const int id = (int)correction_inf.lm_outlier_id[i];
      for (LandMarkInFrame& lm : last_frame->landmarks)
        if (lm.lm_id == id) lm.is_tracking_inlier = false;
    }
  }
  has_localmap_feedback = false;
}

bool F2FTracking::init_frame() {                                   // f2f_tracking.cpp:402-453
  std::vector<P2f> pts2d;
  if (feature_dem->detect(curr_frame->slot0, pts2d)) return false;
  // DEPTH: undistorted = plane.  STEREO_RECT: cv::undistortPoints(K0, D0=0, R0=I, P0) is the identity map up to rounding
  for (const P2f& p : pts2d) {
    P2f u = p;
    if (cam_type == STEREO_UNRECT) undistort_point(d_camera.lens0, p.x, p.y, u.x, u.y);      // :425
    curr_frame->landmarks.push_back(make_landmark(Vec2{p.x, p.y}, Vec2{u.x, u.y}, curr_frame->T_c_w, true));
  }
  if (curr_frame->depthInnovation(iir_ratio, range, enable_dummy)) return false;
  curr_frame->eraseNoDepthPoint();
  if (curr_frame->validLMCount() > 30) {
    pose_records.push_back(ID_POSE{curr_frame->frame_id, curr_frame->T_c_w});
    T_c_w_last_keyframe = curr_frame->T_c_w;
    return true;
  }
  return false;
}

// ---- LKORBTracking ---------------------------------------------------------------------------------
bool LKORBTracking::tracking(CameraFrame& from, CameraFrame& to, SE3 T_c_w_guess, bool use_guess, std::vector<P2f>& lm2d_from,
                             std::vector<P2f>& lm2d_to, std::vector<P2f>& outlier) {   // lkorb_tracking.cpp:9-202
  flv_ctx* ctx_ = from.ctx;
  const int cam_type = d_camera.cam_type;
  lm2d_from.clear(); lm2d_to.clear(); outlier.clear();
  const int n = (int)std::min<size_t>(from.landmarks.size(), MAX_PTS);
  std::vector<P2f> from_plane(n), tracked_plane(n), from_und(n);
  std::vector<P3f> from_p3d(n);
  for (int i = 0; i < n; ++i) {                                  // getAll2dPlaneUndistort3d_cvPf (float copies)
    const LandMarkInFrame& lm = from.landmarks[i];
    from_plane[i] = P2f{(float)lm.lm_2d_plane[0], (float)lm.lm_2d_plane[1]};
    from_und[i] = P2f{(float)lm.lm_2d_undistort[0], (float)lm.lm_2d_undistort[1]};
    from_p3d[i] = P3f{(float)lm.lm_3d_w[0], (float)lm.lm_3d_w[1], (float)lm.lm_3d_w[2]};
  }
  tracked_plane = from_plane;
  if (use_guess) {                                               // :38-63
    if (cam_type == STEREO_UNRECT) {                             // cv::projectPoints(K0, D0): back into the distorted image
      double Rg[9]; q_to_R(T_c_w_guess.q, Rg);
      const double tg[3] = {T_c_w_guess.t[0], T_c_w_guess.t[1], T_c_w_guess.t[2]};
      for (int i = 0; i < n; ++i)
        project_point(d_camera.lens0, Rg, tg, from_p3d[i].x, from_p3d[i].y, from_p3d[i].z, tracked_plane[i].x, tracked_plane[i].y);
    } else {                                                     // pinhole (DEPTH_D435; STEREO_RECT has D0 = 0)
      for (int i = 0; i < n; ++i) {
        const Vec3 pc = DepthCamera::world2cameraT_c_w(Vec3{from_p3d[i].x, from_p3d[i].y, from_p3d[i].z}, T_c_w_guess);
        const Vec2 px = from.d_camera.camera2pixel(pc);
        tracked_plane[i] = P2f{(float)px[0], (float)px[1]};
      }
    }
  }
  std::vector<float> prev(2 * MAX_PTS, 0.f), init(2 * MAX_PTS, 0.f), next(2 * MAX_PTS), err(MAX_PTS);
  std::vector<uint8_t> st(MAX_PTS);
  memcpy(prev.data(), from_plane.data(), (size_t)n * 8);
  memcpy(init.data(), tracked_plane.data(), (size_t)n * 8);
  flv_lk_params lk{31, 10, 30, 1e-3, 1e-4};
  if (flv_lk_track(ctx_, from.slot0, to.slot0, 1, &n, prev.data(), init.data(), next.data(), st.data(), err.data(), &lk, FLV_MEM_HOST))
    return false;
  memcpy(tracked_plane.data(), next.data(), (size_t)n * 8);
  std::vector<P2f> tracked_und = tracked_plane;                  // DEPTH_D435 / STEREO_RECT (:78-85)
  if (cam_type == STEREO_UNRECT)                                 // cv::undistortPoints(K0, D0, R0, P0) (:86-89)
    for (int i = 0; i < n; ++i) undistort_point(d_camera.lens0, tracked_plane[i].x, tracked_plane[i].y, tracked_und[i].x, tracked_und[i].y);
  to.landmarks.clear();
  const int wlim = to.d_camera.img_w - 1, hlim = to.d_camera.img_h - 1;
  int of_inlier_cnt = 0;
  for (int i = n - 1; i >= 0; i--) {                             // :98-119 -- `to.landmarks` ends up REVERSED
    if (st[i] == 1 && tracked_plane[i].x > 0 && tracked_plane[i].y > 0 && tracked_plane[i].x < wlim && tracked_plane[i].y < hlim) {
      of_inlier_cnt++;
      LandMarkInFrame lm = from.landmarks[i];
      lm.lm_2d_plane = Vec2{tracked_plane[i].x, tracked_plane[i].y};
      lm.lm_2d_undistort = Vec2{tracked_und[i].x, tracked_und[i].y};
      to.landmarks.push_back(lm);
    } else {
      from_plane.erase(from_plane.begin() + i); from_p3d.erase(from_p3d.begin() + i);
      tracked_plane.erase(tracked_plane.begin() + i); from_und.erase(from_und.begin() + i);
      tracked_und.erase(tracked_und.begin() + i);
    }
  }
  last_of_inliers = of_inlier_cnt; last_f_inliers = last_pnp_inliers = 0;
  if (of_inlier_cnt < 10) return false;
  // STEP2: F-matrix consistency; mask index i is applied to to.landmarks[i] (mirrored order, kept: :138-149)
  const int m = (int)from_und.size();
  std::vector<uint8_t> maskF(m, 0);
  if (fmat_fn) {
    if (fmat_fn(hook_user, m, &from_und[0].x, &tracked_und[0].x, maskF.data())) return false;
  } else if (host_ransac) {
    double F[9];
    find_fundamental_ransac(from_und, tracked_und, 5.0, 0.99, maskF, F);
  } else {                                                       // K11 on the device (flv_fundamental_ransac)
    std::vector<float> a(2 * MAX_PTS, 0.f), b(2 * MAX_PTS, 0.f);
    std::vector<uint8_t> mk(MAX_PTS, 0);
    int nn = std::min(m, MAX_PTS), ni = 0;
    memcpy(a.data(), from_und.data(), (size_t)nn * 8); memcpy(b.data(), tracked_und.data(), (size_t)nn * 8);
    double F[9];
    const flv_ransac_params rp{5.0, 0.99, 1000};
    if (flv_fundamental_ransac(ctx_, 1, &nn, a.data(), b.data(), &rp, mk.data(), F, &ni, FLV_MEM_HOST)) return false;
    for (int i = 0; i < nn; ++i) maskF[i] = mk[i];
  }
  for (int i = 0; i < m; i++)
    if (maskF[i] == 0) to.landmarks[i].is_tracking_inlier = false;
  int F_inlier_cnt = 0;
  for (const LandMarkInFrame& lm : to.landmarks) F_inlier_cnt += lm.is_tracking_inlier;
  last_f_inliers = F_inlier_cnt;
  if (F_inlier_cnt < 10) return false;
  // STEP3: PnP RANSAC on (has depth && inlier) pairs
  std::vector<P2f> p2d; std::vector<P3f> p3d;
  for (const LandMarkInFrame& lm : to.landmarks)
    if (lm.has_3d && lm.is_tracking_inlier) {
      p2d.push_back(P2f{(float)lm.lm_2d_undistort[0], (float)lm.lm_2d_undistort[1]});
      p3d.push_back(P3f{(float)lm.lm_3d_w[0], (float)lm.lm_3d_w[1], (float)lm.lm_3d_w[2]});
    }
  const double K[4] = {d_camera.cam0_fx, d_camera.cam0_fy, d_camera.cam0_cx, d_camera.cam0_cy};
  Pose7 T = to7(use_guess ? T_c_w_guess : from.T_c_w);
  std::vector<int> inl;
  if (pnp_fn) {
    inl.resize(p2d.size());
    int ninl = 0;
    if (p2d.empty() || pnp_fn(hook_user, (int)p2d.size(), &p3d[0].x, &p2d[0].x, K, use_guess ? 1 : 0, T.data(), inl.data(), &ninl)) return false;
    inl.resize(ninl);
  } else if (host_ransac) {
    solve_pnp_ransac(p3d, p2d, K, T, 100, 3.0, 0.99, inl);
  } else {                                                       // K11 on the device (flv_pnp_ransac), prior = guess / last pose
    std::vector<float> x3(3 * MAX_PTS, 0.f), x2(2 * MAX_PTS, 0.f);
    std::vector<uint8_t> mk(MAX_PTS, 0);
    int nn = std::min((int)p2d.size(), MAX_PTS), ni = 0;
    if (nn > 0) { memcpy(x3.data(), p3d.data(), (size_t)nn * 12); memcpy(x2.data(), p2d.data(), (size_t)nn * 8); }
    Pose7 Tout = T;
    const flv_ransac_params rp{3.0, 0.99, 100};
    if (flv_pnp_ransac(ctx_, 1, &nn, x3.data(), x2.data(), K, T.data(), &rp, Tout.data(), mk.data(), &ni, FLV_MEM_HOST)) return false;
    T = Tout;
    for (int i = 0; i < nn; ++i) if (mk[i]) inl.push_back(i);
  }
  std::vector<uint8_t> mask_pnp(p2d.size(), 0);
  for (int k : inl) mask_pnp[k] = 1;
  to.updateLMState(mask_pnp);
  to.T_c_w = from7(T.data());
  last_pnp_inliers = (int)inl.size();
  for (const LandMarkInFrame& lm : to.landmarks) {           // debug lists (:190-200)
    const P2f p{(float)lm.lm_2d_plane[0], (float)lm.lm_2d_plane[1]};
    if (lm.is_tracking_inlier) lm2d_to.push_back(p); else outlier.push_back(p);
  }
  for (const LandMarkInFrame& lm : from.landmarks) lm2d_from.push_back(P2f{(float)lm.lm_2d_plane[0], (float)lm.lm_2d_plane[1]});
  return inl.size() >= 10;
}

// ---- OptimizeInFrame -------------------------------------------------------------------------------
bool OptimizeInFrame::optimize(CameraFrame& frame) {                // optimize_in_frame.cpp:10-90
  flv_ctx* ctx_ = frame.ctx;
  const DepthCamera& d_camera = frame.d_camera;
  std::vector<const LandMarkInFrame*> lms;
  for (const LandMarkInFrame& lm : frame.landmarks) if (lm.has_3d && lm.is_tracking_inlier) lms.push_back(&lm);
  const int n = (int)lms.size();
  if (n < 10) return false;
  if (flv_ba_reserve(ctx_, 1, MAX_PTS, MAX_PTS)) return false;         // no-op once reserved (F2FTracking::init does it)
  std::vector<double> pts(3 * MAX_PTS, 0.0), uv(2 * MAX_PTS, 0.0);
  std::vector<int> ep(MAX_PTS, 0), el(MAX_PTS, 0);
  std::vector<uint8_t> act(MAX_PTS, 0);
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < 3; ++k) pts[3 * i + k] = lms[i]->lm_3d_w[k];
    uv[2 * i] = lms[i]->lm_2d_undistort[0]; uv[2 * i + 1] = lms[i]->lm_2d_undistort[1];
    el[i] = i; act[i] = 1;
  }
  Pose7 pose = g2o_pose_from_quat(to7(frame.T_c_w));
  flv_ba_problem pb{1, n, n, -1, 1, d_camera.cam0_fx, d_camera.cam0_fy, d_camera.cam0_cx, d_camera.cam0_cy};
  flv_ba_params prm{2, 2, 1.0, 3.0, 10, 0};
  flv_ba_stats st;
  if (flv_ba_optimize(ctx_, 1, &pb, &prm, pose.data(), pts.data(), ep.data(), el.data(), uv.data(), act.data(), &st, FLV_MEM_HOST))
    return false;
  if (!st.ok) return false;                                        // < 10 edges after the chi2 cull: pose left untouched
  frame.T_c_w = from7(pose.data());
  return true;
}

int CameraFrame::calReprjInlierOutlier(double& mean_prjerr, std::vector<Vec2>& outlier, double sh_over_med) {   // camera_frame.cpp:43-91
  CameraFrame& fr = *this;
  flv_ctx* ctx_ = ctx;
  outlier.clear();
  const int n = (int)std::min<size_t>(fr.landmarks.size(), MAX_PTS);
  std::vector<double> und(2 * MAX_PTS, 0.0), p3(3 * MAX_PTS, 0.0);
  for (int i = 0; i < n; ++i) {
    und[2 * i] = fr.landmarks[i].lm_2d_undistort[0]; und[2 * i + 1] = fr.landmarks[i].lm_2d_undistort[1];
    for (int k = 0; k < 3; ++k) p3[3 * i + k] = fr.landmarks[i].lm_3d_w[k];
  }
  flv_camera cam{}; cam.fx = d_camera.cam0_fx; cam.fy = d_camera.cam0_fy; cam.cx = d_camera.cam0_cx; cam.cy = d_camera.cam0_cy;
  const Pose7 T = to7(fr.T_c_w);
  std::vector<uint8_t> inl(MAX_PTS, 0);
  double mean = 0;
  int rc = flv_reprojection_inliers(ctx_, 1, &n, &cam, T.data(), und.data(), p3.data(), sh_over_med, inl.data(), &mean, FLV_MEM_HOST);
  if (rc) return rc;
  for (int i = 0; i < n; ++i) {
    fr.landmarks[i].is_tracking_inlier = inl[i] != 0;
    if (!inl[i]) outlier.push_back(fr.landmarks[i].lm_2d_plane);
  }
  mean_prjerr = mean;
  return FLV_OK;
}

int CameraFrame::depthInnovation(float iir_ratio, float range, bool enable_dummy) {   // camera_frame.cpp:271-330 (+ :93-131 for the stereo LK)
  CameraFrame& fr = *this;
  flv_ctx* ctx_ = ctx;
  const int cam_type = d_camera.cam_type;
  GlibcRand& rand_ = *rand_stream;
  const int n = (int)std::min<size_t>(fr.landmarks.size(), MAX_PTS);
  std::vector<double> plane(2 * MAX_PTS, 0.0), und(2 * MAX_PTS, 0.0), p3w(3 * MAX_PTS, 0.0), p3c(3 * MAX_PTS, 0.0),
      f2(2 * MAX_PTS, 0.0), fp(7 * MAX_PTS, 0.0), pt1(2 * MAX_PTS, 0.0);
  std::vector<uint8_t> has(MAX_PTS, 0), st1(MAX_PTS, 0);
  std::vector<uint16_t> dat(MAX_PTS, 0);
  for (int i = 0; i < n; ++i) {
    const LandMarkInFrame& lm = fr.landmarks[i];
    plane[2 * i] = lm.lm_2d_plane[0]; plane[2 * i + 1] = lm.lm_2d_plane[1];
    und[2 * i] = lm.lm_2d_undistort[0]; und[2 * i + 1] = lm.lm_2d_undistort[1];
    for (int k = 0; k < 3; ++k) { p3w[3 * i + k] = lm.lm_3d_w[k]; p3c[3 * i + k] = lm.lm_3d_c[k]; }
    has[i] = lm.has_3d;
    f2[2 * i] = lm.lm_1st_obs_2d[0]; f2[2 * i + 1] = lm.lm_1st_obs_2d[1];
    const Pose7 p = to7(lm.lm_1st_obs_frame_pose);
    for (int k = 0; k < 7; ++k) fp[7 * i + k] = p[k];
  }
  flv_camera cam{};
  cam.fx = d_camera.cam0_fx; cam.fy = d_camera.cam0_fy; cam.cx = d_camera.cam0_cx; cam.cy = d_camera.cam0_cy;
  memcpy(cam.P0, d_camera.P0_, sizeof(cam.P0)); memcpy(cam.P1, d_camera.P1_, sizeof(cam.P1));
  cam.cam_type = cam_type == DEPTH_D435 ? 0 : 1;
  cam.depth_scale = d_camera.cam_scale_factor;
  if (cam_type == DEPTH_D435) {
    const int w = d_camera.img_w, h = d_camera.img_h;
    for (int i = 0; i < n; ++i) {
      const int px = (int)std::round(plane[2 * i]), py = (int)std::round(plane[2 * i + 1]);
      dat[i] = (px >= 0 && px < w && py >= 0 && py < h) ? fr.d_img[(size_t)py * w + px] : 0;
    }
  } else {
    // project the landmarks that already have depth into cam1 (cv::projectPoints with D1 = 0), else start at the cam0 position
    std::vector<float> prev(2 * MAX_PTS, 0.f), init(2 * MAX_PTS, 0.f), next(2 * MAX_PTS), err(MAX_PTS);
    const SE3 T_c1_w = d_camera.T_cam1_cam0 * fr.T_c_w;
    double R1w[9]; q_to_R(T_c1_w.q, R1w);
    const double t1w[3] = {T_c1_w.t[0], T_c1_w.t[1], T_c1_w.t[2]};
    for (int i = 0; i < n; ++i) {
      const LandMarkInFrame& lm = fr.landmarks[i];
      prev[2 * i] = (float)lm.lm_2d_plane[0]; prev[2 * i + 1] = (float)lm.lm_2d_plane[1];
      init[2 * i] = prev[2 * i]; init[2 * i + 1] = prev[2 * i + 1];
      if (lm.has_3d && cam_type == STEREO_UNRECT) {                // cv::projectPoints(K1, D1) (camera_frame.cpp:113-122)
        project_point(d_camera.lens1, R1w, t1w, (float)lm.lm_3d_w[0], (float)lm.lm_3d_w[1], (float)lm.lm_3d_w[2], init[2 * i], init[2 * i + 1]);
      } else if (lm.has_3d) {
        const Vec3 pc = DepthCamera::world2cameraT_c_w(Vec3{(double)(float)lm.lm_3d_w[0], (double)(float)lm.lm_3d_w[1], (double)(float)lm.lm_3d_w[2]}, T_c1_w);
        init[2 * i] = (float)(d_camera.cam1_fx * pc[0] / pc[2] + d_camera.cam1_cx);
        init[2 * i + 1] = (float)(d_camera.cam1_fy * pc[1] / pc[2] + d_camera.cam1_cy);
      }
    }
    flv_lk_params lk{31, 5, 30, 1e-3, 1e-4};
    int rc = flv_lk_track(ctx_, fr.slot0, fr.slot1, 1, &n, prev.data(), init.data(), next.data(), st1.data(), err.data(), &lk, FLV_MEM_HOST);
    if (rc) return rc;
    for (int i = 0; i < n; ++i) {                                  // cv::undistortPoints(K1, D1, R1, P1) (camera_frame.cpp:130)
      float ux = next[2 * i], uy = next[2 * i + 1];                // rectified input: the map is the identity
      if (cam_type == STEREO_UNRECT) undistort_point(d_camera.lens1, next[2 * i], next[2 * i + 1], ux, uy);
      pt1[2 * i] = ux; pt1[2 * i + 1] = uy;
    }
  }
  // the next MAX_PTS dummy depths of this sequence's rand() stream; only the consumed ones advance the generator
  GlibcRand peek = rand_;
  std::vector<float> rnd(MAX_PTS);
  for (float& v : rnd) v = peek.dummy_depth();
  flv_depth_params prm{iir_ratio, range, enable_dummy ? 1 : 0};
  const Pose7 T = to7(fr.T_c_w);
  int used = 0;
  int rc = flv_depth_innovation(ctx_, 1, &n, &cam, &prm, T.data(), plane.data(), und.data(), p3w.data(), p3c.data(), has.data(),
                                f2.data(), fp.data(), pt1.data(), st1.data(), dat.data(), rnd.data(), &used, FLV_MEM_HOST);
  if (rc) return rc;
  for (int i = 0; i < used; ++i) rand_.rand();
  for (int i = 0; i < n; ++i) {
    LandMarkInFrame& lm = fr.landmarks[i];
    if (has[i]) {
      for (int k = 0; k < 3; ++k) { lm.lm_3d_w[k] = p3w[3 * i + k]; lm.lm_3d_c[k] = p3c[3 * i + k]; }
      lm.has_3d = true;
    }
  }
  return FLV_OK;
}

}  // namespace flv
