// extern "C" handles over the C++ host classes (include/flvis_b200_host.h).
#include "../../include/flvis_b200_host.h"
#include <new>
#include "local_map.h"
#include "vi_motion.h"
#include "f2f_tracking.h"

struct flv_localmap { flv::LocalMap impl; flv_localmap(flv_ctx* c, int w, double fx, double fy, double cx, double cy) : impl(c, w, fx, fy, cx, cy) {} };

extern "C" {

flv_localmap* flv_localmap_create(flv_ctx* ctx, int window_size, double fx, double fy, double cx, double cy) {
  // reference clamps window_size to [3,100] (vo_localmap.cpp:441-447); windows of more than 25 poses run on the global-memory
  // solver (csrc/ba_big.cu)
  if (!ctx || window_size < 3 || window_size > 100) return nullptr;
  return new (std::nothrow) flv_localmap(ctx, window_size, fx, fy, cx, cy);
}
void flv_localmap_destroy(flv_localmap* lm) { delete lm; }
void flv_localmap_reset(flv_localmap* lm) { if (lm) lm->impl.reset(); }

int flv_localmap_add_keyframe(flv_localmap* lm, int64_t frame_id, int n, const int64_t* lm_id, const double* lm_2d,
                              const double* lm_3d, const double* T_c_w, int64_t* out_frame_id, double* out_T_c_w,
                              int* out_lm_count, int64_t* out_lm_id, double* out_lm_3d, int lm_cap,
                              int* out_outlier_count, int64_t* out_outlier_id, int outlier_cap,
                              flv_ba_stats* out_stats) {
  if (!lm || n < 0 || (n > 0 && (!lm_id || !lm_2d || !lm_3d)) || !T_c_w) return FLV_ERR_INVALID;
  flv::KeyFrameStruct kf;
  kf.frame_id = frame_id; kf.lm_count = n;
  kf.lm_id.assign(lm_id, lm_id + n);
  kf.lm_2d.resize(n); kf.lm_3d.resize(n);
  for (int i = 0; i < n; ++i) {
    kf.lm_2d[i] = flv::Vec2{lm_2d[2 * i], lm_2d[2 * i + 1]};
    kf.lm_3d[i] = flv::Vec3{lm_3d[3 * i], lm_3d[3 * i + 1], lm_3d[3 * i + 2]};
  }
  for (int k = 0; k < 7; ++k) kf.T_c_w[k] = T_c_w[k];
  flv::CorrectionInfStruct c;
  const bool was_ready = lm->impl.frame_callback(kf, c);
  if (lm->impl.solve_failed()) return FLV_ERR_CUDA;          // a due solve did not run (flv_last_error of the context says why)
  if (!was_ready) return 0;
  if ((int)c.lm_id.size() > lm_cap || (int)c.lm_outlier_id.size() > outlier_cap) return FLV_ERR_OVERFLOW;
  if (out_frame_id) *out_frame_id = c.frame_id;
  if (out_T_c_w) for (int k = 0; k < 7; ++k) out_T_c_w[k] = c.T_c_w[k];
  if (out_lm_count) *out_lm_count = c.lm_count;
  for (size_t i = 0; i < c.lm_id.size(); ++i) {
    out_lm_id[i] = c.lm_id[i];
    for (int k = 0; k < 3; ++k) out_lm_3d[3 * i + k] = c.lm_3d[i][k];
  }
  if (out_outlier_count) *out_outlier_count = c.lm_outlier_count;
  for (size_t i = 0; i < c.lm_outlier_id.size(); ++i) out_outlier_id[i] = c.lm_outlier_id[i];
  if (out_stats) *out_stats = lm->impl.last_stats();
  return 1;
}

struct flv_vimotion {
  flv::VIMOTION impl;
  flv_vimotion(const flv::SE3& T_i_c, double g, double p1, double p2, double p3, double p4, double p5, double p6)
      : impl(T_i_c, g, p1, p2, p3, p4, p5, p6) {}
};

static flv::SE3 se3_from7(const double* p) { return flv::SE3(flv::Quat{p[3], p[0], p[1], p[2]}, flv::Vec3{p[4], p[5], p[6]}); }
static void se3_to7(const flv::SE3& T, double* p) { p[0] = T.q.x; p[1] = T.q.y; p[2] = T.q.z; p[3] = T.q.w; p[4] = T.t[0]; p[5] = T.t[1]; p[6] = T.t[2]; }

flv_vimotion* flv_vimotion_create(const double* T_i_c, double g, double p1, double p2, double p3, double p4, double p5, double p6) {
  if (!T_i_c) return nullptr;
  return new (std::nothrow) flv_vimotion(se3_from7(T_i_c), g, p1, p2, p3, p4, p5, p6);
}
void flv_vimotion_destroy(flv_vimotion* vm) { delete vm; }
int flv_vimotion_imu_feed(flv_vimotion* vm, double t, const double* acc, const double* gyro, double* q, double* pos, double* vel) {
  if (!vm || !acc || !gyro) return FLV_ERR_INVALID;
  flv::IMUSTATE s; s.timestamp = t; s.acc_raw = flv::Vec3{acc[0], acc[1], acc[2]}; s.gyro_raw = flv::Vec3{gyro[0], gyro[1], gyro[2]};
  flv::Quat qo; flv::Vec3 p, v;
  if (!vm->impl.imu_initialized) vm->impl.viIMUinitialization(s, qo, p, v);      // f2f_tracking.cpp:49-56
  else vm->impl.viIMUPropagation(s, qo, p, v);
  if (q) { q[0] = qo.w; q[1] = qo.x; q[2] = qo.y; q[3] = qo.z; }
  if (pos) for (int k = 0; k < 3; ++k) pos[k] = p[k];
  if (vel) for (int k = 0; k < 3; ++k) vel[k] = v[k];
  return vm->impl.imu_initialized ? 1 : 0;
}
int flv_vimotion_vision_trigger(flv_vimotion* vm, double* q) {
  if (!vm || vm->impl.states.empty()) return FLV_ERR_INVALID;
  flv::Quat qo; vm->impl.viVisiontrigger(qo);
  if (q) { q[0] = qo.w; q[1] = qo.x; q[2] = qo.y; q[3] = qo.z; }
  return FLV_OK;
}
int flv_vimotion_correction(flv_vimotion* vm, double t_curr, const double* Tc, double t_last, const double* Tl) {
  if (!vm || !Tc || !Tl) return FLV_ERR_INVALID;
  vm->impl.viCorrectionFromVision(t_curr, se3_from7(Tc), t_last, se3_from7(Tl), 0.0);
  return FLV_OK;
}
int flv_vimotion_corr_frame_state(flv_vimotion* vm, double t, double* T_c_w) {
  if (!vm || !T_c_w) return FLV_ERR_INVALID;
  flv::SE3 T;
  if (!vm->impl.viGetCorrFrameState(t, T)) return 0;
  se3_to7(T, T_c_w);
  return 1;
}
int flv_vimotion_rp_compensation(flv_vimotion* vm, double t, double* T_c_w) {
  if (!vm || !T_c_w) return FLV_ERR_INVALID;
  flv::SE3 T = se3_from7(T_c_w);
  vm->impl.viVisionRPCompensation(t, T);
  se3_to7(T, T_c_w);
  return FLV_OK;
}
int flv_vimotion_get_bias(flv_vimotion* vm, double* ab, double* gb) {
  if (!vm) return FLV_ERR_INVALID;
  for (int k = 0; k < 3; ++k) { if (ab) ab[k] = vm->impl.acc_bias[k]; if (gb) gb[k] = vm->impl.gyro_bias[k]; }
  return FLV_OK;
}
int flv_vimotion_queue_size(flv_vimotion* vm) { return vm ? (int)vm->impl.states.size() : FLV_ERR_INVALID; }

struct flv_f2f { flv::F2FTracking impl; };

flv_f2f* flv_f2f_create(const flv_f2f_config* c, int device) {
  if (!c) return nullptr;
  flv_f2f* f = new (std::nothrow) flv_f2f();
  if (!f) return nullptr;
  flv::DepthCamera dc;
  dc.cam_type = c->cam_type == 0 ? flv::DEPTH_D435 : c->cam_type == 1 ? flv::STEREO_RECT : flv::STEREO_UNRECT;
  dc.img_w = c->img_w; dc.img_h = c->img_h;
  dc.cam0_fx = c->cam0[0]; dc.cam0_fy = c->cam0[1]; dc.cam0_cx = c->cam0[2]; dc.cam0_cy = c->cam0[3];
  dc.cam1_fx = c->cam1[0]; dc.cam1_fy = c->cam1[1]; dc.cam1_cx = c->cam1[2]; dc.cam1_cy = c->cam1[3];
  dc.cam_scale_factor = c->depth_scale;
  for (int i = 0; i < 12; ++i) { dc.P0_[i] = c->P0[i]; dc.P1_[i] = c->P1[i]; }
  dc.T_cam1_cam0 = se3_from7(c->T_cam1_cam0);
  // lens models default to the rectified pinhole (K from P, D = 0, R = I); flv_f2f_set_lens installs the raw models
  dc.lens0.fx = dc.cam0_fx; dc.lens0.fy = dc.cam0_fy; dc.lens0.cx = dc.cam0_cx; dc.lens0.cy = dc.cam0_cy;
  dc.lens1.fx = dc.cam1_fx; dc.lens1.fy = dc.cam1_fy; dc.lens1.cx = dc.cam1_cx; dc.lens1.cy = dc.cam1_cy;
  for (int i = 0; i < 12; ++i) { dc.lens0.P[i] = c->P0[i]; dc.lens1.P[i] = c->P1[i]; }
  if (f->impl.init(dc, se3_from7(c->T_i_c0), c->feature_para, c->vi_para, c->dc_para, c->skip_first_n_imgs, false, device) != FLV_OK) {
    // keep the object so the caller can read last_error; image_feed will fail
  }
  return f;
}
static flv::DepthCamera stereo_camera(const flv_f2f_stereo_config* c) {
  flv::DepthCamera dc;
  const double zero4[4] = {0, 0, 0, 0};
  const double Kr0[9] = {c->P0[0], c->P0[1], c->P0[2], c->P0[4], c->P0[5], c->P0[6], c->P0[8], c->P0[9], c->P0[10]};
  const double Kr1[9] = {c->P1[0], c->P1[1], c->P1[2], c->P1[4], c->P1[5], c->P1[6], c->P1[8], c->P1[9], c->P1[10]};
  dc.setSteroCamInfo(c->img_w, c->img_h, c->K0, c->D0, 14, Kr0, zero4, c->R0, c->P0, c->K1, c->D1, 14, Kr1, zero4, c->R1, c->P1,
                     se3_from7(c->T_c0_c1), c->cam_type == 1 ? flv::STEREO_RECT : flv::STEREO_UNRECT);
  return dc;
}
flv_f2f* flv_f2f_create_stereo(const flv_f2f_stereo_config* c, int device) {
  if (!c || (c->cam_type != 1 && c->cam_type != 2) || c->D0[12] != 0 || c->D0[13] != 0 || c->D1[12] != 0 || c->D1[13] != 0) return nullptr;
  flv_f2f* f = new (std::nothrow) flv_f2f();
  if (!f) return nullptr;
  f->impl.init(stereo_camera(c), se3_from7(c->T_i_c0), c->feature_para, c->vi_para, c->dc_para, c->skip_first_n_imgs,
               c->need_equal_hist != 0, device);     // on failure the object stays so that the caller can read last_error
  return f;
}
int flv_host_stereo_cam_info(const flv_f2f_stereo_config* c, double* cam0_4, double* cam1_4, double* T_cam1_cam0_7) {
  if (!c || !cam0_4 || !cam1_4 || !T_cam1_cam0_7) return FLV_ERR_INVALID;
  const flv::DepthCamera dc = stereo_camera(c);
  cam0_4[0] = dc.cam0_fx; cam0_4[1] = dc.cam0_fy; cam0_4[2] = dc.cam0_cx; cam0_4[3] = dc.cam0_cy;
  cam1_4[0] = dc.cam1_fx; cam1_4[1] = dc.cam1_fy; cam1_4[2] = dc.cam1_cx; cam1_4[3] = dc.cam1_cy;
  se3_to7(dc.T_cam1_cam0, T_cam1_cam0_7);
  return FLV_OK;
}
void flv_f2f_destroy(flv_f2f* f) { delete f; }
int flv_f2f_set_lens(flv_f2f* f, int cam, const double* K4, const double* D14, const double* R9) {
  if (!f || (cam != 0 && cam != 1) || !K4 || !D14 || !R9) return FLV_ERR_INVALID;
  if (D14[12] != 0 || D14[13] != 0) return FLV_ERR_UNSUPPORTED;      // tilted sensor model
  for (flv::CameraFrame* fr : {f->impl.curr_frame.get(), f->impl.last_frame.get()}) {
    if (!fr) continue;
    flv::LensModel& m = cam == 0 ? fr->d_camera.lens0 : fr->d_camera.lens1;
    m.fx = K4[0]; m.fy = K4[1]; m.cx = K4[2]; m.cy = K4[3];
    for (int i = 0; i < 14; ++i) m.k[i] = D14[i];
    for (int i = 0; i < 9; ++i) m.R[i] = R9[i];
  }
  return f->impl.set_lens(cam, K4, D14, R9);
}
int flv_f2f_set_equalize_hist(flv_f2f* f, int enable) { return f ? f->impl.set_equalize_hist(enable != 0) : FLV_ERR_INVALID; }
int flv_host_undistort_points(const double* K4, const double* D14, const double* R9, const double* P12, int n, const float* in_xy,
                              float* out_xy) {
  if (!K4 || !D14 || !R9 || !P12 || !in_xy || !out_xy || n < 0) return FLV_ERR_INVALID;
  flv::LensModel m;
  m.fx = K4[0]; m.fy = K4[1]; m.cx = K4[2]; m.cy = K4[3];
  for (int i = 0; i < 14; ++i) m.k[i] = D14[i];
  for (int i = 0; i < 9; ++i) m.R[i] = R9[i];
  for (int i = 0; i < 12; ++i) m.P[i] = P12[i];
  for (int i = 0; i < n; ++i) flv::undistort_point(m, in_xy[2 * i], in_xy[2 * i + 1], out_xy[2 * i], out_xy[2 * i + 1]);
  return FLV_OK;
}
int flv_host_fundamental_ransac(int n, const float* from_xy, const float* to_xy, double thr_px, double conf, uint8_t* mask, double* F9) {
  if (n < 0 || !from_xy || !to_xy || !mask || !F9) return FLV_ERR_INVALID;
  std::vector<flv::P2f> a(n), b(n);
  for (int i = 0; i < n; ++i) { a[i] = flv::P2f{from_xy[2 * i], from_xy[2 * i + 1]}; b[i] = flv::P2f{to_xy[2 * i], to_xy[2 * i + 1]}; }
  std::vector<uint8_t> m;
  const bool ok = flv::find_fundamental_ransac(a, b, thr_px, conf, m, F9);
  for (int i = 0; i < n; ++i) mask[i] = ok ? m[i] : 0;
  return ok ? 1 : 0;
}
int flv_host_pnp_ransac(int n, const float* p3d, const float* p2d, const double* K4, double* T_c_w_inout, int iterations, double thr_px,
                        double conf, uint8_t* mask) {
  if (n < 0 || !p3d || !p2d || !K4 || !T_c_w_inout || !mask) return FLV_ERR_INVALID;
  std::vector<flv::P3f> x(n); std::vector<flv::P2f> u(n);
  for (int i = 0; i < n; ++i) { x[i] = flv::P3f{p3d[3 * i], p3d[3 * i + 1], p3d[3 * i + 2]}; u[i] = flv::P2f{p2d[2 * i], p2d[2 * i + 1]}; }
  flv::Pose7 T;
  for (int k = 0; k < 7; ++k) T[k] = T_c_w_inout[k];
  std::vector<int> inl;
  const bool ok = flv::solve_pnp_ransac(x, u, K4, T, iterations, thr_px, conf, inl);
  for (int i = 0; i < n; ++i) mask[i] = 0;
  for (int k : inl) mask[k] = 1;
  for (int k = 0; k < 7; ++k) T_c_w_inout[k] = T[k];
  return ok ? (int)inl.size() : 0;
}
int flv_host_project_points(const double* K4, const double* D14, const double* Rcw9, const double* t3, int n, const float* xyz,
                            float* out_xy) {
  if (!K4 || !D14 || !Rcw9 || !t3 || !xyz || !out_xy || n < 0) return FLV_ERR_INVALID;
  flv::LensModel m;
  m.fx = K4[0]; m.fy = K4[1]; m.cx = K4[2]; m.cy = K4[3];
  for (int i = 0; i < 14; ++i) m.k[i] = D14[i];
  for (int i = 0; i < n; ++i) flv::project_point(m, Rcw9, t3, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], out_xy[2 * i], out_xy[2 * i + 1]);
  return FLV_OK;
}
const char* flv_f2f_last_error(flv_f2f* f) { return f ? f->impl.last_error() : "null"; }
void flv_f2f_set_ransac_hooks(flv_f2f* f, flv_f2f_fmat_fn fmat, flv_f2f_pnp_fn pnp, void* user) {
  if (f) f->impl.set_ransac_hooks(fmat, pnp, user);
}
int flv_f2f_imu_feed(flv_f2f* f, double t, const double* acc, const double* gyro) {
  if (!f || !acc || !gyro || !f->impl.vimotion) return FLV_ERR_INVALID;
  flv::Quat q; flv::Vec3 p, v;
  f->impl.imu_feed(t, flv::Vec3{acc[0], acc[1], acc[2]}, flv::Vec3{gyro[0], gyro[1], gyro[2]}, q, p, v);
  return FLV_OK;
}
int flv_f2f_image_feed(flv_f2f* f, double t, const uint8_t* img0, const void* img1, int* new_keyframe, int* reset_cmd) {
  if (!f || !img0 || !img1 || !f->impl.vimotion) return FLV_ERR_INVALID;
  bool kf = false, rs = false;
  const int rc = f->impl.image_feed(t, img0, img1, kf, rs);
  if (new_keyframe) *new_keyframe = kf;
  if (reset_cmd) *reset_cmd = rs;
  return rc;
}
int flv_f2f_state(flv_f2f* f) { return f ? (int)f->impl.vo_tracking_state : FLV_ERR_INVALID; }
int flv_f2f_get_frame(flv_f2f* f, double* T_c_w, int64_t* lm_id, double* plane_xy, double* undist_xy, double* p3d_w,
                      uint8_t* has_3d, uint8_t* is_inlier, int cap) {
  if (!f || !f->impl.curr_frame) return FLV_ERR_INVALID;
  const flv::CameraFrame& fr = *f->impl.curr_frame;
  if (T_c_w) se3_to7(fr.T_c_w, T_c_w);
  const int n = (int)fr.landmarks.size() < cap ? (int)fr.landmarks.size() : cap;
  for (int i = 0; i < n; ++i) {
    const flv::LandMarkInFrame& lm = fr.landmarks[i];
    if (lm_id) lm_id[i] = lm.lm_id;
    if (plane_xy) { plane_xy[2 * i] = lm.lm_2d_plane[0]; plane_xy[2 * i + 1] = lm.lm_2d_plane[1]; }
    if (undist_xy) { undist_xy[2 * i] = lm.lm_2d_undistort[0]; undist_xy[2 * i + 1] = lm.lm_2d_undistort[1]; }
    if (p3d_w) for (int k = 0; k < 3; ++k) p3d_w[3 * i + k] = lm.lm_3d_w[k];
    if (has_3d) has_3d[i] = lm.has_3d;
    if (is_inlier) is_inlier[i] = lm.is_tracking_inlier;
  }
  return n;
}
int flv_f2f_correction_feed(flv_f2f* f, double t, int64_t frame_id, const double* T_c_w, int lm_count, const int64_t* lm_id,
                            const double* lm_3d, int outlier_count, const int64_t* outlier_id) {
  if (!f || !T_c_w || lm_count < 0 || outlier_count < 0 || (lm_count > 0 && (!lm_id || !lm_3d)) || (outlier_count > 0 && !outlier_id))
    return FLV_ERR_INVALID;
  flv::CorrectionInfStruct c;
  c.frame_id = frame_id;
  for (int k = 0; k < 7; ++k) c.T_c_w[k] = T_c_w[k];
  c.lm_count = lm_count; c.lm_outlier_count = outlier_count;
  for (int i = 0; i < lm_count; ++i) { c.lm_id.push_back(lm_id[i]); c.lm_3d.push_back(flv::Vec3{lm_3d[3 * i], lm_3d[3 * i + 1], lm_3d[3 * i + 2]}); }
  for (int i = 0; i < outlier_count; ++i) c.lm_outlier_id.push_back(outlier_id[i]);
  f->impl.correction_feed(t, c);
  return FLV_OK;
}
int flv_f2f_get_frame_ex(flv_f2f* f, double* p3d_c, double* first_obs_2d, double* first_obs_pose, double* T_c_w_last_keyframe, int cap) {
  if (!f || !f->impl.curr_frame) return FLV_ERR_INVALID;
  const flv::CameraFrame& fr = *f->impl.curr_frame;
  if (T_c_w_last_keyframe) se3_to7(f->impl.last_keyframe_pose(), T_c_w_last_keyframe);
  const int n = (int)fr.landmarks.size() < cap ? (int)fr.landmarks.size() : cap;
  for (int i = 0; i < n; ++i) {
    const flv::LandMarkInFrame& lm = fr.landmarks[i];
    if (p3d_c) for (int k = 0; k < 3; ++k) p3d_c[3 * i + k] = lm.lm_3d_c[k];
    if (first_obs_2d) { first_obs_2d[2 * i] = lm.lm_1st_obs_2d[0]; first_obs_2d[2 * i + 1] = lm.lm_1st_obs_2d[1]; }
    if (first_obs_pose) se3_to7(lm.lm_1st_obs_frame_pose, first_obs_pose + 7 * i);
  }
  return n;
}
int flv_f2f_get_imu_states(flv_f2f* f, double* out11, int cap) {
  if (!f || !f->impl.vimotion || (!out11 && cap > 0)) return FLV_ERR_INVALID;
  return f->impl.vimotion->dump_states(out11, cap);
}
int flv_f2f_get_imu_bias(flv_f2f* f, double* acc_bias, double* gyro_bias) {
  if (!f || !f->impl.vimotion) return FLV_ERR_INVALID;
  for (int k = 0; k < 3; ++k) {
    if (acc_bias) acc_bias[k] = f->impl.vimotion->acc_bias[k];
    if (gyro_bias) gyro_bias[k] = f->impl.vimotion->gyro_bias[k];
  }
  return f->impl.has_imu ? 1 : 0;
}
int flv_f2f_tracking_counts(flv_f2f* f, int* of, int* fi, int* pnp) {
  if (!f) return FLV_ERR_INVALID;
  if (!f->impl.lkorb_tracker) return FLV_ERR_INVALID;
  if (of) *of = f->impl.lkorb_tracker->last_of_inliers;
  if (fi) *fi = f->impl.lkorb_tracker->last_f_inliers;
  if (pnp) *pnp = f->impl.lkorb_tracker->last_pnp_inliers;
  return FLV_OK;
}

}  // extern "C"
