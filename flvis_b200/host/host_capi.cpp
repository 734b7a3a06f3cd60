// extern "C" handles over the C++ host classes (include/flvis_b200_host.h).
#include "../../include/flvis_b200_host.h"
#include <new>
#include "local_map.h"
#include "vi_motion.h"

struct flv_localmap { flv::LocalMap impl; flv_localmap(flv_ctx* c, int w, double fx, double fy, double cx, double cy) : impl(c, w, fx, fy, cx, cy) {} };

extern "C" {

flv_localmap* flv_localmap_create(flv_ctx* ctx, int window_size, double fx, double fy, double cx, double cy) {
  if (!ctx || window_size < 3 || window_size > 32) return nullptr;    // reference clamps to [3,100]; kernel limit 32
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

struct flv_vimotion { flv::VIMOTION impl; explicit flv_vimotion(const flv::VIMOTION& v) : impl(v) {} };

static flv::SE3 se3_from7(const double* p) { return flv::SE3(flv::Quat{p[3], p[0], p[1], p[2]}, flv::Vec3{p[4], p[5], p[6]}); }
static void se3_to7(const flv::SE3& T, double* p) { p[0] = T.q.x; p[1] = T.q.y; p[2] = T.q.z; p[3] = T.q.w; p[4] = T.t[0]; p[5] = T.t[1]; p[6] = T.t[2]; }

flv_vimotion* flv_vimotion_create(const double* T_i_c, double g, double p1, double p2, double p3, double p4, double p5, double p6) {
  if (!T_i_c) return nullptr;
  return new (std::nothrow) flv_vimotion(flv::VIMOTION(se3_from7(T_i_c), g, p1, p2, p3, p4, p5, p6));
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

}  // extern "C"
