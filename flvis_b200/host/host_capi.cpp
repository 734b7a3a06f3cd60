// extern "C" handles over the C++ host classes (include/flvis_b200_host.h).
#include "../../include/flvis_b200_host.h"
#include <new>
#include "local_map.h"

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

}  // extern "C"
