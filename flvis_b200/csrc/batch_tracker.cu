// flv_f2f_batch -- S independent camera sequences advanced together: F2FTracking::image_feed for all of them with one
// launch per stage and device-resident hand-offs (tracker.cu holds the stage kernels, this file the per-stream host side:
// the IMU filter, the UnInit / Tracking / TrackingFail state machine, the keyframe rule and the launch sequence).
//
// Reference: src/frontend/f2f_tracking.cpp:5-453 (init :5-38, imu_feed :46-57, image_feed :59-400, init_frame :402-453).
// One frame costs ONE host synchronisation (the read-back of the per-stream summaries and landmark lists at the end):
// everything the device needs from the host is known before the frame starts -- the IMU pose guess
// (viGetCorrFrameState, :225) and the IMU roll / pitch at the frame time (viVisionRPCompensation, :253) -- and
// everything the host needs from the device is needed only after the frame (viCorrectionFromVision :281, the keyframe
// rule :339-355, the failure counters :229-247).  Tests may install the OpenCV RANSAC hooks; those force two more round
// trips per frame (the correspondences go to the host and the masks come back).
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <new>
#include <vector>
#include "tracker.h"
#include "../../include/flvis_b200_host.h"
#include "../host/glibc_rand.h"
#include "../host/sophus_lite.h"
#include "../host/vi_motion.h"
#include "../host/nvtx_range.h"

namespace {

using flv::Quat; using flv::SE3; using flv::Vec3; using flv::Pose7;

SE3 from7(const double* p) { return SE3(Quat{p[3], p[0], p[1], p[2]}, Vec3{p[4], p[5], p[6]}); }
void to7(const SE3& T, double* p) { p[0] = T.q.x; p[1] = T.q.y; p[2] = T.q.z; p[3] = T.q.w; p[4] = T.t[0]; p[5] = T.t[1]; p[6] = T.t[2]; }
SE3 raw7(const double* p) { SE3 T; T.q = Quat{p[3], p[0], p[1], p[2]}; T.t = Vec3{p[4], p[5], p[6]}; return T; }   // no re-normalisation

Vec3 so3_log(const Quat& q) {          // Sophus SO3::logAndTheta (so3.cpp:127-164)
  const double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z), w = q.w;
  double f;
  if (n < 1e-10) f = 2. / w - 2. * (n * n) / (w * w * w);
  else f = 2 * std::atan(n / w) / n;
  return Vec3{f * q.x, f * q.y, f * q.z};
}

enum { UnInit = 0, Tracking = 1, TrackingFail = 2 };

struct StreamState {
  int state = UnInit;
  bool has_imu = false;
  int frameCount = 0, skip_n_imgs = 0;
  std::unique_ptr<flv::VIMOTION> vim;
  flv::GlibcRand rnd;
  SE3 T_kf;                       // T_c_w_last_keyframe
  int continus_tracking_fail_cnt = 0, fail_cnt = 0;
  double cur_time = 0, last_time = 0;      // frame_time of curr_frame / last_frame (as the reference's two frame objects)
  SE3 cur_T, last_T;
  int of_cnt = 0, f_cnt = 0, pnp_cnt = 0;
  int n_lm = 0;                   // landmarks of curr_frame (host mirror below)
};

}  // namespace

struct flv_f2f_batch {
  flv_ctx* ctx = nullptr;
  flv_f2f_config cfg{};
  int S = 0, M = 512, device = 0;
  bool unrect = false, stereo = false;
  flv_feature_params fprm{};
  flv_depth_params dprm{};
  std::vector<StreamState> st;
  TrkDev d{};
  void* d_block = nullptr; size_t d_bytes = 0;
  // pinned host mirrors
  TrkCtl* h_ctl = nullptr; TrkOut* h_out = nullptr; float* h_rnd = nullptr;
  unsigned char* h_tab = nullptr; size_t tab_bytes = 0, lean_bytes = 0;   // read-back of the L tables
  bool full_readback = true;
  size_t o_id = 0, o_plane = 0, o_und = 0, o_p3w = 0, o_p3c = 0, o_f2d = 0, o_fpose = 0, o_has = 0, o_inl = 0, o_n = 0, o_T = 0;
  void* d_tab = nullptr;                                   // the L table block on the device (contiguous, same offsets)
  int slots[3] = {0, 1, 2};                                // prev0, cur0, cur1
  flv_f2f_fmat_fn fmat_fn = nullptr; flv_f2f_pnp_fn pnp_fn = nullptr; void* hook_user = nullptr;
  // optional per-frame result log for the multi-GPU gather: rlog[frame % rlog_K][stream_offset + s][8] = pose7 + landmark count
  double* rlog = nullptr; int rlog_K = 0, S_total = 0; long long rlog_frame = 0;
  std::vector<int> async_kf, async_rs;                     // flags of the frame started by flv_f2f_batch_frame_async
  std::vector<char> have_last;                             // stream has an accepted "last" frame on the device
  cudaEvent_t ev_done = nullptr;
  // the right image (ingest + pyramid) is only needed by the left->right LK late in the frame: it is prepared on a side
  // stream while the frame->frame stages run
  cudaStream_t side = nullptr; cudaEvent_t ev_fork_r = nullptr, ev_right = nullptr;
  // optional per-stage device timing (flv_f2f_batch_set_profile): events on the compute stream at the stage boundaries
  static constexpr int NSTAGE = 9;
  double host_ms[4] = {0, 0, 0, 0};                        // host wall time: decisions, enqueue, wait for the device, post-frame
  bool profile = false; cudaEvent_t ev_stage[NSTAGE + 1] = {nullptr}; double stage_ms[NSTAGE] = {0}; long long prof_frames = 0;
  // groups > 1: this object only dispatches to `sub` (stream s -> sub[s / per_group], local index s % per_group); every
  // group has its own context / CUDA stream; image_feed enqueues the frame of every group before it waits for the first, so
  // the latency-bound one-CTA-per-stream stages of one group overlap the other groups' kernels and per-frame host work
  std::vector<flv_f2f_batch*> sub; int per_group = 0, stream_offset = 0;
  // frame in flight between feed_begin and feed_end
  struct { int* kf; int* rs; int prev0, cur0, cur1; std::chrono::steady_clock::time_point tp0, tp1, tp2; bool active; } pend{};
  flv_localmap_batch* lmap = nullptr;                      // keyframes go here (flv_f2f_batch_attach_localmap)
  std::vector<int> kf_streams, kf_counts; std::vector<int64_t> kf_frame, kf_ids; std::vector<double> kf_2d, kf_3d, kf_T;
  char err[512] = {0};
};

namespace {

#define B_CUDA(b, call)                                                                                         \
  do {                                                                                                          \
    cudaError_t e__ = (call);                                                                                   \
    if (e__ != cudaSuccess) {                                                                                   \
      snprintf((b)->err, sizeof((b)->err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return FLV_ERR_CUDA;                                                                                      \
    }                                                                                                           \
  } while (0)
#define B_RC(b, call)                                                                              \
  do {                                                                                             \
    const int rc__ = (call);                                                                       \
    if (rc__) { snprintf((b)->err, sizeof((b)->err), "%s: %s", #call, flv_last_error((b)->ctx)); return rc__; } \
  } while (0)

struct Carver {
  size_t off = 0;
  size_t take(size_t bytes) { const size_t r = off; off = (off + bytes + 255) & ~(size_t)255; return r; }
};

int alloc_device(flv_f2f_batch* b) {
  const size_t S = b->S, M = b->M, np = S * M;
  // L table first, contiguous, so one D2H copy brings the whole frame state back
  Carver c;
  // what a keyframe message needs comes first (the lean read-back copies only that part)
  b->o_id = c.take(np * 8); b->o_und = c.take(np * 16); b->o_p3w = c.take(np * 24); b->o_has = c.take(np);
  b->o_inl = c.take(np); b->o_n = c.take(S * 4); b->o_T = c.take(S * 56);
  b->lean_bytes = c.off;
  b->o_plane = c.take(np * 16); b->o_p3c = c.take(np * 24); b->o_f2d = c.take(np * 16); b->o_fpose = c.take(np * 56);
  b->tab_bytes = c.off;
  const size_t o_C = c.take(b->tab_bytes);
  const size_t MP = b->ctx->ba_max_poses, ML = b->ctx->ba_max_lms, ME = b->ctx->ba_max_edges;
  struct { size_t ctl, out, ok, idx, orig, lkp, lki, lkn, lke, lks, nlk, fa, fb, nf, mF, Fm, fni, p3, p2, np_, K4, Tin, Tout, mP, pni,
           bprob, bpose, blm, buv, bep, bel, bact, bst, nrep, rmean, nex, rp, ri, rn, re, rs, nr, pt1, dat, rnd, nru, depth; } o;
  o.ctl = c.take(S * sizeof(TrkCtl)); o.out = c.take(S * sizeof(TrkOut)); o.ok = c.take(S * 4); o.idx = c.take(S * 8); o.orig = c.take(S * 4);
  o.lkp = c.take(np * 8); o.lki = c.take(np * 8); o.lkn = c.take(np * 8); o.lke = c.take(np * 4); o.lks = c.take(np); o.nlk = c.take(S * 4);
  o.fa = c.take(np * 8); o.fb = c.take(np * 8); o.nf = c.take(S * 4); o.mF = c.take(np); o.Fm = c.take(S * 72); o.fni = c.take(S * 4);
  o.p3 = c.take(np * 12); o.p2 = c.take(np * 8); o.np_ = c.take(S * 4); o.K4 = c.take(S * 32); o.Tin = c.take(S * 56); o.Tout = c.take(S * 56);
  o.mP = c.take(np); o.pni = c.take(S * 4);
  o.bprob = c.take(S * sizeof(flv_ba_problem)); o.bpose = c.take(S * MP * 56); o.blm = c.take(S * ML * 24); o.buv = c.take(S * ME * 16);
  o.bep = c.take(S * ME * 4); o.bel = c.take(S * ME * 4); o.bact = c.take(S * ME); o.bst = c.take(S * sizeof(flv_ba_stats));
  o.nrep = c.take(S * 4); o.rmean = c.take(S * 8); o.nex = c.take(S * 4);
  o.rp = c.take(np * 8); o.ri = c.take(np * 8); o.rn = c.take(np * 8); o.re = c.take(np * 4); o.rs = c.take(np); o.nr = c.take(S * 4);
  o.pt1 = c.take(np * 16); o.dat = c.take(np * 2); o.rnd = c.take(np * 4); o.nru = c.take(S * 4);
  o.depth = c.take(b->stereo ? 256 : S * (size_t)b->cfg.img_w * b->cfg.img_h * 2);
  b->d_bytes = c.off;
  B_CUDA(b, cudaMalloc(&b->d_block, b->d_bytes));
  B_CUDA(b, cudaMemset(b->d_block, 0, b->d_bytes));
  char* base = (char*)b->d_block;
  auto table = [&](char* t) {
    TrkTable T;
    T.id = (long long*)(t + b->o_id); T.plane = (double*)(t + b->o_plane); T.undist = (double*)(t + b->o_und);
    T.p3w = (double*)(t + b->o_p3w); T.p3c = (double*)(t + b->o_p3c); T.f2d = (double*)(t + b->o_f2d);
    T.fpose = (double*)(t + b->o_fpose); T.has = (unsigned char*)(t + b->o_has); T.inl = (unsigned char*)(t + b->o_inl);
    T.n = (int*)(t + b->o_n); T.T = (double*)(t + b->o_T);
    return T;
  };
  b->d_tab = base;
  b->d.L = table(base); b->d.C = table(base + o_C);
  TrkBufs& q = b->d.b;
  q.max_pts = (int)M;
  q.ctl = (TrkCtl*)(base + o.ctl); q.out = (TrkOut*)(base + o.out); q.ok = (int*)(base + o.ok); q.id_index = (long long*)(base + o.idx);
  q.orig_size = (int*)(base + o.orig);
  q.lk_prev = (float*)(base + o.lkp); q.lk_init = (float*)(base + o.lki); q.lk_next = (float*)(base + o.lkn); q.lk_err = (float*)(base + o.lke);
  q.lk_st = (unsigned char*)(base + o.lks); q.n_lk = (int*)(base + o.nlk);
  q.fa = (float*)(base + o.fa); q.fb = (float*)(base + o.fb); q.n_f = (int*)(base + o.nf); q.maskF = (unsigned char*)(base + o.mF);
  q.Fm = (double*)(base + o.Fm); q.f_ninl = (int*)(base + o.fni);
  q.pnp3 = (float*)(base + o.p3); q.pnp2 = (float*)(base + o.p2); q.n_pnp = (int*)(base + o.np_); q.pnp_K4 = (double*)(base + o.K4);
  q.pnp_Tin = (double*)(base + o.Tin); q.pnp_Tout = (double*)(base + o.Tout); q.maskP = (unsigned char*)(base + o.mP); q.pnp_ninl = (int*)(base + o.pni);
  q.ba_MP = (int)MP; q.ba_ML = (int)ML; q.ba_ME = (int)ME;
  q.ba_prob = (flv_ba_problem*)(base + o.bprob); q.ba_poses = (double*)(base + o.bpose); q.ba_lms = (double*)(base + o.blm);
  q.ba_uv = (double*)(base + o.buv); q.ba_ep = (int*)(base + o.bep); q.ba_el = (int*)(base + o.bel); q.ba_act = (unsigned char*)(base + o.bact);
  q.ba_stats = (flv_ba_stats*)(base + o.bst);
  q.n_rep = (int*)(base + o.nrep); q.rep_mean = (double*)(base + o.rmean); q.n_exist = (int*)(base + o.nex);
  q.r_prev = (float*)(base + o.rp); q.r_init = (float*)(base + o.ri); q.r_next = (float*)(base + o.rn); q.r_err = (float*)(base + o.re);
  q.r_st = (unsigned char*)(base + o.rs); q.n_r = (int*)(base + o.nr);
  q.pt1 = (double*)(base + o.pt1); q.dat = (unsigned short*)(base + o.dat); q.rnd = (float*)(base + o.rnd); q.n_rand_used = (int*)(base + o.nru);
  b->d.depth = (unsigned short*)(base + o.depth);
  // constants: K per stream, landmark id counters (landmark.cpp:3: ids start at 100, one counter per sequence here)
  std::vector<double> K4(S * 4);
  std::vector<long long> ids(S, 100);
  for (size_t s = 0; s < S; ++s) for (int k = 0; k < 4; ++k) K4[4 * s + k] = b->cfg.cam0[k];
  B_CUDA(b, cudaMemcpy(q.pnp_K4, K4.data(), S * 32, cudaMemcpyHostToDevice));
  B_CUDA(b, cudaMemcpy(q.id_index, ids.data(), S * 8, cudaMemcpyHostToDevice));
  B_CUDA(b, cudaMallocHost(&b->h_ctl, S * sizeof(TrkCtl)));
  B_CUDA(b, cudaMallocHost(&b->h_out, S * sizeof(TrkOut)));
  B_CUDA(b, cudaMallocHost(&b->h_rnd, np * 4));
  B_CUDA(b, cudaMallocHost(&b->h_tab, b->tab_bytes));
  memset(b->h_tab, 0, b->tab_bytes);
  B_CUDA(b, cudaEventCreateWithFlags(&b->ev_done, cudaEventDisableTiming));
  B_CUDA(b, cudaStreamCreateWithFlags(&b->side, cudaStreamNonBlocking));
  B_CUDA(b, cudaEventCreateWithFlags(&b->ev_fork_r, cudaEventDisableTiming));
  B_CUDA(b, cudaEventCreateWithFlags(&b->ev_right, cudaEventDisableTiming));
  return FLV_OK;
}

void fill_cam(flv_f2f_batch* b) {
  TrkCam& c = b->d.cam;
  const flv_f2f_config& g = b->cfg;
  memset(&c.c, 0, sizeof(c.c));
  c.c.fx = g.cam0[0]; c.c.fy = g.cam0[1]; c.c.cx = g.cam0[2]; c.c.cy = g.cam0[3];
  memcpy(c.c.P0, g.P0, sizeof(c.c.P0)); memcpy(c.c.P1, g.P1, sizeof(c.c.P1));
  c.c.cam_type = g.cam_type == 0 ? 0 : 1;
  c.c.depth_scale = g.depth_scale;
  for (int k = 0; k < 4; ++k) c.cam1[k] = g.cam1[k];
  // the reference stores SE3 objects built from (R, t): quaternions are normalised on construction
  double t[7];
  to7(from7(g.T_cam1_cam0), t); memcpy(c.T_c1_c0, t, 56);
  const SE3 Tic = from7(g.T_i_c0);
  to7(Tic, t); memcpy(c.T_i_c, t, 56);
  to7(Tic.inverse(), t); memcpy(c.T_c_i, t, 56);
  c.vi_para2 = g.vi_para[1];
  c.w = g.img_w; c.h = g.img_h; c.unrect = g.cam_type == 2;
  c.lens0 = flv::LensModel(); c.lens1 = flv::LensModel();
  c.lens0.fx = g.cam0[0]; c.lens0.fy = g.cam0[1]; c.lens0.cx = g.cam0[2]; c.lens0.cy = g.cam0[3];
  c.lens1.fx = g.cam1[0]; c.lens1.fy = g.cam1[1]; c.lens1.cx = g.cam1[2]; c.lens1.cy = g.cam1[3];
  for (int i = 0; i < 12; ++i) { c.lens0.P[i] = g.P0[i]; c.lens1.P[i] = g.P1[i]; }
}

// the OpenCV hooks (tests): correspondences to the host, masks back
int run_fmat_hooks(flv_f2f_batch* b) {
  const size_t S = b->S, M = b->M;
  std::vector<int> nf(S); std::vector<float> fa(S * M * 2), fb(S * M * 2); std::vector<unsigned char> mask(S * M, 0);
  B_CUDA(b, cudaStreamSynchronize(b->ctx->stream));
  B_CUDA(b, cudaMemcpy(nf.data(), b->d.b.n_f, S * 4, cudaMemcpyDeviceToHost));
  B_CUDA(b, cudaMemcpy(fa.data(), b->d.b.fa, S * M * 8, cudaMemcpyDeviceToHost));
  B_CUDA(b, cudaMemcpy(fb.data(), b->d.b.fb, S * M * 8, cudaMemcpyDeviceToHost));
  for (size_t s = 0; s < S; ++s)
    if (nf[s] > 0 && b->fmat_fn(b->hook_user, nf[s], &fa[s * M * 2], &fb[s * M * 2], &mask[s * M])) memset(&mask[s * M], 0, M);
  B_CUDA(b, cudaMemcpy(b->d.b.maskF, mask.data(), S * M, cudaMemcpyHostToDevice));
  return FLV_OK;
}
int run_pnp_hooks(flv_f2f_batch* b) {
  const size_t S = b->S, M = b->M;
  std::vector<int> np(S), ninl(S, 0), idx(M); std::vector<float> p3(S * M * 3), p2(S * M * 2);
  std::vector<unsigned char> mask(S * M, 0); std::vector<double> T(S * 7);
  B_CUDA(b, cudaStreamSynchronize(b->ctx->stream));
  B_CUDA(b, cudaMemcpy(np.data(), b->d.b.n_pnp, S * 4, cudaMemcpyDeviceToHost));
  B_CUDA(b, cudaMemcpy(p3.data(), b->d.b.pnp3, S * M * 12, cudaMemcpyDeviceToHost));
  B_CUDA(b, cudaMemcpy(p2.data(), b->d.b.pnp2, S * M * 8, cudaMemcpyDeviceToHost));
  B_CUDA(b, cudaMemcpy(T.data(), b->d.b.pnp_Tin, S * 56, cudaMemcpyDeviceToHost));
  for (size_t s = 0; s < S; ++s) {
    if (np[s] <= 0) continue;
    int n = 0;
    if (b->pnp_fn(b->hook_user, np[s], &p3[s * M * 3], &p2[s * M * 2], b->cfg.cam0, b->h_ctl[s].use_guess ? 1 : 0, &T[7 * s], idx.data(), &n)) n = 0;
    for (int k = 0; k < n; ++k) if (idx[k] >= 0 && idx[k] < (int)M) mask[s * M + idx[k]] = 1;
    ninl[s] = n;
  }
  B_CUDA(b, cudaMemcpy(b->d.b.maskP, mask.data(), S * M, cudaMemcpyHostToDevice));
  B_CUDA(b, cudaMemcpy(b->d.b.pnp_Tout, T.data(), S * 56, cudaMemcpyHostToDevice));
  B_CUDA(b, cudaMemcpy(b->d.b.pnp_ninl, ninl.data(), S * 4, cudaMemcpyHostToDevice));
  return FLV_OK;
}

}  // namespace

namespace {
// dispatch helpers for grouped batches
inline flv_f2f_batch* sub_of(flv_f2f_batch* b, int& stream) {
  if (b->sub.empty()) return b;
  flv_f2f_batch* sb = b->sub[stream / b->per_group];
  stream = stream % b->per_group;
  return sb;
}
}  // namespace

extern "C" {

flv_f2f_batch* flv_f2f_batch_create_grouped(const flv_f2f_config* cfg, int n_streams, int device, int groups) {
  if (!cfg || n_streams < 1 || groups < 1) return nullptr;
  if (groups == 1 || n_streams % groups != 0 || n_streams / groups < 1) return flv_f2f_batch_create(cfg, n_streams, device);
  flv_f2f_batch* b = new (std::nothrow) flv_f2f_batch();
  if (!b) return nullptr;
  b->cfg = *cfg; b->S = n_streams; b->device = device; b->per_group = n_streams / groups;
  b->unrect = cfg->cam_type == 2; b->stereo = cfg->cam_type != 0;
  for (int g = 0; g < groups; ++g) {
    flv_f2f_batch* sb = flv_f2f_batch_create(cfg, b->per_group, device);
    if (!sb) { snprintf(b->err, sizeof(b->err), "group %d: allocation failed", g); return b; }
    sb->stream_offset = g * b->per_group;
    b->sub.push_back(sb);
    if (sb->err[0]) { snprintf(b->err, sizeof(b->err), "group %d: %s", g, sb->err); return b; }
  }
  return b;
}

flv_f2f_batch* flv_f2f_batch_create(const flv_f2f_config* cfg, int n_streams, int device) {
  if (!cfg || n_streams < 1) return nullptr;
  flv_f2f_batch* b = new (std::nothrow) flv_f2f_batch();
  if (!b) return nullptr;
  b->cfg = *cfg; b->S = n_streams; b->device = device;
  b->unrect = cfg->cam_type == 2; b->stereo = cfg->cam_type != 0;
  int rc = flv_create(&b->ctx, device, n_streams, cfg->img_w, cfg->img_h, b->M);
  if (rc) { snprintf(b->err, sizeof(b->err), "flv_create: %s", b->ctx ? flv_last_error(b->ctx) : "failed"); return b; }
  rc = flv_ba_reserve(b->ctx, 1, b->M, b->M);
  if (rc) { snprintf(b->err, sizeof(b->err), "flv_ba_reserve: %s", flv_last_error(b->ctx)); return b; }
  b->fprm.max_region_feature_num = (int)cfg->feature_para[0];
  b->fprm.min_region_feature_num = (int)cfg->feature_para[1];
  b->fprm.boundary_dis = (int)std::floor(cfg->feature_para[2] / 2.0);
  b->fprm.gftt_num = (int)cfg->feature_para[3];
  b->fprm.gftt_ql = cfg->feature_para[4];
  b->fprm.gftt_dis = (int)cfg->feature_para[5];
  b->dprm.iir_ratio = (float)cfg->dc_para[0]; b->dprm.range = (float)cfg->dc_para[1]; b->dprm.dummy_depth = !(cfg->dc_para[2] < 0.5) ? 1 : 0;
  b->st.resize(n_streams);
  b->have_last.assign(n_streams, 0);
  b->async_kf.assign(n_streams, 0); b->async_rs.assign(n_streams, 0);
  for (StreamState& s : b->st) {
    s.vim.reset(new flv::VIMOTION(from7(cfg->T_i_c0), 9.81, cfg->vi_para[0], cfg->vi_para[1], cfg->vi_para[2], cfg->vi_para[3]));
    s.skip_n_imgs = cfg->skip_first_n_imgs;
  }
  fill_cam(b);
  if (alloc_device(b) != FLV_OK) return b;
  return b;
}

void flv_f2f_batch_destroy(flv_f2f_batch* b) {
  if (!b) return;
  if (!b->sub.empty()) {
    for (flv_f2f_batch* sb : b->sub) flv_f2f_batch_destroy(sb);
    delete b;
    return;
  }
  if (b->ctx) { cudaSetDevice(b->device); cudaDeviceSynchronize(); }
  if (b->d_block) cudaFree(b->d_block);
  if (b->h_ctl) cudaFreeHost(b->h_ctl);
  if (b->h_out) cudaFreeHost(b->h_out);
  if (b->h_rnd) cudaFreeHost(b->h_rnd);
  if (b->h_tab) cudaFreeHost(b->h_tab);
  if (b->ev_done) cudaEventDestroy(b->ev_done);
  if (b->ev_fork_r) cudaEventDestroy(b->ev_fork_r);
  if (b->ev_right) cudaEventDestroy(b->ev_right);
  if (b->side) cudaStreamDestroy(b->side);
  for (cudaEvent_t e : b->ev_stage) if (e) cudaEventDestroy(e);
  if (b->ctx) flv_destroy(b->ctx);
  delete b;
}

const char* flv_f2f_batch_last_error(flv_f2f_batch* b) { return b ? b->err : "null"; }
flv_ctx* flv_f2f_batch_context(flv_f2f_batch* b) { return b ? (b->sub.empty() ? b->ctx : b->sub[0]->ctx) : nullptr; }
int flv_f2f_batch_groups(flv_f2f_batch* b) { return b ? (b->sub.empty() ? 1 : (int)b->sub.size()) : 0; }

int flv_f2f_batch_set_lens(flv_f2f_batch* b, int cam, const double* K4, const double* D14, const double* R9) {
  if (!b || (cam != 0 && cam != 1) || !K4 || !D14 || !R9) return FLV_ERR_INVALID;
  if (D14[12] != 0 || D14[13] != 0) return FLV_ERR_UNSUPPORTED;
  if (!b->sub.empty()) { for (flv_f2f_batch* sb : b->sub) { const int rc = flv_f2f_batch_set_lens(sb, cam, K4, D14, R9); if (rc) return rc; } return FLV_OK; }
  flv::LensModel& m = cam == 0 ? b->d.cam.lens0 : b->d.cam.lens1;
  m.fx = K4[0]; m.fy = K4[1]; m.cx = K4[2]; m.cy = K4[3];
  for (int i = 0; i < 14; ++i) m.k[i] = D14[i];
  for (int i = 0; i < 9; ++i) m.R[i] = R9[i];
  return FLV_OK;
}
int flv_f2f_batch_set_equalize_hist(flv_f2f_batch* b, int enable) {
  if (b && !b->sub.empty()) { for (flv_f2f_batch* sb : b->sub) { const int rc = flv_f2f_batch_set_equalize_hist(sb, enable); if (rc) return rc; } return FLV_OK; }
  return b && b->ctx ? flv_set_equalize_hist(b->ctx, enable) : FLV_ERR_INVALID;
}
void flv_f2f_batch_set_ransac_hooks(flv_f2f_batch* b, flv_f2f_fmat_fn fmat, flv_f2f_pnp_fn pnp, void* user) {
  if (b) { b->fmat_fn = fmat; b->pnp_fn = pnp; b->hook_user = user; for (flv_f2f_batch* sb : b->sub) flv_f2f_batch_set_ransac_hooks(sb, fmat, pnp, user); }
}

int flv_f2f_batch_imu_feed(flv_f2f_batch* b, int stream, double t, const double* acc, const double* gyro) {   // f2f_tracking.cpp:46-57
  if (!b || stream < 0 || stream >= b->S || !acc || !gyro) return FLV_ERR_INVALID;
  b = sub_of(b, stream);
  StreamState& s = b->st[stream];
  flv::IMUSTATE im; im.timestamp = t; im.acc_raw = Vec3{acc[0], acc[1], acc[2]}; im.gyro_raw = Vec3{gyro[0], gyro[1], gyro[2]};
  Quat q; Vec3 p, v;
  if (!s.vim->imu_initialized) { s.has_imu = true; s.vim->viIMUinitialization(im, q, p, v); }
  else s.vim->viIMUPropagation(im, q, p, v);
  return FLV_OK;
}

}  // extern "C"

namespace {

// first half of image_feed: host decisions + the whole frame enqueued on the group's stream (no wait)
int feed_begin(flv_f2f_batch* b, const double* t, const uint8_t* img0, const void* img1, flv_memspace mem, int* new_keyframe,
               int* reset_cmd) {
  if (!b || !b->ctx || !b->d_block || !t || !img0 || !img1) return FLV_ERR_INVALID;
  flv_ctx* ctx = b->ctx;
  const int S = b->S, M = b->M;
  const size_t w = b->cfg.img_w, h = b->cfg.img_h;
  const int prev0 = b->slots[0], cur0 = b->slots[1], cur1 = b->slots[2];
  // ---- per-stream decisions that need no image data (f2f_tracking.cpp:59-76, :146-186 entry, :225, :357-375) -----------
  const auto tp0 = std::chrono::steady_clock::now();
  bool any_init = false, any_track = false;
  for (int s = 0; s < S; ++s) {
    StreamState& z = b->st[s];
    TrkCtl& c = b->h_ctl[s];
    memset(&c, 0, sizeof(c));
    z.frameCount++;
    // last_frame.swap(curr_frame): the frame object that was current becomes "last"
    std::swap(z.cur_time, z.last_time); std::swap(z.cur_T, z.last_T);
    z.cur_time = t[s]; z.cur_T = SE3();
    if (new_keyframe) new_keyframe[s] = 0;
    if (reset_cmd) reset_cmd[s] = 0;
    if (z.skip_n_imgs > 0) { z.skip_n_imgs--; c.mode = 0; z.n_lm = 0; continue; }
    switch (z.state) {
      case UnInit: {
        const double R_w_c[9] = {0, 0, 1, -1, 0, 0, 0, -1, 0};
        SE3 T0 = SE3(flv::R_to_q(R_w_c), Vec3{0, 0, 0}).inverse();
        z.cur_T = T0;                                   // assigned before the IMU checks in the reference (:149-151)
        if (z.has_imu) {
          if (z.vim->imu_initialized) {
            Quat q_init;
            z.vim->viVisiontrigger(q_init);
            double Ra[9], Rb[9], Rc[9];
            flv::q_to_R(q_init, Ra); flv::q_to_R(z.vim->T_i_c.q, Rb);
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rc[3 * i + j] = Ra[3 * i] * Rb[j] + Ra[3 * i + 1] * Rb[3 + j] + Ra[3 * i + 2] * Rb[6 + j];
            T0 = SE3(flv::R_to_q(Rc), Vec3{0, 0, 0}).inverse();
          } else { c.mode = 0; z.n_lm = 0; break; }
        }
        c.mode = 2; c.commit_on_init_fail = 1;
        to7(T0, c.init_pose);
        z.cur_T = T0;
        break;
      }
      case Tracking: {
        c.mode = 1;
        SE3 guess;
        if (z.has_imu && z.vim->viGetCorrFrameState(t[s], guess)) { c.use_guess = 1; to7(guess, c.guess); }
        if (z.has_imu) {
          double roll = 0, pitch = 0;
          if (z.vim->viGetIMURollPitchAtTime(t[s], roll, pitch)) { c.rp_found = 1; c.roll = roll; c.pitch = pitch; }
        }
        break;
      }
      case TrackingFail: {
        z.fail_cnt++;
        if ((z.fail_cnt % 3) == 0) {
          SE3 T0;
          if (z.vim->viGetCorrFrameState(t[s], T0)) { c.mode = 2; to7(T0, c.init_pose); z.cur_T = T0; }
          else c.mode = 0;
          z.fail_cnt = 0;
        } else {
          c.mode = 0;
          if ((z.fail_cnt % 2) == 0 && reset_cmd) reset_cmd[s] = 1;
        }
        break;
      }
    }
    any_init |= c.mode == 2; any_track |= c.mode == 1;
    // the next max_pts dummy depths of this sequence's rand() stream (only the consumed ones advance the generator)
    flv::GlibcRand peek = z.rnd;
    float* r = b->h_rnd + (size_t)s * M;
    for (int i = 0; i < M; ++i) r[i] = peek.dummy_depth();
  }
  cudaStream_t cs = ctx->stream;
  const auto tp1 = std::chrono::steady_clock::now();
  int mark_i = 0;
  static const char* const kStage[flv_f2f_batch::NSTAGE] = {
      "flv: ingest + pyramids", "flv: LK frame->frame", "flv: keep rule + F RANSAC", "flv: PnP RANSAC", "flv: pose-only BA",
      "flv: reprojection cull", "flv: FeatureDEM redetect", "flv: LK left->right", "flv: depth innovation + finish"};
  flv::NvtxStages nvtx;                                   // NVTX range per stage (closed on every exit path)
  auto mark = [&]() {                                     // stage boundary: an event when profiling, a counter always
    if (mark_i > flv_f2f_batch::NSTAGE) return;
    if (mark_i < flv_f2f_batch::NSTAGE) nvtx.next(kStage[mark_i]); else nvtx.close();
    if (b->profile) cudaEventRecord(b->ev_stage[mark_i], cs);
    ++mark_i;
  };
  mark();                                                                 // 0: ingest + pyramids
  B_CUDA(b, cudaMemcpyAsync(b->d.b.ctl, b->h_ctl, (size_t)S * sizeof(TrkCtl), cudaMemcpyHostToDevice, cs));
  B_CUDA(b, cudaMemcpyAsync(b->d.b.rnd, b->h_rnd, (size_t)S * M * 4, cudaMemcpyHostToDevice, cs));
  // ---- images: level 0 + pyramids (f2f_tracking.cpp:78-145) -------------------------------------------------------------
  B_RC(b, flv_upload_images(ctx, cur0, S, img0, w, w * h, mem));
  if (b->stereo) {
    // fork after the left ingest (the equalizeHist scratch is shared; the previous frame's readers of the cur1 slot are
    // ordered before this point on the main stream)
    B_CUDA(b, cudaEventRecord(b->ev_fork_r, cs));
    B_CUDA(b, cudaStreamWaitEvent(b->side, b->ev_fork_r, 0));
    ctx->stream = b->side;
    int rc_r = flv_upload_images(ctx, cur1, S, (const uint8_t*)img1, w, w * h, mem);
    if (!rc_r) rc_r = flv_build_pyramid(ctx, cur1, S);
    ctx->stream = cs;
    if (rc_r) { snprintf(b->err, sizeof(b->err), "right image: %s", flv_last_error(ctx)); return rc_r; }
    B_CUDA(b, cudaEventRecord(b->ev_right, b->side));
  } else {
    B_CUDA(b, cudaMemcpyAsync(b->d.depth, img1, (size_t)S * w * h * 2, mem == FLV_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, cs));
  }
  B_RC(b, flv_build_pyramid(ctx, cur0, S));
  if (any_track) B_RC(b, flv_feature_prepare(ctx, cur0, S, &b->fprm, 1));      // Shi-Tomasi of the new image overlaps tracking
  // ---- tracking path (mode 1 streams; the others run on empty counts) ---------------------------------------------------
  B_RC(b, flv_trk_stage_prepare(ctx, b->d, S));
  const flv_lk_params lk_f2f{31, 10, 30, 1e-3, 1e-4}, lk_lr{31, 5, 30, 1e-3, 1e-4};
  const TrkBufs& q = b->d.b;
  mark();                                                                 // 1: frame -> frame LK
  if (any_track) {
    B_RC(b, flv_lk_track(ctx, prev0, cur0, S, q.n_lk, q.lk_prev, q.lk_init, q.lk_next, q.lk_st, /*err: unused by the reference (lkorb_tracking.cpp:36)*/ nullptr, &lk_f2f, FLV_MEM_DEVICE));
    mark();                                                               // 2: keep rule + F RANSAC
    B_RC(b, flv_trk_stage_keep(ctx, b->d, S));
    if (b->fmat_fn) { if (int rc = run_fmat_hooks(b)) return rc; }
    else {
      const flv_ransac_params rp{5.0, 0.99, 1000};
      B_RC(b, flv_fundamental_ransac(ctx, S, q.n_f, q.fa, q.fb, &rp, q.maskF, q.Fm, q.f_ninl, FLV_MEM_DEVICE));
    }
    mark();                                                               // 3: PnP RANSAC
    B_RC(b, flv_trk_stage_after_f(ctx, b->d, S));
    if (b->pnp_fn) { if (int rc = run_pnp_hooks(b)) return rc; }
    else {
      const flv_ransac_params rp{3.0, 0.99, 100};
      B_RC(b, flv_pnp_ransac(ctx, S, q.n_pnp, q.pnp3, q.pnp2, q.pnp_K4, q.pnp_Tin, &rp, q.pnp_Tout, q.maskP, q.pnp_ninl, FLV_MEM_DEVICE));
    }
    mark();                                                               // 4: pose-only BA
    B_RC(b, flv_trk_stage_after_pnp(ctx, b->d, S));
    const flv_ba_params bp{2, 2, 1.0, 3.0, 10, 0};
    B_RC(b, flv_ba_optimize(ctx, S, q.ba_prob, &bp, q.ba_poses, q.ba_lms, q.ba_ep, q.ba_el, q.ba_uv, q.ba_act, q.ba_stats, FLV_MEM_DEVICE));
    mark();                                                               // 5: reprojection cull
    B_RC(b, flv_trk_stage_after_ba(ctx, b->d, S));
    B_RC(b, flv_reprojection_inliers(ctx, S, q.n_rep, &b->d.cam.c, b->d.C.T, b->d.C.undist, b->d.C.p3w, 1.5, b->d.C.inl, q.rep_mean, FLV_MEM_DEVICE));
    B_RC(b, flv_trk_stage_erase_outliers(ctx, b->d, S));
    mark();                                                               // 6: FeatureDEM redetect
    // FeatureDEM::redetect: existing positions were written straight into the context's buffers by the kernel above
    if (ctx->prep_valid) { B_CUDA(b, cudaStreamWaitEvent(cs, ctx->ev_gftt, 0)); ctx->prep_valid = 0; }
    else B_RC(b, flv_launch_gftt(ctx, cur0, S, b->fprm.gftt_num, b->fprm.gftt_ql, (double)b->fprm.gftt_dis));
    B_RC(b, flv_launch_region(ctx, cur0, S, &b->fprm, 1));
    B_RC(b, flv_trk_stage_append(ctx, b->d, S, 1));
  }
  if (any_init) {                                                         // FeatureDEM::detect for the initialising streams
    if (ctx->prep_valid) { B_CUDA(b, cudaStreamWaitEvent(cs, ctx->ev_gftt, 0)); ctx->prep_valid = 0; }
    B_RC(b, flv_launch_gftt(ctx, cur0, S, 2 * b->fprm.gftt_num, b->fprm.gftt_ql, (double)b->fprm.gftt_dis));
    B_RC(b, flv_launch_region(ctx, cur0, S, &b->fprm, 0));
    B_RC(b, flv_trk_stage_append(ctx, b->d, S, 2));
  }
  while (mark_i < 7) mark();
  mark();                                                                 // 7: left -> right LK
  if (any_track || any_init) {
    if (b->stereo) {
      B_CUDA(b, cudaStreamWaitEvent(cs, b->ev_right, 0));
      B_RC(b, flv_lk_track(ctx, cur0, cur1, S, q.n_r, q.r_prev, q.r_init, q.r_next, q.r_st, /*err unused (camera_frame.cpp:124-128)*/ nullptr, &lk_lr, FLV_MEM_DEVICE));
      B_RC(b, flv_trk_stage_pt1(ctx, b->d, S));
    }
    mark();                                                               // 8: depth innovation + finish
    B_RC(b, flv_depth_innovation(ctx, S, q.n_r, &b->d.cam.c, &b->dprm, b->d.C.T, b->d.C.plane, b->d.C.undist, b->d.C.p3w, b->d.C.p3c,
                                 b->d.C.has, b->d.C.f2d, b->d.C.fpose, q.pt1, q.r_st, q.dat, q.rnd, q.n_rand_used, FLV_MEM_DEVICE));
  }
  B_RC(b, flv_trk_stage_finish(ctx, b->d, S));
  // ---- the frame's only synchronisation: summaries + the accepted frame's landmark lists ----------------------------------
  B_CUDA(b, cudaMemcpyAsync(b->h_out, q.out, (size_t)S * sizeof(TrkOut), cudaMemcpyDeviceToHost, cs));
  B_CUDA(b, cudaMemcpyAsync(b->h_tab, b->d_tab, b->full_readback ? b->tab_bytes : b->lean_bytes, cudaMemcpyDeviceToHost, cs));
  while (mark_i <= flv_f2f_batch::NSTAGE) mark();
  B_CUDA(b, cudaEventRecord(b->ev_done, cs));
  b->pend.kf = new_keyframe; b->pend.rs = reset_cmd; b->pend.prev0 = prev0; b->pend.cur0 = cur0; b->pend.cur1 = cur1;
  b->pend.tp0 = tp0; b->pend.tp1 = tp1; b->pend.tp2 = std::chrono::steady_clock::now(); b->pend.active = true;
  return FLV_OK;
}

// second half: wait for the frame, per-stream state machines, keyframe hand-off, slot rotation
int feed_end(flv_f2f_batch* b) {
  if (!b || !b->pend.active) return FLV_ERR_INVALID;
  b->pend.active = false;
  flv_ctx* ctx = b->ctx;
  const int S = b->S, M = b->M;
  int* new_keyframe = b->pend.kf; int* reset_cmd = b->pend.rs;
  (void)reset_cmd;
  const int prev0 = b->pend.prev0, cur0 = b->pend.cur0, cur1 = b->pend.cur1;
  const auto tp0 = b->pend.tp0, tp1 = b->pend.tp1, tp2 = b->pend.tp2;
  cudaStream_t cs = ctx->stream;
  const TrkBufs& q = b->d.b;
  (void)q;
  B_CUDA(b, cudaEventSynchronize(b->ev_done));
  const auto tp3 = std::chrono::steady_clock::now();
  if (b->profile) {
    for (int i = 0; i < flv_f2f_batch::NSTAGE; ++i) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, b->ev_stage[i], b->ev_stage[i + 1]) == cudaSuccess) b->stage_ms[i] += ms;
    }
    b->prof_frames++;
  }
  // ---- per-stream state machine after the frame (f2f_tracking.cpp:229-247, :258-355, :376-394) --------------------------
  bool any_restore = false;
  b->kf_streams.clear(); b->kf_counts.clear(); b->kf_frame.clear(); b->kf_ids.clear(); b->kf_2d.clear(); b->kf_3d.clear(); b->kf_T.clear();
  for (int s = 0; s < S; ++s) {
    StreamState& z = b->st[s];
    const TrkCtl& c = b->h_ctl[s];
    const TrkOut& o = b->h_out[s];
    bool accepted = false, is_kf = false;
    if (c.mode == 2) {
      for (int i = 0; i < o.rand_used; ++i) z.rnd.rand();
      if (o.ok && o.valid_cnt > 30) {                                     // init_frame succeeded (:443-452)
        z.T_kf = z.cur_T;
        if (new_keyframe) new_keyframe[s] = 1;
        is_kf = true;
        z.state = Tracking;
        accepted = true;
      } else if (z.state == TrackingFail) {                               // last_frame.swap(curr_frame)
        std::swap(z.cur_time, z.last_time); std::swap(z.cur_T, z.last_T);
      }
      z.n_lm = (o.committed ? o.n_final : ((const int*)(b->h_tab + b->o_n))[s]);
    } else if (c.mode == 1) {
      z.of_cnt = o.of_cnt; z.f_cnt = o.f_cnt; z.pnp_cnt = o.pnp_cnt;
      if (o.fail_stage >= 1 && o.fail_stage <= 3) {                       // LKORBTracking::tracking returned false (:229-247)
        z.continus_tracking_fail_cnt++;
      } else if (o.fail_stage >= 4) {                                     // OptimizeInFrame::optimize returned false (:258-270)
        z.continus_tracking_fail_cnt = 0;
        z.continus_tracking_fail_cnt++;
      }
      if (o.fail_stage) {
        std::swap(z.cur_time, z.last_time); std::swap(z.cur_T, z.last_T);
        if (z.continus_tracking_fail_cnt >= 2) { z.state = TrackingFail; z.continus_tracking_fail_cnt = 0; }
      } else {
        z.continus_tracking_fail_cnt = 0;
        for (int i = 0; i < o.rand_used; ++i) z.rnd.rand();
        z.cur_T = raw7(o.T);
        if (z.has_imu) z.vim->viCorrectionFromVision(z.cur_time, z.cur_T, z.last_time, z.last_T, o.reproj_err);
        const SE3 T_diff = z.T_kf * z.cur_T.inverse();
        const Vec3 r = so3_log(T_diff.q);
        const double t_norm = std::fabs(T_diff.t[0]) + std::fabs(T_diff.t[1]) + std::fabs(T_diff.t[2]);
        const double r_norm = std::fabs(r[0]) + std::fabs(r[1]) + std::fabs(r[2]);
        bool kf = false;
        if (z.frameCount < 40 && (z.frameCount % 5) == 0) { kf = true; z.T_kf = z.cur_T; }
        if (t_norm >= 0.05 || r_norm >= 0.2) { kf = true; z.T_kf = z.cur_T; }
        if (kf && new_keyframe) new_keyframe[s] = 1;
        is_kf = kf;
        accepted = true;
      }
      z.n_lm = ((const int*)(b->h_tab + b->o_n))[s];
    } else {
      // idle frame: TrackingFail swaps back (:377-393); UnInit / skipped frames leave an empty current frame
      if (z.state == TrackingFail) { std::swap(z.cur_time, z.last_time); std::swap(z.cur_T, z.last_T); z.n_lm = ((const int*)(b->h_tab + b->o_n))[s]; }
    }
    if (is_kf && b->lmap) {                             // CameraFrame::getKeyFrameInf (camera_frame.cpp:515-529)
      const size_t k0 = (size_t)s * M;
      const unsigned char* tb = b->h_tab;
      int cnt = 0;
      for (int i = 0; i < z.n_lm; ++i) {
        const size_t k = k0 + i;
        if (!(tb + b->o_has)[k] || !(tb + b->o_inl)[k]) continue;
        b->kf_ids.push_back(((const long long*)(tb + b->o_id))[k]);
        b->kf_2d.push_back(((const double*)(tb + b->o_und))[2 * k]); b->kf_2d.push_back(((const double*)(tb + b->o_und))[2 * k + 1]);
        for (int c3 = 0; c3 < 3; ++c3) b->kf_3d.push_back(((const double*)(tb + b->o_p3w))[3 * k + c3]);
        ++cnt;
      }
      double T7[7]; to7(z.cur_T, T7);
      b->kf_T.insert(b->kf_T.end(), T7, T7 + 7);
      b->kf_streams.push_back(b->stream_offset + s); b->kf_counts.push_back(cnt); b->kf_frame.push_back(z.frameCount);
    }
    if (b->rlog) {
      double* row = b->rlog + ((size_t)(b->rlog_frame % b->rlog_K) * b->S_total + b->stream_offset + s) * 8;
      to7(z.cur_T, row);
      row[7] = z.n_lm;
    }
    if (accepted) b->have_last[s] = 1;
    else if (b->have_last[s]) {
      // the stream keeps its old "last" frame: its image must survive the slot rotation below
      const size_t stride = ctx->geom.stream_stride;
      B_CUDA(b, cudaMemcpyAsync(ctx->pyr[cur0] + (size_t)s * stride, ctx->pyr[prev0] + (size_t)s * stride, stride, cudaMemcpyDeviceToDevice, cs));
      any_restore = true;
    }
  }
  if (b->lmap && !b->kf_streams.empty()) {
    const int rc = flv_localmap_batch_submit(b->lmap, (int)b->kf_streams.size(), b->kf_streams.data(), b->kf_frame.data(), b->kf_counts.data(),
                                             b->kf_ids.data(), b->kf_2d.data(), b->kf_3d.data(), b->kf_T.data());
    if (rc) { snprintf(b->err, sizeof(b->err), "flv_localmap_batch_submit: %s", flv_localmap_batch_last_error(b->lmap)); return rc; }
  }
  if (b->rlog) b->rlog_frame++;
  if (any_restore) ctx->deriv_streams[cur0] = 0;          // derivative pyramid of that slot: rebuild on next use
  if (b->profile) {
    const auto tp4 = std::chrono::steady_clock::now();
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point c) { return std::chrono::duration<double, std::milli>(c - a).count(); };
    b->host_ms[0] += ms(tp0, tp1); b->host_ms[1] += ms(tp1, tp2); b->host_ms[2] += ms(tp2, tp3); b->host_ms[3] += ms(tp3, tp4);
  }
  b->slots[0] = cur0; b->slots[1] = prev0; b->slots[2] = cur1;
  return FLV_OK;
}

}  // namespace

extern "C" {

// Pipelined form of imu_feed_many + image_feed for grouped batches: for every group in turn, finish the group's previous frame
// (wait, state machines, keyframe hand-off), feed the group's IMU samples, enqueue its new frame -- so while the host works
// on one group the other groups' frames are running on the GPU.  Per stream the order of operations is exactly that of the
// synchronous calls (previous frame finished -> IMU samples -> new frame); only the results become visible one call later
// (flv_f2f_batch_sync finishes the frames in flight).
int flv_f2f_batch_frame_async(flv_f2f_batch* b, const double* t, const uint8_t* img0, const void* img1, flv_memspace mem, int n_imu,
                              const int* imu_streams, const double* imu_t, const double* imu_acc, const double* imu_gyro) {
  if (!b || !t || !img0 || !img1 || n_imu < 0 || (n_imu > 0 && (!imu_streams || !imu_t || !imu_acc || !imu_gyro))) return FLV_ERR_INVALID;
  const size_t w = b->cfg.img_w, h = b->cfg.img_h, px1 = b->stereo ? 1 : 2;
  const size_t G = b->sub.empty() ? 1 : b->sub.size();
  for (size_t g = 0; g < G; ++g) {
    flv_f2f_batch* sb = b->sub.empty() ? b : b->sub[g];
    const int first = b->sub.empty() ? 0 : (int)g * b->per_group, last = first + sb->S;
    if (sb->pend.active) {
      const int rc = feed_end(sb);
      if (rc) { if (sb != b) snprintf(b->err, sizeof(b->err), "group %d: %s", (int)g, sb->err); return rc; }
    }
    for (int i = 0; i < n_imu; ++i) {
      const int s = imu_streams[i];
      if (s < first || s >= last) continue;
      const int rc = flv_f2f_batch_imu_feed(sb, s - first, imu_t[i], imu_acc + 3 * (size_t)i, imu_gyro + 3 * (size_t)i);
      if (rc) return rc;
    }
    const int rc = feed_begin(sb, t + first, img0 + (size_t)first * w * h, (const uint8_t*)img1 + (size_t)first * w * h * px1, mem,
                              sb->async_kf.data(), sb->async_rs.data());
    if (rc) { if (sb != b) snprintf(b->err, sizeof(b->err), "group %d: %s", (int)g, sb->err); return rc; }
  }
  return FLV_OK;
}

// finish every frame in flight; new_keyframe / reset_cmd (may be NULL) receive the flags of the last frame of every stream
int flv_f2f_batch_sync(flv_f2f_batch* b, int* new_keyframe, int* reset_cmd) {
  if (!b) return FLV_ERR_INVALID;
  const size_t G = b->sub.empty() ? 1 : b->sub.size();
  for (size_t g = 0; g < G; ++g) {
    flv_f2f_batch* sb = b->sub.empty() ? b : b->sub[g];
    const int first = b->sub.empty() ? 0 : (int)g * b->per_group;
    if (sb->pend.active) {
      const int rc = feed_end(sb);
      if (rc) { if (sb != b) snprintf(b->err, sizeof(b->err), "group %d: %s", (int)g, sb->err); return rc; }
    }
    for (int s = 0; s < sb->S; ++s) {
      if (new_keyframe) new_keyframe[first + s] = sb->async_kf[s];
      if (reset_cmd) reset_cmd[first + s] = sb->async_rs[s];
    }
  }
  return FLV_OK;
}

int flv_f2f_batch_image_feed(flv_f2f_batch* b, const double* t, const uint8_t* img0, const void* img1, flv_memspace mem,
                             int* new_keyframe, int* reset_cmd) {
  if (!b || !t || !img0 || !img1) return FLV_ERR_INVALID;
  if (int rc0 = flv_f2f_batch_sync(b, nullptr, nullptr)) return rc0;      // a frame started by frame_async finishes first
  if (b->sub.empty()) {
    const int rc = feed_begin(b, t, img0, img1, mem, new_keyframe, reset_cmd);
    return rc ? rc : feed_end(b);
  }
  // grouped: every group's frame is in flight before the first wait
  const size_t w = b->cfg.img_w, h = b->cfg.img_h, px1 = b->stereo ? 1 : 2;
  int rc = FLV_OK;
  size_t started = 0;
  for (size_t g = 0; g < b->sub.size() && !rc; ++g) {
    const size_t first = g * (size_t)b->per_group;
    rc = feed_begin(b->sub[g], t + first, img0 + first * w * h, (const uint8_t*)img1 + first * w * h * px1, mem,
                    new_keyframe ? new_keyframe + first : nullptr, reset_cmd ? reset_cmd + first : nullptr);
    if (rc) snprintf(b->err, sizeof(b->err), "group %d: %s", (int)g, b->sub[g]->err);
    else ++started;
  }
  for (size_t g = 0; g < started; ++g) {
    const int rc2 = feed_end(b->sub[g]);
    if (rc2 && !rc) { rc = rc2; snprintf(b->err, sizeof(b->err), "group %d: %s", (int)g, b->sub[g]->err); }
  }
  return rc;
}

int flv_f2f_batch_state(flv_f2f_batch* b, int stream) {
  if (!b || stream < 0 || stream >= b->S) return FLV_ERR_INVALID;
  b = sub_of(b, stream);
  return b->st[stream].state;
}

int flv_f2f_batch_get_frame(flv_f2f_batch* b, int stream, double* T_c_w, int64_t* lm_id, double* plane_xy, double* undist_xy,
                            double* p3d_w, uint8_t* has_3d, uint8_t* is_inlier, int cap) {
  if (!b || stream < 0 || stream >= b->S) return FLV_ERR_INVALID;
  b = sub_of(b, stream);
  const StreamState& z = b->st[stream];
  if (T_c_w) to7(z.cur_T, T_c_w);
  const size_t k0 = (size_t)stream * b->M;
  const int n = z.n_lm < cap ? z.n_lm : cap;
  const unsigned char* t = b->h_tab;
  for (int i = 0; i < n; ++i) {
    const size_t k = k0 + i;
    if (lm_id) lm_id[i] = ((const long long*)(t + b->o_id))[k];
    if (plane_xy) { plane_xy[2 * i] = ((const double*)(t + b->o_plane))[2 * k]; plane_xy[2 * i + 1] = ((const double*)(t + b->o_plane))[2 * k + 1]; }
    if (undist_xy) { undist_xy[2 * i] = ((const double*)(t + b->o_und))[2 * k]; undist_xy[2 * i + 1] = ((const double*)(t + b->o_und))[2 * k + 1]; }
    if (p3d_w) for (int c = 0; c < 3; ++c) p3d_w[3 * i + c] = ((const double*)(t + b->o_p3w))[3 * k + c];
    if (has_3d) has_3d[i] = (t + b->o_has)[k];
    if (is_inlier) is_inlier[i] = (t + b->o_inl)[k];
  }
  return n;
}

int flv_f2f_batch_get_frame_ex(flv_f2f_batch* b, int stream, double* p3d_c, double* first_obs_2d, double* first_obs_pose,
                               double* T_c_w_last_keyframe, int cap) {
  if (!b || stream < 0 || stream >= b->S) return FLV_ERR_INVALID;
  b = sub_of(b, stream);
  const StreamState& z = b->st[stream];
  if (T_c_w_last_keyframe) to7(z.T_kf, T_c_w_last_keyframe);
  const size_t k0 = (size_t)stream * b->M;
  const int n = z.n_lm < cap ? z.n_lm : cap;
  const unsigned char* t = b->h_tab;
  for (int i = 0; i < n; ++i) {
    const size_t k = k0 + i;
    if (p3d_c) for (int c = 0; c < 3; ++c) p3d_c[3 * i + c] = ((const double*)(t + b->o_p3c))[3 * k + c];
    if (first_obs_2d) { first_obs_2d[2 * i] = ((const double*)(t + b->o_f2d))[2 * k]; first_obs_2d[2 * i + 1] = ((const double*)(t + b->o_f2d))[2 * k + 1]; }
    if (first_obs_pose) for (int c = 0; c < 7; ++c) first_obs_pose[7 * i + c] = ((const double*)(t + b->o_fpose))[7 * k + c];
  }
  return n;
}

int flv_f2f_batch_get_imu_states(flv_f2f_batch* b, int stream, double* out11, int cap) {
  if (!b || stream < 0 || stream >= b->S) return FLV_ERR_INVALID;
  b = sub_of(b, stream);
  return b->st[stream].vim->dump_states(out11, cap);
}
int flv_f2f_batch_get_imu_bias(flv_f2f_batch* b, int stream, double* acc_bias, double* gyro_bias) {
  if (!b || stream < 0 || stream >= b->S) return FLV_ERR_INVALID;
  b = sub_of(b, stream);
  const StreamState& z = b->st[stream];
  for (int k = 0; k < 3; ++k) { if (acc_bias) acc_bias[k] = z.vim->acc_bias[k]; if (gyro_bias) gyro_bias[k] = z.vim->gyro_bias[k]; }
  return z.has_imu ? 1 : 0;
}
int flv_f2f_batch_tracking_counts(flv_f2f_batch* b, int stream, int* of, int* fi, int* pnp) {
  if (!b || stream < 0 || stream >= b->S) return FLV_ERR_INVALID;
  b = sub_of(b, stream);
  const StreamState& z = b->st[stream];
  if (of) *of = z.of_cnt; if (fi) *fi = z.f_cnt; if (pnp) *pnp = z.pnp_cnt;
  return FLV_OK;
}
int flv_f2f_batch_imu_feed_many(flv_f2f_batch* b, int n, const int* streams, const double* t, const double* acc, const double* gyro) {
  if (!b || n < 0 || (n > 0 && (!streams || !t || !acc || !gyro))) return FLV_ERR_INVALID;
  for (int i = 0; i < n; ++i) {
    const int rc = flv_f2f_batch_imu_feed(b, streams[i], t[i], acc + 3 * (size_t)i, gyro + 3 * (size_t)i);
    if (rc) return rc;
  }
  return FLV_OK;
}
int flv_f2f_batch_set_profile(flv_f2f_batch* b, int enable) {
  if (!b) return FLV_ERR_INVALID;
  if (!b->sub.empty()) { for (flv_f2f_batch* sb : b->sub) { const int rc = flv_f2f_batch_set_profile(sb, enable); if (rc) return rc; } return FLV_OK; }
  if (enable && !b->ev_stage[0])
    for (int i = 0; i <= flv_f2f_batch::NSTAGE; ++i) B_CUDA(b, cudaEventCreate(&b->ev_stage[i]));
  b->profile = enable != 0;
  for (double& v : b->stage_ms) v = 0;
  for (double& v : b->host_ms) v = 0;
  b->prof_frames = 0;
  return FLV_OK;
}
int flv_f2f_batch_get_profile(flv_f2f_batch* b, double* stage_ms9, long long* frames) {
  if (!b || !stage_ms9) return FLV_ERR_INVALID;
  if (!b->sub.empty()) {                               // mean over the groups (their stages run concurrently)
    for (int i = 0; i < flv_f2f_batch::NSTAGE; ++i) stage_ms9[i] = 0;
    for (flv_f2f_batch* sb : b->sub) for (int i = 0; i < flv_f2f_batch::NSTAGE; ++i) stage_ms9[i] += sb->stage_ms[i] / (double)b->sub.size();
    if (frames) *frames = b->sub[0]->prof_frames;
    return flv_f2f_batch::NSTAGE;
  }
  for (int i = 0; i < flv_f2f_batch::NSTAGE; ++i) stage_ms9[i] = b->stage_ms[i];
  if (frames) *frames = b->prof_frames;
  return flv_f2f_batch::NSTAGE;
}
int flv_f2f_batch_get_host_profile(flv_f2f_batch* b, double* host_ms4) {
  if (!b || !host_ms4) return FLV_ERR_INVALID;
  if (!b->sub.empty()) {
    for (int i = 0; i < 4; ++i) host_ms4[i] = 0;
    for (flv_f2f_batch* sb : b->sub) for (int i = 0; i < 4; ++i) host_ms4[i] += sb->host_ms[i] / (double)b->sub.size();
    return 4;
  }
  for (int i = 0; i < 4; ++i) host_ms4[i] = b->host_ms[i];
  return 4;
}
long long flv_f2f_batch_launch_count(flv_f2f_batch* b) {
  if (b && !b->sub.empty()) { long long n = 0; for (flv_f2f_batch* sb : b->sub) n += flv_f2f_batch_launch_count(sb); return n; }
  return b && b->ctx ? flv_launch_count(b->ctx) : 0;
}
int flv_f2f_batch_set_result_log(flv_f2f_batch* b, double* buf, int block_frames) {
  if (!b || (buf && block_frames < 1)) return FLV_ERR_INVALID;
  b->rlog = buf; b->rlog_K = block_frames; b->S_total = b->S; b->rlog_frame = 0;
  for (flv_f2f_batch* sb : b->sub) { sb->rlog = buf; sb->rlog_K = block_frames; sb->S_total = b->S; sb->rlog_frame = 0; }
  return FLV_OK;
}
int flv_f2f_batch_set_readback(flv_f2f_batch* b, int full) {
  if (!b) return FLV_ERR_INVALID;
  b->full_readback = full != 0;
  for (flv_f2f_batch* sb : b->sub) sb->full_readback = full != 0;
  return FLV_OK;
}
int flv_f2f_batch_attach_localmap(flv_f2f_batch* b, flv_localmap_batch* lm) {
  if (!b) return FLV_ERR_INVALID;
  b->lmap = lm;
  for (flv_f2f_batch* sb : b->sub) sb->lmap = lm;
  return FLV_OK;
}

}  // extern "C"
