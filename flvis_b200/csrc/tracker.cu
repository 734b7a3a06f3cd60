// Batched, device-resident frame-to-frame tracker: the glue of F2FTracking::image_feed / LKORBTracking::tracking /
// CameraFrame between the heavy kernels, for S independent camera sequences advanced by ONE launch per stage.
//
// Reference (paths relative to /root/reference): src/frontend/f2f_tracking.cpp:59-453 (frame state machine),
// src/processing/lkorb_tracking.cpp:9-202 (projection with the IMU guess :38-63, keep rule and REVERSED landmark order
// :98-119, mirrored F mask :138-149, PnP selection :150-189), src/processing/camera_frame.cpp:18-40, :399-461 (erase /
// updateLMState), :93-131 (projection into cam1 before the left->right LK), src/processing/landmark.cpp:3-39 (ids),
// src/processing/optimize_in_frame.cpp:10-90 (pose-only BA inputs), src/processing/vi_motion.cpp:437-464
// (viVisionRPCompensation).  The per-stream host part (IMU filter, state machine decisions, keyframe rule) lives in
// batch_tracker.cu; the single-sequence host class flv::F2FTracking (flvis_b200/host/f2f_tracking.cpp) is the same
// pipeline with host hand-offs and is the parity twin of this file (tests/test_batch_tracker_gpu.py).
//
// Layout: two landmark tables per stream (L = the last accepted frame, C = the frame being built), structure of arrays
// with a fixed stride of max_pts (512) landmarks; thread i of the stream's CTA owns landmark i.  Every kernel is launched
// with one CTA of 512 threads per stream; streams that are idle or have failed earlier in the frame run on empty counts.
#include "tracker.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;

// ---- SE3 with the reference's Sophus / Eigen semantics (flvis_b200/host/sophus_lite.h), poses stored [qx qy qz qw t] ---
__device__ __forceinline__ void q_rot(const double* q, const double* v, double* o) {
  double ux = q[1] * v[2] - q[2] * v[1], uy = q[2] * v[0] - q[0] * v[2], uz = q[0] * v[1] - q[1] * v[0];
  ux += ux; uy += uy; uz += uz;
  o[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
  o[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
  o[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}
__device__ __forceinline__ void q_to_R(const double* q, double* R) {
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
__device__ void R_to_q(const double* m, double* q) {          // Eigen Quaternion(Matrix3), host/se3.h:R_to_quat
  double t = m[0] + m[4] + m[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t; t = 0.5 / t;
    q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    q[i] = 0.5 * t; t = 0.5 / t;
    q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
    q[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
  }
}
__device__ __forceinline__ void q_normalize(double* q) {     // Quat order irrelevant: sum of squares in w,x,y,z order
  const double n = sqrt(q[3] * q[3] + q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
__device__ __forceinline__ void q_mul(const double* a, const double* b, double* o) {   // Eigen operator*, xyzw storage
  const double aw = a[3], ax = a[0], ay = a[1], az = a[2], bw = b[3], bx = b[0], by = b[1], bz = b[2];
  o[3] = aw * bw - ax * bx - ay * by - az * bz;
  o[0] = aw * bx + ax * bw + ay * bz - az * by;
  o[1] = aw * by + ay * bw + az * bx - ax * bz;
  o[2] = aw * bz + az * bw + ax * by - ay * bx;
}
__device__ void se3_mul(const double* A, const double* B, double* O) {    // SE3::operator*
  double rt[3], q[4];
  q_rot(A, B + 4, rt);
  q_mul(A, B, q);
  q_normalize(q);
  O[0] = q[0]; O[1] = q[1]; O[2] = q[2]; O[3] = q[3];
  O[4] = A[4] + rt[0]; O[5] = A[5] + rt[1]; O[6] = A[6] + rt[2];
}
__device__ void se3_inv(const double* A, double* O) {                     // SE3::inverse
  double q[4] = {-A[0], -A[1], -A[2], A[3]};
  q_normalize(q);
  const double mt[3] = {-A[4], -A[5], -A[6]};
  double t[3];
  q_rot(q, mt, t);
  O[0] = q[0]; O[1] = q[1]; O[2] = q[2]; O[3] = q[3]; O[4] = t[0]; O[5] = t[1]; O[6] = t[2];
}
__device__ __forceinline__ void se3_from7(const double* p, double* O) {   // SE3(Quat, t): normalises
  O[0] = p[0]; O[1] = p[1]; O[2] = p[2]; O[3] = p[3];
  q_normalize(O);
  O[4] = p[4]; O[5] = p[5]; O[6] = p[6];
}
__device__ __forceinline__ void world2camera(const double* T, const double* pw, double* pc) {
  q_rot(T, pw, pc);
  pc[0] += T[4]; pc[1] += T[5]; pc[2] += T[6];
}
__device__ void rpy2R(const double* rpy, double* R) {                    // kinetic_math.h:17-94
  const double r = rpy[0], p = rpy[1], y = rpy[2];
  const double cy = cos(y), sy = sin(y), cp = cos(p), sp = sin(p), cr = cos(r), sr = sin(r);
  R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
  R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
  R[6] = -sp; R[7] = cp * sr; R[8] = cp * cr;
}
__device__ void Q2rpy(const double* q, double* rpy) {
  double R[9];
  q_to_R(q, R);
  rpy[0] = atan2(R[7], R[8]); rpy[1] = atan2(-R[6], sqrt(R[7] * R[7] + R[8] * R[8])); rpy[2] = atan2(R[3], R[0]);
}
__device__ void g2o_pose_from_quat(const double* in, double* out) {      // host/se3.h
  double R[9], q[4];
  q_to_R(in, R);
  R_to_q(R, q);
  if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  out[0] = q[0] / n; out[1] = q[1] / n; out[2] = q[2] / n; out[3] = q[3] / n; out[4] = in[4]; out[5] = in[5]; out[6] = in[6];
}

// ---- block helpers (TRK_THREADS = 512 = 16 warps) ----------------------------------------------------------------------
// exclusive rank of this thread's flag among the block's set flags (thread order), and the block total
__device__ int block_rank(bool flag, int* wsum, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(FULL, flag);
  __syncthreads();                       // wsum may still be read from a previous call
  if (lane == 0) wsum[warp] = __popc(bal);
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < TRK_THREADS / 32; ++w) { const int c = wsum[w]; if (w < warp) base += c; tot += c; }
  total = tot;
  return base + __popc(bal & ((1u << lane) - 1));
}

struct LmReg {
  long long id; double plane[2], und[2], p3w[3], p3c[3], f2d[2], fpose[7]; unsigned char has, inl;
};
__device__ __forceinline__ void lm_load(const TrkTable& t, size_t k, LmReg& r) {
  r.id = t.id[k];
  r.plane[0] = t.plane[2 * k]; r.plane[1] = t.plane[2 * k + 1];
  r.und[0] = t.undist[2 * k]; r.und[1] = t.undist[2 * k + 1];
#pragma unroll
  for (int c = 0; c < 3; ++c) { r.p3w[c] = t.p3w[3 * k + c]; r.p3c[c] = t.p3c[3 * k + c]; }
  r.f2d[0] = t.f2d[2 * k]; r.f2d[1] = t.f2d[2 * k + 1];
#pragma unroll
  for (int c = 0; c < 7; ++c) r.fpose[c] = t.fpose[7 * k + c];
  r.has = t.has[k]; r.inl = t.inl[k];
}
__device__ __forceinline__ void lm_store(const TrkTable& t, size_t k, const LmReg& r) {
  t.id[k] = r.id;
  t.plane[2 * k] = r.plane[0]; t.plane[2 * k + 1] = r.plane[1];
  t.undist[2 * k] = r.und[0]; t.undist[2 * k + 1] = r.und[1];
#pragma unroll
  for (int c = 0; c < 3; ++c) { t.p3w[3 * k + c] = r.p3w[c]; t.p3c[3 * k + c] = r.p3c[c]; }
  t.f2d[2 * k] = r.f2d[0]; t.f2d[2 * k + 1] = r.f2d[1];
#pragma unroll
  for (int c = 0; c < 7; ++c) t.fpose[7 * k + c] = r.fpose[c];
  t.has[k] = r.has; t.inl[k] = r.inl;
}

__device__ __forceinline__ void fail(const TrkBufs& b, int s, int stage) {      // thread 0 only
  b.ok[s] = 0;
  b.out[s].ok = 0;
  b.out[s].fail_stage = stage;
}

// ---- stage kernels -----------------------------------------------------------------------------------------------------
// (1) start of the frame + LK inputs: prev = float copies of the last frame's pixel positions, init = prev or the
//     projection of the landmark with the IMU pose guess (lkorb_tracking.cpp:38-63)
__global__ void __launch_bounds__(TRK_THREADS) trk_prepare_kernel(TrkTable L, TrkTable C, TrkBufs b, TrkCam cam) {
  const int s = blockIdx.x, i = threadIdx.x, M = b.max_pts;
  const TrkCtl ctl = b.ctl[s];
  const int n = ctl.mode == 1 ? min(L.n[s], M) : 0;
  const size_t k = (size_t)s * M + i;
  if (i == 0) {
    TrkOut o;
    o.ok = ctl.mode != 0; o.fail_stage = 0; o.of_cnt = o.f_cnt = o.pnp_cnt = 0; o.n_final = 0; o.valid_cnt = 0;
    o.rand_used = 0; o.n_new = 0; o.committed = 0; o.mode = ctl.mode; o.reserved = 0;
    for (int c = 0; c < 7; ++c) o.T[c] = c == 3 ? 1.0 : 0.0;
    o.reproj_err = 0; o.id_index = b.id_index[s];
    b.out[s] = o;
    b.ok[s] = ctl.mode != 0;
    b.n_lk[s] = n;
    if (ctl.mode == 2) { C.n[s] = 0; for (int c = 0; c < 7; ++c) C.T[7 * s + c] = ctl.init_pose[c]; }
    b.n_f[s] = 0; b.n_pnp[s] = 0; b.n_rep[s] = 0; b.n_exist[s] = 0; b.n_r[s] = 0;
    b.ba_prob[s].n_poses = 0;
  }
  if (i >= n) return;
  const float px = (float)L.plane[2 * k], py = (float)L.plane[2 * k + 1];
  float ix = px, iy = py;
  if (ctl.use_guess) {
    const float X = (float)L.p3w[3 * k], Y = (float)L.p3w[3 * k + 1], Z = (float)L.p3w[3 * k + 2];
    if (cam.unrect) {
      double Rg[9];
      q_to_R(ctl.guess, Rg);
      flv::project_point(cam.lens0, Rg, ctl.guess + 4, X, Y, Z, ix, iy);
    } else {
      const double pw[3] = {(double)X, (double)Y, (double)Z};
      double pc[3];
      world2camera(ctl.guess, pw, pc);
      ix = (float)(cam.c.fx * pc[0] / pc[2] + cam.c.cx);
      iy = (float)(cam.c.fy * pc[1] / pc[2] + cam.c.cy);
    }
  }
  b.lk_prev[2 * k] = px; b.lk_prev[2 * k + 1] = py;
  b.lk_init[2 * k] = ix; b.lk_init[2 * k + 1] = iy;
}

// (2) keep rule (status, strictly inside the image), undistortion of the tracked points, the new frame's landmark list in
//     REVERSED order, forward-ordered point pairs for the fundamental-matrix test (lkorb_tracking.cpp:74-133)
__global__ void __launch_bounds__(TRK_THREADS) trk_keep_kernel(TrkTable L, TrkTable C, TrkBufs b, TrkCam cam) {
  __shared__ int wsum[TRK_THREADS / 32];
  const int s = blockIdx.x, i = threadIdx.x, M = b.max_pts;
  const int n = b.n_lk[s];
  const size_t k = (size_t)s * M + i;
  bool keep = false;
  float tx = 0, ty = 0, ux = 0, uy = 0;
  if (i < n) {
    tx = b.lk_next[2 * k]; ty = b.lk_next[2 * k + 1];
    ux = tx; uy = ty;
    if (cam.unrect) flv::undistort_point(cam.lens0, tx, ty, ux, uy);
    const int wlim = cam.w - 1, hlim = cam.h - 1;
    keep = b.lk_st[k] == 1 && tx > 0 && ty > 0 && tx < wlim && ty < hlim;
  }
  int K;
  const int r = block_rank(keep, wsum, K);
  if (keep) {
    LmReg lm;
    lm_load(L, k, lm);
    b.fa[2 * ((size_t)s * M + r)] = (float)lm.und[0]; b.fa[2 * ((size_t)s * M + r) + 1] = (float)lm.und[1];
    b.fb[2 * ((size_t)s * M + r)] = ux; b.fb[2 * ((size_t)s * M + r) + 1] = uy;
    lm.plane[0] = (double)tx; lm.plane[1] = (double)ty; lm.und[0] = (double)ux; lm.und[1] = (double)uy;
    lm_store(C, (size_t)s * M + (K - 1 - r), lm);
  }
  if (i == 0 && b.ctl[s].mode == 1) {
    C.n[s] = K;
    b.out[s].of_cnt = K;
    if (K < 10) fail(b, s, 1);
    b.n_f[s] = b.ok[s] ? K : 0;
    for (int c = 0; c < 7; ++c) b.pnp_Tin[7 * s + c] = b.ctl[s].use_guess ? b.ctl[s].guess[c] : L.T[7 * s + c];
  }
}

// (3) the F mask is applied to the REVERSED list with the forward index (the reference's mirrored indexing, kept);
//     correspondences with depth that are still inliers go to PnP (lkorb_tracking.cpp:138-169)
__global__ void __launch_bounds__(TRK_THREADS) trk_after_f_kernel(TrkTable C, TrkBufs b) {
  __shared__ int wsum[TRK_THREADS / 32];
  const int s = blockIdx.x, i = threadIdx.x, M = b.max_pts;
  const bool alive = b.ok[s] != 0 && b.ctl[s].mode == 1;
  const int n = alive ? C.n[s] : 0;
  const size_t k = (size_t)s * M + i;
  bool inl = false, sel = false;
  if (i < n) {
    if (b.maskF[k] == 0) C.inl[k] = 0;
    inl = C.inl[k] != 0;
    sel = inl && C.has[k] != 0;
  }
  int fcnt, npnp;
  block_rank(inl, wsum, fcnt);
  const int r = block_rank(sel, wsum, npnp);
  const bool good = alive && fcnt >= 10;
  if (sel && good) {
    const size_t o = (size_t)s * M + r;
    b.pnp2[2 * o] = (float)C.undist[2 * k]; b.pnp2[2 * o + 1] = (float)C.undist[2 * k + 1];
    b.pnp3[3 * o] = (float)C.p3w[3 * k]; b.pnp3[3 * o + 1] = (float)C.p3w[3 * k + 1]; b.pnp3[3 * o + 2] = (float)C.p3w[3 * k + 2];
  }
  if (i == 0 && alive) {
    b.out[s].f_cnt = fcnt;
    if (fcnt < 10) fail(b, s, 2);
    b.n_pnp[s] = good ? npnp : 0;
  }
}

// (4) PnP inlier mask -> landmark flags (updateLMState, camera_frame.cpp:445-461), pose, IMU roll / pitch blend
//     (viVisionRPCompensation, vi_motion.cpp:437-464), inputs of the pose-only BA (optimize_in_frame.cpp:20-63)
__global__ void __launch_bounds__(TRK_THREADS) trk_after_pnp_kernel(TrkTable C, TrkBufs b, TrkCam cam) {
  __shared__ int wsum[TRK_THREADS / 32];
  __shared__ double shT[7];
  __shared__ int sh_good;
  const int s = blockIdx.x, i = threadIdx.x, M = b.max_pts;
  const bool alive = b.ok[s] != 0 && b.ctl[s].mode == 1;
  const int n = alive ? C.n[s] : 0;
  const size_t k = (size_t)s * M + i;
  bool sel = i < n && C.inl[k] != 0 && C.has[k] != 0;
  int nsel;
  const int r = block_rank(sel, wsum, nsel);
  if (sel && b.maskP[(size_t)s * M + r] == 0) { C.inl[k] = 0; sel = false; }
  if (i == 0) {
    sh_good = 0;
    if (alive) {
      const int ninl = b.pnp_ninl[s];
      b.out[s].pnp_cnt = ninl;
      double T[7];
      se3_from7(b.pnp_Tout + 7 * s, T);
      if (ninl < 10) {
        for (int c = 0; c < 7; ++c) C.T[7 * s + c] = T[c];
        fail(b, s, 3);
      } else {
        const TrkCtl& ctl = b.ctl[s];
        if (ctl.rp_found) {
          double Tinv[7], Twi[7], rpy[3], R[9], q[4], Ta[7], Tb[7];
          se3_inv(T, Tinv);
          se3_mul(Tinv, cam.T_c_i, Twi);
          Q2rpy(Twi, rpy);
          const double p2 = cam.vi_para2;
          double ra[3];
          ra[0] = rpy[0] * (1 - p2) + ctl.roll * p2;
          ra[1] = rpy[1] * (1 - p2) + ctl.pitch * p2;
          ra[2] = rpy[2] * (1 - p2) + rpy[2] * p2;
          rpy2R(ra, R);
          R_to_q(R, q);
          Ta[0] = q[0]; Ta[1] = q[1]; Ta[2] = q[2]; Ta[3] = q[3];
          q_normalize(Ta);
          Ta[4] = Twi[4]; Ta[5] = Twi[5]; Ta[6] = Twi[6];
          se3_mul(Ta, cam.T_i_c, Tb);
          se3_inv(Tb, T);
        }
        for (int c = 0; c < 7; ++c) { C.T[7 * s + c] = T[c]; shT[c] = T[c]; }
        sh_good = 1;
      }
    }
  }
  __syncthreads();
  // landmarks that are (still) inliers with depth: edges of the pose-only problem, in landmark order
  int nba;
  const int rb = block_rank(sel, wsum, nba);
  const bool good = sh_good && nba >= 10;
  if (sel && good) {
    const size_t ol = (size_t)s * b.ba_ML + rb, oe = (size_t)s * b.ba_ME + rb;
    b.ba_lms[3 * ol] = C.p3w[3 * k]; b.ba_lms[3 * ol + 1] = C.p3w[3 * k + 1]; b.ba_lms[3 * ol + 2] = C.p3w[3 * k + 2];
    b.ba_uv[2 * oe] = C.undist[2 * k]; b.ba_uv[2 * oe + 1] = C.undist[2 * k + 1];
    b.ba_ep[oe] = 0; b.ba_el[oe] = rb; b.ba_act[oe] = 1;
  }
  if (i == 0 && sh_good) {
    if (nba < 10) fail(b, s, 4);
    else {
      g2o_pose_from_quat(shT, b.ba_poses + 7 * (size_t)s * b.ba_MP);
      flv_ba_problem pb;
      pb.n_poses = 1; pb.n_landmarks = nba; pb.n_edges = nba; pb.fixed_pose = -1; pb.fix_landmarks = 1;
      pb.fx = cam.c.fx; pb.fy = cam.c.fy; pb.cx = cam.c.cx; pb.cy = cam.c.cy;
      b.ba_prob[s] = pb;
    }
  }
}

// (5) pose from the pose-only BA (optimize_in_frame.cpp:79-88: fewer than 10 edges after the cull = failure, pose kept)
__global__ void trk_after_ba_kernel(TrkTable C, TrkBufs b, int S) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  if (!b.ok[s] || b.ctl[s].mode != 1) { b.n_rep[s] = 0; return; }
  if (!b.ba_stats[s].ok) { fail(b, s, 5); b.n_rep[s] = 0; return; }
  double T[7];
  se3_from7(b.ba_poses + 7 * (size_t)s * b.ba_MP, T);
  for (int c = 0; c < 7; ++c) C.T[7 * s + c] = T[c];
  b.n_rep[s] = C.n[s];
}

// (6) eraseReprjOutlier (camera_frame.cpp:18-28) + the position list FeatureDEM::redetect takes (f2f_tracking.cpp:286-291)
__global__ void __launch_bounds__(TRK_THREADS) trk_erase_outliers_kernel(TrkTable C, TrkBufs b, double* exist, int* n_exist) {
  __shared__ int wsum[TRK_THREADS / 32];
  const int s = blockIdx.x, i = threadIdx.x, M = b.max_pts;
  const bool alive = b.ok[s] != 0 && b.ctl[s].mode == 1;
  const int n = alive ? C.n[s] : 0;
  const size_t k = (size_t)s * M + i;
  LmReg lm;
  bool keep = false;
  if (i < n) { lm_load(C, k, lm); keep = lm.inl != 0; }
  int K;
  const int r = block_rank(keep, wsum, K);
  if (keep) {
    const size_t o = (size_t)s * M + r;
    lm_store(C, o, lm);
    exist[2 * o] = lm.plane[0]; exist[2 * o + 1] = lm.plane[1];
  }
  if (i == 0) {
    if (alive) { C.n[s] = K; b.orig_size[s] = K; b.out[s].reproj_err = b.rep_mean[s]; }
    n_exist[s] = alive ? K : 0;
  }
}

// (7) new landmarks from FeatureDEM (ids from the per-sequence counter, landmark.cpp:5-9; f2f_tracking.cpp:294-320 /
//     :418-441), then the inputs of depthInnovation: the left->right LK start positions (projection into cam1 of the
//     landmarks that already have depth, camera_frame.cpp:100-122) or the depth-image samples (:182-234)
__global__ void __launch_bounds__(TRK_THREADS) trk_append_kernel(TrkTable C, TrkBufs b, TrkCam cam, const float* newxy,
                                                                 const int* n_new, int mode_sel, const unsigned short* depth) {
  __shared__ double T1[7], R1[9];
  const int s = blockIdx.x, i = threadIdx.x, M = b.max_pts;
  if (b.ctl[s].mode != mode_sel || !b.ok[s]) return;
  const int n0 = C.n[s];
  int nn = n_new[s];
  if (n0 + nn > M) nn = M - n0;
  const bool as_inlier = mode_sel == 2 ? true : (b.orig_size[s] < 60);
  const double* T = C.T + 7 * s;
  if (i < nn) {
    const size_t src = (size_t)s * M + i, k = (size_t)s * M + n0 + i;
    const float px = newxy[2 * src], py = newxy[2 * src + 1];
    float ux = px, uy = py;
    if (cam.unrect) flv::undistort_point(cam.lens0, px, py, ux, uy);
    LmReg lm;
    lm.id = b.id_index[s] + i;
    lm.plane[0] = (double)px; lm.plane[1] = (double)py; lm.und[0] = (double)ux; lm.und[1] = (double)uy;
    lm.f2d[0] = lm.und[0]; lm.f2d[1] = lm.und[1];
    for (int c = 0; c < 3; ++c) { lm.p3w[c] = 0; lm.p3c[c] = 0; }
    for (int c = 0; c < 7; ++c) lm.fpose[c] = T[c];
    lm.has = 0; lm.inl = as_inlier ? 1 : 0;
    lm_store(C, k, lm);
  }
  if (i == 0) {
    if (cam.c.cam_type != 0) { se3_mul(cam.T_c1_c0, T, T1); q_to_R(T1, R1); }
  }
  __syncthreads();
  const int n = n0 + nn;
  if (i == 0) { C.n[s] = n; b.id_index[s] += nn; b.out[s].n_new = nn; b.out[s].id_index = b.id_index[s]; b.n_r[s] = n; }
  if (i >= n) return;
  const size_t k = (size_t)s * M + i;
  const double plx = C.plane[2 * k], ply = C.plane[2 * k + 1];
  if (cam.c.cam_type == 0) {
    const int px = (int)round(plx), py = (int)round(ply);
    b.dat[k] = (px >= 0 && px < cam.w && py >= 0 && py < cam.h) ? depth[(size_t)s * cam.w * cam.h + (size_t)py * cam.w + px] : 0;
    return;
  }
  const float fx0 = (float)plx, fy0 = (float)ply;
  float ix = fx0, iy = fy0;
  if (C.has[k]) {
    const float X = (float)C.p3w[3 * k], Y = (float)C.p3w[3 * k + 1], Z = (float)C.p3w[3 * k + 2];
    if (cam.unrect) {
      flv::project_point(cam.lens1, R1, T1 + 4, X, Y, Z, ix, iy);
    } else {
      const double pw[3] = {(double)X, (double)Y, (double)Z};
      double pc[3];
      world2camera(T1, pw, pc);
      ix = (float)(cam.cam1[0] * pc[0] / pc[2] + cam.cam1[2]);
      iy = (float)(cam.cam1[1] * pc[1] / pc[2] + cam.cam1[3]);
    }
  }
  b.r_prev[2 * k] = fx0; b.r_prev[2 * k + 1] = fy0;
  b.r_init[2 * k] = ix; b.r_init[2 * k + 1] = iy;
}

// (8) undistortion of the right-image points (camera_frame.cpp:130)
__global__ void __launch_bounds__(TRK_THREADS) trk_pt1_kernel(TrkBufs b, TrkCam cam) {
  const int s = blockIdx.x, i = threadIdx.x, M = b.max_pts;
  if (i >= b.n_r[s]) return;
  const size_t k = (size_t)s * M + i;
  const float x = b.r_next[2 * k], y = b.r_next[2 * k + 1];
  float ux = x, uy = y;
  if (cam.unrect) flv::undistort_point(cam.lens1, x, y, ux, uy);
  b.pt1[2 * k] = (double)ux; b.pt1[2 * k + 1] = (double)uy;
}

// (9) eraseNoDepthPoint (camera_frame.cpp:30-40), validLMCount (:399-413), the frame's summary, and the commit: the frame
//     becomes the stream's "last" frame iff it was accepted (tracking + BA succeeded, or the initialisation found > 30 valid
//     landmarks, f2f_tracking.cpp:443-452) -- the reference's swap-back of curr / last on failure
__global__ void __launch_bounds__(TRK_THREADS) trk_finish_kernel(TrkTable L, TrkTable C, TrkBufs b) {
  __shared__ int wsum[TRK_THREADS / 32];
  const int s = blockIdx.x, i = threadIdx.x, M = b.max_pts;
  const TrkCtl& ctl = b.ctl[s];
  const bool alive = b.ok[s] != 0 && ctl.mode != 0;
  const int n = alive ? C.n[s] : 0;
  const size_t k = (size_t)s * M + i;
  LmReg lm;
  bool keep = false;
  if (i < n) { lm_load(C, k, lm); keep = lm.has != 0; }
  int K, valid;
  const int r = block_rank(keep, wsum, K);
  block_rank(keep && lm.inl != 0, wsum, valid);
  bool commit = alive;
  if (ctl.mode == 2) commit = alive && (valid > 30 || ctl.commit_on_init_fail);
  if (keep && commit) lm_store(L, (size_t)s * M + r, lm);
  if (i == 0) {
    TrkOut& o = b.out[s];
    if (alive) {
      o.n_final = K; o.valid_cnt = valid; o.rand_used = b.n_rand_used[s];
      for (int c = 0; c < 7; ++c) o.T[c] = C.T[7 * s + c];
    } else if (ctl.mode == 1) {
      for (int c = 0; c < 7; ++c) o.T[c] = C.T[7 * s + c];        // pose reached before the failure (diagnostics)
    }
    o.committed = commit ? 1 : 0;
    if (commit) { L.n[s] = K; for (int c = 0; c < 7; ++c) L.T[7 * s + c] = C.T[7 * s + c]; }
  }
}

}  // namespace

// ---- launch sequence of one frame ------------------------------------------------------------------------------------------
#define TRK_LAUNCH(ctx, kernel, grid, block, ...)                 \
  do {                                                            \
    kernel<<<grid, block, 0, (ctx)->stream>>>(__VA_ARGS__);       \
    (ctx)->launches++;                                            \
    FLV_CUDA(ctx, cudaGetLastError());                            \
  } while (0)

int flv_trk_stage_prepare(flv_ctx* ctx, const TrkDev& d, int S) {
  TRK_LAUNCH(ctx, trk_prepare_kernel, S, TRK_THREADS, d.L, d.C, d.b, d.cam);
  return FLV_OK;
}
int flv_trk_stage_keep(flv_ctx* ctx, const TrkDev& d, int S) {
  TRK_LAUNCH(ctx, trk_keep_kernel, S, TRK_THREADS, d.L, d.C, d.b, d.cam);
  return FLV_OK;
}
int flv_trk_stage_after_f(flv_ctx* ctx, const TrkDev& d, int S) {
  TRK_LAUNCH(ctx, trk_after_f_kernel, S, TRK_THREADS, d.C, d.b);
  return FLV_OK;
}
int flv_trk_stage_after_pnp(flv_ctx* ctx, const TrkDev& d, int S) {
  TRK_LAUNCH(ctx, trk_after_pnp_kernel, S, TRK_THREADS, d.C, d.b, d.cam);
  return FLV_OK;
}
int flv_trk_stage_after_ba(flv_ctx* ctx, const TrkDev& d, int S) {
  TRK_LAUNCH(ctx, trk_after_ba_kernel, (S + 63) / 64, 64, d.C, d.b, S);
  return FLV_OK;
}
int flv_trk_stage_erase_outliers(flv_ctx* ctx, const TrkDev& d, int S) {
  TRK_LAUNCH(ctx, trk_erase_outliers_kernel, S, TRK_THREADS, d.C, d.b, ctx->d_exist, ctx->d_nexist);
  return FLV_OK;
}
int flv_trk_stage_append(flv_ctx* ctx, const TrkDev& d, int S, int mode_sel) {
  TRK_LAUNCH(ctx, trk_append_kernel, S, TRK_THREADS, d.C, d.b, d.cam, ctx->d_newxy, ctx->d_nnew, mode_sel, d.depth);
  return FLV_OK;
}
int flv_trk_stage_pt1(flv_ctx* ctx, const TrkDev& d, int S) {
  TRK_LAUNCH(ctx, trk_pt1_kernel, S, TRK_THREADS, d.b, d.cam);
  return FLV_OK;
}
int flv_trk_stage_finish(flv_ctx* ctx, const TrkDev& d, int S) {
  TRK_LAUNCH(ctx, trk_finish_kernel, S, TRK_THREADS, d.L, d.C, d.b);
  return FLV_OK;
}
