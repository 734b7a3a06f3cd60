// K6 -- per-landmark geometry of CameraFrame, batched: one CTA per stream, one thread per landmark.
//
// Reference: CameraFrame::depthInnovation (src/processing/camera_frame.cpp:271-330) with its three measurement
// sources recover3DPts_c_FromTriangulation (:236-270), recover3DPts_c_FromStereo (:133-179, the part after the LK
// call) and recover3DPts_c_FromDepthImg (:182-234); Triangulation::triangulationPt (src/processing/
// triangulation.cpp:9-39, :80-97); CameraFrame::calReprjInlierOutlier (:43-91); DepthCamera (depth_camera.cpp:92-150).
// Oracle: oracle/camera_frame_ref.py.  fp64 like the reference; the 4x4 JacobiSVD null vector is computed with a
// one-sided (Hestenes) Jacobi SVD in registers -- same vector up to rounding, tolerance 1e-9 relative in the tests.
// The reference's `rand()` dummy depths are supplied by the host (glibc sequence, flvis_b200/host) in consumption
// order; the kernel indexes them with an in-order prefix count so landmark i gets exactly the value the
// sequential loop would have drawn.
#include "ctx.h"

namespace {

constexpr int GEO_THREADS = 512;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void q_rot(const double* q, const double* v, double* o) {   // q = x,y,z,w
  double ux = q[1] * v[2] - q[2] * v[1], uy = q[2] * v[0] - q[0] * v[2], uz = q[0] * v[1] - q[1] * v[0];
  ux += ux; uy += uy; uz += uz;
  o[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
  o[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
  o[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}
__device__ __forceinline__ void q_to_R(const double* q, double* R) {
  double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
__device__ __forceinline__ void world2camera(const double* T, const double* pw, double* pc) {   // T = [q xyzw, t]
  q_rot(T, pw, pc);
  pc[0] += T[4]; pc[1] += T[5]; pc[2] += T[6];
}
__device__ __forceinline__ void camera2world(const double* T, const double* pc, double* pw) {   // T_c_w.inverse() * p
  const double n = sqrt(T[0] * T[0] + T[1] * T[1] + T[2] * T[2] + T[3] * T[3]);
  const double qi[4] = {-T[0] / n, -T[1] / n, -T[2] / n, T[3] / n};      // Sophus: conjugate, re-normalised
  const double mt[3] = {-T[4], -T[5], -T[6]};
  double ti[3];
  q_rot(qi, mt, ti);
  q_rot(qi, pc, pw);
  pw[0] += ti[0]; pw[1] += ti[1]; pw[2] += ti[2];
}
__device__ __forceinline__ void proj_matrix(const double* T, double fx, double fy, double cx, double cy, double* P) {
  double R[9];
  q_to_R(T, R);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    P[c] = fx * R[c] + cx * R[6 + c];
    P[4 + c] = fy * R[3 + c] + cy * R[6 + c];
    P[8 + c] = R[6 + c];
  }
  P[3] = fx * T[4] + cx * T[6]; P[7] = fy * T[5] + cy * T[6]; P[11] = T[6];
}

// DLT triangulation: null vector of the 4x4 matrix built from two projections (triangulation.cpp:9-39)
__device__ void triangulation_pt(const double* pt1, const double* pt2, const double* P1, const double* P2, double* X) {
  double A[16], V[16];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    A[c] = pt1[1] * P1[8 + c] - P1[4 + c];
    A[4 + c] = P1[c] - pt1[0] * P1[8 + c];
    A[8 + c] = pt2[1] * P2[8 + c] - P2[4 + c];
    A[12 + c] = P2[c] - pt2[0] * P2[8 + c];
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) V[i] = (i % 5 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0;
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int q = p + 1; q < 4; ++q) {
        double al = 0, be = 0, ga = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { al += A[4 * k + p] * A[4 * k + p]; be += A[4 * k + q] * A[4 * k + q]; ga += A[4 * k + p] * A[4 * k + q]; }
        const double lim = 1e-15 * sqrt(al * be);
        if (fabs(ga) > lim && ga != 0.0) {
          off = fmax(off, fabs(ga) / fmax(sqrt(al * be), 1e-300));
          const double zeta = (be - al) / (2.0 * ga);
          const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const double ap = A[4 * k + p], aq = A[4 * k + q];
            A[4 * k + p] = c * ap - s * aq; A[4 * k + q] = s * ap + c * aq;
            const double vp = V[4 * k + p], vq = V[4 * k + q];
            V[4 * k + p] = c * vp - s * vq; V[4 * k + q] = s * vp + c * vq;
          }
        }
      }
    if (off < 1e-15) break;
  }
  int best = 0;
  double bn = 1e300;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    double nn = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) nn += A[4 * k + c] * A[4 * k + c];
    if (nn < bn) { bn = nn; best = c; }
  }
  const double w = V[12 + best];
  X[0] = V[best] / w; X[1] = V[4 + best] / w; X[2] = V[8 + best] / w;
}

struct GeoArgs {
  const int* n_lms; flv_camera cam; flv_depth_params prm;
  const double* T_c_w; const double* plane; const double* undist; double* p3d_w; double* p3d_c; uint8_t* has_3d;
  const double* first_2d; const double* first_pose; const double* pt1_undist; const uint8_t* lk_status;
  const uint16_t* depth_at_pts; const float* dummy_rand; int* n_rand_used; int max_pts;
};

__global__ void __launch_bounds__(GEO_THREADS) depth_innovation_kernel(GeoArgs a) {
  __shared__ int wsum[GEO_THREADS / 32];
  const int s = blockIdx.x, i = threadIdx.x, lane = i & 31, warp = i >> 5;
  const int n = a.n_lms[s];
  const size_t k = (size_t)s * a.max_pts + i;
  const bool on = i < n;
  const double* T = a.T_c_w + 7 * (size_t)s;
  const flv_camera& cam = a.cam;
  const double range = (double)a.prm.range, iir = (double)a.prm.iir_ratio;
  // ---- measurement from the camera (stereo or depth image) -------------------------------------------
  double cam_pt[3] = {0, 0, 0};
  bool cam_ok = false;
  if (on) {
    if (cam.cam_type == 0) {
      const float px = (float)round(a.plane[2 * k]), py = (float)round(a.plane[2 * k + 1]);
      const float z = (float)a.depth_at_pts[k] / (float)cam.depth_scale;
      if ((double)z >= 0.3 && z <= a.prm.range) {
        cam_pt[2] = z; cam_pt[0] = ((double)px - cam.cx) * (double)z / cam.fx; cam_pt[1] = ((double)py - cam.cy) * (double)z / cam.fy;
        cam_ok = true;
      }
    } else if (a.lk_status[k] == 1) {
      // the reference hands cv::Point2f (float) copies of the undistorted points to the stereo triangulation
      const double u0[2] = {(double)(float)a.undist[2 * k], (double)(float)a.undist[2 * k + 1]};
      const double u1[2] = {(double)(float)a.pt1_undist[2 * k], (double)(float)a.pt1_undist[2 * k + 1]};
      triangulation_pt(u0, u1, cam.P0, cam.P1, cam_pt);
      cam_ok = !(cam_pt[2] < 0 || cam_pt[2] > range);
    }
  }
  // in-order index of this landmark's rand() draw (every landmark without a camera measurement draws one)
  const bool draws = on && !cam_ok;
  const unsigned bal = __ballot_sync(FULL, draws);
  if (lane == 0) wsum[warp] = __popc(bal);
  __syncthreads();
  int base = 0;
  for (int wv = 0; wv < warp; ++wv) base += wsum[wv];
  if (draws) {
    const double d = (double)a.dummy_rand[(size_t)s * a.max_pts + base + __popc(bal & ((1u << lane) - 1))];
    const double sx = cam.cam_type == 0 ? a.plane[2 * k] : (double)(float)a.undist[2 * k];
    const double sy = cam.cam_type == 0 ? a.plane[2 * k + 1] : (double)(float)a.undist[2 * k + 1];
    cam_pt[0] = (sx - cam.cx) * d / cam.fx; cam_pt[1] = (sy - cam.cy) * d / cam.fy; cam_pt[2] = d;
  }
  if (i == GEO_THREADS - 1) {
    int tot = 0;
    for (int wv = 0; wv < GEO_THREADS / 32; ++wv) tot += wsum[wv];
    a.n_rand_used[s] = tot;
  }
  if (!on) return;
  // ---- two-view triangulation against the first observation (camera_frame.cpp:236-270) ----------------
  double tri_pt[3] = {0, 0, 0};
  bool tri_ok = false;
  {
    const double* T1 = a.first_pose + 7 * k;
    const double bx = T1[4] - T[4], by = T1[5] - T[5], bz = T1[6] - T[6];
    if (sqrt(bx * bx + by * by + bz * bz) >= 0.2) {
      double P1[12], P2[12], pw[3];
      proj_matrix(T1, cam.fx, cam.fy, cam.cx, cam.cy, P1);
      proj_matrix(T, cam.fx, cam.fy, cam.cx, cam.cy, P2);
      triangulation_pt(a.first_2d + 2 * k, a.undist + 2 * k, P1, P2, pw);
      world2camera(T, pw, tri_pt);
      tri_ok = tri_pt[2] >= 0.5 && tri_pt[2] <= range;
      if (!tri_ok) { tri_pt[0] = tri_pt[1] = tri_pt[2] = 0; }
    }
  }
  // ---- fusion (camera_frame.cpp:285-329) ---------------------------------------------------------------
  const bool had = a.has_3d[k] != 0;
  double* pw = a.p3d_w + 3 * k;
  double* pc = a.p3d_c + 3 * k;
  if (!cam_ok && !tri_ok && cam.cam_type != 0) {
    if (!had && a.prm.dummy_depth) {
      pc[0] = cam_pt[0]; pc[1] = cam_pt[1]; pc[2] = cam_pt[2];
      camera2world(T, cam_pt, pw);
      a.has_3d[k] = 1;
    }
    return;
  }
  const double* meas = cam_ok ? cam_pt : tri_pt;
  if (cam.cam_type == 0 && !cam_ok && !tri_ok) meas = tri_pt;      // depth sensor, nothing valid: the reference falls through with (0,0,0)
  if (had) {
    double lm_c[3], upd[3];
    world2camera(T, pw, lm_c);
#pragma unroll
    for (int c = 0; c < 3; ++c) upd[c] = lm_c[c] * iir + meas[c] * (1 - iir);
    pc[0] = upd[0]; pc[1] = upd[1]; pc[2] = upd[2];
    camera2world(T, upd, pw);
  } else {
    pc[0] = meas[0]; pc[1] = meas[1]; pc[2] = meas[2];
    camera2world(T, meas, pw);
    a.has_3d[k] = 1;
  }
}

// calReprjInlierOutlier (camera_frame.cpp:43-91): distance of every landmark, median of those < 3 px, threshold
__global__ void __launch_bounds__(GEO_THREADS)
reprj_inlier_kernel(const int* __restrict__ n_lms, flv_camera cam, const double* __restrict__ T_c_w,
                    const double* __restrict__ undist, const double* __restrict__ p3d_w, double sh_over_med,
                    uint8_t* __restrict__ inlier, double* __restrict__ mean_err, int max_pts) {
  __shared__ double dist[GEO_THREADS];
  __shared__ double red[GEO_THREADS / 32];
  __shared__ int cnt[GEO_THREADS / 32];
  __shared__ double sh_med;
  const int s = blockIdx.x, i = threadIdx.x, lane = i & 31, warp = i >> 5;
  const int n = n_lms[s];
  const size_t k = (size_t)s * max_pts + i;
  double d = 1e300;
  if (i < n) {
    double pc[3];
    world2camera(T_c_w + 7 * (size_t)s, p3d_w + 3 * k, pc);
    const double ex = undist[2 * k] - (cam.fx * pc[0] / pc[2] + cam.cx), ey = undist[2 * k + 1] - (cam.fy * pc[1] / pc[2] + cam.cy);
    d = sqrt(ex * ex + ey * ey);
  }
  dist[i] = d;
  const bool valid = i < n && d < 3.0;
  // mean of the valid distances (summed in landmark order per warp, then across warps)
  double v = valid ? d : 0.0;
  int c = valid ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(FULL, v, o); c += __shfl_xor_sync(FULL, c, o); }
  if (lane == 0) { red[warp] = v; cnt[warp] = c; }
  if (i == 0) sh_med = 3.0;
  __syncthreads();
  double sum = 0; int nv = 0;
  for (int wv = 0; wv < GEO_THREADS / 32; ++wv) { sum += red[wv]; nv += cnt[wv]; }
  // sorted(valid).at(floor(nv/2)): the valid element whose rank (ties broken by index) is nv/2
  if (valid) {
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const double dj = dist[j];
      if (dj < 3.0 && (dj < d || (dj == d && j < i))) ++rank;
    }
    if (rank == nv / 2) sh_med = d;
  }
  __syncthreads();
  double sh = sh_over_med * sh_med;
  if (sh >= 3.0) sh = 3.0;
  if (i < n) inlier[k] = d > sh ? 0 : 1;
  if (i == 0) mean_err[s] = nv ? sum / (double)nv : 0.0 / 0.0;
}

}  // namespace

extern "C" {

int flv_depth_innovation(flv_ctx* ctx, int n_streams, const int* n_lms, const flv_camera* cam,
                         const flv_depth_params* prm, const double* T_c_w, const double* lm_2d_plane,
                         const double* lm_2d_undist, double* lm_3d_w, double* lm_3d_c, uint8_t* has_3d,
                         const double* first_obs_2d, const double* first_obs_pose, const double* stereo_pt1_undist,
                         const uint8_t* stereo_status, const uint16_t* depth_at_pts, const float* dummy_rand,
                         int* n_rand_used, flv_memspace mem) {
  if (!ctx || !n_lms || !cam || !prm || !T_c_w || !lm_2d_plane || !lm_2d_undist || !lm_3d_w || !lm_3d_c || !has_3d ||
      !first_obs_2d || !first_obs_pose || !dummy_rand || !n_rand_used || n_streams < 1 || n_streams > ctx->S)
    return FLV_ERR_INVALID;
  if (cam->cam_type == 0 ? !depth_at_pts : (!stereo_pt1_undist || !stereo_status)) return FLV_ERR_INVALID;
  if (ctx->max_pts > GEO_THREADS) FLV_FAIL(ctx, FLV_ERR_UNSUPPORTED, "max_pts %d > %d", ctx->max_pts, GEO_THREADS);
  const size_t S = n_streams, M = ctx->max_pts, np = S * M;
  GeoArgs a;
  a.cam = *cam; a.prm = *prm; a.max_pts = ctx->max_pts;
  if (mem == FLV_MEM_DEVICE) {
    a.n_lms = n_lms; a.T_c_w = T_c_w; a.plane = lm_2d_plane; a.undist = lm_2d_undist; a.p3d_w = lm_3d_w; a.p3d_c = lm_3d_c;
    a.has_3d = has_3d; a.first_2d = first_obs_2d; a.first_pose = first_obs_pose; a.pt1_undist = stereo_pt1_undist;
    a.lk_status = stereo_status; a.depth_at_pts = depth_at_pts; a.dummy_rand = dummy_rand; a.n_rand_used = n_rand_used;
    depth_innovation_kernel<<<n_streams, GEO_THREADS, 0, ctx->stream>>>(a);
    ctx->launches++;
    FLV_CUDA(ctx, cudaGetLastError());
    return FLV_OK;
  }
  // host arrays: pack into the staging block, run, copy the three outputs back
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = (o + bytes + 255) & ~(size_t)255; return r; };
  const size_t o_n = take(S * 4), o_T = take(S * 56), o_pl = take(np * 16), o_un = take(np * 16), o_w = take(np * 24),
               o_c = take(np * 24), o_h = take(np), o_f2 = take(np * 16), o_fp = take(np * 56), o_p1 = take(np * 16),
               o_st = take(np), o_d = take(np * 2), o_r = take(np * 4), o_u = take(S * 4);
  int rc = flv_stage_reserve(ctx, o);
  if (rc) return rc;
  char* hs = (char*)ctx->h_stage; char* ds = (char*)ctx->d_stage;
  memcpy(hs + o_n, n_lms, S * 4); memcpy(hs + o_T, T_c_w, S * 56); memcpy(hs + o_pl, lm_2d_plane, np * 16);
  memcpy(hs + o_un, lm_2d_undist, np * 16); memcpy(hs + o_w, lm_3d_w, np * 24); memcpy(hs + o_c, lm_3d_c, np * 24);
  memcpy(hs + o_h, has_3d, np); memcpy(hs + o_f2, first_obs_2d, np * 16); memcpy(hs + o_fp, first_obs_pose, np * 56);
  if (stereo_pt1_undist) memcpy(hs + o_p1, stereo_pt1_undist, np * 16);
  if (stereo_status) memcpy(hs + o_st, stereo_status, np);
  if (depth_at_pts) memcpy(hs + o_d, depth_at_pts, np * 2);
  memcpy(hs + o_r, dummy_rand, np * 4);
  FLV_CUDA(ctx, cudaMemcpyAsync(ds, hs, o, cudaMemcpyHostToDevice, ctx->stream));
  a.n_lms = (const int*)(ds + o_n); a.T_c_w = (const double*)(ds + o_T); a.plane = (const double*)(ds + o_pl);
  a.undist = (const double*)(ds + o_un); a.p3d_w = (double*)(ds + o_w); a.p3d_c = (double*)(ds + o_c);
  a.has_3d = (uint8_t*)(ds + o_h); a.first_2d = (const double*)(ds + o_f2); a.first_pose = (const double*)(ds + o_fp);
  a.pt1_undist = (const double*)(ds + o_p1); a.lk_status = (const uint8_t*)(ds + o_st);
  a.depth_at_pts = (const uint16_t*)(ds + o_d); a.dummy_rand = (const float*)(ds + o_r); a.n_rand_used = (int*)(ds + o_u);
  depth_innovation_kernel<<<n_streams, GEO_THREADS, 0, ctx->stream>>>(a);
  ctx->launches++;
  FLV_CUDA(ctx, cudaGetLastError());
  FLV_CUDA(ctx, cudaMemcpyAsync(hs + o_w, ds + o_w, o_f2 - o_w, cudaMemcpyDeviceToHost, ctx->stream));
  FLV_CUDA(ctx, cudaMemcpyAsync(hs + o_u, ds + o_u, S * 4, cudaMemcpyDeviceToHost, ctx->stream));
  FLV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(lm_3d_w, hs + o_w, np * 24); memcpy(lm_3d_c, hs + o_c, np * 24); memcpy(has_3d, hs + o_h, np);
  memcpy(n_rand_used, hs + o_u, S * 4);
  return FLV_OK;
}

int flv_reprojection_inliers(flv_ctx* ctx, int n_streams, const int* n_lms, const flv_camera* cam, const double* T_c_w,
                             const double* lm_2d_undist, const double* lm_3d_w, double sh_over_med,
                             uint8_t* is_inlier, double* mean_prjerr, flv_memspace mem) {
  if (!ctx || !n_lms || !cam || !T_c_w || !lm_2d_undist || !lm_3d_w || !is_inlier || !mean_prjerr || n_streams < 1 ||
      n_streams > ctx->S)
    return FLV_ERR_INVALID;
  if (ctx->max_pts > GEO_THREADS) FLV_FAIL(ctx, FLV_ERR_UNSUPPORTED, "max_pts %d > %d", ctx->max_pts, GEO_THREADS);
  const size_t S = n_streams, M = ctx->max_pts, np = S * M;
  if (mem == FLV_MEM_DEVICE) {
    reprj_inlier_kernel<<<n_streams, GEO_THREADS, 0, ctx->stream>>>(n_lms, *cam, T_c_w, lm_2d_undist, lm_3d_w, sh_over_med,
                                                                    is_inlier, mean_prjerr, ctx->max_pts);
    ctx->launches++;
    FLV_CUDA(ctx, cudaGetLastError());
    return FLV_OK;
  }
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = (o + bytes + 255) & ~(size_t)255; return r; };
  const size_t o_n = take(S * 4), o_T = take(S * 56), o_un = take(np * 16), o_w = take(np * 24), o_i = take(np), o_m = take(S * 8);
  int rc = flv_stage_reserve(ctx, o);
  if (rc) return rc;
  char* hs = (char*)ctx->h_stage; char* ds = (char*)ctx->d_stage;
  memcpy(hs + o_n, n_lms, S * 4); memcpy(hs + o_T, T_c_w, S * 56); memcpy(hs + o_un, lm_2d_undist, np * 16);
  memcpy(hs + o_w, lm_3d_w, np * 24);
  FLV_CUDA(ctx, cudaMemcpyAsync(ds, hs, o_i, cudaMemcpyHostToDevice, ctx->stream));
  reprj_inlier_kernel<<<n_streams, GEO_THREADS, 0, ctx->stream>>>((const int*)(ds + o_n), *cam, (const double*)(ds + o_T),
                                                                  (const double*)(ds + o_un), (const double*)(ds + o_w),
                                                                  sh_over_med, (uint8_t*)(ds + o_i), (double*)(ds + o_m),
                                                                  ctx->max_pts);
  ctx->launches++;
  FLV_CUDA(ctx, cudaGetLastError());
  FLV_CUDA(ctx, cudaMemcpyAsync(hs + o_i, ds + o_i, o - o_i, cudaMemcpyDeviceToHost, ctx->stream));
  FLV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(is_inlier, hs + o_i, np); memcpy(mean_prjerr, hs + o_m, S * 8);
  return FLV_OK;
}

}  // extern "C"
