// K11 -- batched RANSAC for the two geometric verification calls of LKORBTracking::tracking:
//   cv::findFundamentalMat(from, to, FM_RANSAC, 5.0, 0.99)              src/processing/lkorb_tracking.cpp:134-135
//   cv::solvePnPRansac(p3d, p2d, K, D, r, t, guess, 100, 3.0, 0.99, ...) src/processing/lkorb_tracking.cpp:170-177
// OpenCV's results depend on its private RNG stream and on LAPACK-backed solvers, so they cannot be reproduced bit for
// bit; this kernel is a GPU-shaped RANSAC with the same models, thresholds and error measures, checked against cv2
// statistically (tests/test_ransac_gpu.py: inlier-set overlap, epipolar / reprojection residuals, pose distance):
//   * all hypotheses of a stream are generated and scored in parallel by one CTA (no adaptive early exit: a fixed
//     hypothesis budget costs less than the control flow), counter-based sample streams => run-to-run deterministic;
//   * F: normalised 8-point on 8 samples (9x9 Gram matrix, cyclic Jacobi), rank-2 projection (3x3 Jacobi SVD),
//     error = max of the two squared point-to-epipolar-line distances (OpenCV's FMEstimatorCallback::computeError);
//     F is refit on the best sample model's inliers (least squares) and, if the refit explains at least as many
//     points, F and the mask are those of the refit (one local-optimisation step), else those of the sample model;
//   * PnP: 5-point samples refined from the initial pose by Gauss-Newton on the reprojection error (the tracker always
//     has a pose prior: the IMU prediction or the previous frame), error = squared reprojection distance; the best
//     model is refined on all its inliers (block-parallel normal equations) and re-scored.
// One CTA per stream; fp64 throughout.
#include "ctx.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int F_HYP = 256;          // hypotheses per stream (8-point samples)
constexpr int P_HYP = 128;          // hypotheses per stream (5-point samples)
constexpr int P_SAMPLE = 5;

struct Lcg {
  unsigned long long s;
  __device__ explicit Lcg(unsigned long long seed) : s(seed) {}
  __device__ unsigned next() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (unsigned)(s >> 33); }
  __device__ int below(int n) { return (int)(next() % (unsigned)n); }
};

__device__ void sample_distinct(Lcg& rng, int n, int k, int* out) {
  for (int i = 0; i < k; ++i) {
    for (;;) {
      const int v = rng.below(n);
      bool dup = false;
      for (int j = 0; j < i; ++j) dup |= (out[j] == v);
      if (!dup) { out[i] = v; break; }
    }
  }
}

// cyclic Jacobi eigen-decomposition of a symmetric N x N matrix (row-major, destroyed); V columns = eigenvectors
template <int N>
__device__ void jacobi_eigen(double* A, double* V) {
  for (int i = 0; i < N * N; ++i) V[i] = 0;
  for (int i = 0; i < N; ++i) V[i * N + i] = 1;
  for (int sweep = 0; sweep < 40; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < N; ++i) {
      diag += A[i * N + i] * A[i * N + i];
      for (int j = i + 1; j < N; ++j) off += A[i * N + j] * A[i * N + j];
    }
    if (off <= 1e-30 * (diag + 1e-300)) break;
    for (int p = 0; p < N - 1; ++p)
      for (int q = p + 1; q < N; ++q) {
        const double apq = A[p * N + q];
        if (fabs(apq) < 1e-300) continue;
        const double theta = (A[q * N + q] - A[p * N + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < N; ++k) {
          const double akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = c * akp - s * akq; A[k * N + q] = s * akp + c * akq;
        }
        for (int k = 0; k < N; ++k) {
          const double apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = c * apk - s * aqk; A[q * N + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < N; ++k) {
          const double vkp = V[k * N + p], vkq = V[k * N + q];
          V[k * N + p] = c * vkp - s * vkq; V[k * N + q] = s * vkp + c * vkq;
        }
      }
  }
}

__device__ void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

// project a 3x3 matrix onto rank 2 (zero the smallest singular value): M <- U diag(s0, s1, 0) V^T
__device__ void rank2(double* M) {
  double G[9], V[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) G[3 * i + j] = M[i] * M[j] + M[3 + i] * M[3 + j] + M[6 + i] * M[6 + j];
  jacobi_eigen<3>(G, V);
  int lo = 0;
  for (int i = 1; i < 3; ++i) if (G[4 * i] < G[4 * lo]) lo = i;
  // M <- M (I - v v^T), v = right singular vector of the smallest singular value
  const double v[3] = {V[lo], V[3 + lo], V[6 + lo]};
  for (int r = 0; r < 3; ++r) {
    const double d = M[3 * r] * v[0] + M[3 * r + 1] * v[1] + M[3 * r + 2] * v[2];
    for (int c = 0; c < 3; ++c) M[3 * r + c] -= d * v[c];
  }
}

// Hartley normalisation + Gram matrix of the epipolar constraint rows over the points selected by `sel`
// (sel(i) -> bool, evaluated for i in [i0, i1) with stride); returns partial sums for a later reduction
struct NormT { double cax, cay, cbx, cby, sa, sb; };

// null vector of an 8 x 9 matrix (row-major, destroyed) by Gaussian elimination with complete pivoting: the minimal
// 8-point sample determines F up to scale exactly, so no eigen-decomposition is needed (a 9 x 9 Jacobi costs ~50x more
// serial fp64 work per hypothesis thread).  Returns false for a rank-deficient sample.
__device__ bool null_vector_8x9(double* A, double* x) {
  // every index below is a compile-time constant after unrolling (pivot row / column swaps are predicated), so the
  // matrix never needs dynamically indexed local memory
  int swp[8];
  double amax0 = 0;
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int bi = k, bj = k;
    double best = -1;
#pragma unroll
    for (int i = k; i < 8; ++i)
#pragma unroll
      for (int j = k; j < 9; ++j) { const double v = fabs(A[9 * i + j]); if (v > best) { best = v; bi = i; bj = j; } }
    if (k == 0) amax0 = best;
    if (!(best > 1e-12 * amax0) || best == 0) ok = false;
    swp[k] = bj;
#pragma unroll
    for (int i = k + 1; i < 8; ++i)
      if (i == bi) {
#pragma unroll
        for (int j = 0; j < 9; ++j) { const double t = A[9 * k + j]; A[9 * k + j] = A[9 * i + j]; A[9 * i + j] = t; }
      }
#pragma unroll
    for (int j = k + 1; j < 9; ++j)
      if (j == bj) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const double t = A[9 * i + k]; A[9 * i + k] = A[9 * i + j]; A[9 * i + j] = t; }
      }
    const double piv = A[9 * k + k];
    const double ip = piv != 0 ? 1.0 / piv : 0.0;
#pragma unroll
    for (int i = k + 1; i < 8; ++i) {
      const double f = A[9 * i + k] * ip;
#pragma unroll
      for (int j = k; j < 9; ++j) A[9 * i + j] -= f * A[9 * k + j];
    }
  }
  if (!ok) return false;
  double y[9];
  y[8] = 1.0;
#pragma unroll
  for (int k = 7; k >= 0; --k) {
    double sum = 0;
#pragma unroll
    for (int j = k + 1; j < 9; ++j) sum += A[9 * k + j] * y[j];
    y[k] = -sum / A[9 * k + k];
  }
  // undo the column swaps, last first
#pragma unroll
  for (int k = 7; k >= 0; --k)
#pragma unroll
    for (int j = k + 1; j < 9; ++j)
      if (j == swp[k]) { const double t = y[k]; y[k] = y[j]; y[j] = t; }
  double n2 = 0;
#pragma unroll
  for (int j = 0; j < 9; ++j) n2 += y[j] * y[j];
  const double in = 1.0 / sqrt(n2);
#pragma unroll
  for (int j = 0; j < 9; ++j) x[j] = y[j] * in;
  return true;
}

// eigenvector of the smallest eigenvalue of a symmetric positive semi-definite 9 x 9 matrix (row-major, destroyed) by
// inverse iteration on an LDL^T factorisation: the least-squares F of an over-determined inlier set
__device__ void smallest_eigvec_9(double* G, double* x) {
  double tr = 0;
  for (int i = 0; i < 9; ++i) tr += G[10 * i];
  const double eps = 1e-13 * tr + 1e-300;
  for (int i = 0; i < 9; ++i) G[10 * i] += eps;
  // in-place LDL^T: L below the diagonal (unit), D on it
  for (int j = 0; j < 9; ++j) {
    double d = G[10 * j];
    for (int k = 0; k < j; ++k) d -= G[9 * j + k] * G[9 * j + k] * G[10 * k];
    if (!(d > 1e-300)) d = 1e-300;
    G[10 * j] = d;
    for (int i = j + 1; i < 9; ++i) {
      double v = G[9 * i + j];
      for (int k = 0; k < j; ++k) v -= G[9 * i + k] * G[9 * j + k] * G[10 * k];
      G[9 * i + j] = v / d;
    }
  }
  for (int i = 0; i < 9; ++i) x[i] = 1.0 / 3.0 + 0.01 * i;      // generic start: not orthogonal to anything in particular
  for (int it = 0; it < 10; ++it) {
    for (int i = 0; i < 9; ++i) { double v = x[i]; for (int k = 0; k < i; ++k) v -= G[9 * i + k] * x[k]; x[i] = v; }
    for (int i = 0; i < 9; ++i) x[i] /= G[10 * i];
    for (int i = 8; i >= 0; --i) { double v = x[i]; for (int k = i + 1; k < 9; ++k) v -= G[9 * k + i] * x[k]; x[i] = v; }
    double n2 = 0;
    for (int i = 0; i < 9; ++i) n2 += x[i] * x[i];
    const double in = 1.0 / sqrt(n2);
    for (int i = 0; i < 9; ++i) x[i] *= in;
  }
}

__device__ void f_from_vec(double* Fn /*unit 9-vector, destroyed*/, const NormT& T, double* F) {
  rank2(Fn);
  const double Ta[9] = {T.sa, 0, -T.sa * T.cax, 0, T.sa, -T.sa * T.cay, 0, 0, 1};
  const double TbT[9] = {T.sb, 0, 0, 0, T.sb, 0, -T.sb * T.cbx, -T.sb * T.cby, 1};
  double tmp[9];
  mat3_mul(Fn, Ta, tmp);
  mat3_mul(TbT, tmp, F);
}

// inlier test of the symmetric epipolar distance, max(e1, e2) <= thr2, without the two fp64 divisions of f_error:
// d^2 / (la^2 + lb^2) <= thr2  <=>  d^2 <= thr2 (la^2 + lb^2).  The scoring loops run this 256 x n times per sequence.
__device__ __forceinline__ bool f_inlier(const double* F, double x1, double y1, double x2, double y2, double thr2) {
  double la = F[0] * x1 + F[1] * y1 + F[2], lb = F[3] * x1 + F[4] * y1 + F[5], lc = F[6] * x1 + F[7] * y1 + F[8];
  const double d2 = x2 * la + y2 * lb + lc;
  const bool ok2 = d2 * d2 <= thr2 * (la * la + lb * lb + 1e-300);
  la = F[0] * x2 + F[3] * y2 + F[6]; lb = F[1] * x2 + F[4] * y2 + F[7]; lc = F[2] * x2 + F[5] * y2 + F[8];
  const double d1 = x1 * la + y1 * lb + lc;
  return ok2 && d1 * d1 <= thr2 * (la * la + lb * lb + 1e-300);
}

__global__ void __launch_bounds__(RS_THREADS) fmat_ransac_kernel(const int* __restrict__ npts, const float* __restrict__ from_xy,
                                                                 const float* __restrict__ to_xy, int max_pts, double thr2,
                                                                 uint8_t* __restrict__ mask_out, double* __restrict__ F_out,
                                                                 int* __restrict__ n_inl) {
  __shared__ double sF[F_HYP][9];
  __shared__ int score[F_HYP];
  __shared__ double sG[81];
  __shared__ double red[RS_WARPS][6];
  __shared__ int s_best;
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = npts[s];
  const float* A = from_xy + (size_t)s * max_pts * 2;
  const float* B = to_xy + (size_t)s * max_pts * 2;
  uint8_t* mask = mask_out + (size_t)s * max_pts;
  if (n < 8) {
    for (int i = tid; i < max_pts; i += RS_THREADS) mask[i] = 0;
    if (tid == 0) { n_inl[s] = 0; for (int k = 0; k < 9; ++k) F_out[9 * s + k] = 0; }
    return;
  }
  // ---- hypotheses: thread h <-> sample h ----------------------------------------------------------------------------
  for (int h = tid; h < F_HYP; h += RS_THREADS) {
    Lcg rng(0x9E3779B97F4A7C15ULL + 0xD1B54A32D192ED03ULL * (unsigned long long)(h + 1));
    int idx[8];
    sample_distinct(rng, n, 8, idx);
    NormT T;
    T.cax = T.cay = T.cbx = T.cby = 0;
    for (int i = 0; i < 8; ++i) { T.cax += A[2 * idx[i]]; T.cay += A[2 * idx[i] + 1]; T.cbx += B[2 * idx[i]]; T.cby += B[2 * idx[i] + 1]; }
    T.cax /= 8; T.cay /= 8; T.cbx /= 8; T.cby /= 8;
    double da = 0, db = 0;
    for (int i = 0; i < 8; ++i) {
      const double ax = A[2 * idx[i]] - T.cax, ay = A[2 * idx[i] + 1] - T.cay, bx = B[2 * idx[i]] - T.cbx, by = B[2 * idx[i] + 1] - T.cby;
      da += sqrt(ax * ax + ay * ay); db += sqrt(bx * bx + by * by);
    }
    double Fh[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (da > 1e-9 && db > 1e-9) {
      T.sa = sqrt(2.0) * 8 / da; T.sb = sqrt(2.0) * 8 / db;
      double R8[72], Fn[9];
      for (int i = 0; i < 8; ++i) {
        const double x1 = (A[2 * idx[i]] - T.cax) * T.sa, y1 = (A[2 * idx[i] + 1] - T.cay) * T.sa;
        const double x2 = (B[2 * idx[i]] - T.cbx) * T.sb, y2 = (B[2 * idx[i] + 1] - T.cby) * T.sb;
        double* r = R8 + 9 * i;
        r[0] = x2 * x1; r[1] = x2 * y1; r[2] = x2; r[3] = y2 * x1; r[4] = y2 * y1; r[5] = y2; r[6] = x1; r[7] = y1; r[8] = 1;
      }
      if (null_vector_8x9(R8, Fn)) f_from_vec(Fn, T, Fh);
    }
    for (int k = 0; k < 9; ++k) sF[h][k] = Fh[k];
  }
  __syncthreads();
  // ---- scoring: warp per hypothesis, lanes over points ------------------------------------------------------------------
  for (int h = warp; h < F_HYP; h += RS_WARPS) {
    double Fh[9];
    for (int k = 0; k < 9; ++k) Fh[k] = sF[h][k];
    int c = 0;
    for (int i = lane; i < n; i += 32) c += f_inlier(Fh, A[2 * i], A[2 * i + 1], B[2 * i], B[2 * i + 1], thr2);
    c = __reduce_add_sync(FULL, c);
    if (lane == 0) score[h] = c;
  }
  __syncthreads();
  if (tid == 0) {
    int best = 0;
    for (int h = 1; h < F_HYP; ++h) if (score[h] > score[best]) best = h;      // first maximum: deterministic
    s_best = best;
  }
  __syncthreads();
  const int best = s_best;
  double Fb[9];
  for (int k = 0; k < 9; ++k) Fb[k] = sF[best][k];
  const int cnt = score[best];
  for (int i = tid; i < max_pts; i += RS_THREADS)
    mask[i] = (i < n && f_inlier(Fb, A[2 * i], A[2 * i + 1], B[2 * i], B[2 * i + 1], thr2)) ? 1 : 0;
  if (tid == 0) n_inl[s] = cnt >= 8 ? cnt : 0;
  if (cnt < 8) {
    __syncthreads();
    for (int i = tid; i < max_pts; i += RS_THREADS) mask[i] = 0;
    if (tid == 0) for (int k = 0; k < 9; ++k) F_out[9 * s + k] = 0;
    return;
  }
  // ---- least-squares refit on the inliers (block-parallel normalisation + Gram matrix) ---------------------------------
  double acc[6] = {0, 0, 0, 0, 0, 0};     // sums of ax, ay, bx, by
  for (int i = tid; i < n; i += RS_THREADS)
    if (f_inlier(Fb, A[2 * i], A[2 * i + 1], B[2 * i], B[2 * i + 1], thr2)) { acc[0] += A[2 * i]; acc[1] += A[2 * i + 1]; acc[2] += B[2 * i]; acc[3] += B[2 * i + 1]; }
  for (int k = 0; k < 4; ++k) { double v = acc[k]; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o); if (lane == 0) red[warp][k] = v; }
  __syncthreads();
  NormT T;
  { double t[4] = {0, 0, 0, 0}; for (int w = 0; w < RS_WARPS; ++w) for (int k = 0; k < 4; ++k) t[k] += red[w][k];
    T.cax = t[0] / cnt; T.cay = t[1] / cnt; T.cbx = t[2] / cnt; T.cby = t[3] / cnt; }
  __syncthreads();
  acc[0] = acc[1] = 0;
  for (int i = tid; i < n; i += RS_THREADS)
    if (f_inlier(Fb, A[2 * i], A[2 * i + 1], B[2 * i], B[2 * i + 1], thr2)) {
      const double ax = A[2 * i] - T.cax, ay = A[2 * i + 1] - T.cay, bx = B[2 * i] - T.cbx, by = B[2 * i + 1] - T.cby;
      acc[0] += sqrt(ax * ax + ay * ay); acc[1] += sqrt(bx * bx + by * by);
    }
  for (int k = 0; k < 2; ++k) { double v = acc[k]; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o); if (lane == 0) red[warp][k] = v; }
  __syncthreads();
  { double da = 0, db = 0; for (int w = 0; w < RS_WARPS; ++w) { da += red[w][0]; db += red[w][1]; }
    T.sa = sqrt(2.0) * cnt / fmax(da, 1e-300); T.sb = sqrt(2.0) * cnt / fmax(db, 1e-300); }
  for (int k = tid; k < 81; k += RS_THREADS) sG[k] = 0;
  __syncthreads();
  // Gram matrix: thread (a, b) sums its entry over the inliers in index order (deterministic, 81 threads busy)
  if (tid < 81) {
    const int a = tid / 9, b = tid - 9 * a;
    double g = 0;
    for (int i = 0; i < n; ++i) {
      if (!mask[i]) continue;
      const double x1 = (A[2 * i] - T.cax) * T.sa, y1 = (A[2 * i + 1] - T.cay) * T.sa;
      const double x2 = (B[2 * i] - T.cbx) * T.sb, y2 = (B[2 * i + 1] - T.cby) * T.sb;
      const double r[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1};
      g += r[a] * r[b];
    }
    sG[tid] = g;
  }
  __syncthreads();
  if (tid == 0) {
    double G[81], Fn[9], Fr[9];
    for (int k = 0; k < 81; ++k) G[k] = sG[k];
    smallest_eigvec_9(G, Fn);
    f_from_vec(Fn, T, Fr);
    for (int k = 0; k < 9; ++k) sF[0][k] = Fr[k];
  }
  __syncthreads();
  // local optimisation step: keep the refit model (and its inlier set) if it explains at least as many points
  double Fr[9];
  for (int k = 0; k < 9; ++k) Fr[k] = sF[0][k];
  int c2 = 0;
  for (int i = tid; i < n; i += RS_THREADS) c2 += f_inlier(Fr, A[2 * i], A[2 * i + 1], B[2 * i], B[2 * i + 1], thr2);
  c2 = __reduce_add_sync(FULL, c2);
  if (lane == 0) score[warp] = c2;
  __syncthreads();
  int tot = 0;
  for (int w = 0; w < RS_WARPS; ++w) tot += score[w];
  const bool take = tot >= cnt;
  if (take)
    for (int i = tid; i < n; i += RS_THREADS) mask[i] = f_inlier(Fr, A[2 * i], A[2 * i + 1], B[2 * i], B[2 * i + 1], thr2) ? 1 : 0;
  if (tid == 0) {
    n_inl[s] = take ? tot : cnt;
    for (int k = 0; k < 9; ++k) F_out[9 * s + k] = take ? Fr[k] : Fb[k];
  }
}

// ---- PnP ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void q_rot(const double* q, const double* v, double* o) {
  double ux = q[1] * v[2] - q[2] * v[1], uy = q[2] * v[0] - q[0] * v[2], uz = q[0] * v[1] - q[1] * v[0];
  ux += ux; uy += uy; uz += uz;
  o[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
  o[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
  o[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}

__device__ void R_to_q(const double* m, double* q) {
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
    double qq[3];
    qq[i] = 0.5 * t; t = 0.5 / t;
    q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
    qq[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    qq[k] = (m[3 * k + i] + m[3 * i + k]) * t;
    q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2];
  }
}

__device__ void se3_oplus(double* pose, const double* u) {      // pose <- exp(u) * pose, u = [omega, upsilon] (g2o SE3Quat)
  const double wx = u[0], wy = u[1], wz = u[2];
  const double theta = sqrt(wx * wx + wy * wy + wz * wz);
  const double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
  double O2[9];
  mat3_mul(O, O, O2);
  double a, b, d;
  if (theta < 0.00001) { a = 1.0; b = 0.5; d = 1.0 / 6.0; }
  else { a = sin(theta) / theta; b = (1 - cos(theta)) / (theta * theta); d = (theta - sin(theta)) / (theta * theta * theta); }
  double R[9], V[9];
  for (int i = 0; i < 9; ++i) {
    const double id = (i % 4 == 0) ? 1.0 : 0.0;
    R[i] = id + a * O[i] + b * O2[i];
    V[i] = id + b * O[i] + d * O2[i];
  }
  double qe[4], rt[3];
  R_to_q(R, qe);
  const double te[3] = {V[0] * u[3] + V[1] * u[4] + V[2] * u[5], V[3] * u[3] + V[4] * u[4] + V[5] * u[5], V[6] * u[3] + V[7] * u[4] + V[8] * u[5]};
  q_rot(qe, pose + 4, rt);
  const double* p = pose;
  double x = qe[3] * p[0] + qe[0] * p[3] + qe[1] * p[2] - qe[2] * p[1];
  double y = qe[3] * p[1] + qe[1] * p[3] + qe[2] * p[0] - qe[0] * p[2];
  double z = qe[3] * p[2] + qe[2] * p[3] + qe[0] * p[1] - qe[1] * p[0];
  double w = qe[3] * p[3] - qe[0] * p[0] - qe[1] * p[1] - qe[2] * p[2];
  if (w < 0) { x = -x; y = -y; z = -z; w = -w; }
  const double nn = sqrt(x * x + y * y + z * z + w * w);
  pose[0] = x / nn; pose[1] = y / nn; pose[2] = z / nn; pose[3] = w / nn;
  pose[4] = te[0] + rt[0]; pose[5] = te[1] + rt[1]; pose[6] = te[2] + rt[2];
}

// accumulate the normal equations of one observation into H (21 upper-triangular entries) and g (6)
__device__ __forceinline__ bool pnp_accumulate(const double* T, const double* K, const float* X3, const float* uv, double* H21, double* g) {
  const double X[3] = {X3[0], X3[1], X3[2]};
  double Xc[3];
  q_rot(T, X, Xc);
  const double x = Xc[0] + T[4], y = Xc[1] + T[5], z = Xc[2] + T[6];
  if (z < 1e-6) return false;
  const double iz = 1.0 / z, xz = x * iz, yz = y * iz;
  const double r0 = uv[0] - (xz * K[0] + K[2]), r1 = uv[1] - (yz * K[1] + K[3]);
  const double Bm[12] = {xz * yz * K[0], -(1 + xz * xz) * K[0], yz * K[0], -iz * K[0], 0, xz * iz * K[0],
                         (1 + yz * yz) * K[1], -xz * yz * K[1], -xz * K[1], 0, -iz * K[1], yz * iz * K[1]};
  int k2 = 0;
  for (int a = 0; a < 6; ++a) {
    g[a] += -(Bm[a] * r0 + Bm[6 + a] * r1);
    for (int b = a; b < 6; ++b) H21[k2++] += Bm[a] * Bm[b] + Bm[6 + a] * Bm[6 + b];
  }
  return true;
}

__device__ bool solve6(const double* H21, const double* g, double* u) {   // symmetric 6x6, Gaussian elimination with pivoting
  double A[6][7];
  int k2 = 0;
  for (int a = 0; a < 6; ++a) for (int b = a; b < 6; ++b) { A[a][b] = A[b][a] = H21[k2++]; }
  for (int a = 0; a < 6; ++a) { A[a][a] += 1e-9; A[a][6] = g[a]; }
  for (int c = 0; c < 6; ++c) {
    int p = c;
    for (int r = c + 1; r < 6; ++r) if (fabs(A[r][c]) > fabs(A[p][c])) p = r;
    if (fabs(A[p][c]) < 1e-14) return false;
    if (p != c) for (int k = 0; k < 7; ++k) { const double t = A[p][k]; A[p][k] = A[c][k]; A[c][k] = t; }
    for (int r = c + 1; r < 6; ++r) {
      const double f = A[r][c] / A[c][c];
      for (int k = c; k < 7; ++k) A[r][k] -= f * A[c][k];
    }
  }
  for (int r = 5; r >= 0; --r) {
    double sres = A[r][6];
    for (int k = r + 1; k < 6; ++k) sres -= A[r][k] * u[k];
    u[r] = sres / A[r][r];
  }
  return true;
}

__device__ __forceinline__ bool pnp_inlier(const double* T, const double* K, const float* X3, const float* uv, double thr2) {
  const double X[3] = {X3[0], X3[1], X3[2]};
  double Xc[3];
  q_rot(T, X, Xc);
  const double z = Xc[2] + T[6];
  if (z < 1e-6) return false;
  // (u - (fx x / z + cx))^2 + (v - (fy y / z + cy))^2 <= thr2, multiplied through by z^2 > 0: no fp64 division
  const double ex = (uv[0] - K[2]) * z - (Xc[0] + T[4]) * K[0], ey = (uv[1] - K[3]) * z - (Xc[1] + T[5]) * K[1];
  return ex * ex + ey * ey <= thr2 * (z * z);
}

__global__ void __launch_bounds__(RS_THREADS) pnp_ransac_kernel(const int* __restrict__ npts, const float* __restrict__ p3d,
                                                                const float* __restrict__ p2d, const double* __restrict__ K4,
                                                                const double* __restrict__ T_in, int max_pts, double thr2,
                                                                double* __restrict__ T_out, uint8_t* __restrict__ mask_out,
                                                                int* __restrict__ n_inl) {
  __shared__ double sT[P_HYP + 1][7];
  __shared__ int score[P_HYP + 1];
  __shared__ double red[RS_WARPS][27];
  __shared__ double sTb[7];
  __shared__ int s_best, s_stop;
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = npts[s];
  const float* X = p3d + (size_t)s * max_pts * 3;
  const float* U = p2d + (size_t)s * max_pts * 2;
  uint8_t* mask = mask_out + (size_t)s * max_pts;
  const double K[4] = {K4[4 * s], K4[4 * s + 1], K4[4 * s + 2], K4[4 * s + 3]};
  double T0[7];
  for (int k = 0; k < 7; ++k) T0[k] = T_in[7 * s + k];
  if (n < 4) {
    for (int i = tid; i < max_pts; i += RS_THREADS) mask[i] = 0;
    if (tid == 0) { n_inl[s] = 0; for (int k = 0; k < 7; ++k) T_out[7 * s + k] = T0[k]; }
    return;
  }
  // hypothesis P_HYP is the prior itself
  for (int h = tid; h <= P_HYP; h += RS_THREADS) {
    double T[7];
    for (int k = 0; k < 7; ++k) T[k] = T0[k];
    if (h < P_HYP) {
      Lcg rng(0xD1B54A32D192ED03ULL + 0x9E3779B97F4A7C15ULL * (unsigned long long)(h + 1));
      int idx[P_SAMPLE];
      const int m = n < P_SAMPLE ? n : P_SAMPLE;
      sample_distinct(rng, n, m, idx);
      for (int it = 0; it < 8; ++it) {
        double H[21], g[6], u[6];
        for (int k = 0; k < 21; ++k) H[k] = 0;
        for (int k = 0; k < 6; ++k) g[k] = 0;
        for (int j = 0; j < m; ++j) pnp_accumulate(T, K, X + 3 * idx[j], U + 2 * idx[j], H, g);
        if (!solve6(H, g, u)) break;
        se3_oplus(T, u);
        double n2 = 0;
        for (int k = 0; k < 6; ++k) n2 += u[k] * u[k];
        if (n2 < 1e-20) break;
      }
    }
    for (int k = 0; k < 7; ++k) sT[h][k] = T[k];
  }
  __syncthreads();
  for (int h = warp; h <= P_HYP; h += RS_WARPS) {
    double T[7];
    for (int k = 0; k < 7; ++k) T[k] = sT[h][k];
    int c = 0;
    for (int i = lane; i < n; i += 32) c += pnp_inlier(T, K, X + 3 * i, U + 2 * i, thr2);
    c = __reduce_add_sync(FULL, c);
    if (lane == 0) score[h] = c;
  }
  __syncthreads();
  if (tid == 0) {
    int best = P_HYP;                                        // the prior wins ties
    for (int h = 0; h < P_HYP; ++h) if (score[h] > score[best]) best = h;
    s_best = best; s_stop = 0;
    for (int k = 0; k < 7; ++k) sTb[k] = sT[best][k];
  }
  __syncthreads();
  // refinement on all inliers of the best model: block-parallel normal equations, 10 Gauss-Newton steps
  for (int i = tid; i < max_pts; i += RS_THREADS) mask[i] = (i < n && pnp_inlier(sTb, K, X + 3 * i, U + 2 * i, thr2)) ? 1 : 0;
  __syncthreads();
  if (score[s_best] >= 4) {
    for (int it = 0; it < 10; ++it) {
      double T[7], H[21], g[6];
      for (int k = 0; k < 7; ++k) T[k] = sTb[k];
      for (int k = 0; k < 21; ++k) H[k] = 0;
      for (int k = 0; k < 6; ++k) g[k] = 0;
      for (int i = tid; i < n; i += RS_THREADS) if (mask[i]) pnp_accumulate(T, K, X + 3 * i, U + 2 * i, H, g);
      for (int k = 0; k < 27; ++k) {
        double v = k < 21 ? H[k] : g[k - 21];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == 0) red[warp][k] = v;
      }
      __syncthreads();
      if (tid == 0) {
        double Hs[21], gs[6], u[6];
        for (int k = 0; k < 27; ++k) { double v = 0; for (int w = 0; w < RS_WARPS; ++w) v += red[w][k]; if (k < 21) Hs[k] = v; else gs[k - 21] = v; }
        if (solve6(Hs, gs, u)) {
          double Tn[7];
          for (int k = 0; k < 7; ++k) Tn[k] = sTb[k];
          se3_oplus(Tn, u);
          for (int k = 0; k < 7; ++k) sTb[k] = Tn[k];
          double n2 = 0;
          for (int k = 0; k < 6; ++k) n2 += u[k] * u[k];
          if (n2 < 1e-20) s_stop = 1;
        } else s_stop = 1;
      }
      __syncthreads();
      if (s_stop) break;
    }
  }
  // final inlier set of the refined pose
  int c = 0;
  for (int i = tid; i < max_pts; i += RS_THREADS) {
    const bool in = i < n && pnp_inlier(sTb, K, X + 3 * i, U + 2 * i, thr2);
    mask[i] = in ? 1 : 0;
    c += in;
  }
  c = __reduce_add_sync(FULL, c);
  __syncthreads();
  if (lane == 0) score[warp] = c;
  __syncthreads();
  if (tid == 0) {
    int t = 0;
    for (int w = 0; w < RS_WARPS; ++w) t += score[w];
    n_inl[s] = t;
    for (int k = 0; k < 7; ++k) T_out[7 * s + k] = sTb[k];
  }
}

}  // namespace

int flv_launch_fmat_ransac(flv_ctx* ctx, int n_streams, const int* d_npts, const float* d_from, const float* d_to, double thr_px,
                           uint8_t* d_mask, double* d_F, int* d_ninl) {
  fmat_ransac_kernel<<<n_streams, RS_THREADS, 0, ctx->stream>>>(d_npts, d_from, d_to, ctx->max_pts, thr_px * thr_px, d_mask, d_F, d_ninl);
  ctx->launches++;
  FLV_CUDA(ctx, cudaGetLastError());
  return FLV_OK;
}

int flv_launch_pnp_ransac(flv_ctx* ctx, int n_streams, const int* d_npts, const float* d_p3d, const float* d_p2d, const double* d_K4,
                          const double* d_Tin, double thr_px, double* d_Tout, uint8_t* d_mask, int* d_ninl) {
  pnp_ransac_kernel<<<n_streams, RS_THREADS, 0, ctx->stream>>>(d_npts, d_p3d, d_p2d, d_K4, d_Tin, ctx->max_pts, thr_px * thr_px, d_Tout,
                                                               d_mask, d_ninl);
  ctx->launches++;
  FLV_CUDA(ctx, cudaGetLastError());
  return FLV_OK;
}
