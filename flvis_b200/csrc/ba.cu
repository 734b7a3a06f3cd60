// K7-K10 -- sliding-window bundle adjustment, whole Levenberg-Marquardt loop on the device.
// One persistent CTA per stream (= one independent window / one g2o SparseOptimizer); all streams
// of a batch run concurrently on different SMs.  No atomics on floating-point data: every sum has a
// fixed order, so results are run-to-run deterministic.
//
// Reference (paths under 3rdPartLib/g2o/g2o/ unless noted; restated in oracle/ba_ref.c):
//   residual / Jacobians   types/sba/types_six_dof_expmap.h:209-214, types_six_dof_expmap.cpp:389-433
//   Huber + quadratic form core/robust_kernel_impl.cpp:65-78, core/base_binary_edge.hpp:62-134
//   Schur + back-subst     core/block_solver.hpp:328-447;  lambda handling :525-565
//   LM control             core/optimization_algorithm_levenberg.cpp:58-175
//   outer loop / chi2      core/sparse_optimizer.cpp:366-430, :102-116;  active sets :168-272
//   callers                src/backend/vo_localmap.cpp:292-319 (12, cull chi2>3, 8),
//                          src/processing/optimize_in_frame.cpp:64-80 (2, cull, <10 edges => fail, 2)
//
// Data layout (per stream, fp64, global memory that stays L2-resident; S/b/x in shared memory):
//   eidx[p][l]   edge id observing landmark l from pose p (or -1): built once per optimize() call.
//   lmask[l]     bitmask of poses observing l (P <= 32).
//   W[e][6][3]   rho' * B^T A  (pose-landmark Hessian block of edge e).
//   Hll[l][6], bl[l][3], Dinv[l][6], xl[l][3];  Hd[pi][21], bp[6 pi] (shared).
// Passes per LM trial:
//   build   thread per landmark (Hll, bl, W) ; warp per pose (Hpp diagonal blocks, bp)
//   Schur   warp per pose pair (a<=b): members = landmarks seen by both (bitmask test + warp queue),
//           lane-private 6x6 accumulators, shuffle tree reduction, one writer per block of S
//   solve   dense Cholesky of the reduced camera system in shared memory (n = 6*(free poses) <= 144)
//   update  thread per landmark (back-substitution), thread per pose (exp map), chi2 block reduction
// Algorithmic bytes: build 168E+392P+120L, Schur 144E+96L+288Pf^2+48Pf per trial.
#include "ctx.h"

namespace {

constexpr int BA_THREADS = 512;
constexpr int BA_WARPS = BA_THREADS / 32;
constexpr int BA_MAX_POSES = 32;
constexpr int BA_MAX_FREE = 24;          // reduced system n <= 144 -> S fits shared memory
constexpr unsigned FULL = 0xffffffffu;

struct BAArgs {
  const flv_ba_problem* problems;
  flv_ba_params prm;
  double* poses; double* lms;
  const int* ep; const int* el; const double* uv; uint8_t* active;
  flv_ba_stats* stats;
  int max_poses, max_lms, max_edges;
  unsigned char* ws; size_t ws_stride;
  long long* prof;   // [S][8] cycle counters or nullptr
  int stage_doubles, stage_offset_doubles;   // Schur staging area inside the dynamic shared memory (0 = not available)
};

// ---- SE3 helpers (g2o SE3Quat semantics, quaternion stored x,y,z,w) ----------------------------
__device__ __forceinline__ void q_rotate(const double* q, const double* v, double* o) {
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
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    double qq[3];
    qq[i] = 0.5 * t; t = 0.5 / t;
    q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
    qq[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    qq[k] = (m[3 * k + i] + m[3 * i + k]) * t;
    q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2];
  }
}
__device__ void pose_oplus(double* pose, const double* u) {   // pose <- exp(u) * pose  (se3quat.h:218-260, :99-105)
  const double wx = u[0], wy = u[1], wz = u[2];
  const double theta = sqrt(wx * wx + wy * wy + wz * wz);
  const double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
  double O2[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) O2[3 * r + c] = O[3 * r] * O[c] + O[3 * r + 1] * O[3 + c] + O[3 * r + 2] * O[6 + c];
  double a, b, d;
  if (theta < 0.00001) { a = 1.0; b = 0.5; d = 1.0 / 6.0; }
  else { a = sin(theta) / theta; b = (1 - cos(theta)) / (theta * theta); d = (theta - sin(theta)) / (theta * theta * theta); }
  double R[9], V[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const double id = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
    R[i] = id + a * O[i] + b * O2[i];
    V[i] = id + b * O[i] + d * O2[i];
  }
  double qe[4], te[3], rt[3];
  R_to_q(R, qe);
#pragma unroll
  for (int r = 0; r < 3; ++r) te[r] = V[3 * r] * u[3] + V[3 * r + 1] * u[4] + V[3 * r + 2] * u[5];
  q_rotate(qe, pose + 4, rt);
  const double* b4 = pose;
  double x = qe[3] * b4[0] + qe[0] * b4[3] + qe[1] * b4[2] - qe[2] * b4[1];
  double y = qe[3] * b4[1] + qe[1] * b4[3] + qe[2] * b4[0] - qe[0] * b4[2];
  double z = qe[3] * b4[2] + qe[2] * b4[3] + qe[0] * b4[1] - qe[1] * b4[0];
  double w = qe[3] * b4[3] - qe[0] * b4[0] - qe[1] * b4[1] - qe[2] * b4[2];
  if (w < 0) { x = -x; y = -y; z = -z; w = -w; }
  const double nn = sqrt(x * x + y * y + z * z + w * w);
  pose[0] = x / nn; pose[1] = y / nn; pose[2] = z / nn; pose[3] = w / nn;
  pose[4] = te[0] + rt[0]; pose[5] = te[1] + rt[1]; pose[6] = te[2] + rt[2];
}

// W and Y (6x3 blocks per edge) are stored as two 16-byte aligned halves of 9 (+1 pad) doubles: 20 doubles per edge,
// so rows 0..2 / 3..5 can each be moved with 128-bit accesses.
#define WOFF(t) ((t) < 9 ? (t) : (t) + 1)
constexpr int WSTRIDE = 20;
constexpr int STG_STRIDE = 42;          // doubles per staged member (Ya 20 | Wb 20 | 2 pad: 2-way bank conflicts at most)

struct Cam { double fx, fy, cx, cy; };

// residual r (2), optional A = d r / d point (2x3), B = d r / d pose (2x6).  One fp64 division per edge:
// the reference's x/z, y/z, 1/z, x*y/z^2 ... are evaluated with iz = 1/z (differences ~1 ulp, tolerance-checked).
template <bool JAC>
__device__ __forceinline__ void edge_eval(const double* pose, const double* X, const double* uv, const Cam& c,
                                          double* r, double* A, double* B) {
  double Xc[3];
  q_rotate(pose, X, Xc);
  const double x = Xc[0] + pose[4], y = Xc[1] + pose[5], z = Xc[2] + pose[6];
  const double iz = 1.0 / z;
  const double xz = x * iz, yz = y * iz;
  r[0] = uv[0] - (xz * c.fx + c.cx);
  r[1] = uv[1] - (yz * c.fy + c.cy);
  if (JAC) {
    double R[9];
    q_to_R(pose, R);
    const double t02 = -xz * c.fx, t12 = -yz * c.fy, miz = -iz;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      A[k] = miz * (c.fx * R[k] + t02 * R[6 + k]);
      A[3 + k] = miz * (c.fy * R[3 + k] + t12 * R[6 + k]);
    }
    B[0] = xz * yz * c.fx; B[1] = -(1 + xz * xz) * c.fx; B[2] = yz * c.fx;
    B[3] = miz * c.fx; B[4] = 0; B[5] = xz * iz * c.fx;
    B[6] = (1 + yz * yz) * c.fy; B[7] = -xz * yz * c.fy; B[8] = -xz * c.fy;
    B[9] = 0; B[10] = miz * c.fy; B[11] = yz * iz * c.fy;
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// deterministic block sum; result valid in all threads.  red: shared double[BA_WARPS]
__device__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
#pragma unroll
  for (int i = 0; i < BA_WARPS; ++i) t += red[i];
  return t;
}
__device__ double block_max(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
#pragma unroll
  for (int i = 0; i < BA_WARPS; ++i) t = fmax(t, red[i]);
  return t;
}

constexpr int BA_MAX_PAIRS = BA_MAX_FREE * (BA_MAX_FREE + 1) / 2;   // 300

struct Sh {   // fixed-size shared state
  double red[BA_WARPS];
  double Hd[BA_MAX_FREE][21];
  double part[2 * BA_MAX_FREE][27];     // pose-pass partial sums (two halves per pose)
  double bp[6 * BA_MAX_FREE];
  double x[6 * BA_MAX_FREE];
  double piv;
  int pidx[BA_MAX_POSES];
  int pose_of[BA_MAX_FREE];
  int pcount[BA_MAX_POSES];
  int pstart[BA_MAX_POSES + 1];
  int pair_off[BA_MAX_PAIRS + 1];
  int np, fail, nact, overflow;
  long long prof[8], tlast;   // cycle counters: 0 chi2, 1 build, 2 schur, 3 cholesky, 4 substitution, 5 update, 6 setup
};

__device__ __forceinline__ void mark(Sh& sh, int slot) {
  if (threadIdx.x == 0) { const long long now = clock64(); sh.prof[slot] += now - sh.tlast; sh.tlast = now; }
}

struct Ws {   // per-stream global workspace views
  double *pbk, *lbk, *W, *Y, *Bw, *g, *Hll, *bl, *Dinv, *xl;
  int *eidx; unsigned* lmask; int* plist; int* pairs; int pair_cap;
};

__device__ __forceinline__ int sym21(int i, int j) {   // index into upper-triangular 6x6 (i<=j)
  return i * 6 - (i * (i - 1)) / 2 + (j - i);
}
__device__ __forceinline__ void pair_of(int blk, int np, int& a, int& b) {
  a = 0;
  int rem = blk;
  while (rem >= np - a) { rem -= np - a; ++a; }
  b = a + rem;
}

__device__ double robust_chi2(const flv_ba_problem& pb, const Cam& cam, const double* poses, const double* lms,
                              const int* ep, const int* el, const double* uv, const uint8_t* act, double delta,
                              double* red) {
  double acc = 0;
  const double d2 = delta * delta;
  for (int e = threadIdx.x; e < pb.n_edges; e += BA_THREADS) {
    if (!act[e]) continue;
    double r[2];
    edge_eval<false>(poses + 7 * ep[e], lms + 3 * el[e], uv + 2 * e, cam, r, nullptr, nullptr);
    const double c = r[0] * r[0] + r[1] * r[1];
    acc += (c <= d2) ? c : 2 * sqrt(c) * delta - d2;
  }
  return block_sum(acc, red);
}

// active sets + lookup tables (sparse_optimizer.cpp:168-272 semantics); built once per optimize() call:
//   eidx[p][l], lmask[l]; per-pose edge lists (landmark order); per pose-pair member lists (landmark order).
__device__ void setup_active(const flv_ba_problem& pb, const int* ep, const int* el, const uint8_t* act, Ws& ws, Sh& sh) {
  const int P = pb.n_poses, L = pb.n_landmarks, E = pb.n_edges, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < P * L; i += BA_THREADS) ws.eidx[i] = -1;
  for (int i = tid; i < L; i += BA_THREADS) ws.lmask[i] = 0u;
  if (tid < BA_MAX_POSES) sh.pcount[tid] = 0;
  if (tid == 0) sh.overflow = 0;
  __syncthreads();
  for (int e = tid; e < E; e += BA_THREADS) {
    if (!act[e]) continue;
    const int p = ep[e], l = el[e];
    ws.eidx[p * L + l] = e;
    atomicOr(&ws.lmask[l], 1u << p);
    atomicAdd(&sh.pcount[p], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int np = 0, nact = 0;
    for (int p = 0; p < P; ++p) {
      sh.pstart[p] = nact;
      nact += sh.pcount[p];
      if (p != pb.fixed_pose && sh.pcount[p] > 0) { sh.pidx[p] = np; sh.pose_of[np] = p; ++np; }
      else sh.pidx[p] = -1;
    }
    sh.pstart[P] = nact;
    sh.np = np; sh.nact = nact;
  }
  __syncthreads();
  // per-pose edge lists in landmark order (deterministic): warp per pose, ballot compaction
  for (int p = warp; p < P; p += BA_WARPS) {
    int base = sh.pstart[p];
    for (int l0 = 0; l0 < L; l0 += 32) {
      const int l = l0 + lane;
      const int e = l < L ? ws.eidx[p * L + l] : -1;
      const unsigned bal = __ballot_sync(FULL, e >= 0);
      if (e >= 0) ws.plist[base + __popc(bal & ((1u << lane) - 1))] = e;
      base += __popc(bal);
    }
  }
  if (pb.fix_landmarks || sh.np > BA_MAX_FREE) { __syncthreads(); return; }
  // pose-pair member lists: count, prefix, fill (two scans over lmask)
  const int np = sh.np, nblk = np * (np + 1) / 2;
  for (int blk = warp; blk < nblk; blk += BA_WARPS) {
    int a, b; pair_of(blk, np, a, b);
    const unsigned need = (1u << sh.pose_of[a]) | (1u << sh.pose_of[b]);
    int cnt = 0;
    for (int l0 = 0; l0 < L; l0 += 32) {
      const int l = l0 + lane;
      cnt += __popc(__ballot_sync(FULL, l < L && (ws.lmask[l] & need) == need));
    }
    if (lane == 0) sh.pair_off[blk + 1] = cnt;
  }
  __syncthreads();
  if (tid == 0) {
    sh.pair_off[0] = 0;
    for (int i = 0; i < nblk; ++i) sh.pair_off[i + 1] += sh.pair_off[i];
    if (sh.pair_off[nblk] > ws.pair_cap) sh.overflow = 1;
  }
  __syncthreads();
  if (sh.overflow) return;
  for (int blk = warp; blk < nblk; blk += BA_WARPS) {
    int a, b; pair_of(blk, np, a, b);
    const int pa = sh.pose_of[a], pb_ = sh.pose_of[b];
    const unsigned need = (1u << pa) | (1u << pb_);
    int base = sh.pair_off[blk];
    for (int l0 = 0; l0 < L; l0 += 32) {
      const int l = l0 + lane;
      const bool mem = l < L && (ws.lmask[l] & need) == need;
      const unsigned bal = __ballot_sync(FULL, mem);
      if (mem) {
        int* it = ws.pairs + 3 * (size_t)(base + __popc(bal & ((1u << lane) - 1)));
        it[0] = ws.eidx[pa * L + l]; it[1] = ws.eidx[pb_ * L + l]; it[2] = l;
      }
      base += __popc(bal);
    }
  }
  __syncthreads();
}

__device__ void build_system(const flv_ba_problem& pb, const Cam& cam, const double* poses, const double* lms,
                             const int* el, const double* uv, double delta, Ws& ws, Sh& sh) {
  const int L = pb.n_landmarks, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double d2 = delta * delta;
  if (!pb.fix_landmarks) {
    // thread per landmark: Hll, bl, and per edge W = rho' B^T A, Bw = sqrt(rho') B, g = -sqrt(rho') r
    for (int l = tid; l < L; l += BA_THREADS) {
      unsigned m = ws.lmask[l];
      if (!m) continue;
      double H[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
      const double X[3] = {lms[3 * l], lms[3 * l + 1], lms[3 * l + 2]};
      while (m) {
        const int p = __ffs(m) - 1; m &= m - 1;
        const int e = ws.eidx[p * L + l];
        double r[2], A[6], B[12];
        edge_eval<true>(poses + 7 * p, X, uv + 2 * e, cam, r, A, B);
        const double c = r[0] * r[0] + r[1] * r[1];
        const double rho1 = (c <= d2) ? 1.0 : delta / sqrt(c);
        const double o0 = -r[0] * rho1, o1 = -r[1] * rho1;
#pragma unroll
        for (int i = 0; i < 3; ++i) b[i] += A[i] * o0 + A[3 + i] * o1;
        H[0] += rho1 * (A[0] * A[0] + A[3] * A[3]); H[1] += rho1 * (A[0] * A[1] + A[3] * A[4]);
        H[2] += rho1 * (A[0] * A[2] + A[3] * A[5]); H[3] += rho1 * (A[1] * A[1] + A[4] * A[4]);
        H[4] += rho1 * (A[1] * A[2] + A[4] * A[5]); H[5] += rho1 * (A[2] * A[2] + A[5] * A[5]);
        if (sh.pidx[p] >= 0) {
          double* W = ws.W + WSTRIDE * (size_t)e;
#pragma unroll
          for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) W[WOFF(3 * i + j)] = rho1 * (B[i] * A[j] + B[6 + i] * A[3 + j]);
          const double sr = sqrt(rho1);
          double* Bw = ws.Bw + 12 * (size_t)e;
#pragma unroll
          for (int i = 0; i < 12; ++i) Bw[i] = sr * B[i];
          ws.g[2 * (size_t)e] = -sr * r[0]; ws.g[2 * (size_t)e + 1] = -sr * r[1];
        }
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) ws.Hll[6 * (size_t)l + i] = H[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) ws.bl[3 * (size_t)l + i] = b[i];
    }
    __syncthreads();
  }
  // pose diagonal blocks: task = (free pose, half of its edge list); lanes stride over the list
  for (int task = warp; task < 2 * sh.np; task += BA_WARPS) {
    const int pi = task >> 1, p = sh.pose_of[pi];
    const int b0 = sh.pstart[p], cnt = sh.pstart[p + 1] - b0;
    const int mid = (cnt + 1) >> 1;
    const int lo = (task & 1) ? mid : 0, hi = (task & 1) ? cnt : mid;
    double H[21], b[6];
#pragma unroll
    for (int i = 0; i < 21; ++i) H[i] = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) b[i] = 0;
    for (int k = lo + lane; k < hi; k += 32) {
      const int e = ws.plist[b0 + k];
      double B[12], g0, g1;
      if (pb.fix_landmarks) {
        double r[2], A[6];
        edge_eval<true>(poses + 7 * p, lms + 3 * el[e], uv + 2 * e, cam, r, A, B);
        const double c = r[0] * r[0] + r[1] * r[1];
        const double sr = (c <= d2) ? 1.0 : sqrt(delta / sqrt(c));
#pragma unroll
        for (int i = 0; i < 12; ++i) B[i] *= sr;
        g0 = -sr * r[0]; g1 = -sr * r[1];
      } else {
        const double* Bw = ws.Bw + 12 * (size_t)e;
#pragma unroll
        for (int i = 0; i < 12; ++i) B[i] = Bw[i];
        g0 = ws.g[2 * (size_t)e]; g1 = ws.g[2 * (size_t)e + 1];
      }
      int k2 = 0;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        b[i] += B[i] * g0 + B[6 + i] * g1;
#pragma unroll
        for (int j = i; j < 6; ++j) H[k2++] += B[i] * B[j] + B[6 + i] * B[6 + j];
      }
    }
#pragma unroll
    for (int i = 0; i < 21; ++i) { const double v = warp_sum(H[i]); if (lane == 0) sh.part[task][i] = v; }
#pragma unroll
    for (int i = 0; i < 6; ++i) { const double v = warp_sum(b[i]); if (lane == 0) sh.part[task][21 + i] = v; }
  }
  __syncthreads();
  for (int i = tid; i < sh.np * 27; i += BA_THREADS) {
    const int pi = i / 27, k = i - 27 * pi;
    const double v = sh.part[2 * pi][k] + sh.part[2 * pi + 1][k];
    if (k < 21) sh.Hd[pi][k] = v; else sh.bp[6 * pi + k - 21] = v;
  }
  __syncthreads();
}

// (H + lambda I) x = b through the Schur complement: S (n x ld, shared, lower triangle used), y, x in shared
__device__ void solve_system(const flv_ba_problem& pb, double lambda, double* S, double* y, int ld, double* stage, Ws& ws,
                             Sh& sh) {
  const int L = pb.n_landmarks, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np = sh.np, n = 6 * np;
  if (!pb.fix_landmarks) {
    for (int l = tid; l < L; l += BA_THREADS) {
      if (!ws.lmask[l]) continue;
      const double* H = ws.Hll + 6 * (size_t)l;
      const double a = H[0] + lambda, b = H[1], c = H[2], d = H[3] + lambda, e = H[4], f = H[5] + lambda;
      const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
      const double id = 1.0 / (a * c00 + b * c01 + c * c02);
      double Di[6] = {c00 * id, c01 * id, c02 * id, (a * f - c * c) * id, (b * c - a * e) * id, (a * d - b * b) * id};
#pragma unroll
      for (int i = 0; i < 6; ++i) ws.Dinv[6 * (size_t)l + i] = Di[i];
      // Y_e = W_e * Dinv for every edge of this landmark with a free pose (used by all pose pairs of the landmark)
      unsigned m = ws.lmask[l];
      while (m) {
        const int p = __ffs(m) - 1; m &= m - 1;
        if (sh.pidx[p] < 0) continue;
        const size_t e = (size_t)ws.eidx[p * L + l];
        const double* W = ws.W + WSTRIDE * e;
        double* Y = ws.Y + WSTRIDE * e;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const double w0 = W[WOFF(3 * i)], w1 = W[WOFF(3 * i + 1)], w2 = W[WOFF(3 * i + 2)];
          Y[WOFF(3 * i)] = w0 * Di[0] + w1 * Di[1] + w2 * Di[2];
          Y[WOFF(3 * i + 1)] = w0 * Di[1] + w1 * Di[3] + w2 * Di[4];
          Y[WOFF(3 * i + 2)] = w0 * Di[2] + w1 * Di[4] + w2 * Di[5];
        }
      }
    }
    __syncthreads();
  }
  const int nblk = np * (np + 1) / 2;
  // One warp per pose pair.  TWO lanes share a member: lane parity h owns columns 3h..3h+2 of the 6x6 block, so a
  // lane keeps 18 accumulators + Y_a (18) + half of W_b (9) in registers -- the full 36 + 18 + 18 layout spilled
  // under the 128-register cap and made this pass local-memory bound (profiles/r01_ba_notes.md).
  const int half = lane & 1, mslot = lane >> 1;
  for (int blk = warp; blk < nblk; blk += BA_WARPS) {
    int a, b; pair_of(blk, np, a, b);
    double acc[18], cf[6];
#pragma unroll
    for (int i = 0; i < 18; ++i) acc[i] = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) cf[i] = 0;
    if (!pb.fix_landmarks) {
      const int i0 = sh.pair_off[blk], i1 = sh.pair_off[blk + 1];
      if (stage) {
        // 16 members per round: the warp copies their Y_a | W_b blocks (2 x 160 B each) into its shared-memory
        // staging area with fully coalesced 128-bit loads (every 32-byte sector is requested once; the direct
        // per-lane 8-byte loads asked for each sector ~3x and made this pass LSU-bound), then computes from there.
        double* stg = stage + (size_t)warp * 16 * STG_STRIDE;
        for (int k0 = i0; k0 < i1; k0 += 16) {
          const int cntm = i1 - k0 < 16 ? i1 - k0 : 16;
          int ea = 0, eb = 0, ll = 0;
          if (lane < cntm) { const int* it = ws.pairs + 3 * (size_t)(k0 + lane); ea = it[0]; eb = it[1]; ll = it[2]; }
#pragma unroll
          for (int pss = 0; pss < 10; ++pss) {
            const int c = pss * 32 + lane, mem = c / 20, ch = c - 20 * mem;
            const int sa = __shfl_sync(FULL, ea, mem), sb = __shfl_sync(FULL, eb, mem);
            if (mem < cntm) {
              const double* base = ch < 10 ? ws.Y + WSTRIDE * (size_t)sa : ws.W + WSTRIDE * (size_t)sb;
              const double2 v = *reinterpret_cast<const double2*>(base + 2 * (ch < 10 ? ch : ch - 10));
              *reinterpret_cast<double2*>(stg + (size_t)mem * STG_STRIDE + 2 * ch) = v;
            }
          }
          const int myl = __shfl_sync(FULL, ll, mslot);
          __syncwarp();
          if (mslot < cntm) {
            const double* m = stg + (size_t)mslot * STG_STRIDE;
            double ya[18], wb[9];
#pragma unroll
            for (int i = 0; i < 18; ++i) ya[i] = m[WOFF(i)];
#pragma unroll
            for (int i = 0; i < 9; ++i) wb[i] = m[20 + 10 * half + i];
            if (a == b && half == 0) {
              const double* bl = ws.bl + 3 * (size_t)myl;
              const double b0 = bl[0], b1 = bl[1], b2 = bl[2];
#pragma unroll
              for (int i = 0; i < 6; ++i) cf[i] += ya[3 * i] * b0 + ya[3 * i + 1] * b1 + ya[3 * i + 2] * b2;
            }
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j)
                acc[3 * i + j] += ya[3 * i] * wb[3 * j] + ya[3 * i + 1] * wb[3 * j + 1] + ya[3 * i + 2] * wb[3 * j + 2];
          }
          __syncwarp();
        }
      } else {
      for (int k = i0 + mslot; k < i1; k += 16) {
        const int* it = ws.pairs + 3 * (size_t)k;
        const double* Ya = ws.Y + WSTRIDE * (size_t)it[0];
        const double* Wb = ws.W + WSTRIDE * (size_t)it[1] + 10 * half;  // rows 3h..3h+2 of W_b
        double ya[18], wb[9];
#pragma unroll
        for (int i = 0; i < 18; ++i) ya[i] = Ya[WOFF(i)];
#pragma unroll
        for (int i = 0; i < 9; ++i) wb[i] = Wb[i];
        if (a == b && half == 0) {
          const double* bl = ws.bl + 3 * (size_t)it[2];
          const double b0 = bl[0], b1 = bl[1], b2 = bl[2];
#pragma unroll
          for (int i = 0; i < 6; ++i) cf[i] += ya[3 * i] * b0 + ya[3 * i + 1] * b1 + ya[3 * i + 2] * b2;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j)
            acc[3 * i + j] += ya[3 * i] * wb[3 * j] + ya[3 * i + 1] * wb[3 * j + 1] + ya[3 * i + 2] * wb[3 * j + 2];
      }
      }
    }
    // reduce over the 16 member slots (lanes of equal parity): xor 2, 4, 8, 16
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      double v = acc[i];
#pragma unroll
      for (int o = 2; o < 32; o <<= 1) v += __shfl_xor_sync(FULL, v, o);
      acc[i] = v;
    }
    if (a == b) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        double v = cf[i];
#pragma unroll
        for (int o = 2; o < 32; o <<= 1) v += __shfl_xor_sync(FULL, v, o);
        cf[i] = v;
      }
    }
    // lanes 0 and 1 write their three columns: S block (b,a) of the lower triangle = -(acc)^T (+ Hpp + lambda)
    if (lane < 2) {
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {
          const int j = 3 * half + jj;
          double v = -acc[3 * i + jj];
          if (a == b) {
            v += sh.Hd[a][i <= j ? sym21(i, j) : sym21(j, i)];
            if (i == j) v += lambda;
          }
          S[(6 * b + j) * ld + 6 * a + i] = v;            // row index from pose b >= a: lower triangle
          if (a == b) S[(6 * a + i) * ld + 6 * b + j] = v;
        }
      if (a == b && lane == 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) y[6 * a + i] = sh.bp[6 * a + i] - cf[i];
      }
    }
  }
  if (tid == 0) sh.fail = 0;
  __syncthreads();
  mark(sh, 2);
  // left-looking Cholesky on the augmented matrix [S ; y^T]: row n is the right-hand side, so the forward
  // substitution comes for free.  T threads per row split each dot product.
  int T = 1;
  while ((n + 1) * (T << 1) <= BA_THREADS && T < 32) T <<= 1;
  const int row = tid / T, t = tid - row * T;
  for (int j = 0; j < n; ++j) {
    double sres = 0;
    const bool mine = row >= j && row <= n;
    const double* Li = row < n ? S + row * ld : y;         // row n: y holds b (entries < j already y_k)
    double acc = 0;
    if (mine) {
      const double* Lj = S + j * ld;
      for (int k = t; k < j; k += T) acc += Li[k] * Lj[k];
    }
    for (int o = T >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);   // whole warp, uniform
    if (mine) {
      sres = Li[j] - acc;
      if (row == j && t == 0) sh.piv = sres;
    }
    __syncthreads();
    const double d = sh.piv;
    if (!(d > 0)) { if (tid == 0) sh.fail = 1; }
    const double isd = rsqrt(d > 0 ? d : 1.0);
    if (mine && t == 0) {
      if (row == j) S[j * ld + j] = d * isd;
      else if (row < n) S[row * ld + j] = sres * isd;
      else y[j] = sres * isd;
    }
    __syncthreads();
  }
  mark(sh, 3);
  // back substitution x = L^-T y, column oriented
  for (int j = n - 1; j >= 0; --j) {
    const double xj = y[j] / S[j * ld + j];
    __syncthreads();
    if (tid < j) y[tid] -= S[j * ld + tid] * xj;
    if (tid == j) sh.x[j] = xj;
    __syncthreads();
  }
  mark(sh, 4);
}

// state update (sparse_optimizer.cpp:433-446) incl. landmark back-substitution (block_solver.hpp:422-444).
// Returns sum_j x_j (lambda x_j + b_j) (computeScale, optimization_algorithm_levenberg.cpp:168-175).
__device__ double apply_update(const flv_ba_problem& pb, double lambda, double* poses, double* lms, Ws& ws, Sh& sh) {
  const int P = pb.n_poses, L = pb.n_landmarks, tid = threadIdx.x;
  double sc = 0;
  for (int i = tid; i < 7 * P; i += BA_THREADS) ws.pbk[i] = poses[i];
  if (!pb.fix_landmarks) {
    for (int l = tid; l < L; l += BA_THREADS) {
      double* X = lms + 3 * (size_t)l;
      ws.lbk[3 * (size_t)l] = X[0]; ws.lbk[3 * (size_t)l + 1] = X[1]; ws.lbk[3 * (size_t)l + 2] = X[2];
      unsigned m = ws.lmask[l];
      if (!m) continue;
      const double* bl = ws.bl + 3 * (size_t)l;
      double c0 = bl[0], c1 = bl[1], c2 = bl[2];
      while (m) {
        const int p = __ffs(m) - 1; m &= m - 1;
        const int pi = sh.pidx[p];
        if (pi < 0) continue;
        const double* W = ws.W + WSTRIDE * (size_t)ws.eidx[p * L + l];
        const double* xp = sh.x + 6 * pi;
#pragma unroll
        for (int i = 0; i < 6; ++i) { c0 -= W[WOFF(3 * i)] * xp[i]; c1 -= W[WOFF(3 * i + 1)] * xp[i]; c2 -= W[WOFF(3 * i + 2)] * xp[i]; }
      }
      const double* Di = ws.Dinv + 6 * (size_t)l;
      const double x0 = Di[0] * c0 + Di[1] * c1 + Di[2] * c2, x1 = Di[1] * c0 + Di[3] * c1 + Di[4] * c2,
                   x2 = Di[2] * c0 + Di[4] * c1 + Di[5] * c2;
      sc += x0 * (lambda * x0 + bl[0]) + x1 * (lambda * x1 + bl[1]) + x2 * (lambda * x2 + bl[2]);
      X[0] += x0; X[1] += x1; X[2] += x2;
    }
  }
  if (tid < 6 * sh.np) sc += sh.x[tid] * (lambda * sh.x[tid] + sh.bp[tid]);
  __syncthreads();     // backups of poses complete before anyone overwrites
  if (tid < sh.np) pose_oplus(poses + 7 * sh.pose_of[tid], sh.x + 6 * tid);
  return block_sum(sc, sh.red);
}

__device__ void restore_state(const flv_ba_problem& pb, double* poses, double* lms, Ws& ws) {
  for (int i = threadIdx.x; i < 7 * pb.n_poses; i += BA_THREADS) poses[i] = ws.pbk[i];
  if (!pb.fix_landmarks)
    for (int i = threadIdx.x; i < 3 * pb.n_landmarks; i += BA_THREADS) lms[i] = ws.lbk[i];
  __syncthreads();
}

__host__ __device__ inline size_t pair_capacity(int max_poses, int max_edges) {
  return (size_t)max_edges * ((max_poses + 2) / 2);
}

__global__ void __launch_bounds__(BA_THREADS, 1) ba_kernel(BAArgs a) {
  extern __shared__ double dyn[];
  __shared__ Sh sh;
  const int s = blockIdx.x, tid = threadIdx.x;
  const flv_ba_problem pb = a.problems[s];
  const Cam cam = {pb.fx, pb.fy, pb.cx, pb.cy};
  double* poses = a.poses + (size_t)s * a.max_poses * 7;
  double* lms = a.lms + (size_t)s * a.max_lms * 3;
  const int* ep = a.ep + (size_t)s * a.max_edges;
  const int* el = a.el + (size_t)s * a.max_edges;
  const double* uv = a.uv + (size_t)s * a.max_edges * 2;
  uint8_t* act = a.active + (size_t)s * a.max_edges;
  const int P = pb.n_poses, L = pb.n_landmarks, E = pb.n_edges;
  // carve the per-stream workspace
  Ws ws;
  {
    double* d = (double*)(a.ws + (size_t)s * a.ws_stride);
    ws.pbk = d; d += 7 * a.max_poses;
    ws.lbk = d; d += 3 * a.max_lms;
    ws.W = d; d += WSTRIDE * (size_t)a.max_edges;
    ws.Bw = d; d += 12 * (size_t)a.max_edges;
    ws.g = d; d += 2 * (size_t)a.max_edges;
    ws.Hll = d; d += 6 * a.max_lms;
    ws.bl = d; d += 3 * a.max_lms;
    ws.Dinv = d; d += 6 * a.max_lms;
    ws.xl = d; d += 3 * a.max_lms;
    ws.Y = d; d += WSTRIDE * (size_t)a.max_edges;
    ws.eidx = (int*)d;
    ws.lmask = (unsigned*)(ws.eidx + (size_t)a.max_poses * a.max_lms);
    ws.plist = (int*)(ws.lmask + a.max_lms);
    ws.pairs = ws.plist + a.max_edges;
    ws.pair_cap = (int)pair_capacity(a.max_poses, a.max_edges);
  }
  const double delta = a.prm.huber_delta;
  flv_ba_stats st;
  st.iterations_run = 0; st.n_culled = 0; st.ok = 1; st.reserved = 0;
  st.chi2_initial = st.chi2_after1 = st.chi2_final = 0; st.lambda_final = 0;
  if (P < 1 || P > BA_MAX_POSES || E < 0) {
    if (tid == 0) { st.ok = 0; st.reserved = 1; a.stats[s] = st; }
    return;
  }
  if (tid == 0) { for (int i = 0; i < 8; ++i) sh.prof[i] = 0; sh.tlast = clock64(); }
  st.chi2_initial = robust_chi2(pb, cam, poses, lms, ep, el, uv, act, delta, sh.red);
  double lambda = 0;
  for (int phase = 0; phase < 2; ++phase) {
    const int iters = phase == 0 ? a.prm.iters1 : a.prm.iters2;
    setup_active(pb, ep, el, act, ws, sh);
    mark(sh, 6);
    if (sh.np > BA_MAX_FREE) { st.ok = 0; st.reserved = 2; break; }
    if (sh.overflow) { st.ok = 0; st.reserved = 3; break; }
    const int n = 6 * sh.np, ld = n + 1;
    double* S = dyn;
    double* y = dyn + (size_t)n * ld;
    // staging area of the Schur pass sits behind S | y when the launch reserved room for it (a.stage_doubles > 0)
    double* stage = a.stage_doubles ? dyn + a.stage_offset_doubles : nullptr;
    double ni = 2;
    for (int it = 0; it < iters; ++it) {
      double currentChi = robust_chi2(pb, cam, poses, lms, ep, el, uv, act, delta, sh.red);
      mark(sh, 0);
      build_system(pb, cam, poses, lms, el, uv, delta, ws, sh);
      mark(sh, 1);
      if (it == 0) {
        double md = 0;
        if (tid < sh.np) {
#pragma unroll
          for (int i = 0; i < 6; ++i) md = fmax(md, fabs(sh.Hd[tid][sym21(i, i)]));
        }
        if (!pb.fix_landmarks)
          for (int l = tid; l < L; l += BA_THREADS)
            if (ws.lmask[l]) md = fmax(md, fmax(fabs(ws.Hll[6 * (size_t)l]), fmax(fabs(ws.Hll[6 * (size_t)l + 3]), fabs(ws.Hll[6 * (size_t)l + 5]))));
        lambda = 1e-5 * block_max(md, sh.red);
        ni = 2;
      }
      double rho = 0;
      int qmax = 0;
      do {
        solve_system(pb, lambda, S, y, ld, stage, ws, sh);
        const int ok2 = !sh.fail;
        double scale = 0, tempChi;
        if (ok2) {
          scale = apply_update(pb, lambda, poses, lms, ws, sh);
          __syncthreads();
          mark(sh, 5);
          tempChi = robust_chi2(pb, cam, poses, lms, ep, el, uv, act, delta, sh.red);
          mark(sh, 0);
        } else {
          tempChi = 1.7976931348623157e308;
        }
        rho = (currentChi - tempChi) / (scale + 1e-3);
        if (rho > 0 && isfinite(tempChi)) {
          double alpha = 1. - (2 * rho - 1) * (2 * rho - 1) * (2 * rho - 1);
          alpha = fmin(alpha, 2. / 3.);
          lambda *= fmax(1. / 3., alpha);
          ni = 2;
          currentChi = tempChi;
        } else {
          lambda *= ni; ni *= 2;
          if (ok2) restore_state(pb, poses, lms, ws);
          if (!isfinite(lambda)) break;
        }
        ++qmax;
      } while (rho < 0 && qmax < 10);
      ++st.iterations_run;
      if (qmax == 10 || rho == 0 || !isfinite(lambda)) break;
    }
    if (phase == 0) {
      st.chi2_after1 = robust_chi2(pb, cam, poses, lms, ep, el, uv, act, delta, sh.red);
      // cull: un-robustified chi2 > threshold (vo_localmap.cpp:303-316, optimize_in_frame.cpp:67-74)
      int culled = 0, remaining = 0;
      for (int e = tid; e < E; e += BA_THREADS) {
        if (!act[e]) continue;
        double r[2];
        edge_eval<false>(poses + 7 * ep[e], lms + 3 * el[e], uv + 2 * e, cam, r, nullptr, nullptr);
        if (r[0] * r[0] + r[1] * r[1] > a.prm.cull_chi2) { act[e] = 0; ++culled; } else ++remaining;
      }
      st.n_culled = (int)(block_sum((double)culled, sh.red) + 0.5);
      const int rem = (int)(block_sum((double)remaining, sh.red) + 0.5);
      __syncthreads();
      if (rem < a.prm.min_edges_after_cull) { st.ok = 0; break; }
    }
  }
  st.chi2_final = robust_chi2(pb, cam, poses, lms, ep, el, uv, act, delta, sh.red);
  st.lambda_final = lambda;
  if (tid == 0) {
    a.stats[s] = st;
    if (a.prof) for (int i = 0; i < 8; ++i) a.prof[8 * s + i] = sh.prof[i];
  }
}

size_t ws_stride_bytes(int max_poses, int max_lms, int max_edges) {
  size_t d = 7 * (size_t)max_poses + 3 * (size_t)max_lms + (2 * WSTRIDE + 12 + 2) * (size_t)max_edges + 6 * (size_t)max_lms +
             3 * (size_t)max_lms + 6 * (size_t)max_lms + 3 * (size_t)max_lms;
  size_t ints = (size_t)max_poses * max_lms + max_lms + max_edges + 3 * pair_capacity(max_poses, max_edges);
  size_t b = d * 8 + ints * 4;
  return (b + 255) & ~(size_t)255;
}

size_t ba_sys_doubles(int nfree) {
  const size_t n = 6 * (size_t)nfree;
  return ((n * (n + 1) + n + 8) + 1) & ~(size_t)1;          // S | y, rounded to 16 bytes
}
constexpr size_t BA_STAGE_DOUBLES = (size_t)BA_WARPS * 16 * STG_STRIDE;
// dynamic shared memory of a launch: the reduced system plus, when it still fits next to the static state, the staging area
size_t ba_dyn_smem(int nfree, bool* with_stage) {
  const size_t sys = ba_sys_doubles(nfree) * 8, stg = BA_STAGE_DOUBLES * 8;
  const bool fits = sys + stg + sizeof(Sh) + 1024 <= 227 * 1024;
  if (with_stage) *with_stage = fits;
  return sys + (fits ? stg : 0);
}

}  // namespace

int flv_ba_free(flv_ctx* ctx) {
  if (ctx->ba_ws) cudaFree(ctx->ba_ws);
  ctx->ba_ws = nullptr; ctx->ba_ws_bytes = 0;
  return FLV_OK;
}

extern "C" {

int flv_ba_reserve(flv_ctx* ctx, int max_poses, int max_landmarks, int max_edges) {
  if (!ctx || max_poses < 1 || max_landmarks < 1 || max_edges < 1) return FLV_ERR_INVALID;
  if (max_poses > BA_MAX_POSES)
    FLV_FAIL(ctx, FLV_ERR_UNSUPPORTED, "window of %d poses: this build supports <= %d (<= %d free poses)", max_poses,
             BA_MAX_POSES, BA_MAX_FREE);
  flv_ba_free(ctx);
  size_t stride = ws_stride_bytes(max_poses, max_landmarks, max_edges);
  // tail: device copies of problems / stats / staging are carved after the per-stream blocks
  size_t total = stride * ctx->S + (size_t)ctx->S * (sizeof(flv_ba_problem) + sizeof(flv_ba_stats) + 64) + 512;
  FLV_CUDA(ctx, cudaMalloc(&ctx->ba_ws, total));
  ctx->ba_ws_bytes = total;
  ctx->ba_max_poses = max_poses; ctx->ba_max_lms = max_landmarks; ctx->ba_max_edges = max_edges;
  int nfree = max_poses < BA_MAX_FREE ? max_poses : BA_MAX_FREE;
  size_t smem = ba_dyn_smem(nfree, nullptr);
  FLV_CUDA(ctx, cudaFuncSetAttribute(ba_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return FLV_OK;
}

int flv_ba_optimize(flv_ctx* ctx, int n_streams, const flv_ba_problem* problems, const flv_ba_params* prm,
                    double* poses, double* landmarks, const int* edge_pose, const int* edge_lm,
                    const double* edge_uv, uint8_t* edge_active, flv_ba_stats* stats, flv_memspace mem) {
  if (!ctx || !problems || !prm || !poses || !landmarks || !edge_pose || !edge_lm || !edge_uv || !edge_active ||
      !stats || n_streams < 1 || n_streams > ctx->S)
    return FLV_ERR_INVALID;
  if (!ctx->ba_ws) FLV_FAIL(ctx, FLV_ERR_INVALID, "flv_ba_reserve has not been called");
  const int MP = ctx->ba_max_poses, ML = ctx->ba_max_lms, ME = ctx->ba_max_edges;
  const size_t stride = ws_stride_bytes(MP, ML, ME);
  unsigned char* tail = (unsigned char*)ctx->ba_ws + stride * ctx->S;
  const int slot0 = prm->ws_slot0;
  if (slot0 < 0 || slot0 + n_streams > ctx->S) FLV_FAIL(ctx, FLV_ERR_INVALID, "ws_slot0 %d + %d streams exceeds %d slots", slot0, n_streams, ctx->S);
  flv_ba_problem* d_prob = (flv_ba_problem*)tail + slot0;
  flv_ba_stats* d_stats = (flv_ba_stats*)(tail + (size_t)ctx->S * sizeof(flv_ba_problem)) + slot0;
  BAArgs a;
  a.prof = (long long*)(tail + (size_t)ctx->S * (sizeof(flv_ba_problem) + sizeof(flv_ba_stats))) + 8 * slot0;
  a.prm = *prm; a.max_poses = MP; a.max_lms = ML; a.max_edges = ME;
  a.ws = (unsigned char*)ctx->ba_ws + stride * slot0; a.ws_stride = stride;
  const int nfree = MP < BA_MAX_FREE ? MP : BA_MAX_FREE;
  bool with_stage = false;
  const size_t smem = ba_dyn_smem(nfree, &with_stage);
  a.stage_doubles = with_stage ? (int)BA_STAGE_DOUBLES : 0;
  a.stage_offset_doubles = (int)ba_sys_doubles(nfree);
  const size_t S = n_streams;
  cudaStream_t stream = ctx->ba_stream_set ? ctx->ba_stream : ctx->stream;
  if (mem == FLV_MEM_DEVICE) {
    a.problems = problems; a.poses = poses; a.lms = landmarks; a.ep = edge_pose; a.el = edge_lm; a.uv = edge_uv;
    a.active = edge_active; a.stats = stats;
    ba_kernel<<<n_streams, BA_THREADS, smem, stream>>>(a);
    ctx->launches++;
    FLV_CUDA(ctx, cudaGetLastError());
    return FLV_OK;
  }
  for (int s = 0; s < n_streams; ++s) {
    const flv_ba_problem& p = problems[s];
    if (p.n_poses < 1 || p.n_poses > MP || p.n_landmarks < 0 || p.n_landmarks > ML || p.n_edges < 0 || p.n_edges > ME)
      FLV_FAIL(ctx, FLV_ERR_INVALID, "stream %d: problem (P=%d L=%d E=%d) exceeds reserved (%d,%d,%d)", s, p.n_poses,
               p.n_landmarks, p.n_edges, MP, ML, ME);
  }
  // staging layout: poses | lms | uv | ep | el | active
  const size_t b_pose = S * MP * 7 * 8, b_lm = S * ML * 3 * 8, b_uv = S * ME * 2 * 8, b_i = S * ME * 4, b_a = S * ME;
  const size_t o_pose = 0, o_lm = o_pose + b_pose, o_uv = o_lm + b_lm, o_ep = o_uv + b_uv, o_el = o_ep + b_i,
               o_act = o_el + b_i, total = ((o_act + b_a + 255) & ~(size_t)255);
  int rc = flv_stage_reserve(ctx, total);
  if (rc) return rc;
  char* hs = (char*)ctx->h_stage; char* ds = (char*)ctx->d_stage;
  memcpy(hs + o_pose, poses, b_pose); memcpy(hs + o_lm, landmarks, b_lm); memcpy(hs + o_uv, edge_uv, b_uv);
  memcpy(hs + o_ep, edge_pose, b_i); memcpy(hs + o_el, edge_lm, b_i); memcpy(hs + o_act, edge_active, b_a);
  FLV_CUDA(ctx, cudaMemcpyAsync(ds, hs, total, cudaMemcpyHostToDevice, stream));
  FLV_CUDA(ctx, cudaMemcpyAsync(d_prob, problems, S * sizeof(flv_ba_problem), cudaMemcpyHostToDevice, stream));
  a.problems = d_prob; a.poses = (double*)(ds + o_pose); a.lms = (double*)(ds + o_lm); a.uv = (const double*)(ds + o_uv);
  a.ep = (const int*)(ds + o_ep); a.el = (const int*)(ds + o_el); a.active = (uint8_t*)(ds + o_act); a.stats = d_stats;
  ba_kernel<<<n_streams, BA_THREADS, smem, stream>>>(a);
  ctx->launches++;
  FLV_CUDA(ctx, cudaGetLastError());
  FLV_CUDA(ctx, cudaMemcpyAsync(hs + o_pose, ds + o_pose, b_pose + b_lm, cudaMemcpyDeviceToHost, stream));
  FLV_CUDA(ctx, cudaMemcpyAsync(hs + o_act, ds + o_act, b_a, cudaMemcpyDeviceToHost, stream));
  FLV_CUDA(ctx, cudaMemcpyAsync(stats, d_stats, S * sizeof(flv_ba_stats), cudaMemcpyDeviceToHost, stream));
  FLV_CUDA(ctx, cudaStreamSynchronize(stream));
  memcpy(poses, hs + o_pose, b_pose); memcpy(landmarks, hs + o_lm, b_lm); memcpy(edge_active, hs + o_act, b_a);
  for (int s = 0; s < n_streams; ++s)
    if (stats[s].reserved)
      FLV_FAIL(ctx, FLV_ERR_UNSUPPORTED, "stream %d: %s", s,
               stats[s].reserved == 1 ? "pose count outside [1,32]" : stats[s].reserved == 2 ? "more than 24 free poses (reduced system > 144)" : "pose-pair list capacity exceeded");
  return FLV_OK;
}

int flv_set_ba_stream(flv_ctx* ctx, void* cuda_stream, int enable) {
  if (!ctx) return FLV_ERR_INVALID;
  ctx->ba_stream = (cudaStream_t)cuda_stream;
  ctx->ba_stream_set = enable ? 1 : 0;
  return FLV_OK;
}

/* debug: cycle counters of the last flv_ba_optimize for `stream` (8 values, see Sh::prof) */
int flv_ba_profile(flv_ctx* ctx, int stream, long long* out8) {
  if (!ctx || !ctx->ba_ws || !out8 || stream < 0 || stream >= ctx->S) return FLV_ERR_INVALID;
  const size_t stride = ws_stride_bytes(ctx->ba_max_poses, ctx->ba_max_lms, ctx->ba_max_edges);
  unsigned char* tail = (unsigned char*)ctx->ba_ws + stride * ctx->S;
  long long* d = (long long*)(tail + (size_t)ctx->S * (sizeof(flv_ba_problem) + sizeof(flv_ba_stats)));
  FLV_CUDA(ctx, cudaDeviceSynchronize());
  FLV_CUDA(ctx, cudaMemcpy(out8, d + 8 * stream, 64, cudaMemcpyDeviceToHost));
  return FLV_OK;
}

}  // extern "C"
