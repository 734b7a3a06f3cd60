#include "ransac.h"
#include <algorithm>
#include <cmath>
#include "small_linalg.h"

namespace flv {

namespace {

struct Lcg {   // deterministic sample stream (the reference's OpenCV RNG is also fixed per call)
  uint64_t s;
  explicit Lcg(uint64_t seed) : s(seed) {}
  uint32_t next() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (uint32_t)(s >> 33); }
  int below(int n) { return (int)(next() % (uint32_t)n); }
};

void sample_distinct(Lcg& rng, int n, int k, int* out) {
  for (int i = 0; i < k; ++i) {
    for (;;) {
      int v = rng.below(n);
      bool dup = false;
      for (int j = 0; j < i; ++j) dup |= (out[j] == v);
      if (!dup) { out[i] = v; break; }
    }
  }
}

// normalised 8-point on the index set idx (size m >= 8); returns false on degenerate input
bool eight_point(const std::vector<P2f>& a, const std::vector<P2f>& b, const int* idx, int m, double* F) {
  double ca[2] = {0, 0}, cb[2] = {0, 0};
  for (int i = 0; i < m; ++i) { ca[0] += a[idx[i]].x; ca[1] += a[idx[i]].y; cb[0] += b[idx[i]].x; cb[1] += b[idx[i]].y; }
  ca[0] /= m; ca[1] /= m; cb[0] /= m; cb[1] /= m;
  double da = 0, db = 0;
  for (int i = 0; i < m; ++i) {
    da += std::hypot(a[idx[i]].x - ca[0], a[idx[i]].y - ca[1]);
    db += std::hypot(b[idx[i]].x - cb[0], b[idx[i]].y - cb[1]);
  }
  if (da < 1e-9 || db < 1e-9) return false;
  const double sa = std::sqrt(2.0) * m / da, sb = std::sqrt(2.0) * m / db;
  std::vector<double> M((size_t)m * 9);
  for (int i = 0; i < m; ++i) {
    const double x1 = (a[idx[i]].x - ca[0]) * sa, y1 = (a[idx[i]].y - ca[1]) * sa;
    const double x2 = (b[idx[i]].x - cb[0]) * sb, y2 = (b[idx[i]].y - cb[1]) * sb;
    double* r = &M[(size_t)i * 9];
    r[0] = x2 * x1; r[1] = x2 * y1; r[2] = x2; r[3] = y2 * x1; r[4] = y2 * y1; r[5] = y2; r[6] = x1; r[7] = y1; r[8] = 1;
  }
  std::vector<double> f = null_vector(M, m, 9);
  double U[9], s[3], V[9], Fn[9];
  svd3(f.data(), U, s, V);
  s[2] = 0;                                               // rank 2
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Fn[3 * i + j] = U[3 * i] * s[0] * V[3 * j] + U[3 * i + 1] * s[1] * V[3 * j + 1];
  // denormalise: F = Tb^T Fn Ta
  const double Ta[9] = {sa, 0, -sa * ca[0], 0, sa, -sa * ca[1], 0, 0, 1};
  const double TbT[9] = {sb, 0, 0, 0, sb, 0, -sb * cb[0], -sb * cb[1], 1};
  double tmp[9];
  mat3_mul(Fn, Ta, tmp);
  mat3_mul(TbT, tmp, F);
  return true;
}

int count_f_inliers(const std::vector<P2f>& a, const std::vector<P2f>& b, const double* F, double thr2, std::vector<uint8_t>* mask) {
  int cnt = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    const double x1 = a[i].x, y1 = a[i].y, x2 = b[i].x, y2 = b[i].y;
    double la = F[0] * x1 + F[1] * y1 + F[2], lb = F[3] * x1 + F[4] * y1 + F[5], lc = F[6] * x1 + F[7] * y1 + F[8];
    double d2 = x2 * la + y2 * lb + lc;
    const double e2 = d2 * d2 / (la * la + lb * lb + 1e-300);
    la = F[0] * x2 + F[3] * y2 + F[6]; lb = F[1] * x2 + F[4] * y2 + F[7]; lc = F[2] * x2 + F[5] * y2 + F[8];
    double d1 = x1 * la + y1 * lb + lc;
    const double e1 = d1 * d1 / (la * la + lb * lb + 1e-300);
    const bool in = std::fmax(e1, e2) <= thr2;
    if (mask) (*mask)[i] = in ? 1 : 0;
    cnt += in;
  }
  return cnt;
}

int adaptive_iters(double conf, double inlier_ratio, int sample, int max_iters) {
  const double w = std::pow(std::fmin(std::fmax(inlier_ratio, 1e-6), 1.0), sample);
  if (w >= 1.0 - 1e-12) return 0;
  const double n = std::log(1.0 - conf) / std::log(1.0 - w);
  return n < 0 || n > max_iters ? max_iters : (int)std::ceil(n);
}

void q_rot(const double* q, const double* v, double* o) {
  double ux = q[1] * v[2] - q[2] * v[1], uy = q[2] * v[0] - q[0] * v[2], uz = q[0] * v[1] - q[1] * v[0];
  ux += ux; uy += uy; uz += uz;
  o[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
  o[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
  o[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}

// Gauss-Newton on the reprojection error over idx (same residual / Jacobian as EdgeSE3ProjectXYZ w.r.t. the pose)
void refine_pose(const std::vector<P3f>& p3d, const std::vector<P2f>& p2d, const double K[4], const int* idx, int m,
                 Pose7& T, int iters) {
  for (int it = 0; it < iters; ++it) {
    double H[36] = {0}, g[6] = {0};
    for (int k = 0; k < m; ++k) {
      const int i = idx[k];
      const double X[3] = {p3d[i].x, p3d[i].y, p3d[i].z};
      double Xc[3];
      q_rot(T.data(), X, Xc);
      const double x = Xc[0] + T[4], y = Xc[1] + T[5], z = Xc[2] + T[6];
      if (z < 1e-6) continue;
      const double iz = 1.0 / z, xz = x * iz, yz = y * iz;
      const double r0 = p2d[i].x - (xz * K[0] + K[2]), r1 = p2d[i].y - (yz * K[1] + K[3]);
      const double B[12] = {xz * yz * K[0], -(1 + xz * xz) * K[0], yz * K[0], -iz * K[0], 0, xz * iz * K[0],
                            (1 + yz * yz) * K[1], -xz * yz * K[1], -xz * K[1], 0, -iz * K[1], yz * iz * K[1]};
      for (int a = 0; a < 6; ++a) {
        g[a] += -(B[a] * r0 + B[6 + a] * r1);
        for (int b = 0; b < 6; ++b) H[6 * a + b] += B[a] * B[b] + B[6 + a] * B[6 + b];
      }
    }
    for (int a = 0; a < 6; ++a) H[7 * a] += 1e-9;
    // solve H u = g (Gaussian elimination with partial pivoting)
    double A[6][7];
    for (int a = 0; a < 6; ++a) { for (int b = 0; b < 6; ++b) A[a][b] = H[6 * a + b]; A[a][6] = g[a]; }
    bool ok = true;
    for (int c = 0; c < 6 && ok; ++c) {
      int p = c;
      for (int r = c + 1; r < 6; ++r) if (std::fabs(A[r][c]) > std::fabs(A[p][c])) p = r;
      if (std::fabs(A[p][c]) < 1e-14) { ok = false; break; }
      if (p != c) for (int k = 0; k < 7; ++k) std::swap(A[p][k], A[c][k]);
      for (int r = c + 1; r < 6; ++r) {
        const double f = A[r][c] / A[c][c];
        for (int k = c; k < 7; ++k) A[r][k] -= f * A[c][k];
      }
    }
    if (!ok) return;
    double u[6];
    for (int r = 5; r >= 0; --r) {
      double sres = A[r][6];
      for (int k = r + 1; k < 6; ++k) sres -= A[r][k] * u[k];
      u[r] = sres / A[r][r];
    }
    se3_oplus(T, u);
    double n2 = 0;
    for (double v : u) n2 += v * v;
    if (n2 < 1e-20) break;
  }
}

int count_pnp_inliers(const std::vector<P3f>& p3d, const std::vector<P2f>& p2d, const double K[4], const Pose7& T, double thr2,
                      std::vector<int>* out) {
  int cnt = 0;
  if (out) out->clear();
  for (size_t i = 0; i < p3d.size(); ++i) {
    const double X[3] = {p3d[i].x, p3d[i].y, p3d[i].z};
    double Xc[3];
    q_rot(T.data(), X, Xc);
    const double z = Xc[2] + T[6];
    if (z < 1e-6) continue;
    const double ex = p2d[i].x - ((Xc[0] + T[4]) / z * K[0] + K[2]), ey = p2d[i].y - ((Xc[1] + T[5]) / z * K[1] + K[3]);
    if (ex * ex + ey * ey <= thr2) { ++cnt; if (out) out->push_back((int)i); }
  }
  return cnt;
}

}  // namespace

void se3_oplus(Pose7& pose, const double u[6]) {      // g2o SE3Quat::exp(u) * pose (se3quat.h:218-260, :99-105)
  const double wx = u[0], wy = u[1], wz = u[2];
  const double theta = std::sqrt(wx * wx + wy * wy + wz * wz);
  const double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
  double O2[9];
  mat3_mul(O, O, O2);
  double a, b, d;
  if (theta < 0.00001) { a = 1.0; b = 0.5; d = 1.0 / 6.0; }
  else { a = std::sin(theta) / theta; b = (1 - std::cos(theta)) / (theta * theta); d = (theta - std::sin(theta)) / (theta * theta * theta); }
  double R[9], V[9];
  for (int i = 0; i < 9; ++i) {
    const double id = (i % 4 == 0) ? 1.0 : 0.0;
    R[i] = id + a * O[i] + b * O2[i];
    V[i] = id + b * O[i] + d * O2[i];
  }
  double qe[4];
  R_to_quat(R, qe);
  const double te[3] = {V[0] * u[3] + V[1] * u[4] + V[2] * u[5], V[3] * u[3] + V[4] * u[4] + V[5] * u[5], V[6] * u[3] + V[7] * u[4] + V[8] * u[5]};
  double rt[3];
  q_rot(qe, pose.data() + 4, rt);
  const double* p = pose.data();
  double x = qe[3] * p[0] + qe[0] * p[3] + qe[1] * p[2] - qe[2] * p[1];
  double y = qe[3] * p[1] + qe[1] * p[3] + qe[2] * p[0] - qe[0] * p[2];
  double z = qe[3] * p[2] + qe[2] * p[3] + qe[0] * p[1] - qe[1] * p[0];
  double w = qe[3] * p[3] - qe[0] * p[0] - qe[1] * p[1] - qe[2] * p[2];
  if (w < 0) { x = -x; y = -y; z = -z; w = -w; }
  const double n = std::sqrt(x * x + y * y + z * z + w * w);
  pose = Pose7{x / n, y / n, z / n, w / n, te[0] + rt[0], te[1] + rt[1], te[2] + rt[2]};
}

bool find_fundamental_ransac(const std::vector<P2f>& from, const std::vector<P2f>& to, double thr_px, double conf,
                             std::vector<uint8_t>& mask, double F[9]) {
  const int n = (int)from.size();
  mask.assign(n, 0);
  if (n < 8) return false;
  const double thr2 = thr_px * thr_px;
  Lcg rng(0x9E3779B97F4A7C15ULL);
  int best = -1, iters = 1000;
  double Fb[9] = {0};
  for (int it = 0; it < iters; ++it) {
    int idx[8];
    sample_distinct(rng, n, 8, idx);
    double Fc[9];
    if (!eight_point(from, to, idx, 8, Fc)) continue;
    const int c = count_f_inliers(from, to, Fc, thr2, nullptr);
    if (c > best) {
      best = c;
      for (int k = 0; k < 9; ++k) Fb[k] = Fc[k];
      iters = std::min(iters, std::max(it + 1, adaptive_iters(conf, (double)c / n, 8, 1000)));
    }
  }
  if (best < 8) return false;
  count_f_inliers(from, to, Fb, thr2, &mask);
  // least-squares refit on the inliers (the mask stays that of the RANSAC model, like OpenCV)
  std::vector<int> in;
  for (int i = 0; i < n; ++i) if (mask[i]) in.push_back(i);
  double Fr[9];
  if ((int)in.size() >= 8 && eight_point(from, to, in.data(), (int)in.size(), Fr)) for (int k = 0; k < 9; ++k) Fb[k] = Fr[k];
  for (int k = 0; k < 9; ++k) F[k] = Fb[k];
  return true;
}

bool solve_pnp_ransac(const std::vector<P3f>& p3d, const std::vector<P2f>& p2d, const double K[4], Pose7& T_c_w,
                      int iterations, double thr_px, double conf, std::vector<int>& inliers) {
  const int n = (int)p3d.size();
  inliers.clear();
  if (n < 4) return false;
  const double thr2 = thr_px * thr_px;
  Lcg rng(0xD1B54A32D192ED03ULL);
  const Pose7 T0 = T_c_w;
  Pose7 best_T = T0;
  int best = count_pnp_inliers(p3d, p2d, K, T0, thr2, nullptr);
  int iters = iterations;
  const int SAMPLE = 5;
  for (int it = 0; it < iters; ++it) {
    int idx[SAMPLE];
    sample_distinct(rng, n, std::min(SAMPLE, n), idx);
    Pose7 T = T0;
    refine_pose(p3d, p2d, K, idx, std::min(SAMPLE, n), T, 8);
    const int c = count_pnp_inliers(p3d, p2d, K, T, thr2, nullptr);
    if (c > best) {
      best = c; best_T = T;
      iters = std::min(iters, std::max(it + 1, adaptive_iters(conf, (double)c / n, SAMPLE, iterations)));
    }
  }
  count_pnp_inliers(p3d, p2d, K, best_T, thr2, &inliers);
  if (inliers.size() >= 4) {
    refine_pose(p3d, p2d, K, inliers.data(), (int)inliers.size(), best_T, 10);
    count_pnp_inliers(p3d, p2d, K, best_T, thr2, &inliers);
  }
  T_c_w = best_T;
  return inliers.size() >= 4;
}

}  // namespace flv
