// SE3 / projection-edge arithmetic shared by the BA kernels (ba.cu: windows of <= 25 poses in shared memory; ba_big.cu: windows
// of up to 100 poses in global memory).  g2o semantics: types/sba/types_six_dof_expmap.{h,cpp}, se3quat.h.  Host + device: the
// big-window solver also runs as a sequential host emulation in the tests.
#pragma once
#include <math.h>

#ifndef BA_HD
#ifdef __CUDACC__
#define BA_HD __host__ __device__ __forceinline__
#define BA_HDN __host__ __device__ inline
#else
#define BA_HD inline
#define BA_HDN inline
#endif
#endif

// ---- SE3 helpers (g2o SE3Quat semantics, quaternion stored x,y,z,w) ----------------------------
BA_HD void q_rotate(const double* q, const double* v, double* o) {
  double ux = q[1] * v[2] - q[2] * v[1], uy = q[2] * v[0] - q[0] * v[2], uz = q[0] * v[1] - q[1] * v[0];
  ux += ux; uy += uy; uz += uz;
  o[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
  o[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
  o[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}
BA_HD void q_to_R(const double* q, double* R) {
  double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
BA_HDN void R_to_q(const double* m, double* q) {
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
BA_HDN void pose_oplus(double* pose, const double* u) {   // pose <- exp(u) * pose  (se3quat.h:218-260, :99-105)
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

struct Cam { double fx, fy, cx, cy; };

// residual r (2), optional A = d r / d point (2x3), B = d r / d pose (2x6).  One fp64 division per edge:
// the reference's x/z, y/z, 1/z, x*y/z^2 ... are evaluated with iz = 1/z (differences ~1 ulp, tolerance-checked).
template <bool JAC>
BA_HD void edge_eval(const double* pose, const double* X, const double* uv, const Cam& c,
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

// W = rho' B^T A of one edge (6x3, w[3 i + j]): recomputed where it is needed instead of being stored -- every pass of this
// kernel waits on memory, not on the fp64 pipe, and 150 flops cost less than nine 16-byte round trips per edge
BA_HD void edge_W(const double* pose, const double* X, const double* uv, const Cam& cam, double delta,
                                       double d2, double* w) {
  double r[2], A[6], B[12];
  edge_eval<true>(pose, X, uv, cam, r, A, B);
  const double c = r[0] * r[0] + r[1] * r[1];
  const double rho1 = (c <= d2) ? 1.0 : delta / sqrt(c);
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) w[3 * i + j] = rho1 * (B[i] * A[j] + B[6 + i] * A[3 + j]);
}

