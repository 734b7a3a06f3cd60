// Minimal SE3 / quaternion helpers for the host layer (the reference uses Sophus + Eigen, neither is a
// dependency here).  Poses are T_c_w stored g2o-style: [qx qy qz qw tx ty tz].
#pragma once
#include <array>
#include <cmath>
#include <cstdint>

namespace flv {

using Vec2 = std::array<double, 2>;
using Vec3 = std::array<double, 3>;
using Pose7 = std::array<double, 7>;

inline void quat_to_R(const double* q, double* R) {   // Eigen toRotationMatrix, row-major
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

inline void R_to_quat(const double* m, double* q) {   // Eigen Quaternion(Matrix3)
  double t = m[0] + m[4] + m[8];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t; t = 0.5 / t;
    q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    q[i] = 0.5 * t; t = 0.5 / t;
    q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
    q[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
  }
}

// g2o::SE3Quat(R, t): quaternion from the rotation matrix, then normalizeRotation() (w >= 0, unit norm)
inline Pose7 g2o_pose_from_quat(const Pose7& in) {
  double R[9], q[4];
  quat_to_R(in.data(), R);
  R_to_quat(R, q);
  if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  return Pose7{q[0] / n, q[1] / n, q[2] / n, q[3] / n, in[4], in[5], in[6]};
}

}  // namespace flv
