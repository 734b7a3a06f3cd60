// Quaternion / SE3 with the semantics of the vendored (old, non-templated) Sophus + Eigen the reference uses
// (3rdPartLib/Sophus/sophus/so3.cpp:36-90, se3.cpp:36-86): SO3(q) normalises, products normalise, rotation is
// Eigen's _transformVector, inverse is the conjugate.  Quaternions are stored w,x,y,z like Eigen's ctor order.
#pragma once
#include <cmath>
#include "se3.h"

namespace flv {

struct Quat { double w = 1, x = 0, y = 0, z = 0; };

inline Quat q_normalized(Quat q) {
  const double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  return Quat{q.w / n, q.x / n, q.y / n, q.z / n};
}
inline Quat q_mul(const Quat& a, const Quat& b) {       // Eigen operator*
  return Quat{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
inline Quat q_conj(const Quat& q) { return Quat{q.w, -q.x, -q.y, -q.z}; }
inline Vec3 q_rot(const Quat& q, const Vec3& v) {       // Eigen _transformVector
  double ux = q.y * v[2] - q.z * v[1], uy = q.z * v[0] - q.x * v[2], uz = q.x * v[1] - q.y * v[0];
  ux += ux; uy += uy; uz += uz;
  return Vec3{v[0] + q.w * ux + (q.y * uz - q.z * uy), v[1] + q.w * uy + (q.z * ux - q.x * uz),
              v[2] + q.w * uz + (q.x * uy - q.y * ux)};
}
inline void q_to_R(const Quat& q, double* R) { const double a[4] = {q.x, q.y, q.z, q.w}; quat_to_R(a, R); }
inline Quat R_to_q(const double* R) { double a[4]; R_to_quat(R, a); return Quat{a[3], a[0], a[1], a[2]}; }

struct SE3 {
  Quat q; Vec3 t{0, 0, 0};
  SE3() {}
  SE3(const Quat& q_in, const Vec3& t_in) : q(q_normalized(q_in)), t(t_in) {}
  SE3 operator*(const SE3& o) const {
    SE3 r;
    const Vec3 rt = q_rot(q, o.t);
    r.t = Vec3{t[0] + rt[0], t[1] + rt[1], t[2] + rt[2]};
    r.q = q_normalized(q_mul(q, o.q));
    return r;
  }
  SE3 inverse() const {
    SE3 r;
    r.q = q_normalized(q_conj(q));
    r.t = q_rot(r.q, Vec3{-t[0], -t[1], -t[2]});
    return r;
  }
};

// kinetic_math.h:17-94 (ENU, R = Rz*Ry*Rx)
inline void rpy2R(const Vec3& rpy, double* R) {
  const double r = rpy[0], p = rpy[1], y = rpy[2];
  const double cy = std::cos(y), sy = std::sin(y), cp = std::cos(p), sp = std::sin(p), cr = std::cos(r), sr = std::sin(r);
  R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
  R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
  R[6] = -sp; R[7] = cp * sr; R[8] = cp * cr;
}
inline Vec3 R2rpy(const double* R) {
  return Vec3{std::atan2(R[7], R[8]), std::atan2(-R[6], std::sqrt(R[7] * R[7] + R[8] * R[8])), std::atan2(R[3], R[0])};
}
inline Quat rpy2Q(const Vec3& rpy) { double R[9]; rpy2R(rpy, R); return R_to_q(R); }
inline Vec3 Q2rpy(const Quat& q) { double R[9]; q_to_R(q, R); return R2rpy(R); }
// kinetic_math.h:112-141
inline Quat q1_multi_q2(const Quat& q1, const Quat& q2) {
  return Quat{q2.w * q1.w - q2.x * q1.x - q2.y * q1.y - q2.z * q1.z, q2.x * q1.w + q2.w * q1.x + q2.z * q1.y - q2.y * q1.z,
              q2.y * q1.w - q2.z * q1.x + q2.w * q1.y + q2.x * q1.z, q2.z * q1.w + q2.y * q1.x - q2.x * q1.y + q2.w * q1.z};
}
inline Quat scalar_multi_q(const float a, const Quat& b) { return Quat{a * b.w, a * b.x, a * b.y, a * b.z}; }   // float scalar, as the reference
inline Quat q_plus_q(const Quat& a, const Quat& b) { return Quat{a.w + b.w, a.x + b.x, a.y + b.y, a.z + b.z}; }

}  // namespace flv
