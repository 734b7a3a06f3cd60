// Lens model helpers of the STEREO_UNRECT path: restatements of the two OpenCV calls FLVIS makes per point,
//   cv::undistortPoints(src, dst, K, D, R, P)      src/processing/lkorb_tracking.cpp:87, src/frontend/f2f_tracking.cpp:301,
//                                                   :425, src/processing/camera_frame.cpp:130
//   cv::projectPoints(pts3d, rvec, tvec, K, D, out) src/processing/lkorb_tracking.cpp:59, camera_frame.cpp:116
// (OpenCV is an external dependency of the reference, not vendored: the arithmetic follows calib3d's
// cvUndistortPointsInternal / cvProjectPoints2Internal for the plumb-bob model with up to 14 coefficients, zero tilt;
// tests/test_undistort_cpu.py pins it to cv2 4.13.0.)  Plain double arithmetic in the order OpenCV uses, float in / out.
#pragma once
#include <cmath>
#ifdef __CUDACC__
#define FLV_HD __host__ __device__
#else
#define FLV_HD
#endif

namespace flv {

struct LensModel {
  double fx = 1, fy = 1, cx = 0, cy = 0;                 // K (raw camera matrix)
  double k[14] = {0};                                    // D: k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 taux tauy (tilt must be 0)
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};             // rectification rotation
  double P[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};   // rectified projection (row-major 3x4); only the 3x3 part is used
};

// cv::undistortPoints with the default criteria (COUNT, 5 iterations)
FLV_HD inline void undistort_point(const LensModel& m, float u, float v, float& uo, float& vo) {
  const double ifx = 1. / m.fx, ify = 1. / m.fy;
  double x = ((double)u - m.cx) * ifx, y = ((double)v - m.cy) * ify;
  const double x0 = x, y0 = y;
  const double* k = m.k;
  for (int j = 0; j < 5; ++j) {
    const double r2 = x * x + y * y;
    const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
    if (icdist < 0) { x = ((double)u - m.cx) * ifx; y = ((double)v - m.cy) * ify; break; }
    const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
    const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
    x = (x0 - deltaX) * icdist;
    y = (y0 - deltaY) * icdist;
  }
  const double xx = m.R[0] * x + m.R[1] * y + m.R[2], yy = m.R[3] * x + m.R[4] * y + m.R[5];
  const double ww = 1. / (m.R[6] * x + m.R[7] * y + m.R[8]);
  x = xx * ww; y = yy * ww;
  uo = (float)(x * m.P[0] + m.P[2]);
  vo = (float)(y * m.P[5] + m.P[6]);
}

// cv::projectPoints for one point: X_c = Rcw X + t, plumb-bob distortion, K
FLV_HD inline void project_point(const LensModel& m, const double Rcw[9], const double t[3], float X, float Y, float Z, float& u, float& v) {
  const double Xc = Rcw[0] * X + Rcw[1] * Y + Rcw[2] * Z + t[0];
  const double Yc = Rcw[3] * X + Rcw[4] * Y + Rcw[5] * Z + t[1];
  double z = Rcw[6] * X + Rcw[7] * Y + Rcw[8] * Z + t[2];
  z = z ? 1. / z : 1;
  const double x = Xc * z, y = Yc * z;
  const double* k = m.k;
  const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
  const double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
  const double cdist = 1 + k[0] * r2 + k[1] * r4 + k[4] * r6;
  const double icdist2 = 1. / (1 + k[5] * r2 + k[6] * r4 + k[7] * r6);
  const double xd = x * cdist * icdist2 + k[2] * a1 + k[3] * a2 + k[8] * r2 + k[9] * r4;
  const double yd = y * cdist * icdist2 + k[2] * a3 + k[3] * a1 + k[10] * r2 + k[11] * r4;
  u = (float)(xd * m.fx + m.cx);
  v = (float)(yd * m.fy + m.cy);
}

}  // namespace flv
