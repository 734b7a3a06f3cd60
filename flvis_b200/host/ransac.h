// Host stand-ins for the two OpenCV RANSAC calls of LKORBTracking::tracking
// (src/processing/lkorb_tracking.cpp:134-135 cv::findFundamentalMat(FM_RANSAC, 5.0, 0.99) and :170-177
// cv::solvePnPRansac(100 it, 3.0 px, 0.99)).  SURVEY.md 8(f): these stay on the host for now ("next" row).
// They are written from the published algorithms (normalised 8-point, symmetric epipolar distance; RANSAC over
// small samples with Gauss-Newton pose refinement) and are NOT bit-compatible with OpenCV's internal RNG / solver
// choices: parity tests inject OpenCV's masks through the callbacks of LKORBTracking instead (tests/test_pipeline*).
#pragma once
#include <cstdint>
#include <vector>
#include "se3.h"

namespace flv {

struct P2f { float x, y; };
struct P3f { float x, y, z; };

// returns false if fewer than 8 correspondences; mask[i] = 1 for inliers of the best model
bool find_fundamental_ransac(const std::vector<P2f>& from, const std::vector<P2f>& to, double thr_px, double conf,
                             std::vector<uint8_t>& mask, double F[9]);

// T_c_w: in = initial pose (IMU guess or the previous frame's pose), out = refined pose.  inliers = indices.
bool solve_pnp_ransac(const std::vector<P3f>& p3d, const std::vector<P2f>& p2d, const double K[4], Pose7& T_c_w,
                      int iterations, double thr_px, double conf, std::vector<int>& inliers);

// g2o-style pose update used by the refinement: pose <- exp(u) * pose, u = [omega, upsilon]
void se3_oplus(Pose7& pose, const double u[6]);

}  // namespace flv
