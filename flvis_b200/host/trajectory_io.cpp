// Trajectory text files in the two formats of the reference's recorder node (src/independ_modules/vo_repub_rec.cpp):
//   * process()          :80-101  one line per pose  "stamp x y z qw qx qy qz"  (note: qw FIRST, unlike the TUM benchmark's
//                                  qx qy qz qw), stream precision 6, stamp printed like ros::Time (sec.nsec, 9 digits);
//   * writetokittifile() :103-126 KITTI odometry format, the 3x4 [R | t] row-major, precision 6.
// The recorder subscribes to the pose FLVIS publishes, i.e. the camera pose in the world T_w_c = T_c_w^-1
// (vo_tracking.cpp:432-434 -> rviz path publishers); these writers take the tracker's T_c_w and invert it.
// Used to export both paths' trajectories identically for ATE evaluation.
#include <cmath>
#include <fstream>
#include <iomanip>
#include <new>
#include <sstream>
#include <string>
#include "../../include/flvis_b200_host.h"
#include "sophus_lite.h"

struct flv_traj { std::ofstream fd; int format; };

extern "C" {

flv_traj* flv_traj_open(const char* path, int format) {
  if (!path || (format != 0 && format != 1)) return nullptr;
  flv_traj* t = new (std::nothrow) flv_traj();
  if (!t) return nullptr;
  t->format = format;
  t->fd.open(path);
  if (!t->fd.is_open()) { delete t; return nullptr; }
  return t;
}

int flv_traj_write(flv_traj* t, double stamp, const double* T_c_w) {
  if (!t || !T_c_w) return FLV_ERR_INVALID;
  const flv::SE3 Tcw(flv::Quat{T_c_w[3], T_c_w[0], T_c_w[1], T_c_w[2]}, flv::Vec3{T_c_w[4], T_c_w[5], T_c_w[6]});
  const flv::SE3 Twc = Tcw.inverse();
  std::ofstream& fd = t->fd;
  if (t->format == 0) {
    // ros::Time operator<< : sec "." nsec (9 digits, zero padded)
    const double fl = std::floor(stamp);
    long long sec = (long long)fl, nsec = (long long)std::llround((stamp - fl) * 1e9);
    if (nsec >= 1000000000LL) { sec += 1; nsec -= 1000000000LL; }
    std::ostringstream ts;
    ts << sec << "." << std::setw(9) << std::setfill('0') << nsec;
    fd << std::setprecision(6) << ts.str() << " " << std::setprecision(6) << Twc.t[0] << " " << Twc.t[1] << " " << Twc.t[2] << " "
       << Twc.q.w << " " << Twc.q.x << " " << Twc.q.y << " " << Twc.q.z << std::endl;
  } else {
    double R[9];
    flv::q_to_R(Twc.q, R);
    fd << std::setprecision(6) << R[0] << " " << R[1] << " " << R[2] << " " << Twc.t[0] << " ";
    fd << std::setprecision(6) << R[3] << " " << R[4] << " " << R[5] << " " << Twc.t[1] << " ";
    fd << std::setprecision(6) << R[6] << " " << R[7] << " " << R[8] << " " << Twc.t[2] << std::endl;
  }
  return fd.good() ? FLV_OK : FLV_ERR_INVALID;
}

void flv_traj_close(flv_traj* t) {
  if (!t) return;
  t->fd.close();
  delete t;
}

}  // extern "C"
