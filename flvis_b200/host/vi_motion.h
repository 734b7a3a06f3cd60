// VIMOTION -- IMU attitude filter / dead reckoning / vision-driven bias feedback.
// Same public surface and arithmetic as the reference's src/processing/include/vi_motion.h and
// src/processing/vi_motion.cpp:3-464 (Madgwick gradient step with `s *= s.norm()`, float-cast scalar in
// scalar_multi_q, forward-Euler p/v, 400-sample state queue, gyro clamp testing ba_est_norm, para_3 on both
// decay terms).  Sequential, a few hundred flops per 200 Hz sample: it stays on the host.
#pragma once
#include <deque>
#include <mutex>
#include "sophus_lite.h"

namespace flv {

constexpr size_t STATES_QUEUE_SIZE = 400;

struct IMUSTATE { Vec3 acc_raw{0, 0, 0}, gyro_raw{0, 0, 0}; double timestamp = 0; };
struct MOTION_STATE { Vec3 pos{0, 0, 0}, vel{0, 0, 0}; Quat q_w_i; IMUSTATE imu_data; };

class VIMOTION {
 public:
  bool imu_initialized = false, is_first_data = true;
  SE3 T_i_c, T_c_i;
  Vec3 acc_bias{0, 0, 0}, gyro_bias{0, 0, 0};
  std::deque<MOTION_STATE> states;
  MOTION_STATE init_state;
  double magnitude_g; Vec3 gravity;
  double para_1, para_2, para_3, para_4, ba_sat, bw_sat;

  VIMOTION(SE3 T_i_c_fromCalibration, double magnitude_g_in = 9.81, double para_1_in = 0.1, double para_2_in = 0.05,
           double para_3_in = 0.01, double para_4_in = 0.01, double para_5_in = 0.5, double para_6_in = 0.1);
  void viIMUinitialization(const IMUSTATE imu_read, Quat& q_w_i, Vec3& pos_w_i, Vec3& vel_w_i);
  void viVisiontrigger(Quat& init_orientation);
  void viIMUPropagation(const IMUSTATE imu_read, Quat& q_w_i, Vec3& pos_w_i, Vec3& vel_w_i);
  void viCorrectionFromVision(const double t_curr, const SE3 Tcw_curr, const double t_last, const SE3 Tcw_last, const double err);
  bool viFindStateIdx(const double time, int& idx_in_q);
  bool viGetIMURollPitchAtTime(const double time, double& roll, double& pitch);
  void viGetLatestImuState(SE3& T_w_i, Vec3& vel);
  bool viGetCorrFrameState(const double time, SE3& T_c_w);
  void viVisionRPCompensation(const double time, SE3& T_c_w);
  int dump_states(double* out11, int cap) const;     // {t, q wxyz, pos, vel} per queued state (tests / diagnostics)

 private:
  // imu_feed runs on other threads than image_feed (ROS callback threads in the reference): `states`, the biases and
  // the init flags are guarded like the reference's mtx_states_RW (vi_motion.cpp:119-131, :150-205, :220-339, :390-433);
  // the reference leaves viIMUinitialization unguarded (:34-115), here it takes the lock too.  Recursive because
  // viVisionRPCompensation -> viGetIMURollPitchAtTime nests.
  mutable std::recursive_mutex mtx_states_RW;
  Quat madgwick_qdot(const Quat& q_prev, const Vec3& acc, const Vec3& gyro, double gain) const;
};

}  // namespace flv
