#include "vi_motion.h"
#include <cmath>

namespace flv {

static inline double norm3(const Vec3& v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
static inline Vec3 sub3(const Vec3& a, const Vec3& b) { return Vec3{a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
static inline Vec3 add3(const Vec3& a, const Vec3& b) { return Vec3{a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
static inline Vec3 mul3(const Vec3& a, double s) { return Vec3{a[0] * s, a[1] * s, a[2] * s}; }

VIMOTION::VIMOTION(SE3 T_i_c_fromCalibration, double magnitude_g_in, double para_1_in, double para_2_in, double para_3_in,
                   double para_4_in, double para_5_in, double para_6_in) {       // vi_motion.cpp:3-32
  T_i_c = T_i_c_fromCalibration;
  T_c_i = T_i_c.inverse();
  init_state.q_w_i = Quat{1, 0, 0, 0};
  magnitude_g = magnitude_g_in;
  gravity = Vec3{0, 0, -magnitude_g};
  para_1 = para_1_in; para_2 = para_2_in; para_3 = para_3_in; para_4 = para_4_in; ba_sat = para_5_in; bw_sat = para_6_in;
}

// the Madgwick gradient step shared by initialisation (:69-98, gain 10*beta) and propagation (:160-187, gain beta)
Quat VIMOTION::madgwick_qdot(const Quat& q_prev, const Vec3& acc, const Vec3& gyro, double gain) const {
  const Quat omega{0, gyro[0], gyro[1], gyro[2]};
  Quat qdot = scalar_multi_q(0.5, q1_multi_q2(q_prev, omega));
  const double acc_norm = norm3(acc);
  if ((acc_norm - magnitude_g) < 0.3) {
    const double ax = acc[0] / acc_norm, ay = acc[1] / acc_norm, az = acc[2] / acc_norm;
    const double qw = q_prev.w, qx = q_prev.x, qy = q_prev.y, qz = q_prev.z;
    double s[4];
    s[0] = 2 * qx * (ay + 2 * qw * qx + 2 * qy * qz) - 2 * qy * (ax - 2 * qw * qy + 2 * qx * qz);
    s[1] = 2 * qw * (ay + 2 * qw * qx + 2 * qy * qz) + 2 * qz * (ax - 2 * qw * qy + 2 * qx * qz) - 4 * qx * (-2 * qx * qx - 2 * qy * qy + az + 1);
    s[2] = 2 * qz * (ay + 2 * qw * qx + 2 * qy * qz) - 2 * qw * (ax - 2 * qw * qy + 2 * qx * qz) - 4 * qy * (-2 * qx * qx - 2 * qy * qy + az + 1);
    s[3] = 2 * qx * (ax - 2 * qw * qy + 2 * qx * qz) + 2 * qy * (ay + 2 * qw * qx + 2 * qy * qz);
    const double sn = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2] + s[3] * s[3]);
    for (double& v : s) v *= sn;                       // `s*=s.norm()` (multiplies; kept)
    qdot.w -= gain * s[0]; qdot.x -= gain * s[1]; qdot.y -= gain * s[2]; qdot.z -= gain * s[3];
  }
  return qdot;
}

void VIMOTION::viIMUinitialization(const IMUSTATE imu_read, Quat& q_w_i, Vec3& pos_w_i, Vec3& vel_w_i) {   // :34-115
  std::lock_guard<std::recursive_mutex> lock(mtx_states_RW);
  q_w_i = Quat{1, 0, 0, 0};
  pos_w_i = vel_w_i = Vec3{0, 0, 0};
  init_state.imu_data = imu_read;
  init_state.pos = pos_w_i; init_state.vel = vel_w_i;
  const Vec3 acc = sub3(imu_read.acc_raw, acc_bias), gyro = sub3(imu_read.gyro_raw, gyro_bias);
  if (is_first_data) {
    if ((norm3(acc) - magnitude_g) < 0.3) {
      const Vec3 rpy{std::atan2(-acc[1], -acc[2]), std::atan2(acc[0], -acc[2]), 0};
      init_state.q_w_i = rpy2Q(rpy);
      states.push_back(init_state);
      if (states.size() >= STATES_QUEUE_SIZE) states.pop_front();
      is_first_data = false;
      q_w_i = rpy2Q(rpy);
    }
  } else {
    const double dt = imu_read.timestamp - states.back().imu_data.timestamp;
    const Quat q_prev = states.back().q_w_i;
    const Quat qdot = madgwick_qdot(q_prev, acc, gyro, 10 * para_1);
    const Quat q_new = q_normalized(q_plus_q(q_prev, scalar_multi_q((float)dt, qdot)));
    init_state.q_w_i = q_new;
    states.push_back(init_state);
    if (states.size() >= STATES_QUEUE_SIZE) states.pop_front();
    if (states.size() > 30) imu_initialized = true;
  }
}

void VIMOTION::viVisiontrigger(Quat& init_orientation) {                                                    // :117-137
  std::lock_guard<std::recursive_mutex> lock(mtx_states_RW);
  MOTION_STATE state = states.back();
  state.pos = Vec3{0, 0, 0}; state.vel = Vec3{0, 0, 0};
  Vec3 rpy = Q2rpy(state.q_w_i);
  rpy[2] = 0;
  state.q_w_i = q_normalized(rpy2Q(rpy));
  states.clear();
  states.push_back(state);
  init_orientation = state.q_w_i;
}

void VIMOTION::viIMUPropagation(const IMUSTATE imu_read, Quat& q_w_i, Vec3& pos_w_i, Vec3& vel_w_i) {      // :139-209
  std::lock_guard<std::recursive_mutex> lock(mtx_states_RW);
  MOTION_STATE s_new;
  const Vec3 acc = sub3(imu_read.acc_raw, acc_bias), gyro = sub3(imu_read.gyro_raw, gyro_bias);
  const MOTION_STATE s_prev = states.back();
  const double dt = imu_read.timestamp - s_prev.imu_data.timestamp;
  const Quat q_prev = s_prev.q_w_i;
  double R[9]; q_to_R(q_prev, R);
  const Quat qdot = madgwick_qdot(q_prev, acc, gyro, para_1);
  s_new.q_w_i = q_normalized(q_plus_q(q_prev, scalar_multi_q((float)dt, qdot)));
  s_new.pos = add3(s_prev.pos, mul3(s_prev.vel, dt));
  const Vec3 Ra{R[0] * acc[0] + R[1] * acc[1] + R[2] * acc[2], R[3] * acc[0] + R[4] * acc[1] + R[5] * acc[2],
                R[6] * acc[0] + R[7] * acc[1] + R[8] * acc[2]};
  s_new.vel = add3(s_prev.vel, mul3(sub3(Ra, gravity), dt));
  s_new.imu_data = imu_read;
  states.push_back(s_new);
  if (states.size() >= STATES_QUEUE_SIZE) states.pop_front();
  q_w_i = s_new.q_w_i; pos_w_i = s_new.pos; vel_w_i = s_new.vel;
}

bool VIMOTION::viFindStateIdx(const double time, int& idx_in_q) {                                           // :348-383
  int idx = 9999;
  for (int i = (int)states.size() - 1; i >= 0; i--) {
    if ((states.at(i).imu_data.timestamp - time) > 0) idx = i;
    else { idx = i; break; }
  }
  if (idx > 0 && idx != 9999) { idx_in_q = idx; return true; }
  return false;
}

void VIMOTION::viCorrectionFromVision(const double t_curr, const SE3 Tcw_curr, const double t_last, const SE3 Tcw_last,
                                      const double /*err*/) {                                               // :212-342
  std::lock_guard<std::recursive_mutex> lock(mtx_states_RW);
  Vec3 acc_bias_est{0, 0, 0}, gyro_bias_est{0, 0, 0};
  int idx_curr, idx_last, idx_mid;
  if (viFindStateIdx(t_last, idx_last) && viFindStateIdx(t_curr, idx_curr)) {
    if (idx_last == idx_curr) return;
    const double dt = t_curr - t_last;
    idx_mid = idx_last + (int)std::floor((idx_curr - idx_last) / 2);
    const SE3 T_w_iA = Tcw_last.inverse() * T_c_i, T_w_iB = Tcw_curr.inverse() * T_c_i;
    const SE3 T_w_ia(states.at(idx_last).q_w_i, states.at(idx_last).pos);
    const SE3 T_w_ib(states.at(idx_curr).q_w_i, states.at(idx_curr).pos);
    const SE3 T_w_im(states.at(idx_mid).q_w_i, states.at(idx_mid).pos);
    const SE3 T_iB_iA = T_w_iB.inverse() * T_w_iA, T_ib_ia = T_w_ib.inverse() * T_w_ia;
    // Eigen Quaternion::inverse() = conjugate / squaredNorm
    const Quat qb = T_ib_ia.q;
    const double n2 = qb.w * qb.w + qb.x * qb.x + qb.y * qb.y + qb.z * qb.z;
    const Quat qbi{qb.w / n2, -qb.x / n2, -qb.y / n2, -qb.z / n2};
    const Quat Q_B_b = q_mul(T_iB_iA.q, qbi);
    gyro_bias_est = Vec3{Q_B_b.x / dt, Q_B_b.y / dt, Q_B_b.z / dt};
    const int cnt = idx_curr - idx_last + 1;
    Vec3 vel_imu{0, 0, 0};
    for (int i = idx_last; i <= idx_curr; i++) vel_imu = add3(vel_imu, states.at(i).vel);
    vel_imu = mul3(vel_imu, 1.0 / cnt);
    const Vec3 dp = sub3(T_w_iB.t, T_w_iA.t);
    const Vec3 vel_vision_world{dp[0] / dt, dp[1] / dt, dp[2] / dt};
    const Vec3 diff_vel_world = sub3(vel_vision_world, vel_imu);
    // (q.inverse().toRotationMatrix()) * v
    const Quat qm = T_w_im.q;
    const double m2 = qm.w * qm.w + qm.x * qm.x + qm.y * qm.y + qm.z * qm.z;
    double Rm[9]; q_to_R(Quat{qm.w / m2, -qm.x / m2, -qm.y / m2, -qm.z / m2}, Rm);
    const Vec3 diff_vel_local{Rm[0] * diff_vel_world[0] + Rm[1] * diff_vel_world[1] + Rm[2] * diff_vel_world[2],
                              Rm[3] * diff_vel_world[0] + Rm[4] * diff_vel_world[1] + Rm[5] * diff_vel_world[2],
                              Rm[6] * diff_vel_world[0] + Rm[7] * diff_vel_world[1] + Rm[8] * diff_vel_world[2]};
    acc_bias_est = Vec3{-diff_vel_local[0] / dt, -diff_vel_local[1] / dt, -diff_vel_local[2] / dt};
    const SE3 T_diff = T_w_iB * T_w_ib.inverse();
    for (size_t i = idx_curr; i < states.size(); i++) {
      const SE3 newT = T_diff * SE3(states.at(i).q_w_i, states.at(i).pos);
      states.at(i).q_w_i = newT.q;
      states.at(i).pos = newT.t;
      states.at(i).vel = add3(states.at(i).vel, diff_vel_world);
    }
    if (std::isnan(acc_bias_est[0])) acc_bias_est = Vec3{0, 0, 0};
    if (std::isnan(gyro_bias_est[0])) gyro_bias_est = Vec3{0, 0, 0};
    const double ba_est_norm = norm3(acc_bias_est);
    if (ba_est_norm > ba_sat) acc_bias_est = mul3(acc_bias_est, ba_sat / ba_est_norm);
    const double bw_est_norm = norm3(gyro_bias_est);
    if (ba_est_norm > bw_sat) gyro_bias_est = mul3(gyro_bias_est, bw_sat / bw_est_norm);   // tests ba_est_norm (:312), kept
    if (dt < 0.1) {
      acc_bias = add3(mul3(acc_bias, 1 - para_3), mul3(acc_bias_est, para_3));
      gyro_bias = add3(mul3(gyro_bias, 1 - para_3), mul3(gyro_bias_est, para_4));          // (1-para_3) on the gyro term (:328), kept
    }
  }
}

bool VIMOTION::viGetIMURollPitchAtTime(const double time, double& roll, double& pitch) {                    // :386-406
  std::lock_guard<std::recursive_mutex> lock(mtx_states_RW);
  int idx;
  if (viFindStateIdx(time, idx)) {
    const SE3 T_w_i(states.at(idx).q_w_i, states.at(idx).pos);
    const Vec3 rpy = Q2rpy(T_w_i.q);
    roll = rpy[0]; pitch = rpy[1];
    return true;
  }
  return false;
}

void VIMOTION::viGetLatestImuState(SE3& T_w_i, Vec3& vel) {
  std::lock_guard<std::recursive_mutex> lock(mtx_states_RW);
  T_w_i = SE3(states.back().q_w_i, states.back().pos);
  vel = states.back().vel;
}

bool VIMOTION::viGetCorrFrameState(const double time, SE3& T_c_w) {                                         // :416-435
  std::lock_guard<std::recursive_mutex> lock(mtx_states_RW);
  int idx;
  if (viFindStateIdx(time, idx)) {
    const SE3 T_w_i(states.at(idx).q_w_i, states.at(idx).pos);
    T_c_w = (T_w_i * T_i_c).inverse();
    return true;
  }
  return false;
}

int VIMOTION::dump_states(double* out, int cap) const {
  std::lock_guard<std::recursive_mutex> lock(mtx_states_RW);
  const int n = (int)states.size() < cap ? (int)states.size() : cap;
  for (int i = 0; i < n; ++i) {
    const MOTION_STATE& s = states[i];
    double* o = out + 11 * (size_t)i;
    o[0] = s.imu_data.timestamp; o[1] = s.q_w_i.w; o[2] = s.q_w_i.x; o[3] = s.q_w_i.y; o[4] = s.q_w_i.z;
    for (int k = 0; k < 3; ++k) { o[5 + k] = s.pos[k]; o[8 + k] = s.vel[k]; }
  }
  return n;
}

void VIMOTION::viVisionRPCompensation(const double time, SE3& T_c_w) {                                       // :437-464
  const SE3 T_w_i_before = T_c_w.inverse() * T_c_i;
  const Vec3 rpy_before = Q2rpy(T_w_i_before.q);
  Vec3 rpy_vimotion{0, 0, 0};
  if (viGetIMURollPitchAtTime(time, rpy_vimotion[0], rpy_vimotion[1])) {
    rpy_vimotion[2] = rpy_before[2];
    const Vec3 rpy_after = add3(mul3(rpy_before, 1 - para_2), mul3(rpy_vimotion, para_2));
    const SE3 T_w_i_after(rpy2Q(rpy_after), T_w_i_before.t);
    T_c_w = (T_w_i_after * T_i_c).inverse();
  }
}

}  // namespace flv
