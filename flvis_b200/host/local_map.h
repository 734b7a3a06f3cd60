// LocalMap -- the sliding-window bundle-adjustment state machine of
// LocalMapNodeletClass::frame_callback (src/backend/vo_localmap.cpp:87-380) without ROS: same states
// (UN_INITIALIZED -> OPTIMIZING <-> SLIDING_WINDOW), same graph edits, same quirks, with the
// g2o SparseOptimizer replaced by flat vertex / edge tables and the solve done by flv_ba_optimize on the GPU.
#pragma once
#include <deque>
#include <map>
#include <vector>
#include "../../include/flvis_b200.h"
#include "poselmbag.h"

namespace flv {

struct KeyFrameStruct {           // src/utils/include/keyframe_msg.h:14-24 minus the images
  int64_t frame_id = 0;
  int lm_count = 0;
  std::vector<int64_t> lm_id;
  std::vector<Vec2> lm_2d;        // undistorted pixel coordinates
  std::vector<Vec3> lm_3d;        // world frame
  Pose7 T_c_w{0, 0, 0, 1, 0, 0, 0};
};

struct CorrectionInfStruct {      // src/utils/include/correction_inf_msg.h:10-18
  int64_t frame_id = 0;
  Pose7 T_c_w{0, 0, 0, 1, 0, 0, 0};
  int lm_count = 0;
  std::vector<int64_t> lm_id;
  std::vector<Vec3> lm_3d;
  int lm_outlier_count = 0;
  std::vector<int64_t> lm_outlier_id;
};

class LocalMap {
 public:
  enum State { UN_INITIALIZED = 0, OPTIMIZING, SLIDING_WINDOW, FAIL };
  LocalMap(flv_ctx* ctx, int window_size, double fx, double fy, double cx, double cy);
  void reset();
  // returns true when a solve ran and `out` was filled (the reference publishes CorrectionInf then)
  bool frame_callback(const KeyFrameStruct& kf, CorrectionInfStruct& out);
  State state() const { return optimizer_state; }
  const flv_ba_stats& last_stats() const { return stats_; }

 private:
  struct Edge { int64_t lm_id; int pose_slot; Vec2 uv; };
  flv_ctx* ctx_;
  int W_;
  double fx_, fy_, cx_, cy_;
  State optimizer_state = UN_INITIALIZED;
  PoseLMBag bag;
  std::deque<KeyFrameStruct> kfs;
  // the "graph": pose vertex id == bag slot; landmark vertex id == lm_id; edges in insertion order
  std::vector<Pose7> pose_est; std::vector<char> pose_present;
  int fixed_slot = -1;
  std::map<int64_t, Vec3> lm_est;
  std::vector<Edge> edges;
  flv_ba_stats stats_{};
  int reserved_P = 0, reserved_L = 0, reserved_E = 0;

  void remove_pose_vertex(int slot);
  void remove_lm_vertex(int64_t id);
  bool solve(CorrectionInfStruct& out);
};

}  // namespace flv
