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
  // The same callback split around the solver call, for callers that batch the solves of several sequences into one
  // launch (flv_localmap_batch): begin() edits the graph and, when a solve is due (returns true), flattens it into `in`;
  // the caller runs flv_ba_optimize on those arrays and hands them back to end().  begin() -> [solve] -> end() must
  // alternate per LocalMap; begin() returning false needs no end().
  struct SolveArrays {
    int P = 0, L = 0, E = 0, fixed = -1;
    std::vector<int64_t> ids;                 // landmark id of landmark vertex l
    std::vector<double> poses, lms, uv;       // [P][7], [L][3], [E][2]
    std::vector<int> ep, el;
    std::vector<uint8_t> active;
  };
  bool begin(const KeyFrameStruct& kf, SolveArrays& in);
  void end(const SolveArrays& solved, const flv_ba_stats& st, CorrectionInfStruct& out);
  void intrinsics(double K4[4]) const { K4[0] = fx_; K4[1] = fy_; K4[2] = cx_; K4[3] = cy_; }
  int window() const { return W_; }
  State state() const { return optimizer_state; }
  const flv_ba_stats& last_stats() const { return stats_; }
  bool solve_failed() const { return solve_failed_; }   // the last frame_callback had a solve due and the solver call failed

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
  bool solve_failed_ = false;
  int reserved_P = 0, reserved_L = 0, reserved_E = 0;

  void remove_pose_vertex(int slot);
  void remove_lm_vertex(int64_t id);
  bool edit_graph(const KeyFrameStruct& kf);          // the state machine up to OPTIMIZING; true = solve now
  void flatten(SolveArrays& in) const;
  bool solve(CorrectionInfStruct& out);
};

}  // namespace flv
