// PoseLMBag -- ring buffer of window poses + landmark reference counts.
// Same public surface and semantics as the reference's src/backend/include/poselmbag.h:24-63 /
// src/backend/poselmbag.cpp:5-208 (slot index == optimizer vertex id; a new pose overwrites the oldest slot;
// addLMObservation keeps a running mean, the sliding variant only counts).
#pragma once
#include <unordered_map>
#include <vector>
#include "se3.h"

namespace flv {

struct LM_ITEM { int64_t id; int count; Vec3 p3d_w; };
struct POSE_ITEM { int64_t relevent_frame_id; int64_t pose_id; Pose7 pose; };

class PoseLMBag {
 public:
  std::vector<LM_ITEM> lm_sub_bag;
  std::vector<POSE_ITEM> pose_sub_bag;
  int pose_buffer_size;
  int newest = 0, oldest = 0, wp_init = 0, pose_cnt_init = 0;
  bool pose_sub_bag_initialized = false;

  explicit PoseLMBag(int pose_buffer_size_in);
  void reset();
  bool hasTheLM(int64_t id_in, int& idx);
  bool addLMObservation(int64_t id_in, Vec3 p3d_w_in);
  bool addLMObservationSlidingWindow(int64_t id_in, Vec3 p3d_w_in);
  bool removeLMObservation(int64_t id_in);
  void addPose(int64_t id_in, Pose7 pose_in);
  void getAllLMs(std::vector<LM_ITEM>& lms_out);
  void getMultiViewLMs(std::vector<LM_ITEM>& lms_out, int view_cnt = 3);
  void getAllPoses(std::vector<POSE_ITEM>& poses_out);
  int getNewestPoseInOptimizerIdx() { return newest; }
  int getOldestPoseInOptimizerIdx() { return oldest; }
  int64_t getPoseIdByReleventFrameId(int64_t frame_id);

 private:
  // The reference searches lm_sub_bag linearly (poselmbag.cpp:34-46) and erases from the middle of the vector: O(N L) per
  // keyframe.  Same contents and order here, but the lookup goes through an id -> position index and an erased entry is a
  // tombstone until the next compaction (relative order of the live entries, which getMultiViewLMs exposes, is unchanged).
  std::unordered_map<int64_t, int> index_;
  int n_dead_ = 0;
  void compact();
};

}  // namespace flv
