#include "poselmbag.h"
#include <cstdint>

namespace flv {

PoseLMBag::PoseLMBag(int pose_buffer_size_in) : pose_buffer_size(pose_buffer_size_in) { reset(); }

static constexpr int64_t DEAD = INT64_MIN;

void PoseLMBag::compact() {
  if (n_dead_ == 0) return;
  size_t w = 0;
  for (size_t i = 0; i < lm_sub_bag.size(); ++i)
    if (lm_sub_bag[i].id != DEAD) { if (w != i) lm_sub_bag[w] = lm_sub_bag[i]; index_[lm_sub_bag[w].id] = (int)w; ++w; }
  lm_sub_bag.resize(w);
  n_dead_ = 0;
}

void PoseLMBag::reset() {
  index_.clear(); n_dead_ = 0;
  lm_sub_bag.clear();
  pose_sub_bag.assign(pose_buffer_size, POSE_ITEM{0, 0, Pose7{0, 0, 0, 1, 0, 0, 0}});
  wp_init = 0; pose_cnt_init = 0; pose_sub_bag_initialized = false;
  newest = oldest = 0;
}

bool PoseLMBag::hasTheLM(int64_t id_in, int& idx) {       // poselmbag.cpp:34-46 (ids are unique in the bag)
  idx = 0;
  const auto it = index_.find(id_in);
  if (it == index_.end()) return false;
  idx = it->second;
  return true;
}

bool PoseLMBag::addLMObservationSlidingWindow(int64_t id_in, Vec3 p3d_w_in) {   // :48-67 -- count only
  int idx;
  if (hasTheLM(id_in, idx)) { lm_sub_bag[idx].count++; return false; }
  index_[id_in] = (int)lm_sub_bag.size();
  lm_sub_bag.push_back(LM_ITEM{id_in, 1, p3d_w_in});
  return true;
}

bool PoseLMBag::addLMObservation(int64_t id_in, Vec3 p3d_w_in) {                // :69-91 -- running mean
  int idx;
  if (hasTheLM(id_in, idx)) {
    LM_ITEM& it = lm_sub_bag[idx];
    int cnt = it.count;
    Vec3 p;
    for (int k = 0; k < 3; ++k) p[k] = static_cast<double>(cnt) * it.p3d_w[k] + p3d_w_in[k];
    cnt++;
    for (int k = 0; k < 3; ++k) p[k] = (1.0 / static_cast<double>(cnt)) * p[k];
    it.count = cnt; it.p3d_w = p;
    return false;
  }
  index_[id_in] = (int)lm_sub_bag.size();
  lm_sub_bag.push_back(LM_ITEM{id_in, 1, p3d_w_in});
  return true;
}

bool PoseLMBag::removeLMObservation(int64_t id_in) {                             // :93-108
  int idx;
  if (hasTheLM(id_in, idx)) {
    lm_sub_bag[idx].count--;
    if (lm_sub_bag[idx].count == 0) {
      lm_sub_bag[idx].id = DEAD; index_.erase(id_in); ++n_dead_;
      if (n_dead_ > 1024) compact();
      return true;
    }
  }
  return false;
}

void PoseLMBag::addPose(int64_t id_in, Pose7 pose_in) {                          // :110-136
  if (pose_sub_bag_initialized) {
    newest = oldest;
    pose_sub_bag[newest].relevent_frame_id = id_in;
    pose_sub_bag[newest].pose = pose_in;
    oldest++;
    if (oldest == pose_buffer_size) oldest = 0;
  } else {
    pose_sub_bag[wp_init].relevent_frame_id = id_in;
    pose_sub_bag[wp_init].pose = pose_in;
    pose_sub_bag[wp_init].pose_id = wp_init;
    wp_init++;
    if (wp_init == pose_buffer_size) { pose_sub_bag_initialized = true; oldest = 0; newest = pose_buffer_size - 1; }
  }
}

void PoseLMBag::getAllLMs(std::vector<LM_ITEM>& lms_out) { compact(); lms_out = lm_sub_bag; }

void PoseLMBag::getMultiViewLMs(std::vector<LM_ITEM>& lms_out, int view_cnt) {
  lms_out.clear();
  compact();
  for (const LM_ITEM& lm : lm_sub_bag)
    if (lm.count >= view_cnt) lms_out.push_back(lm);
}

void PoseLMBag::getAllPoses(std::vector<POSE_ITEM>& poses_out) { poses_out = pose_sub_bag; }

int64_t PoseLMBag::getPoseIdByReleventFrameId(int64_t frame_id) {
  for (int i = 0; i < pose_buffer_size; ++i)
    if (pose_sub_bag[i].relevent_frame_id == frame_id) return i;
  return -1;
}

}  // namespace flv
