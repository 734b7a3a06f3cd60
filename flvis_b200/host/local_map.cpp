#include "local_map.h"
#include <algorithm>
#include <set>

namespace flv {

LocalMap::LocalMap(flv_ctx* ctx, int window_size, double fx, double fy, double cx, double cy)
    : ctx_(ctx), W_(window_size), fx_(fx), fy_(fy), cx_(cx), cy_(cy), bag(window_size) {
  reset();
}

void LocalMap::reset() {                         // KFMSG_CMD_RESET_LM branch, vo_localmap.cpp:89-98
  optimizer_state = UN_INITIALIZED;
  bag.reset();
  kfs.clear();
  pose_est.assign(W_, Pose7{0, 0, 0, 1, 0, 0, 0});
  pose_present.assign(W_, 0);
  fixed_slot = -1;
  lm_est.clear();
  edges.clear();
}

void LocalMap::remove_pose_vertex(int slot) {    // g2o removeVertex deletes the incident edges (hyper_graph.cpp:214-220)
  pose_present[slot] = 0;
  if (fixed_slot == slot) fixed_slot = -1;
  edges.erase(std::remove_if(edges.begin(), edges.end(), [slot](const Edge& e) { return e.pose_slot == slot; }),
              edges.end());
}

void LocalMap::remove_lm_vertex(int64_t id) {
  lm_est.erase(id);
  edges.erase(std::remove_if(edges.begin(), edges.end(), [id](const Edge& e) { return e.lm_id == id; }), edges.end());
}

bool LocalMap::edit_graph(const KeyFrameStruct& kf) {
  kfs.push_back(kf);
  switch (optimizer_state) {
    case OPTIMIZING: break;
    case UN_INITIALIZED:
      if ((int)kfs.size() >= W_) {                                     // vo_localmap.cpp:125-209
        for (int f = 0; f < W_; ++f) {
          bag.addPose(kfs[f].frame_id, kfs[f].T_c_w);
          for (int i = 0; i < kfs[f].lm_count; ++i) bag.addLMObservation(kfs[f].lm_id[i], kfs[f].lm_3d[i]);
        }
        const int oldest = bag.getOldestPoseInOptimizerIdx();
        std::vector<POSE_ITEM> poses; bag.getAllPoses(poses);
        for (const POSE_ITEM& it : poses) {
          pose_est[it.pose_id] = g2o_pose_from_quat(it.pose);
          pose_present[it.pose_id] = 1;
          if (it.pose_id == oldest) fixed_slot = (int)it.pose_id;
        }
        std::vector<LM_ITEM> lms; bag.getAllLMs(lms);
        for (const LM_ITEM& it : lms) lm_est[it.id] = it.p3d_w;
        edges.clear();
        for (int f = 0; f < W_; ++f) {
          const int slot = (int)bag.getPoseIdByReleventFrameId(kfs[f].frame_id);
          for (int i = 0; i < kfs[f].lm_count; ++i) edges.push_back(Edge{kfs[f].lm_id[i], slot, kfs[f].lm_2d[i]});
        }
        optimizer_state = OPTIMIZING;
      } else {
        return false;                                                  // :211-214 (no pop_front)
      }
      break;
    case SLIDING_WINDOW: {                                             // :218-284
      remove_pose_vertex(bag.getOldestPoseInOptimizerIdx());
      {                                                                // off-by-one keyframe on purpose (:226-232)
        std::set<int64_t> gone;                                        // removeVertex per landmark == one pass over the edges
        for (int64_t id : kfs.at(0).lm_id)
          if (bag.removeLMObservation(id)) { gone.insert(id); lm_est.erase(id); }
        if (!gone.empty())
          edges.erase(std::remove_if(edges.begin(), edges.end(), [&gone](const Edge& e) { return gone.count(e.lm_id) != 0; }), edges.end());
      }
      bag.addPose(kfs.back().frame_id, kfs.back().T_c_w);
      const int newest = bag.getNewestPoseInOptimizerIdx();
      pose_est[newest] = g2o_pose_from_quat(kfs.back().T_c_w);
      pose_present[newest] = 1;
      fixed_slot = bag.getOldestPoseInOptimizerIdx();                  // :241
      for (int i = 0; i < kfs.back().lm_count; ++i)
        if (bag.addLMObservationSlidingWindow(kfs.back().lm_id[i], kfs.back().lm_3d[i]))
          lm_est[kfs.back().lm_id[i]] = kfs.back().lm_3d[i];
      for (int i = 0; i < kfs.back().lm_count; ++i) {
        // g2o refuses an edge whose landmark vertex is missing (setVertex(nullptr) then addEdge fails)
        if (lm_est.count(kfs.back().lm_id[i])) edges.push_back(Edge{kfs.back().lm_id[i], newest, kfs.back().lm_2d[i]});
      }
      optimizer_state = OPTIMIZING;
    } break;
    default: break;
  }
  return optimizer_state == OPTIMIZING;
}

bool LocalMap::frame_callback(const KeyFrameStruct& kf, CorrectionInfStruct& out) {
  bool solved = false;
  solve_failed_ = false;
  if (!edit_graph(kf)) return false;
  solved = solve(out);
  solve_failed_ = !solved;
  optimizer_state = SLIDING_WINDOW;
  kfs.pop_front();                                                     // :379
  return solved;
}

bool LocalMap::begin(const KeyFrameStruct& kf, SolveArrays& in) {
  if (!edit_graph(kf)) return false;
  flatten(in);
  return true;
}

// flatten: poses by slot id, landmarks by id (g2o orders vertices by id), edges in insertion order
void LocalMap::flatten(SolveArrays& in) const {
  const int P = W_;
  in.ids.clear(); in.ids.reserve(lm_est.size());
  in.lms.clear(); in.lms.reserve(3 * lm_est.size());
  for (const auto& kv : lm_est) { in.ids.push_back(kv.first); in.lms.insert(in.lms.end(), kv.second.begin(), kv.second.end()); }
  const int L = (int)in.ids.size(), E = (int)edges.size();
  in.P = P; in.L = L; in.E = E; in.fixed = fixed_slot;
  in.poses.assign(7 * (size_t)P, 0.0); in.uv.assign(2 * (size_t)E, 0.0);
  for (int p = 0; p < P; ++p) std::copy(pose_est[p].begin(), pose_est[p].end(), in.poses.begin() + 7 * p);
  in.ep.assign(E, 0); in.el.assign(E, 0);
  in.active.assign(E, 1);
  for (int e = 0; e < E; ++e) {
    in.ep[e] = edges[e].pose_slot;
    in.el[e] = (int)(std::lower_bound(in.ids.begin(), in.ids.end(), edges[e].lm_id) - in.ids.begin());
    in.uv[2 * e] = edges[e].uv[0]; in.uv[2 * e + 1] = edges[e].uv[1];
  }
}

void LocalMap::end(const SolveArrays& r, const flv_ba_stats& st, CorrectionInfStruct& out) {
  stats_ = st;
  const int P = r.P, L = r.L, E = r.E;
  // estimates persist in the graph
  for (int p = 0; p < P; ++p) std::copy(r.poses.begin() + 7 * p, r.poses.begin() + 7 * p + 7, pose_est[p].begin());
  for (int l = 0; l < L; ++l) lm_est[r.ids[l]] = Vec3{r.lms[3 * l], r.lms[3 * l + 1], r.lms[3 * l + 2]};
  // culled edges leave the graph for good; the reference walks `edges` from the back (:303-316)
  out = CorrectionInfStruct();
  for (int e = E - 1; e >= 0; --e)
    if (!r.active[e]) { out.lm_outlier_id.push_back(edges[e].lm_id); }
  out.lm_outlier_count = (int)out.lm_outlier_id.size();
  {
    std::vector<Edge> kept; kept.reserve(E);
    for (int e = 0; e < E; ++e) if (r.active[e]) kept.push_back(edges[e]);
    edges.swap(kept);
  }
  out.frame_id = kfs.back().frame_id;
  out.T_c_w = pose_est[bag.getNewestPoseInOptimizerIdx()];
  std::vector<LM_ITEM> mv; bag.getMultiViewLMs(mv, 4);                 // :329-357
  out.lm_count = (int)mv.size();
  for (const LM_ITEM& lm : mv) { out.lm_id.push_back(lm.id); out.lm_3d.push_back(lm_est[lm.id]); }
  optimizer_state = SLIDING_WINDOW;
  kfs.pop_front();                                                     // :379
}

bool LocalMap::solve(CorrectionInfStruct& out) {                       // vo_localmap.cpp:292-366
  SolveArrays in;
  flatten(in);
  const int P = in.P, L = in.L, E = in.E;
  if (P > reserved_P || L > reserved_L || E > reserved_E) {
    reserved_P = std::max(P, reserved_P); reserved_L = std::max(L + L / 2 + 64, reserved_L);
    reserved_E = std::max(E + E / 2 + 64, reserved_E);
    if (flv_ba_reserve(ctx_, reserved_P, reserved_L, reserved_E) != FLV_OK) return false;
  }
  // the C ABI takes per-stream strides = the reserved sizes
  std::vector<double> poses_s(7 * (size_t)reserved_P, 0.0), lms_s(3 * (size_t)reserved_L, 0.0), uv_s(2 * (size_t)reserved_E, 0.0);
  std::vector<int> ep_s(reserved_E, 0), el_s(reserved_E, 0);
  std::vector<uint8_t> act_s(reserved_E, 0);
  std::copy(in.poses.begin(), in.poses.end(), poses_s.begin()); std::copy(in.lms.begin(), in.lms.end(), lms_s.begin());
  std::copy(in.uv.begin(), in.uv.end(), uv_s.begin()); std::copy(in.ep.begin(), in.ep.end(), ep_s.begin());
  std::copy(in.el.begin(), in.el.end(), el_s.begin()); std::copy(in.active.begin(), in.active.end(), act_s.begin());
  flv_ba_problem pb{P, L, E, in.fixed, 0, fx_, fy_, cx_, cy_};
  flv_ba_params prm{12, 8, 1.0, 3.0, 0, 0};
  flv_ba_stats st{};
  if (flv_ba_optimize(ctx_, 1, &pb, &prm, poses_s.data(), lms_s.data(), ep_s.data(), el_s.data(), uv_s.data(),
                      act_s.data(), &st, FLV_MEM_HOST) != FLV_OK)
    return false;
  std::copy(poses_s.begin(), poses_s.begin() + 7 * (size_t)P, in.poses.begin());
  std::copy(lms_s.begin(), lms_s.begin() + 3 * (size_t)L, in.lms.begin());
  std::copy(act_s.begin(), act_s.begin() + E, in.active.begin());
  // end() without its state / queue epilogue (frame_callback does that for this path)
  stats_ = st;
  for (int p = 0; p < P; ++p) std::copy(in.poses.begin() + 7 * p, in.poses.begin() + 7 * p + 7, pose_est[p].begin());
  for (int l = 0; l < L; ++l) lm_est[in.ids[l]] = Vec3{in.lms[3 * l], in.lms[3 * l + 1], in.lms[3 * l + 2]};
  out = CorrectionInfStruct();
  for (int e = E - 1; e >= 0; --e)
    if (!in.active[e]) { out.lm_outlier_id.push_back(edges[e].lm_id); }
  out.lm_outlier_count = (int)out.lm_outlier_id.size();
  {
    std::vector<Edge> kept; kept.reserve(E);
    for (int e = 0; e < E; ++e) if (in.active[e]) kept.push_back(edges[e]);
    edges.swap(kept);
  }
  out.frame_id = kfs.back().frame_id;
  out.T_c_w = pose_est[bag.getNewestPoseInOptimizerIdx()];
  std::vector<LM_ITEM> mv; bag.getMultiViewLMs(mv, 4);                 // :329-357
  out.lm_count = (int)mv.size();
  for (const LM_ITEM& lm : mv) { out.lm_id.push_back(lm.id); out.lm_3d.push_back(lm_est[lm.id]); }
  return true;
}

}  // namespace flv
