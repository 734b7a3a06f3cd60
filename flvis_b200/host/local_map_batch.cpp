// flv_localmap_batch -- the local-map threads for S sequences (LocalMapNodeletClass, src/backend/vo_localmap.cpp:87-380).
//
// FLVIS runs the sliding-window BA in its own nodelet thread: keyframes arrive through a queue (depth 10,
// vo_localmap.cpp:464-467) and never block tracking.  Here ONE worker thread serves all S sequences: the tracker thread
// submits the keyframes of a frame, the worker edits each sequence's graph (flv::LocalMap::begin, the reference's state
// machine with its quirks), solves ALL windows that are due with ONE flv_ba_optimize launch on its own context / CUDA
// stream (so the kernels overlap the tracker's), and stores the CorrectionInf results (LocalMap::end).
// Results are identical to S separate flv_localmap handles fed the same keyframes (tests/test_localmap.py).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#include "../../include/flvis_b200_host.h"
#include "local_map.h"

namespace {
struct KfMsg { int stream; flv::KeyFrameStruct kf; };

// The graph editing of different sequences is independent: a few helper threads share the loop over a submission's
// keyframes (the solver launch itself stays one call).  Persistent threads, work handed out through an atomic counter.
class Helpers {
 public:
  explicit Helpers(int n) {
    for (int i = 0; i < n; ++i) th_.emplace_back([this] { loop(); });
  }
  ~Helpers() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; ++gen_; }
    cv_.notify_all();
    for (std::thread& t : th_) t.join();
  }
  template <class F>
  void parallel_for(int n, F&& f) {
    if (n <= 1 || th_.empty()) { for (int i = 0; i < n; ++i) f(i); return; }
    fn_ = [&f](int i) { f(i); };
    n_ = n; next_.store(0); done_.store(0);
    { std::lock_guard<std::mutex> lk(mu_); ++gen_; }
    cv_.notify_all();
    work();                                     // the caller takes part
    while (done_.load(std::memory_order_acquire) < n_) std::this_thread::yield();
  }

 private:
  void work() {
    for (;;) {
      const int i = next_.fetch_add(1);
      if (i >= n_) break;
      fn_(i);
      done_.fetch_add(1, std::memory_order_release);
    }
  }
  void loop() {
    unsigned long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
      }
      work();
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_;
  unsigned long gen_ = 0;
  bool stop_ = false;
  std::function<void(int)> fn_;
  int n_ = 0;
  std::atomic<int> next_{0}, done_{0};
};
}

struct LmShard {
  int device = 0, S = 0, W = 0;
  double K[4] = {0, 0, 0, 0};
  flv_ctx* ctx = nullptr;                       // created and used by the worker thread only
  std::vector<std::unique_ptr<flv::LocalMap>> maps;
  std::vector<flv::CorrectionInfStruct> latest; // last CorrectionInf per sequence
  std::vector<long long> n_results;
  std::thread worker;
  std::mutex mu;
  std::condition_variable cv_work, cv_idle;
  std::deque<std::vector<KfMsg>> queue;
  bool stop = false, busy = false, failed = false, started = false;
  long long n_keyframes = 0, n_solves = 0, n_launches = 0;
  double solve_ms = 0, host_ms = 0;            // solver calls / graph editing + packing on the worker thread
  double sum_edges = 0, sum_lms = 0, sum_poses = 0, sum_iters = 0;   // problem sizes / LM iterations of the solved windows
  int rP = 0, rL = 0, rE = 0;
  std::unique_ptr<Helpers> helpers;
  std::vector<double> poses, lms, uv; std::vector<int> ep, el; std::vector<uint8_t> act;
  char err[256] = {0};

  void run();
  bool process(std::vector<KfMsg>& batch);
};

bool LmShard::process(std::vector<KfMsg>& batch) {
  // a sequence may contribute several keyframes to one submission only if the caller batches frames; solve in rounds so
  // that every LocalMap sees begin -> solve -> end in order
  size_t done = 0;
  const auto th0 = std::chrono::steady_clock::now();
  const double solve_before = solve_ms;
  std::vector<char> used(batch.size(), 0);
  while (done < batch.size()) {
    std::vector<int> round;                      // indices of this round: at most one keyframe per sequence
    std::vector<char> seen(S, 0);
    for (size_t i = 0; i < batch.size(); ++i)
      if (!used[i] && !seen[batch[i].stream]) { seen[batch[i].stream] = 1; round.push_back((int)i); used[i] = 1; }
    done += round.size();
    std::vector<flv::LocalMap::SolveArrays> arr(round.size());
    std::vector<char> is_due(round.size(), 0);
    helpers->parallel_for((int)round.size(), [&](int r) {
      const KfMsg& m = batch[round[r]];
      is_due[r] = maps[m.stream]->begin(m.kf, arr[r]) ? 1 : 0;
    });
    std::vector<int> due;
    for (size_t r = 0; r < round.size(); ++r) if (is_due[r]) due.push_back((int)r);
    n_keyframes += (long long)round.size();
    if (due.empty()) continue;
    int P = W, L = 0, E = 0;
    for (int r : due) { L = std::max(L, arr[r].L); E = std::max(E, arr[r].E); }
    if (P > rP || L > rL || E > rE) {
      rP = std::max(P, rP); rL = std::max(L + L / 2 + 64, rL); rE = std::max(E + E / 2 + 64, rE);
      if (flv_ba_reserve(ctx, rP, rL, rE) != FLV_OK) { snprintf(err, sizeof(err), "flv_ba_reserve: %s", flv_last_error(ctx)); return false; }
    }
    const size_t n = due.size();
    // launch arrays persist across submissions (strides = the reserved sizes; entries past a window's own counts are never read)
    if (poses.size() < n * rP * 7) poses.resize(n * rP * 7);
    if (lms.size() < n * (size_t)rL * 3) lms.resize(n * (size_t)rL * 3);
    if (uv.size() < n * (size_t)rE * 2) uv.resize(n * (size_t)rE * 2);
    if (ep.size() < n * (size_t)rE) { ep.resize(n * (size_t)rE); el.resize(n * (size_t)rE); act.resize(n * (size_t)rE); }
    std::vector<flv_ba_problem> pb(n);
    std::vector<flv_ba_stats> st(n);
    helpers->parallel_for((int)n, [&](int jj) {
      const size_t j = (size_t)jj;
      const flv::LocalMap::SolveArrays& a = arr[due[j]];
      std::copy(a.poses.begin(), a.poses.end(), poses.begin() + j * rP * 7);
      std::copy(a.lms.begin(), a.lms.end(), lms.begin() + j * (size_t)rL * 3);
      std::copy(a.uv.begin(), a.uv.end(), uv.begin() + j * (size_t)rE * 2);
      std::copy(a.ep.begin(), a.ep.end(), ep.begin() + j * (size_t)rE);
      std::copy(a.el.begin(), a.el.end(), el.begin() + j * (size_t)rE);
      std::copy(a.active.begin(), a.active.end(), act.begin() + j * (size_t)rE);
      pb[j] = flv_ba_problem{a.P, a.L, a.E, a.fixed, 0, K[0], K[1], K[2], K[3]};
    });
    const flv_ba_params prm{12, 8, 1.0, 3.0, 0, 0};
    const auto t0 = std::chrono::steady_clock::now();
    if (flv_ba_optimize(ctx, (int)n, pb.data(), &prm, poses.data(), lms.data(), ep.data(), el.data(), uv.data(), act.data(), st.data(),
                        FLV_MEM_HOST) != FLV_OK) {
      snprintf(err, sizeof(err), "flv_ba_optimize: %s", flv_last_error(ctx));
      return false;
    }
    solve_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    n_solves += (long long)n; n_launches++;
    for (size_t j = 0; j < n; ++j) { sum_edges += pb[j].n_edges; sum_lms += pb[j].n_landmarks; sum_poses += pb[j].n_poses; sum_iters += st[j].iterations_run; }
    std::vector<flv::CorrectionInfStruct> outs(n);
    helpers->parallel_for((int)n, [&](int jj) {
      const size_t j = (size_t)jj;
      flv::LocalMap::SolveArrays& a = arr[due[j]];
      std::copy(poses.begin() + j * rP * 7, poses.begin() + j * rP * 7 + 7 * (size_t)a.P, a.poses.begin());
      std::copy(lms.begin() + j * (size_t)rL * 3, lms.begin() + j * (size_t)rL * 3 + 3 * (size_t)a.L, a.lms.begin());
      std::copy(act.begin() + j * (size_t)rE, act.begin() + j * (size_t)rE + a.E, a.active.begin());
      maps[batch[round[due[j]]].stream]->end(a, st[j], outs[j]);
    });
    {
      std::lock_guard<std::mutex> lk(mu);
      for (size_t j = 0; j < n; ++j) {
        const int s = batch[round[due[j]]].stream;
        latest[s] = std::move(outs[j]);
        n_results[s]++;
      }
    }
  }
  host_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - th0).count() - (solve_ms - solve_before);
  return true;
}

void LmShard::run() {
  // the context belongs to this thread: flv_create selects the device for the thread
  const int rc = flv_create(&ctx, device, S, 64, 64, 64);
  {
    std::lock_guard<std::mutex> lk(mu);
    started = true;
    if (rc != FLV_OK) { failed = true; snprintf(err, sizeof(err), "flv_create failed (%d)", rc); }
  }
  cv_idle.notify_all();
  for (;;) {
    std::vector<KfMsg> batch;
    {
      std::unique_lock<std::mutex> lk(mu);
      cv_work.wait(lk, [&] { return stop || !queue.empty(); });
      if (queue.empty()) break;                  // stop requested and nothing left
      // everything that is queued goes into ONE launch: the solver's latency (a few ms) is paid once per drain, however
      // many submissions (frames, stream groups) piled up behind the previous one
      batch = std::move(queue.front());
      queue.pop_front();
      while (!queue.empty()) {
        for (KfMsg& m : queue.front()) batch.push_back(std::move(m));
        queue.pop_front();
      }
      busy = true;
    }
    const bool ok = failed ? false : process(batch);
    {
      std::lock_guard<std::mutex> lk(mu);
      if (!ok) failed = true;
      busy = false;
    }
    cv_idle.notify_all();
  }
  if (ctx) { flv_destroy(ctx); ctx = nullptr; }
}

// The public object: `n_shards` independent shards (stream s -> shard s % n_shards, local index s / n_shards), each with its
// own worker thread, kernel context and CUDA stream, so several solver launches are in flight at once and the graph editing
// of one shard overlaps the solve of another.
struct flv_localmap_batch {
  int S = 0;
  std::vector<std::unique_ptr<LmShard>> shards;
  char err[256] = {0};
};

namespace {
LmShard* make_shard(int device, int n_streams, int window_size, const double* K) {
  LmShard* b = new (std::nothrow) LmShard();
  if (!b) return nullptr;
  b->device = device; b->S = n_streams; b->W = window_size;
  for (int k = 0; k < 4; ++k) b->K[k] = K[k];
  b->latest.resize(n_streams); b->n_results.assign(n_streams, 0);
  for (int s = 0; s < n_streams; ++s) b->maps.emplace_back(new flv::LocalMap(nullptr, window_size, K[0], K[1], K[2], K[3]));
  const unsigned hw = std::thread::hardware_concurrency();
  b->helpers.reset(new Helpers(n_streams > 1 ? (int)std::min<unsigned>(3, hw > 8 ? hw / 8 : 1) : 0));
  b->worker = std::thread([b] { b->run(); });
  std::unique_lock<std::mutex> lk(b->mu);
  b->cv_idle.wait(lk, [&] { return b->started; });
  return b;
}
void destroy_shard(LmShard* b) {
  { std::lock_guard<std::mutex> lk(b->mu); b->stop = true; }
  b->cv_work.notify_all();
  if (b->worker.joinable()) b->worker.join();
}
}  // namespace

extern "C" {

flv_localmap_batch* flv_localmap_batch_create(int device, int n_streams, int window_size, double fx, double fy, double cx, double cy) {
  if (n_streams < 1 || window_size < 3 || window_size > 100) return nullptr;
  flv_localmap_batch* b = new (std::nothrow) flv_localmap_batch();
  if (!b) return nullptr;
  b->S = n_streams;
  const double K[4] = {fx, fy, cx, cy};
  int n_shards = n_streams >= 8 ? 2 : 1;
  if (const char* e = getenv("FLV_LM_SHARDS")) { const int v = atoi(e); if (v >= 1 && v <= n_streams) n_shards = v; }   // tuning knob
  for (int i = 0; i < n_shards; ++i) {
    const int n_local = (n_streams - i + n_shards - 1) / n_shards;
    LmShard* sh = make_shard(device, n_local, window_size, K);
    if (!sh) { flv_localmap_batch_destroy(b); return nullptr; }
    b->shards.emplace_back(sh);
  }
  return b;
}

void flv_localmap_batch_destroy(flv_localmap_batch* b) {
  if (!b) return;
  for (auto& sh : b->shards) destroy_shard(sh.get());
  delete b;
}

const char* flv_localmap_batch_last_error(flv_localmap_batch* b) {
  if (!b) return "null";
  for (auto& sh : b->shards) if (sh->err[0]) return sh->err;
  return b->err;
}

int flv_localmap_batch_submit(flv_localmap_batch* b, int n_kf, const int* streams, const int64_t* frame_ids, const int* lm_counts,
                              const int64_t* lm_id, const double* lm_2d, const double* lm_3d, const double* T_c_w) {
  if (!b || n_kf < 0 || (n_kf > 0 && (!streams || !frame_ids || !lm_counts || !lm_id || !lm_2d || !lm_3d || !T_c_w))) return FLV_ERR_INVALID;
  if (n_kf == 0) return FLV_OK;
  const int W = (int)b->shards.size();
  std::vector<std::vector<KfMsg>> per(W);
  size_t off = 0;
  for (int i = 0; i < n_kf; ++i) {
    if (streams[i] < 0 || streams[i] >= b->S || lm_counts[i] < 0) return FLV_ERR_INVALID;
    per[streams[i] % W].emplace_back();
    KfMsg& m = per[streams[i] % W].back();
    m.stream = streams[i] / W;
    m.kf.frame_id = frame_ids[i]; m.kf.lm_count = lm_counts[i];
    m.kf.lm_id.assign(lm_id + off, lm_id + off + lm_counts[i]);
    m.kf.lm_2d.resize(lm_counts[i]); m.kf.lm_3d.resize(lm_counts[i]);
    for (int k = 0; k < lm_counts[i]; ++k) {
      m.kf.lm_2d[k] = flv::Vec2{lm_2d[2 * (off + k)], lm_2d[2 * (off + k) + 1]};
      m.kf.lm_3d[k] = flv::Vec3{lm_3d[3 * (off + k)], lm_3d[3 * (off + k) + 1], lm_3d[3 * (off + k) + 2]};
    }
    for (int k = 0; k < 7; ++k) m.kf.T_c_w[k] = T_c_w[7 * i + k];
    off += lm_counts[i];
  }
  for (int w = 0; w < W; ++w) {
    if (per[w].empty()) continue;
    LmShard* sh = b->shards[w].get();
    {
      std::lock_guard<std::mutex> lk(sh->mu);
      if (sh->failed) return FLV_ERR_CUDA;
      sh->queue.push_back(std::move(per[w]));
    }
    sh->cv_work.notify_one();
  }
  return FLV_OK;
}

int flv_localmap_batch_wait(flv_localmap_batch* b) {
  if (!b) return FLV_ERR_INVALID;
  int rc = FLV_OK;
  for (auto& shp : b->shards) {
    LmShard* sh = shp.get();
    std::unique_lock<std::mutex> lk(sh->mu);
    sh->cv_idle.wait(lk, [&] { return (sh->queue.empty() && !sh->busy) || sh->failed; });
    if (sh->failed) rc = FLV_ERR_CUDA;
  }
  return rc;
}

int flv_localmap_batch_stats(flv_localmap_batch* b, long long* n_keyframes, long long* n_solves, long long* n_launches, double* solve_ms) {
  if (!b) return FLV_ERR_INVALID;
  long long nk = 0, ns = 0, nl = 0; double ms = 0, hms = 0;
  for (auto& shp : b->shards) {
    LmShard* sh = shp.get();
    std::lock_guard<std::mutex> lk(sh->mu);
    nk += sh->n_keyframes; ns += sh->n_solves; nl += sh->n_launches; ms += sh->solve_ms; hms += sh->host_ms;
  }
  if (n_keyframes) *n_keyframes = nk;
  if (n_solves) *n_solves = ns;
  if (n_launches) *n_launches = nl;
  if (solve_ms) { solve_ms[0] = ms; solve_ms[1] = hms; }   // solve_ms is double[2]: {solver calls, graph editing + packing}, summed over shards
  return FLV_OK;
}

int flv_localmap_batch_problem_totals(flv_localmap_batch* b, double* totals4) {
  if (!b || !totals4) return FLV_ERR_INVALID;
  for (int i = 0; i < 4; ++i) totals4[i] = 0;
  for (auto& shp : b->shards) {
    LmShard* sh = shp.get();
    std::lock_guard<std::mutex> lk(sh->mu);
    totals4[0] += sh->sum_edges; totals4[1] += sh->sum_lms; totals4[2] += sh->sum_poses; totals4[3] += sh->sum_iters;
  }
  return FLV_OK;
}

int flv_localmap_batch_result(flv_localmap_batch* b, int stream, int64_t* out_frame_id, double* out_T_c_w, int* out_lm_count,
                              int64_t* out_lm_id, double* out_lm_3d, int lm_cap, int* out_outlier_count, int64_t* out_outlier_id,
                              int outlier_cap) {
  if (!b || stream < 0 || stream >= b->S) return FLV_ERR_INVALID;
  const int W = (int)b->shards.size();
  LmShard* sh = b->shards[stream % W].get();
  const int ls = stream / W;
  std::lock_guard<std::mutex> lk(sh->mu);
  if (sh->n_results[ls] == 0) return 0;
  const flv::CorrectionInfStruct& c = sh->latest[ls];
  if ((int)c.lm_id.size() > lm_cap || (int)c.lm_outlier_id.size() > outlier_cap) return FLV_ERR_OVERFLOW;
  if (out_frame_id) *out_frame_id = c.frame_id;
  if (out_T_c_w) for (int k = 0; k < 7; ++k) out_T_c_w[k] = c.T_c_w[k];
  if (out_lm_count) *out_lm_count = c.lm_count;
  for (size_t i = 0; i < c.lm_id.size(); ++i) {
    if (out_lm_id) out_lm_id[i] = c.lm_id[i];
    if (out_lm_3d) for (int k = 0; k < 3; ++k) out_lm_3d[3 * i + k] = c.lm_3d[i][k];
  }
  if (out_outlier_count) *out_outlier_count = c.lm_outlier_count;
  if (out_outlier_id) for (size_t i = 0; i < c.lm_outlier_id.size(); ++i) out_outlier_id[i] = c.lm_outlier_id[i];
  return (int)sh->n_results[ls];
}

}  // extern "C"
