// glibc's rand() (random_r TYPE_3, default seed 1) restated, so the dummy depths of
// CameraFrame::recover3DPts_c_FromStereo / FromDepthImg (src/processing/camera_frame.cpp:153,168,198,222:
// d_rand = 0.3 + (float)rand() / (float)(RAND_MAX/0.4)) follow the sequence the reference process would draw.
// One generator per tracked sequence (the reference has one process-global stream per process = per sequence).
#pragma once
#include <cstdint>

namespace flv {

class GlibcRand {
 public:
  explicit GlibcRand(unsigned seed = 1) {
    int32_t r[34];
    r[0] = (int32_t)seed;
    for (int i = 1; i < 31; ++i) {
      const int64_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
      int64_t w = 16807 * lo - 2836 * hi;
      if (w < 0) w += 2147483647;
      r[i] = (int32_t)w;
    }
    for (int i = 31; i < 34; ++i) r[i] = r[i - 31];
    for (int i = 0; i < 34; ++i) st_[i] = (uint32_t)r[i];
    head_ = 0;
    for (int i = 0; i < 310; ++i) step();
  }
  int rand() { return (int)(step() >> 1); }
  float dummy_depth() { return (float)(0.3 + (double)((float)rand() / (float)(2147483647 / 0.4))); }

 private:
  uint32_t st_[34];
  int head_;     // index of the oldest element (r[i-34]); the ring always holds the last 34 outputs
  uint32_t step() {
    // o_i = o_{i-31} + o_{i-3}
    const uint32_t v = st_[(head_ + 34 - 31) % 34] + st_[(head_ + 34 - 3) % 34];
    st_[head_] = v;
    head_ = (head_ + 1) % 34;
    return v;
  }
};

}  // namespace flv
