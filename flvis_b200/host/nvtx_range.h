// NVTX ranges for the tracker's steps (tracing hooks).  Header-only NVTX v3: a no-op unless a tool
// (Nsight Systems / Compute) is attached.
#pragma once
#include <nvtx3/nvToolsExt.h>

namespace flv {

struct NvtxRange {                       // RAII: one named range
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

struct NvtxStages {                      // consecutive stages of one function: next() closes the open range and opens the next
  int open = 0;
  void next(const char* name) { if (open) nvtxRangePop(); nvtxRangePushA(name); open = 1; }
  void close() { if (open) { nvtxRangePop(); open = 0; } }
  ~NvtxStages() { close(); }
};

}  // namespace flv
