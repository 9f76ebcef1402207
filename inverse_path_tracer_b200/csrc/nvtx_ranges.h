// nvtx_ranges.h -- NVTX ranges around the stages of the hot path (tracing): visible in Nsight
// Systems / ncu --nvtx timelines, one range per wavefront stage and per persistent-kernel launch.  NVTX v3 is
// header-only (the tool injects itself at run time); without a tool attached a range is one indirect call.
#pragma once
#include <nvtx3/nvToolsExt.h>

namespace ipt {
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace ipt
