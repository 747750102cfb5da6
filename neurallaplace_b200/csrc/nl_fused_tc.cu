// tcgen05 / TMEM fused path (placeholder until the tensor-core kernels land).
#include "nl_kernels.h"

namespace nl {
TcPlan tc_plan(const nl_desc*, int) { return TcPlan{}; }
cudaError_t launch_fused_tc(const nl_desc*, bool, const void*, const void*, const void* const*, const void*,
                            const void*, void*, const void*, void* const*, void*, cudaStream_t, int*) {
  return cudaErrorNotSupported;
}
}  // namespace nl
