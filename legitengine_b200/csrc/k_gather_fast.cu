// k_gather_fast.cu — screen-space GI gather (K5), throughput variant. (placeholder: forwards to the strict kernel)
#include "lgcu_kernels.h"

namespace lgcu {
cudaError_t launchGatherFast(const GatherArgs &a, const GatherTables &t, cudaStream_t s) { return launchGatherStrict(a, t, s); }
} // namespace lgcu
