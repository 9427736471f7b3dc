// dense_tc.cu -- K2: tcgen05 TF32 candidate generation + exact fp32 rescoring.
// (placeholder translation unit until the tensor-core kernel lands; the dense path then
// runs on K1, which is still a GPU kernel -- there is no CPU fallback anywhere)
#include "engine.h"

namespace krag {
bool dense_tc_supported(const DeviceInfo&, int) { return false; }
size_t dense_tc_workspace_bytes(const DeviceInfo&, int, int) { return 0; }
bool launch_dense_tc(const DeviceInfo&, const float*, int64_t, int, const uint32_t*, const float*, int, int, uint32_t,
                     void*, size_t, uint64_t*, uint64_t*, cudaStream_t)
{
    return false;
}
}  // namespace krag
