// build.cu — GPU builders (placeholder until the builder kernels land; every entry fails loudly).
#include <string>

#include "build.cuh"

namespace rtb {
ResultCode gpu_build_bvh(const RTAabb*, size_t, const float*, size_t, size_t, uint32_t, HostBvh*) {
    return fail("gpu_build_bvh: not implemented yet");
}
ResultCode gpu_collapse(const HostBvh&, HostMbvh*) { return fail("gpu_collapse: not implemented yet"); }
ResultCode gpu_refit(HostBvh*, const RTAabb*) { return fail("gpu_refit: not implemented yet"); }
}  // namespace rtb
