// build.cuh — GPU builders behind create_bvh / create_mbvh / refit (build.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace rtb {

struct HostBvh {   // host mirror handed out through RTBvh (rtbvh::Bvh, src/bvh.rs:143-147)
    std::vector<RTBvhNode> nodes;
    std::vector<uint32_t> indices;
    int build_type = 0;  // src/bvh.rs:18-23
};
struct HostMbvh {  // rtbvh::Mbvh, src/bvh.rs:320-324
    std::vector<RTBvhNode> nodes;
    std::vector<RTMbvhNode> m_nodes;
    std::vector<uint32_t> indices;
};

extern thread_local std::string g_last_error;
ResultCode fail(const char* what, cudaError_t e);
ResultCode fail(const char* what);

// Builder::construct_binned_sah / construct_locally_ordered_clustered (src/bvh.rs:87-137) on the GPU.
ResultCode gpu_build_bvh(const RTAabb* aabbs, size_t prim_count, const float* centers, size_t center_stride,
                         size_t prims_per_leaf, uint32_t bvh_type, HostBvh* out);
// Mbvh::construct (src/bvh.rs:381-404) on the GPU.
ResultCode gpu_collapse(const HostBvh& bvh, HostMbvh* out);
// Bvh::refit (src/bvh.rs:176-205) on the GPU.
ResultCode gpu_refit(HostBvh* bvh, const RTAabb* aabbs);

}  // namespace rtb
