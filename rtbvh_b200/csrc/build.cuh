// build.cuh — GPU builders behind create_bvh / create_mbvh / refit (build.cu).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "hostmem.cuh"

namespace rtb {

struct HostBvh {   // host mirror handed out through RTBvh (rtbvh::Bvh, src/bvh.rs:143-147); pooled page-locked arrays
    HostArray<RTBvhNode> nodes;
    HostArray<uint32_t> indices;
    int build_type = 0;  // src/bvh.rs:18-23
    uint64_t serial = 0;  // process-wide build number (identifies the device copy a collapse may reuse)
};
// rtbvh::Mbvh (src/bvh.rs:320-324) keeps clones of the binary nodes and of prim_indices next to m_nodes.  Here the Mbvh shares
// them with the Bvh it was collapsed from (`base`: both tables hold shared ownership, so free_bvh never invalidates the
// pointers in an RTMbvh) instead of copying 68 MB per Mtri on the host; RTMbvh exposes only m_nodes and indices.
struct HostMbvh {
    std::shared_ptr<const HostBvh> base;
    HostArray<RTMbvhNode> m_nodes;
    const uint32_t* indices() const { return base ? base->indices.data() : nullptr; }
    size_t index_count() const { return base ? base->indices.size() : 0; }
};

struct BuildStats {   // of the last build / collapse / refit on this thread
    double device_ms = 0;   // kernels only (inputs resident -> tree resident), CUDA events
    double total_ms = 0;    // incl. H2D of the inputs and D2H of the host mirror
    uint32_t iterations = 0;  // LOCB clustering iterations
    uint32_t node_count = 0;
};
extern thread_local BuildStats g_build_stats;
extern thread_local std::string g_last_error;
ResultCode fail(const char* what, cudaError_t e);
ResultCode fail(const char* what);

// Builder::construct_binned_sah / construct_locally_ordered_clustered (src/bvh.rs:87-137) on the GPU.
ResultCode gpu_build_bvh(const RTAabb* aabbs, size_t prim_count, const float* centers, size_t center_stride,
                         size_t prims_per_leaf, uint32_t bvh_type, HostBvh* out);
// Same, with Primitive::aabb / Primitive::center of the bench Triangle computed on the device from the
// vertices (shared/src/lib.rs:27-39): the "AABB/centroid reduction" stage of the GPU builder.
ResultCode gpu_build_bvh_triangles(const float* vertices, size_t vertex_stride, size_t tri_count, size_t prims_per_leaf,
                                   uint32_t bvh_type, HostBvh* out);
// Mbvh::construct (src/bvh.rs:381-404) on the GPU.
ResultCode gpu_collapse(const HostBvh& bvh, HostMbvh* out);  // fills out->m_nodes (out->base is the caller's business)
// Bvh::refit (src/bvh.rs:176-205) on the GPU.
ResultCode gpu_refit(HostBvh* bvh, const RTAabb* aabbs);

// Builder::construct_* + Mbvh::construct with the results LEFT ON THE DEVICE (cudaFree-able allocations): what
// rtbvh_gpu_scene_build hands to a scene.  d_vertices is the uploaded copy of host vertices (null if they were resident).
struct ResidentTrees {
    void* d_nodes = nullptr;
    uint32_t* d_indices = nullptr;
    void* d_mnodes = nullptr;
    float* d_vertices = nullptr;
    uint32_t node_count = 0, index_count = 0, m_count = 0;
};
ResultCode gpu_build_resident(const float* vertices, bool vertices_on_device, size_t vertex_stride, size_t tri_count,
                              size_t prims_per_leaf, uint32_t bvh_type, bool want_mbvh, ResidentTrees* out);

// Device blocks of resident scenes (trees, triangle records) come from a small process-wide cache: a scene that is freed
// leaves its large cudaMalloc'ed blocks there and the next scene build / replication of a similar size takes them again —
// a rebuild per frame then costs no cudaMalloc / cudaFree (measured: 95 of 104 ms of a 10 M-triangle scene build were those).
// Blocks are plain cudaMalloc allocations (cudaIpc-exportable); the cache keeps at most kDevCacheBlocks blocks and
// RTBVH_SCENE_CACHE_MB megabytes per device (default 24 GiB); rtbvh_gpu_trim_workspace empties it.
cudaError_t dev_block_alloc(void** p, size_t bytes);
void dev_block_free(void* p);  // null-safe; pointers the cache does not know are cudaFree'd
void dev_block_trim();

// Frees the calling thread's builder workspace (it is otherwise kept between builds and only ever grows) and the cached
// page-locked host blocks.
ResultCode gpu_trim_workspace();

// Dynamic scenes: refit of a device-resident Bvh (+ refresh of its Mbvh) from new vertex positions,
// everything on `stream`.  The cache holds what does not change under refit (parent links, the binary node behind every
// MbvhNode) plus scratch; it belongs to one tree pair.
struct ResidentRefit {
    uint32_t n_nodes = 0;
    int32_t* parent = nullptr;
    uint32_t* arrived = nullptr;
    uint8_t* is_mroot = nullptr;
    uint32_t* mindex = nullptr;
    float4* bb = nullptr;
    ~ResidentRefit();
};
ResultCode gpu_refit_resident(ResidentRefit* cache, float4* d_nodes, uint32_t n_nodes, const uint32_t* d_indices, uint32_t n_prims,
                              const float* d_vertices, uint32_t vstride, uint32_t tri_count, float4* d_mnodes, uint32_t m_count,
                              cudaStream_t stream);

}  // namespace rtb
