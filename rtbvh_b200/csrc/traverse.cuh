// traverse.cuh — launchers of the traversal kernels (traverse.cu).
#pragma once
#include "common.cuh"

namespace rtb {

// Kernel variants of the single-ray path (RTBVH_TRACE_MODE = static | persistent | phased | coop; default persistent):
enum TraceMode {
    kTraceStatic = 0,      // one thread per ray, no refill (A/B only)
    kTracePersistent = 1,  // persistent warps + ray refill, every lane fetches its own node
    kTraceCoop = 2,        // same + lane-cooperative node fetch through shared memory (Mbvh; Bvh falls back to 1)
    kTraceLane = 4,        // Mbvh packets: one lane per RayPacket4 (phased, persistent); other trees fall back to 1
    kTracePhased = 3,      // persistent warps + refill, node visits and triangle tests as separate warp-wide phases (measured 1-6 % behind 1 on single rays)
};
// d_counter: one 64-bit work counter owned by this launch (zeroed on `stream` by the launcher).
// Destinations of the fused multi-GPU gather: up to 8 buffers (own + cudaIpc-mapped peers); record i goes to
// p[k][offset + i] for every k.  count == 0: plain single-GPU store.
struct PeerDests {
    void* p[8];
    int count;
    size_t offset;
    // Chunk-wise push (0: every record is stored to every destination as its ray finishes).  Finished rays store their
    // record locally; the warp that traced a chunk of kRayChunk rays copies the chunk to all destinations with full-width
    // stores once its last ray has finished (needs a local result buffer, an even `offset`, the caller's ray order).
    int push;
    // Input gate of the host-buffer pipeline (null: all rays are resident).  *ready = number of rays of this launch
    // whose H2D copy has completed (written by the copy engine, stream-ordered behind each sub-chunk): a warp that
    // reserves rays [a, b) waits until *ready >= b, so ONE launch can start while its input is still arriving.
    const unsigned long long* ready;
    // Split ray input (the argument shape of the reference's FFI intersect: origin[3], direction[3], t; rtbvh_ffi/src/lib.rs:
    // 551-581): when `directions` is set, the kernel's ray pointer addresses 3 floats of origin per ray, `directions` 3 floats
    // of direction per ray, and t_min / t_max apply to every ray — 24 instead of 32 bytes per ray across PCIe.
    const float* directions;
    float t_min, t_max;
    // Image-ordered batches (rtbvh_gpu_scene_set_ray_tiling): rows of `tile_w` rays; the persistent kernels hand out the rays
    // of the first `tile_n` indices as 8x8 pixel tiles instead of 64-ray row segments (work order only; 0 = off).
    uint32_t tile_w;
    unsigned long long tile_n;
};
// sort_bounds: null = trace in the caller's order; else {min xyz, max xyz} of the scene: the batch is traced in
// Morton order of (origin, direction) and results are scattered back (same results, better coherence).
cudaError_t launch_trace_single(const DeviceTree& tree, int tree_kind, bool any, const RTRay* d_rays, size_t n,
                                RTHit* d_hits, uint8_t* d_occluded, unsigned long long* d_counter,
                                uint32_t* d_overflow, int mode, const float* sort_bounds, const PeerDests* peers,
                                cudaStream_t stream);
cudaError_t launch_trace_packets(const DeviceTree& tree, int tree_kind, bool any, const RTRayPacket4* d_packets,
                                 size_t n_packets, float t_min, RTHitPacket4* d_hits, uint8_t* d_occluded,
                                 unsigned long long* d_counter, uint32_t* d_overflow, int mode, cudaStream_t stream);
// Staged top of an Mbvh (DeviceTree::top): capacity in nodes (0: this build does not stage), and the one-warp builder.
int top_table_capacity();
cudaError_t launch_build_top_table(const float4* d_nodes, uint32_t node_count, float4* d_top, uint32_t* d_top_count,
                                   cudaStream_t stream);
cudaError_t launch_gather_tris(const float* d_verts, uint32_t stride_floats, const uint32_t* d_indices,
                               uint32_t index_count, uint32_t tri_count, TriRec* d_out, cudaStream_t stream);
cudaError_t launch_camera_rays(const float pos[3], const float p1[3], const float right[3], const float up[3],
                               uint32_t width, uint32_t height, uint32_t row0, uint32_t rows, uint64_t seed,
                               uint64_t frame, RTRay* d_rays, cudaStream_t stream);

}  // namespace rtb
