// legacy.cu — the ten rtbvh_ffi entry points of include/rtbvh.h.
//
// Table semantics follow rtbvh_ffi's StructureManager (rtbvh_ffi/src/lib.rs:12-127): two process-global
// tables (Bvh, Mbvh) behind reader/writer locks; `store` appends and returns raw pointers into the stored
// tree; ids are never reused; `free_*` replaces the entry by an empty tree.  create_* and refit run the GPU
// builders (build.cu) and mirror the result to the host so the returned pointers are host pointers like the
// reference's.  intersect* take no lock and do no lookup (lib.rs:551-581): they trust the struct.
#include <cstdio>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <vector>

#include "build.cuh"
#include "../../include/rtbvh_iter.hpp"
#include "../../include/rtbvh_gpu.h"

using namespace rtb;
namespace hi = rtbvh_host;

namespace {
struct Manager {
    std::shared_mutex mu_bvh, mu_mbvh;
    std::vector<std::shared_ptr<HostBvh>> bvhs;  // shared: an Mbvh keeps the Bvh it was collapsed from alive (its indices)
    std::vector<std::unique_ptr<HostMbvh>> mbvhs;

    RTBvh store(std::shared_ptr<HostBvh> b) {  // lib.rs:46-60
        std::unique_lock<std::shared_mutex> lk(mu_bvh);
        bvhs.push_back(std::move(b));
        const HostBvh& s = *bvhs.back();
        return RTBvh{(uint32_t)(bvhs.size() - 1), (uint32_t)s.nodes.size(), s.nodes.data(), (uint32_t)s.indices.size(),
                     s.indices.data()};
    }
    RTMbvh store_mbvh(std::unique_ptr<HostMbvh> m) {  // lib.rs:62-79
        std::unique_lock<std::shared_mutex> lk(mu_mbvh);
        mbvhs.push_back(std::move(m));
        const HostMbvh& s = *mbvhs.back();
        return RTMbvh{(uint32_t)(mbvhs.size() - 1), (uint32_t)s.m_nodes.size(), s.m_nodes.data(),
                      (uint32_t)s.index_count(), s.indices()};
    }
} g_manager;

hi::RayPacket4 make_packet(const float* ox, const float* oy, const float* oz, const float* dx, const float* dy,
                           const float* dz, const float* t) {  // lib.rs:611-668
    hi::RayPacket4 p;
    const float* o[3] = {ox, oy, oz};
    const float* d[3] = {dx, dy, dz};
    for (int k = 0; k < 3; k++)
        for (int l = 0; l < 4; l++) {
            p.origin[k][l] = o[k][l];
            p.direction[k][l] = d[k][l];
            p.inv_direction[k][l] = 1.0f / d[k][l];
        }
    for (int l = 0; l < 4; l++) p.t[l] = t[l];
    return p;
}
}  // namespace

extern "C" {

ResultCode create_spatial_Bvh(const RTAabb*, size_t, const float*, size_t, const float*, size_t, size_t, uint32_t,
                              RTBvh*) {
    return fail("create_spatial_Bvh: the spatial-split builder is outside the GPU hot path (SURVEY.md section 2)");
}

ResultCode create_bvh(const RTAabb* aabbs, size_t prim_count, const float* centers, size_t center_stride,
                      size_t prims_per_leaf, BvhType bvh_type, RTBvh* result) {
    if (!centers || !result) return Error;                       // lib.rs:437-439
    if (center_stride != 12 && center_stride != 16)              // lib.rs:441-449: assert! (panic) in the reference
        return fail("create_bvh: center_stride must be 12 or 16 bytes");
    if (prim_count == 0) return NoPrimitives;                    // src/bvh.rs:88-90
    auto b = std::make_shared<HostBvh>();
    const ResultCode rc = gpu_build_bvh(aabbs, prim_count, centers, center_stride, prims_per_leaf, (uint32_t)bvh_type, b.get());
    if (rc != Ok) return rc;
    *result = g_manager.store(std::move(b));
    return Ok;
}

// rtbvh_gpu.h: Builder::construct_* for triangle primitives whose aabb()/center() are computed on the device
ResultCode rtbvh_gpu_create_bvh_triangles(const float* vertices, size_t vertex_stride, size_t triangle_count,
                                          size_t prims_per_leaf, BvhType bvh_type, RTBvh* result) {
    if (!vertices || !result) return Error;
    if (vertex_stride != 12 && vertex_stride != 16) return fail("vertex_stride must be 12 or 16 bytes");
    if (triangle_count == 0) return NoPrimitives;
    auto b = std::make_shared<HostBvh>();
    const ResultCode rc = gpu_build_bvh_triangles(vertices, vertex_stride, triangle_count, prims_per_leaf, (uint32_t)bvh_type, b.get());
    if (rc != Ok) return rc;
    *result = g_manager.store(std::move(b));
    return Ok;
}

// rtbvh_gpu.h: Mbvh::construct for a tree that does not live in this library's table (e.g. a reference-built one):
// the struct's pointers are trusted like intersect() trusts them; the result is stored like create_mbvh's.
ResultCode rtbvh_gpu_create_mbvh_from(const RTBvh* bvh, RTMbvh* mbvh) {
    if (!bvh || !mbvh || !bvh->nodes || !bvh->indices) return Error;
    auto copy = std::make_shared<HostBvh>();  // the Mbvh's own clone of nodes and prim_indices (src/bvh.rs:399-403)
    if (!copy->nodes.assign(bvh->nodes, bvh->nodes + bvh->node_count) ||
        !copy->indices.assign(bvh->indices, bvh->indices + bvh->index_count))
        return fail("rtbvh_gpu_create_mbvh_from: out of memory");
    auto m = std::make_unique<HostMbvh>();
    const ResultCode rc = gpu_collapse(*copy, m.get());
    if (rc != Ok) return rc;
    m->base = copy;
    *mbvh = g_manager.store_mbvh(std::move(m));
    return Ok;
}

ResultCode rtbvh_gpu_last_build_stats(double* device_ms, double* total_ms, uint32_t* iterations) {
    if (device_ms) *device_ms = g_build_stats.device_ms;
    if (total_ms) *total_ms = g_build_stats.total_ms;
    if (iterations) *iterations = g_build_stats.iterations;
    return Ok;
}

ResultCode create_mbvh(RTBvh bvh, RTMbvh* mbvh) {
    if (!mbvh || !bvh.nodes || !bvh.indices) return Error;       // lib.rs:500-502
    auto m = std::make_unique<HostMbvh>();
    {
        std::shared_lock<std::shared_mutex> lk(g_manager.mu_bvh);
        if (bvh.id >= g_manager.bvhs.size()) return Error;       // MANAGER.get(..) == None
        const ResultCode rc = gpu_collapse(*g_manager.bvhs[bvh.id], m.get());
        if (rc != Ok) return rc;
        m->base = g_manager.bvhs[bvh.id];
    }
    *mbvh = g_manager.store_mbvh(std::move(m));
    return Ok;
}

ResultCode refit(const RTAabb* aabbs, RTBvh bvh) {
    if (!aabbs || !bvh.nodes || !bvh.indices) return Error;      // lib.rs:520-522
    std::unique_lock<std::shared_mutex> lk(g_manager.mu_bvh);
    if (bvh.id >= g_manager.bvhs.size()) return Error;
    return gpu_refit(g_manager.bvhs[bvh.id].get(), aabbs);
}

ResultCode intersect(RTBvh bvh, const float* origin, const float* direction, float* t, void* user_data,
                     RTIntersectCallback cb) {
    hi::Ray ray = hi::Ray::make(origin, direction);
    if (ray.has_nan()) return Nan;                               // lib.rs:562-564
    ray.t = *t;
    hi::BvhIndexIterator it(&ray, bvh.nodes, bvh.node_count, bvh.indices);
    uint32_t prim;
    while (it.next(&prim))
        if (cb(prim, &ray.t, user_data)) break;
    *t = ray.t;
    return Ok;
}

ResultCode intersect_packet(RTBvh bvh, const float* origin_x, const float* origin_y, const float* origin_z,
                            const float* direction_x, const float* direction_y, const float* direction_z, float* t,
                            void* user_data, RTIntersectCallback cb) {
    hi::RayPacket4 p = make_packet(origin_x, origin_y, origin_z, direction_x, direction_y, direction_z, t);
    if (p.has_nan()) return Nan;                                 // lib.rs:647-655
    hi::BvhPacketIndexIterator it(&p, bvh.nodes, bvh.node_count, bvh.indices);
    uint32_t prim;
    while (it.next(&prim))
        if (cb(prim, p.t, user_data)) break;
    for (int l = 0; l < 4; l++) t[l] = p.t[l];
    return Ok;
}

ResultCode intersect_mbvh(RTMbvh bvh, const float* origin, const float* direction, float* t, void* user_data,
                          RTIntersectCallback cb) {
    hi::Ray ray = hi::Ray::make(origin, direction);
    if (ray.has_nan()) return Nan;                               // lib.rs:711-713
    ray.t = *t;
    hi::MbvhIndexIterator it(&ray, bvh.nodes, bvh.node_count, bvh.indices);
    uint32_t prim;
    while (it.next(&prim))
        if (cb(prim, &ray.t, user_data)) break;
    *t = ray.t;
    return Ok;
}

ResultCode intersect_mbvh_packet(RTMbvh bvh, const float* origin_x, const float* origin_y, const float* origin_z,
                                 const float* direction_x, const float* direction_y, const float* direction_z,
                                 float* t, void* user_data, RTIntersectCallback cb) {
    hi::RayPacket4 p = make_packet(origin_x, origin_y, origin_z, direction_x, direction_y, direction_z, t);
    if (p.has_nan()) return Nan;                                 // lib.rs:797-805
    hi::MbvhPacketIndexIterator it(&p, bvh.nodes, bvh.node_count, bvh.indices);
    uint32_t prim;
    while (it.next(&prim))
        if (cb(prim, p.t, user_data)) break;
    for (int l = 0; l < 4; l++) t[l] = p.t[l];
    return Ok;
}

void free_bvh(RTBvh bvh) {                                        // lib.rs:838-842, :108-116
    std::unique_lock<std::shared_mutex> lk(g_manager.mu_bvh);
    if (bvh.id < g_manager.bvhs.size())
        g_manager.bvhs[bvh.id] = std::make_shared<HostBvh>();
    else
        std::fprintf(stderr, "Could not free bvh with id: %u\n", bvh.id);
}

void free_mbvh(RTMbvh bvh) {                                      // lib.rs:845-849, :118-126
    std::unique_lock<std::shared_mutex> lk(g_manager.mu_mbvh);
    if (bvh.id < g_manager.mbvhs.size())
        g_manager.mbvhs[bvh.id] = std::make_unique<HostMbvh>();
    else
        std::fprintf(stderr, "Could not free bvh with id: %u\n", bvh.id);
}

}  // extern "C"
