// oracle_capi.cpp — plain-C entry points over rtbvh_oracle.hpp for ctypes.
//
// TEST INFRASTRUCTURE ONLY (see the header of rtbvh_oracle.hpp).  Loaded by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the
// checker / the timed CPU baseline; never by the product.
//
// Threading of the batch calls mirrors examples/benchmark.rs:25,55: rayon
// par_chunks_mut(1000) -> OpenMP schedule(dynamic) over chunks of 1000 rays / packets.
#include "rtbvh_oracle.hpp"

#include <chrono>
#include <cstdio>
#ifdef _OPENMP
#include <omp.h>
#include <parallel/algorithm>
#endif

using namespace rto;

std::vector<uint32_t> LocbBuilder::sorted_indices(const MortonEncoder& enc, std::vector<uint32_t>* codes_out,
                                                  bool parallel) {
    std::vector<uint32_t> indices(n), codes(n);
    for (size_t i = 0; i < n; i++) indices[i] = (uint32_t)i;
#pragma omp parallel for if (parallel) schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) codes[i] = enc.encode(centers[i]);
    auto cmp = [&](uint32_t a, uint32_t b) { return codes[a] < codes[b]; };
#ifdef _OPENMP
    if (parallel)
        __gnu_parallel::stable_sort(indices.begin(), indices.end(), cmp);
    else
#endif
        std::stable_sort(indices.begin(), indices.end(), cmp);
    if (codes_out) *codes_out = std::move(codes);
    return indices;
}

namespace {
struct RTRay {  // first 32 bytes of ray.rs:9-16
    float origin[3];
    float t_min;
    float direction[3];
    float t;
};
struct RTHit {
    float t;
    uint32_t prim;
};
struct RTRayPacket4 {  // the seven SoA inputs of rtbvh_ffi intersect_packet (lib.rs:599-608)
    float origin_x[4], origin_y[4], origin_z[4], direction_x[4], direction_y[4], direction_z[4], t[4];
};
struct RTHitPacket4 {
    float t[4];
    uint32_t prim[4];
};
static_assert(sizeof(RTRay) == 32 && sizeof(RTHit) == 8 && sizeof(RTRayPacket4) == 112 && sizeof(RTHitPacket4) == 32, "");

constexpr uint32_t kNoHit = 0xFFFFFFFFu;

inline Tri load_tri(const float* v, size_t id) {
    const float* p = v + id * 9;
    return Tri{{p[0], p[1], p[2]}, {p[3], p[4], p[5]}, {p[6], p[7], p[8]}};
}
inline Ray make_ray(const RTRay& in) {
    Ray r = ray_new(v3(in.origin[0], in.origin[1], in.origin[2]), v3(in.direction[0], in.direction[1], in.direction[2]));
    r.t_min = in.t_min;
    r.t = in.t;  // rtbvh_ffi/src/lib.rs:566
    return r;
}
double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

extern "C" {

int rto_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---- KAT helpers -----------------------------------------------------------------
uint32_t rto_morton_split(uint32_t v) { return morton_split(v); }
uint32_t rto_prefix_sum_u32(const uint32_t* in, size_t n, uint32_t* out) { return prefix_sum(in, n, out); }
int32_t rto_prefix_sum_i32(const int32_t* in, size_t n, int32_t* out) { return prefix_sum(in, n, out); }
uint64_t rto_prefix_sum_u64(const uint64_t* in, size_t n, uint64_t* out) { return prefix_sum(in, n, out); }
size_t rto_sizeof_aabb() { return sizeof(Aabb); }
size_t rto_sizeof_bvh_node() { return sizeof(BvhNode); }
size_t rto_sizeof_mbvh_node() { return sizeof(MbvhNode); }
// utils.rs:59-72 move_backward on u32
void rto_move_backward_u32(uint32_t* first, uint32_t* last, uint32_t* d_last) {
    while (first != last) *(--d_last) = *(--last);
}
size_t rto_partition_lt(uint32_t* slice, size_t n, uint32_t pivot) {
    return partition(slice, n, [&](uint32_t v) { return v < pivot; });
}
uint32_t rto_morton_encode(const float* world_aabb32, const float* p) {
    Aabb bb;
    std::memcpy(&bb, world_aabb32, 32);
    return MortonEncoder(bb, 1024).encode(v3(p[0], p[1], p[2]));
}

// ---- primitives from triangles (shared/src/lib.rs:27-39; `aabb!` macro aabb.rs:472-482) --------
// verts: n x 9 floats.  pad = 0 for Primitive::aabb of the bench Triangle, 1e-4 for the aabb! macro.
void rto_prims_from_triangles(const float* verts, size_t n, float pad, void* aabbs_out, float* centers_out) {
    Aabb* bbs = (Aabb*)aabbs_out;
    for (size_t i = 0; i < n; i++) {
        Tri t = load_tri(verts, i);
        Aabb bb = aabb_new();
        grow(bb, t.v0);
        grow(bb, t.v1);
        grow(bb, t.v2);
        if (pad != 0.0f) offset_by(bb, pad);
        bbs[i] = bb;
        Vec3 c = (t.v0 + t.v1 + t.v2) * (1.0f / 3.0f);
        centers_out[i * 3 + 0] = c.x;
        centers_out[i * 3 + 1] = c.y;
        centers_out[i * 3 + 2] = c.z;
    }
}
// aabb.rs:325-327 (what the FFI tests use as centers)
void rto_aabb_centers(const void* aabbs, size_t n, float* centers_out) {
    const Aabb* bbs = (const Aabb*)aabbs;
    for (size_t i = 0; i < n; i++) {
        Vec3 c = center(bbs[i]);
        centers_out[i * 3 + 0] = c.x;
        centers_out[i * 3 + 1] = c.y;
        centers_out[i * 3 + 2] = c.z;
    }
}

// ---- Bvh handles -----------------------------------------------------------------
// Builder::construct_* (bvh.rs:58-138) + rtbvh_ffi create_bvh (lib.rs:428-493).
// result codes follow rtbvh_ffi ResultCode (lib.rs:17-25): 0 Ok, 1 Error, 2 NoPrimitives,
// 3 InequalAabbsAndPrimitives.  aabbs may be null: the centers then act as point primitives
// (lib.rs:396-422).  type: 0 LocallyOrderedClustered, 1 BinnedSAH (lib.rs:129-133).
int rto_bvh_build(int type, const void* aabbs, size_t aabb_count, const float* centers, size_t center_stride_bytes,
                  size_t prim_count, size_t prims_per_leaf, int parallel, Bvh** out, double* build_ms, double* kappa) {
    if (!centers || !out) return 1;
    if (center_stride_bytes != 12 && center_stride_bytes != 16) return 1;  // reference: assert! panic
    if (prim_count == 0) return 2;
    if (aabbs && aabb_count != prim_count) return 3;
    std::vector<Vec3> c(prim_count);
    size_t sf = center_stride_bytes / 4;
    for (size_t i = 0; i < prim_count; i++) c[i] = v3(centers[i * sf], centers[i * sf + 1], centers[i * sf + 2]);
    std::vector<Aabb> own;
    const Aabb* bbs = (const Aabb*)aabbs;
    if (!bbs) {
        own.resize(prim_count);
        for (size_t i = 0; i < prim_count; i++) {
            own[i] = aabb_new();
            grow(own[i], c[i]);
        }
        bbs = own.data();
    }
    double t0 = now_ms();
    Bvh* b = new Bvh();
    if (type == 0) {
        LocbBuilder lb{bbs, c.data(), prim_count};
        *b = lb.build(parallel != 0);
        if (kappa) *kappa = prim_count ? (double)lb.cluster_sum / (double)prim_count : 0.0;
    } else {
        BinnedSahBuilder sb{bbs, c.data(), prim_count, prims_per_leaf ? prims_per_leaf : 1, {}, {}, 1};
        // parallel: the reference's threaded scheduling (subtrees > 1024 primitives on other threads) with `parallel` threads
        // (1 = all cores); node numbering then depends on the thread interleaving, like the reference's.  0: deterministic.
        *b = parallel != 0 ? sb.build_parallel(parallel == 1 ? omp_get_max_threads() : parallel) : sb.build();
        if (kappa) *kappa = 0.0;
    }
    if (build_ms) *build_ms = now_ms() - t0;
    *out = b;
    return 0;
}
// Builder::construct_spatial_sah (bvh.rs:58-85) for triangles given as n x 9 floats; aabb = Primitive::aabb
// (un-padded), center = (v0+v1+v2)*(1/3) like the bench Triangle.  fix_child_ranges: see rtbvh_oracle.hpp.
int rto_bvh_build_spatial(const float* verts, size_t n, size_t prims_per_leaf, int fix_child_ranges, Bvh** out,
                          double* build_ms, uint64_t* stats /* spatial splits, object splits, references */) {
    if (!verts || !out) return 1;
    if (n == 0) return 2;
    std::vector<Tri> tris(n);
    std::vector<Aabb> bbs(n);
    std::vector<Vec3> c(n);
    for (size_t i = 0; i < n; i++) {
        tris[i] = load_tri(verts, i);
        Aabb bb = aabb_new();
        grow(bb, tris[i].v0);
        grow(bb, tris[i].v1);
        grow(bb, tris[i].v2);
        bbs[i] = bb;
        c[i] = (tris[i].v0 + tris[i].v1 + tris[i].v2) * (1.0f / 3.0f);
    }
    double t0 = now_ms();
    SpatialSahBuilder sb;
    sb.aabbs = bbs.data();
    sb.tris = tris.data();
    sb.centers = c.data();
    sb.n = n;
    sb.max_leaf_size = prims_per_leaf ? prims_per_leaf : 1;
    sb.fix_child_ranges = fix_child_ranges != 0;
    Bvh* b = new Bvh(sb.build());
    if (build_ms) *build_ms = now_ms() - t0;
    if (stats) {
        stats[0] = sb.spatial_splits;
        stats[1] = sb.object_splits;
        stats[2] = sb.reference_count;
    }
    *out = b;
    return 0;
}
Bvh* rto_bvh_from_raw(const void* nodes, size_t n_nodes, const uint32_t* indices, size_t n_idx) {
    Bvh* b = new Bvh();
    b->nodes.resize(n_nodes);
    std::memcpy(b->nodes.data(), nodes, n_nodes * sizeof(BvhNode));
    b->prim_indices.assign(indices, indices + n_idx);
    return b;
}
void rto_bvh_free(Bvh* b) { delete b; }
size_t rto_bvh_node_count(const Bvh* b) { return b->nodes.size(); }
const void* rto_bvh_nodes(const Bvh* b) { return b->nodes.data(); }
size_t rto_bvh_index_count(const Bvh* b) { return b->prim_indices.size(); }
const uint32_t* rto_bvh_indices(const Bvh* b) { return b->prim_indices.data(); }
int rto_bvh_validate(const Bvh* b, size_t prim_count) { return validate(*b, prim_count) ? 1 : 0; }
double rto_bvh_sah_cost(const Bvh* b) { return sah_cost(b->nodes.data(), b->nodes.size()); }
double rto_sah_cost_raw(const void* nodes, size_t n) { return sah_cost((const BvhNode*)nodes, n); }
void rto_bvh_refit(Bvh* b, const void* aabbs) { refit(*b, (const Aabb*)aabbs); }
// mean / max leaf depth and leaf count (for D-bar in the build roofline)
void rto_bvh_depth_stats(const Bvh* b, double* mean_leaf_depth, uint32_t* max_depth, uint64_t* leaves) {
    uint64_t nl = 0, sum = 0;
    uint32_t md = 0;
    if (!b->nodes.empty()) {
        std::vector<std::pair<int32_t, uint32_t>> st{{0, 0}};
        while (!st.empty()) {
            auto [k, d] = st.back();
            st.pop_back();
            const BvhNode& nd = b->nodes[k];
            if (nd.extra1 >= 0) {
                nl++;
                sum += (uint64_t)d * (uint64_t)std::max(nd.extra1, 1);
                md = std::max(md, d);
            } else if (nd.extra2 >= 0) {
                st.push_back({nd.extra2, d + 1});
                st.push_back({nd.extra2 + 1, d + 1});
            }
        }
    }
    uint64_t prims = b->prim_indices.size();
    *mean_leaf_depth = prims ? (double)sum / (double)prims : 0.0;
    *max_depth = md;
    *leaves = nl;
}

// ---- Mbvh handles (bvh.rs:381-404) -------------------------------------------------
Mbvh* rto_mbvh_construct(const Bvh* b, double* ms) {
    double t0 = now_ms();
    Mbvh* m = new Mbvh(mbvh_construct(*b));
    if (ms) *ms = now_ms() - t0;
    return m;
}
void rto_mbvh_free(Mbvh* m) { delete m; }
size_t rto_mbvh_node_count(const Mbvh* m) { return m->m_nodes.size(); }
const void* rto_mbvh_nodes(const Mbvh* m) { return m->m_nodes.data(); }
size_t rto_mbvh_index_count(const Mbvh* m) { return m->prim_indices.size(); }
const uint32_t* rto_mbvh_indices(const Mbvh* m) { return m->prim_indices.data(); }

// ---- batch traversal: what examples/benchmark.rs:15-71 does per ray, plus hit-id tracking ----
// tree: 0 = Bvh (nodes are 32-byte BvhNode), 1 = Mbvh (128-byte MbvhNode)
// mode: 0 = closest hit (id = lowest id among exactly-equal t), 1 = any hit
// counters_out: 5 x u64 {node_visits, inner_visits, prim_tests, max_stack, overflow32} or null
// returns elapsed milliseconds of the traversal loop.
double rto_trace(int tree, int mode, const void* nodes, size_t n_nodes, const uint32_t* indices, const float* verts,
                 const void* rays_v, size_t n_rays, void* hits_v, uint8_t* occluded, uint64_t* counters_out, int threads) {
    const RTRay* rays = (const RTRay*)rays_v;
    RTHit* hits = (RTHit*)hits_v;
    Counters total;
    const bool count = counters_out != nullptr;
    (void)threads;
    double t0 = now_ms();
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
    {
        Counters local;
        Counters* lc = count ? &local : nullptr;
#pragma omp for schedule(dynamic, 1000) nowait
        for (int64_t i = 0; i < (int64_t)n_rays; i++) {
            Ray ray = make_ray(rays[i]);
            uint32_t best = kNoHit;
            bool occ = false;
            auto body = [&](uint32_t id) -> bool {
                Tri tr = load_tri(verts, id);
                float t;
                if (!tri_geom(tr, ray.origin, ray.direction, &t)) return false;
                if (!(t > ray.t_min)) return false;
                if (t < ray.t) {  // spatial_sah.rs:156-158
                    ray.t = t;
                    best = id;
                    if (mode == 1) {
                        occ = true;
                        return true;  // callback returned true -> break (lib.rs:573)
                    }
                } else if (mode == 0 && best != kNoHit && t == ray.t && id < best) {
                    best = id;  // north-star tie rule; ray.t unchanged so visitation is unchanged
                }
                return false;
            };
            if (tree == 0)
                bvh_traverse((const BvhNode*)nodes, n_nodes, indices, ray, body, lc);
            else
                mbvh_traverse((const MbvhNode*)nodes, n_nodes, indices, ray, body, lc);
            if (hits) hits[i] = RTHit{ray.t, best};
            if (occluded) occluded[i] = occ ? 1 : 0;
        }
        if (count) {
#pragma omp critical
            total.merge(local);
        }
    }
    double ms = now_ms() - t0;
    if (counters_out) {
        counters_out[0] = total.node_visits;
        counters_out[1] = total.inner_visits;
        counters_out[2] = total.prim_tests;
        counters_out[3] = total.max_stack;
        counters_out[4] = total.overflow32;
    }
    return ms;
}

// Packet flavour (benchmark.rs:43-71: intersect4(packet, splat(1e-4))).
// mode 1 (any hit): a lane that finds a hit is retired by writing t = -1e34 through the callback's
// `t` pointer (the FFI lets the callback mutate t, lib.rs:677-681) and the callback returns true
// once all four lanes are retired.  occluded: 4 bytes per packet.
double rto_trace_packet(int tree, int mode, const void* nodes, size_t n_nodes, const uint32_t* indices,
                        const float* verts, const void* packets_v, size_t n_packets, float t_min_f, void* hits_v,
                        uint8_t* occluded, uint64_t* counters_out, int threads) {
    const RTRayPacket4* packets = (const RTRayPacket4*)packets_v;
    RTHitPacket4* hits = (RTHitPacket4*)hits_v;
    Counters total;
    const bool count = counters_out != nullptr;
    const __m128 t_min = _mm_set1_ps(t_min_f);
    double t0 = now_ms();
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
    {
        Counters local;
        Counters* lc = count ? &local : nullptr;
#pragma omp for schedule(dynamic, 250) nowait
        for (int64_t i = 0; i < (int64_t)n_packets; i++) {
            const RTRayPacket4& in = packets[i];
            RayPacket4 p = packet_new(in.origin_x, in.origin_y, in.origin_z, in.direction_x, in.direction_y,
                                      in.direction_z, in.t);
            uint32_t best[4] = {kNoHit, kNoHit, kNoHit, kNoHit};
            int retired = 0;
            auto body = [&](uint32_t id) -> bool {
                Tri tr = load_tri(verts, id);
                __m128 t;
                int m = tri_geom4(tr, p, t_min, &t);
                if (!m) return false;
                alignas(16) float tv[4], pt[4];
                _mm_store_ps(tv, t);
                _mm_store_ps(pt, p.t);
                for (int l = 0; l < 4; l++) {
                    if (!(m & (1 << l))) continue;
                    if (tv[l] < pt[l]) {  // spatial_sah.rs:229-236
                        pt[l] = tv[l];
                        best[l] = id;
                        if (mode == 1) {
                            pt[l] = -1e34f;
                            retired |= 1 << l;
                        }
                    } else if (mode == 0 && best[l] != kNoHit && tv[l] == pt[l] && id < best[l]) {
                        best[l] = id;
                    }
                }
                p.t = _mm_load_ps(pt);
                return mode == 1 && retired == 0xF;
            };
            if (tree == 0)
                bvh_traverse_packet((const BvhNode*)nodes, n_nodes, indices, p, body, lc);
            else
                mbvh_traverse_packet((const MbvhNode*)nodes, n_nodes, indices, p, body, lc);
            alignas(16) float pt[4];
            _mm_store_ps(pt, p.t);
            if (hits)
                for (int l = 0; l < 4; l++) {
                    hits[i].t[l] = pt[l];
                    hits[i].prim[l] = best[l];
                }
            if (occluded)
                for (int l = 0; l < 4; l++) occluded[i * 4 + l] = (retired >> l) & 1;
        }
        if (count) {
#pragma omp critical
            total.merge(local);
        }
    }
    double ms = now_ms() - t0;
    if (counters_out) {
        counters_out[0] = total.node_visits;
        counters_out[1] = total.inner_visits;
        counters_out[2] = total.prim_tests;
        counters_out[3] = total.max_stack;
        counters_out[4] = total.overflow32;
    }
    return ms;
}

// Brute-force arbiter: every triangle in ascending id with the single-ray test (strict <, so the
// lowest id wins exact ties).  Third opinion for trees whose boxes are non-conservative (quirk Q3).
void rto_brute_force(const float* verts, size_t n_tris, const void* rays_v, size_t n_rays, void* hits_v, int threads) {
    const RTRay* rays = (const RTRay*)rays_v;
    RTHit* hits = (RTHit*)hits_v;
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)n_rays; i++) {
        Ray ray = make_ray(rays[i]);
        uint32_t best = kNoHit;
        if (!(is_nan(ray.origin) || is_nan(ray.direction))) {
            for (size_t id = 0; id < n_tris; id++) {
                Tri tr = load_tri(verts, id);
                if (tri_intersect(tr, ray)) best = (uint32_t)id;
            }
        }
        hits[i] = RTHit{ray.t, best};
    }
}

// ---- the legacy per-candidate callback ABI, as the reference's own FFI tests drive it ----------
// rtbvh_ffi/src/lib.rs:551-581 / :700-731.  Returns ResultCode (4 = Nan).
typedef bool (*rto_cb)(uint32_t, float*, void*);
int rto_intersect_cb(int tree, const void* nodes, size_t n_nodes, const uint32_t* indices, const float* origin,
                     const float* direction, float* t, void* user, rto_cb cb) {
    Vec3 o = v3(origin[0], origin[1], origin[2]), d = v3(direction[0], direction[1], direction[2]);
    if (is_nan(o) || is_nan(d)) return 4;
    Ray ray = ray_new(o, d);
    ray.t = *t;
    auto body = [&](uint32_t id) -> bool { return cb(id, &ray.t, user); };
    if (tree == 0)
        bvh_traverse((const BvhNode*)nodes, n_nodes, indices, ray, body, nullptr);
    else
        mbvh_traverse((const MbvhNode*)nodes, n_nodes, indices, ray, body, nullptr);
    *t = ray.t;
    return 0;
}

}  // extern "C"
