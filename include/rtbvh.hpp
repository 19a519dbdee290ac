// rtbvh.hpp — C++ host mirror of the reference's Rust surface over the C ABI (rtbvh.h / rtbvh_gpu.h).
//
// The reference's toolchain (Rust) is absent from this image, so this header plays the role of the `rtbvh`
// crate's public API for compiled callers: same names, argument meaning and error behaviour.
//
//   rtbvh::Primitive concept : T::center() -> Vec3, T::aabb() -> Aabb                      src/bvh.rs:10-14
//   rtbvh::Builder<T>{aabbs, primitives, primitives_per_leaf}
//        .construct_binned_sah() / .construct_locally_ordered_clustered() -> Result<Bvh>      src/bvh.rs:49-138
//   rtbvh::Bvh  : nodes(), indices(), prim_count(), refit(), validate(), bounds(), traverse_iter_indices*   src/bvh.rs:143-284
//   rtbvh::Mbvh : construct(bvh) / Mbvh(Bvh), nodes(), quad_nodes(), indices(), traverse_iter_indices*       src/bvh.rs:320-450
//   rtbvh::SpatialTriangle helpers : intersect(tri, ray) / intersect4(tri, packet, t_min)        src/builders/spatial_sah.rs:131-244
//   traverse_iter / traverse_iter_packet : the `&T` iterators of src/iter.rs over the index iterators
//   Bvh::from_raw / into_raw, Mbvh::from_raw / construct_from_raw / into_raw : trees that live in caller memory   src/bvh.rs:246-256, 353-416
//   rtbvh::Scene : the batched GPU traversal (closest hit / any hit) that replaces the per-ray iterator loops
//
// Header only; link with -lrtbvh_rs (rtbvh_b200/librtbvh_rs.so).  Builds run on the GPU; the iterators walk the
// host mirror exactly like the reference's (rtbvh_iter.hpp).
#pragma once
#ifndef RTBVH_HPP  // the include guard of the reference's generated C++ header (rtbvh_ffi/build.rs:37)
#define RTBVH_HPP
#endif
#include <cmath>
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <utility>
#include <variant>
#include <vector>

#include "rtbvh_gpu.h"
#include "rtbvh_iter.hpp"

namespace rtbvh {

// The reference also generates a C++ flavour of its FFI header (rtbvh_ffi/build.rs:33-45: cbindgen Language::Cxx,
// namespace `rtbvh`, guard RTBVH_HPP), whose callers spell the C ABI as rtbvh::create_bvh(..), rtbvh::RTBvh,
// rtbvh::ResultCode::Ok, rtbvh::BvhType::BinnedSAH.  The same spellings resolve here (the enums are the C header's
// unscoped ones, so the qualified names work and the values are the same).
using ::BvhType;
using ::ResultCode;
using ::RTAabb;
using ::RTBvh;
using ::RTBvhNode;
using ::RTMbvh;
using ::RTMbvhNode;
using ::create_bvh;
using ::create_mbvh;
using ::create_spatial_Bvh;
using ::free_bvh;
using ::free_mbvh;
using ::intersect;
using ::intersect_mbvh;
using ::intersect_mbvh_packet;
using ::intersect_packet;
using ::refit;

struct Vec3 {
    float x = 0, y = 0, z = 0;
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }

// rtbvh::Aabb<i32> (src/aabb.rs:13-20): the C POD plus the reference's helpers
struct alignas(16) Aabb : RTAabb {
    Aabb() {  // Aabb::new / empty (aabb.rs:50-57)
        for (int k = 0; k < 3; k++) {
            min[k] = 1e34f;
            max[k] = -1e34f;
        }
        count = 0;
        left_first = 0;
    }
    static Aabb empty() { return Aabb(); }
    void grow(Vec3 p) {  // aabb.rs:252-262
        const float v[3] = {p.x, p.y, p.z};
        for (int k = 0; k < 3; k++) {
            min[k] = std::fmin(min[k], v[k]);
            max[k] = std::fmax(max[k], v[k]);
        }
    }
    void grow_bb(const RTAabb& b) {  // aabb.rs:264-273
        for (int k = 0; k < 3; k++) {
            min[k] = std::fmin(min[k], b.min[k]);
            max[k] = std::fmax(max[k], b.max[k]);
        }
    }
    void offset_by(float d) {  // aabb.rs:313-321
        for (int k = 0; k < 3; k++) {
            min[k] -= d;
            max[k] += d;
        }
    }
    Vec3 center() const { return Vec3{(min[0] + max[0]) * 0.5f, (min[1] + max[1]) * 0.5f, (min[2] + max[2]) * 0.5f}; }
    float half_area() const {  // aabb.rs:343-346
        const float dx = max[0] - min[0], dy = max[1] - min[1], dz = max[2] - min[2];
        return (dx + dy) * dz + dx * dy;
    }
    bool is_valid() const { return min[0] <= max[0] && min[1] <= max[1] && min[2] <= max[2]; }
    bool contains(Vec3 p) const {  // strict, aabb.rs:246-249
        return p.x > min[0] && p.y > min[1] && p.z > min[2] && p.x < max[0] && p.y < max[1] && p.z < max[2];
    }
};
static_assert(sizeof(Aabb) == 32, "Aabb == RTAabb (rtbvh_ffi same_size test)");

// aabb!(v0, v1, ...) macro of the reference (aabb.rs:472-482): grow + 1e-4 pad
template <class... V>
inline Aabb aabb_of(V... v) {
    Aabb bb;
    (bb.grow(v), ...);
    bb.offset_by(1e-4f);
    return bb;
}

using Ray = rtbvh_host::Ray;                // src/ray.rs:9-16 (Ray::make == Ray::new)
using RayPacket4 = rtbvh_host::RayPacket4;  // src/ray.rs:47-61

// RayPacket4::new (src/ray.rs:64-107) from SoA lanes; t starts at 1e34 like the reference's
inline RayPacket4 make_packet(const float ox[4], const float oy[4], const float oz[4], const float dx[4], const float dy[4],
                              const float dz[4]) {
    RayPacket4 p;
    const float* o[3] = {ox, oy, oz};
    const float* d[3] = {dx, dy, dz};
    for (int k = 0; k < 3; k++)
        for (int l = 0; l < 4; l++) {
            p.origin[k][l] = o[k][l];
            p.direction[k][l] = d[k][l];
            p.inv_direction[k][l] = 1.0f / d[k][l];
        }
    for (int l = 0; l < 4; l++) p.t[l] = 1e34f;
    return p;
}

// The `&T` iterators of src/iter.rs (BvhIterator :22-129, BvhPacketIterator :131-244, MbvhIterator :246-356,
// MbvhPacketIterator :358-469): the index iterator with the primitive looked up.  `bool next(const T** prim)` ==
// `next() -> Option<(&T, &mut Ray)>`.  REJECT_EMPTY: the Bvh flavours yield nothing for an empty primitive slice
// (iter.rs:36-44, :145-158); the Mbvh flavours do not look at it (iter.rs:262-280).
template <class IndexIt, class T, bool REJECT_EMPTY>
class PrimIterator {
  public:
    PrimIterator(IndexIt it, const T* prims, size_t count) : it_(it), prims_(prims), live_(!REJECT_EMPTY || count > 0) {}
    bool next(const T** prim) {
        uint32_t id;
        if (!live_ || !it_.next(&id)) return false;
        *prim = prims_ + id;
        return true;
    }

  private:
    IndexIt it_;
    const T* prims_;
    bool live_;
};

enum class BuildType { None, LocallyOrderedClustered, BinnedSAH, Spatial };  // src/bvh.rs:18-23

struct BuildError {  // src/bvh.rs:26-29
    enum Kind { NoPrimitives, InequalAabbsAndPrimitives, Device } kind;
    size_t aabbs = 0, primitives = 0;
    bool operator==(const BuildError& o) const { return kind == o.kind && aabbs == o.aabbs && primitives == o.primitives; }
};

template <class T>
class Result {  // Result<T, BuildError>
  public:
    Result(T v) : v_(std::move(v)) {}
    Result(BuildError e) : v_(e) {}
    bool is_ok() const { return std::holds_alternative<T>(v_); }
    T unwrap() {
        if (!is_ok()) throw std::runtime_error("called unwrap() on an Err value");
        return std::move(std::get<T>(v_));
    }
    BuildError unwrap_err() const {
        if (is_ok()) throw std::runtime_error("called unwrap_err() on an Ok value");
        return std::get<BuildError>(v_);
    }

  private:
    std::variant<T, BuildError> v_;
};

class Mbvh;

// rtbvh::Bvh (src/bvh.rs:143-284): owns an entry of the library's table (freed on destruction).
class Bvh {
  public:
    Bvh() = default;
    Bvh(RTBvh rt, BuildType t) : rt_(rt), type_(t), owned_(true) {}
    Bvh(Bvh&& o) noexcept { *this = std::move(o); }
    Bvh& operator=(Bvh&& o) noexcept {
        release();
        rt_ = o.rt_;
        type_ = o.type_;
        owned_ = o.owned_;
        o.owned_ = false;
        o.rt_ = RTBvh{UINT32_MAX, 0, nullptr, 0, nullptr};
        return *this;
    }
    Bvh(const Bvh&) = delete;
    Bvh& operator=(const Bvh&) = delete;
    ~Bvh() { release(); }

    const RTBvhNode* nodes() const { return rt_.nodes; }
    size_t node_count() const { return rt_.node_count; }
    const uint32_t* indices() const { return rt_.indices; }
    size_t prim_count() const { return rt_.index_count; }
    BuildType build_type() const { return type_; }
    RTBvh raw() const { return rt_; }
    Aabb bounds() const {  // Bounds for Bvh (bvh.rs:452-459)
        Aabb b;
        if (rt_.node_count) std::memcpy(static_cast<RTAabb*>(&b), &rt_.nodes[0].aabb, sizeof(RTAabb));
        return b;
    }
    void refit(const Aabb* new_aabbs) {  // bvh.rs:176-205 (GPU); topology fields are kept, see DESIGN.md
        if (::refit(new_aabbs, rt_) != Ok) throw std::runtime_error("refit failed");
    }
    bool validate(size_t prims) const {  // bvh.rs:232-244
        if (!rt_.node_count) return false;
        std::vector<uint8_t> found(prims, 0);
        std::vector<int32_t> st{0};
        while (!st.empty()) {
            const RTAabb& n = rt_.nodes[st.back()].aabb;
            st.pop_back();
            if (n.left_first < 0) continue;
            if (n.count >= 0) {
                for (int32_t i = 0; i < n.count; i++) {
                    const uint32_t p = rt_.indices[n.left_first + i];
                    if (p >= prims) return false;
                    found[p] = 1;
                }
            } else {
                st.push_back(n.left_first);
                st.push_back(n.left_first + 1);
            }
        }
        for (uint8_t f : found)
            if (!f) return false;
        return true;
    }
    // IntoRayIndexIterator / IntoPacketIndexIterator (iter_indices.rs:6-16): `while (it.next(&prim)) { ... }`
    rtbvh_host::BvhIndexIterator traverse_iter_indices(Ray& ray) const { return {&ray, rt_.nodes, rt_.node_count, rt_.indices}; }
    rtbvh_host::BvhPacketIndexIterator traverse_iter_indices_packet(RayPacket4& p) const {
        return {&p, rt_.nodes, rt_.node_count, rt_.indices};
    }
    // IntoRayIterator / IntoPacketIterator (iter.rs:9-19)
    template <class T>
    PrimIterator<rtbvh_host::BvhIndexIterator, T, true> traverse_iter(Ray& ray, const T* prims, size_t count) const {
        return {traverse_iter_indices(ray), prims, count};
    }
    template <class T>
    PrimIterator<rtbvh_host::BvhPacketIndexIterator, T, true> traverse_iter_packet(RayPacket4& p, const T* prims, size_t count) const {
        return {traverse_iter_indices_packet(p), prims, count};
    }
    // A tree that lives in the caller's memory (deserialised, or built by the reference): not owned, never freed here.
    // The struct is trusted like the reference's intersect* trust theirs (rtbvh_ffi/src/lib.rs:551-581).
    static Bvh from_raw(const RTBvhNode* nodes, size_t node_count, const uint32_t* indices, size_t index_count,
                        BuildType type = BuildType::None) {
        Bvh b;
        b.rt_ = RTBvh{UINT32_MAX, (uint32_t)node_count, nodes, (uint32_t)index_count, indices};
        b.type_ = type;
        return b;
    }
    // into_raw (bvh.rs:246-256): the arrays, copied out of the library's mirror
    std::pair<std::vector<RTBvhNode>, std::vector<uint32_t>> into_raw() const {
        return {std::vector<RTBvhNode>(rt_.nodes, rt_.nodes + rt_.node_count),
                std::vector<uint32_t>(rt_.indices, rt_.indices + rt_.index_count)};
    }

  private:
    void release() {
        if (owned_) free_bvh(rt_);
        owned_ = false;
    }
    RTBvh rt_{UINT32_MAX, 0, nullptr, 0, nullptr};
    BuildType type_ = BuildType::None;
    bool owned_ = false;
};

// rtbvh::Mbvh (src/bvh.rs:320-450)
class Mbvh {
  public:
    Mbvh() = default;
    explicit Mbvh(const Bvh& bvh) { *this = construct(bvh); }  // From<Bvh>
    static Mbvh construct(const Bvh& bvh) {                    // bvh.rs:381-404
        Mbvh m;
        if (bvh.node_count() == 0) return m;
        if (create_mbvh(bvh.raw(), &m.rt_) != Ok) throw std::runtime_error("create_mbvh failed");
        m.owned_ = true;
        return m;
    }
    Mbvh(Mbvh&& o) noexcept { *this = std::move(o); }
    Mbvh& operator=(Mbvh&& o) noexcept {
        release();
        rt_ = o.rt_;
        owned_ = o.owned_;
        o.owned_ = false;
        o.rt_ = RTMbvh{UINT32_MAX, 0, nullptr, 0, nullptr};
        return *this;
    }
    Mbvh(const Mbvh&) = delete;
    Mbvh& operator=(const Mbvh&) = delete;
    ~Mbvh() { release(); }

    const RTMbvhNode* quad_nodes() const { return rt_.nodes; }
    size_t quad_node_count() const { return rt_.node_count; }
    const uint32_t* indices() const { return rt_.indices; }
    size_t prim_count() const { return rt_.index_count; }
    RTMbvh raw() const { return rt_; }
    rtbvh_host::MbvhIndexIterator traverse_iter_indices(Ray& ray) const { return {&ray, rt_.nodes, rt_.node_count, rt_.indices}; }
    rtbvh_host::MbvhPacketIndexIterator traverse_iter_indices_packet(RayPacket4& p) const {
        return {&p, rt_.nodes, rt_.node_count, rt_.indices};
    }
    template <class T>
    PrimIterator<rtbvh_host::MbvhIndexIterator, T, false> traverse_iter(Ray& ray, const T* prims, size_t count) const {
        return {traverse_iter_indices(ray), prims, count};
    }
    template <class T>
    PrimIterator<rtbvh_host::MbvhPacketIndexIterator, T, false> traverse_iter_packet(RayPacket4& p, const T* prims, size_t count) const {
        return {traverse_iter_indices_packet(p), prims, count};
    }
    // 4-wide nodes that live in the caller's memory: not owned
    static Mbvh from_raw(const RTMbvhNode* nodes, size_t node_count, const uint32_t* indices, size_t index_count) {
        Mbvh m;
        m.rt_ = RTMbvh{UINT32_MAX, (uint32_t)node_count, nodes, (uint32_t)index_count, indices};
        return m;
    }
    // Mbvh::construct_from_raw (bvh.rs:353-379) for binary nodes in caller memory, collapsed on the GPU.  Always the
    // merge_nodes path: the reference's special case for <= 4 nodes mislabels inner nodes.
    static Mbvh construct_from_raw(const RTBvhNode* nodes, size_t node_count, const uint32_t* indices, size_t index_count) {
        Mbvh m;
        if (node_count == 0) return m;
        const RTBvh src{UINT32_MAX, (uint32_t)node_count, nodes, (uint32_t)index_count, indices};
        if (rtbvh_gpu_create_mbvh_from(&src, &m.rt_) != Ok) throw std::runtime_error(rtbvh_gpu_last_error());
        m.owned_ = true;
        return m;
    }
    std::pair<std::vector<RTMbvhNode>, std::vector<uint32_t>> into_raw() const {  // bvh.rs:406-416
        return {std::vector<RTMbvhNode>(rt_.nodes, rt_.nodes + rt_.node_count),
                std::vector<uint32_t>(rt_.indices, rt_.indices + rt_.index_count)};
    }
  private:
    void release() {
        if (owned_) free_mbvh(rt_);
        owned_ = false;
    }
    RTMbvh rt_{UINT32_MAX, 0, nullptr, 0, nullptr};
    bool owned_ = false;
};

// rtbvh::Builder (src/bvh.rs:49-138).  `aabbs` is Option<&[Aabb]>: std::nullopt -> computed from Primitive::aabb().
template <class T>
struct Builder {
    std::optional<std::pair<const Aabb*, size_t>> aabbs;
    const T* primitives = nullptr;
    size_t primitive_count = 0;
    size_t primitives_per_leaf = 0;  // Option<NonZeroUsize>: 0 == None

    Result<Bvh> construct_binned_sah() const { return construct(BinnedSAH, BuildType::BinnedSAH); }
    Result<Bvh> construct_locally_ordered_clustered() const {
        return construct(LocallyOrderedClustered, BuildType::LocallyOrderedClustered);
    }

  private:
    Result<Bvh> construct(uint32_t kind, BuildType type) const {
        if (primitive_count == 0) return BuildError{BuildError::NoPrimitives};
        if (aabbs && aabbs->second != primitive_count)
            return BuildError{BuildError::InequalAabbsAndPrimitives, aabbs->second, primitive_count};
        // gather Primitive::center() (and aabb() when none were given) into the flat arrays the C ABI takes
        std::vector<float> centers(primitive_count * 3);
        std::vector<Aabb> own;
        if (!aabbs) own.resize(primitive_count);
        for (size_t i = 0; i < primitive_count; i++) {
            const Vec3 c = primitives[i].center();
            centers[3 * i] = c.x;
            centers[3 * i + 1] = c.y;
            centers[3 * i + 2] = c.z;
            if (!aabbs) own[i] = primitives[i].aabb();
        }
        RTBvh out{UINT32_MAX, 0, nullptr, 0, nullptr};
        const RTAabb* bb = aabbs ? aabbs->first : own.data();
        const ResultCode rc = create_bvh(bb, primitive_count, centers.data(), 12, primitives_per_leaf, (BvhType)kind, &out);
        if (rc == NoPrimitives) return BuildError{BuildError::NoPrimitives};
        if (rc != Ok) return BuildError{BuildError::Device};
        return Bvh(out, type);
    }
};

// SpatialTriangle::intersect (src/builders/spatial_sah.rs:131-163) for any T with vertex0/1/2() -> Vec3
template <class T>
inline bool intersect(const T& tri, Ray& ray) {
    const Vec3 v0 = tri.vertex0(), v1 = tri.vertex1(), v2 = tri.vertex2();
    const Vec3 d{ray.direction[0], ray.direction[1], ray.direction[2]}, o{ray.origin[0], ray.origin[1], ray.origin[2]};
    auto cross = [](Vec3 a, Vec3 b) { return Vec3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; };
    auto dot = [](Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; };
    const Vec3 e1 = v1 - v0, e2 = v2 - v0, h = cross(d, e2);
    const float a = dot(e1, h);
    if (a > -1e-5f && a < 1e-5f) return false;
    const float f = 1.0f / a;
    const Vec3 s = o - v0;
    const float u = f * dot(s, h);
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    const Vec3 q = cross(s, e1);
    const float v = f * dot(d, q);
    if (v < 0.0f || (u + v) > 1.0f) return false;
    const float t = f * dot(e2, q);
    if (t > ray.t_min && t < ray.t) {
        ray.t = t;
        return true;
    }
    return false;
}

// SpatialTriangle::intersect4 (src/builders/spatial_sah.rs:165-244): four lanes, its own constants (determinant eps 1e-6
// with a <= -eps | a >= eps; t >= t_min, non-strict; t < packet.t), `packet.t = select(mask, t, packet.t)`.  Returns the
// lane mask (0 == None).  The lanes are independent, so the per-lane loop equals the reference's SSE code bit for bit
// (compile without FMA contraction: -ffp-contract=off).
template <class T>
inline unsigned intersect4(const T& tri, RayPacket4& p, const float t_min[4]) {
    const Vec3 v0 = tri.vertex0(), v1 = tri.vertex1(), v2 = tri.vertex2();
    const float e1x = v1.x - v0.x, e1y = v1.y - v0.y, e1z = v1.z - v0.z;
    const float e2x = v2.x - v0.x, e2y = v2.y - v0.y, e2z = v2.z - v0.z;
    unsigned mask = 0;
    for (int l = 0; l < 4; l++) {
        const float dx = p.direction[0][l], dy = p.direction[1][l], dz = p.direction[2][l];
        const float hx = (dy * e2z) - (dz * e2y), hy = (dz * e2x) - (dx * e2z), hz = (dx * e2y) - (dy * e2x);
        const float a = ((e1x * hx) + (e1y * hy)) + (e1z * hz);
        if (!(a <= -1e-6f || a >= 1e-6f)) continue;
        const float f = 1.0f / a;
        const float sx = p.origin[0][l] - v0.x, sy = p.origin[1][l] - v0.y, sz = p.origin[2][l] - v0.z;
        const float u = f * (((sx * hx) + (sy * hy)) + (sz * hz));
        if (!(u >= 0.0f && u <= 1.0f)) continue;
        const float qx = sy * e1z - sz * e1y, qy = sz * e1x - sx * e1z, qz = sx * e1y - sy * e1x;
        const float v = f * (((dx * qx) + (dy * qy)) + (dz * qz));
        if (!(v >= 0.0f && (u + v) <= 1.0f)) continue;
        const float t = f * (((e2x * qx) + (e2y * qy)) + (e2z * qz));
        if (!(t >= t_min[l] && t < p.t[l])) continue;
        p.t[l] = t;
        mask |= 1u << l;
    }
    return mask;
}

// The batched GPU traversal: what `for (prim, ray) in tree.iter(ray) { prim.intersect(ray) }` becomes on a B200.
class Scene {
  public:
    // vertices: 3 * triangle_count vertices, vertex_stride bytes apart (12 or 16)
    Scene(const Bvh* bvh, const Mbvh* mbvh, const float* vertices, size_t vertex_stride, size_t triangle_count) {
        RTBvh b{};
        RTMbvh m{};
        if (bvh) b = bvh->raw();
        if (mbvh) m = mbvh->raw();
        if (rtbvh_gpu_scene_create(bvh ? &b : nullptr, mbvh ? &m : nullptr, vertices, vertex_stride, triangle_count, &h_) != Ok)
            throw std::runtime_error(rtbvh_gpu_last_error());
    }
    // Builder{aabbs: None, primitives}.construct_* + Mbvh::from, built and kept on the device (no host mirror)
    static Scene build(const float* vertices, size_t vertex_stride, size_t triangle_count, BvhType type = BinnedSAH,
                       size_t primitives_per_leaf = 1, bool with_mbvh = true) {
        Scene s;
        if (rtbvh_gpu_scene_build(vertices, vertex_stride, triangle_count, primitives_per_leaf, type, with_mbvh ? 1 : 0, &s.h_) != Ok)
            throw std::runtime_error(rtbvh_gpu_last_error());
        return s;
    }
    ~Scene() {
        if (h_) rtbvh_gpu_scene_free(h_);
    }
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;
    Scene(Scene&& o) noexcept : h_(o.h_) { o.h_ = 0; }
    // Multi-GPU: build once, replicate device to device.  export_handle() / import_handle(): one process per GPU (ship the 512
    // bytes over MPI / a pipe; keep this scene alive until every importer has returned); clone_to(): one process, several GPUs.
    RTGpuSceneExport export_handle() const {
        RTGpuSceneExport x;
        if (rtbvh_gpu_scene_export(h_, &x) != Ok) throw std::runtime_error(rtbvh_gpu_last_error());
        return x;
    }
    static Scene import_handle(const RTGpuSceneExport& x) {
        Scene s;
        if (rtbvh_gpu_scene_import(&x, &s.h_) != Ok) throw std::runtime_error(rtbvh_gpu_last_error());
        return s;
    }
    Scene clone_to(int device) const {
        Scene s;
        if (rtbvh_gpu_scene_clone(h_, device, &s.h_) != Ok) throw std::runtime_error(rtbvh_gpu_last_error());
        return s;
    }
    // Bvh::refit for a resident scene: same triangles, new positions; also refreshes the Mbvh and the triangle records
    void refit(const float* vertices, size_t vertex_stride, size_t triangle_count) {
        if (rtbvh_gpu_scene_refit(h_, vertices, vertex_stride, triangle_count) != Ok) throw std::runtime_error(rtbvh_gpu_last_error());
    }
    // submit / wait: the buffers must stay alive and untouched until wait(ticket) returns (page-locked ones make it asynchronous)
    uint64_t intersect_async(const RTRay* rays, size_t n, RTHit* hits, RTTreeKind tree = RT_TREE_MBVH) {
        uint64_t ticket = 0;
        if (rtbvh_gpu_intersect_async(h_, tree, rays, n, hits, &ticket) != Ok) throw std::runtime_error(rtbvh_gpu_last_error());
        return ticket;
    }
    void wait(uint64_t ticket = 0) {
        if (rtbvh_gpu_wait(h_, ticket) != Ok) throw std::runtime_error(rtbvh_gpu_last_error());
    }
    std::vector<RTMbvhNode> read_mbvh_nodes() const {
        uint32_t n = 0;
        if (rtbvh_gpu_scene_tree_size(h_, RT_TREE_MBVH, &n, nullptr) != Ok) throw std::runtime_error(rtbvh_gpu_last_error());
        std::vector<RTMbvhNode> out(n);
        if (rtbvh_gpu_scene_read_nodes(h_, RT_TREE_MBVH, out.data(), out.size() * sizeof(RTMbvhNode)) != Ok)
            throw std::runtime_error(rtbvh_gpu_last_error());
        return out;
    }
    std::vector<RTHit> intersect(const std::vector<RTRay>& rays, RTTreeKind tree = RT_TREE_MBVH) const {
        std::vector<RTHit> hits(rays.size());
        if (rtbvh_gpu_intersect(h_, tree, rays.data(), rays.size(), hits.data()) != Ok) throw std::runtime_error(rtbvh_gpu_last_error());
        return hits;
    }
    std::vector<uint8_t> occluded(const std::vector<RTRay>& rays, RTTreeKind tree = RT_TREE_MBVH) const {
        std::vector<uint8_t> occ(rays.size());
        if (rtbvh_gpu_occluded(h_, tree, rays.data(), rays.size(), occ.data()) != Ok) throw std::runtime_error(rtbvh_gpu_last_error());
        return occ;
    }
    // packets of four rays, SpatialTriangle::intersect4 semantics; pass t_min = 1e-4f for examples/benchmark.rs:58
    std::vector<RTHitPacket4> intersect_packets(const std::vector<RTRayPacket4>& packets, float t_min = 1e-4f,
                                                RTTreeKind tree = RT_TREE_MBVH) const {
        std::vector<RTHitPacket4> hits(packets.size());
        if (rtbvh_gpu_intersect_packets(h_, tree, packets.data(), packets.size(), t_min, hits.data()) != Ok)
            throw std::runtime_error(rtbvh_gpu_last_error());
        return hits;
    }
    std::vector<uint8_t> occluded_packets(const std::vector<RTRayPacket4>& packets, float t_min = 1e-4f,
                                          RTTreeKind tree = RT_TREE_MBVH) const {
        std::vector<uint8_t> occ(4 * packets.size());
        if (rtbvh_gpu_occluded_packets(h_, tree, packets.data(), packets.size(), t_min, occ.data()) != Ok)
            throw std::runtime_error(rtbvh_gpu_last_error());
        return occ;
    }
    // split input: packed origins / directions (3 floats per ray), one t_min / initial t for the batch
    std::vector<RTHit> intersect_od(const float* origins, const float* directions, size_t n, float t_min = 1e-4f,
                                    float t_max = 1e34f, RTTreeKind tree = RT_TREE_MBVH) const {
        std::vector<RTHit> hits(n);
        if (rtbvh_gpu_intersect_od(h_, tree, origins, directions, n, t_min, t_max, hits.data()) != Ok)
            throw std::runtime_error(rtbvh_gpu_last_error());
        return hits;
    }
    void set_ray_sorting(bool on) { rtbvh_gpu_scene_set_ray_sorting(h_, on ? 1 : 0); }

  private:
    Scene() = default;
    RTGpuScene h_ = 0;
};

}  // namespace rtbvh
