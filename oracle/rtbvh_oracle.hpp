// rtbvh_oracle.hpp — CPU restatement of meirbon/rtbvh's hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.  The product
// (rtbvh_b200/csrc) never includes, links or calls anything in oracle/.
//
// Why a restatement: the reference is Rust (crate rtbvh 0.6.2) and this image has no
// rustc/cargo, so the reference cannot be compiled here (oracle/_ref is therefore
// absent).  Every function below cites the reference file:line it
// follows.  Arithmetic is fp32, one rounding per operation: build with
// -ffp-contract=off (Rust never contracts to FMA).  Vec4 paths use SSE intrinsics
// because the reference's glam::Vec4 *is* __m128 on x86-64, including the
// min/max NaN rule (returns the 2nd operand).  glam itself is an un-vendored,
// unpinned dependency (Cargo.toml:14, ">=0.14"); its cross/dot/normalize are restated
// from its published scalar definitions:
//     cross = (y*rz - ry*z, z*rx - rz*x, x*ry - rx*y),  dot = (x*rx + y*ry) + z*rz.
//
// Parity pins: the reference's own KATs (morton_split, prefix_sum, same_size, the FFI
// `intersect` t==1.0 vector, create_delete / invalid-input codes, teapot structural
// checks, five_triangle no-panic) are ported in tests/test_oracle_kats.py.  Hit ids,
// packets, any-hit, SAH values and refit are NOT pinned by any reference test:
// for those "parity unpinned" — this restatement is the definition, cross-checked
// against brute force and against a second restatement written separately from the Rust
// sources in plain Python (tests/test_oracle_second_opinion*.py: bit-identical t / ids for the
// eight traversal flavours, byte-identical trees for binned SAH, LOCB and the collapse).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <emmintrin.h>
#include <limits>
#include <vector>
#include <xmmintrin.h>

namespace rto {

// ----------------------------------------------------------------------------------
// L0 value types
// ----------------------------------------------------------------------------------
struct Vec3 {
    float x, y, z;
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
static inline Vec3 v3(float x, float y, float z) { return Vec3{x, y, z}; }
static inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline Vec3 operator*(Vec3 a, Vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }
// glam scalar Vec3::cross / dot (see header comment)
static inline Vec3 cross(Vec3 a, Vec3 b) {
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
static inline float dot(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline bool is_nan(Vec3 a) { return std::isnan(a.x) || std::isnan(a.y) || std::isnan(a.z); }
// Rust f32::min / f32::max: if one argument is NaN the other is returned (== fminf/fmaxf)
static inline float f32_min(float a, float b) { return fminf(a, b); }
static inline float f32_max(float a, float b) { return fmaxf(a, b); }

// Aabb<i32>: aabb.rs:13-20.  32 bytes, 16-aligned.  As a BvhNode (bvh_node.rs:11-14):
// extra1 = count (>=0 leaf, -1 inner), extra2 = left_first (-1 invalid).
struct alignas(16) Aabb {
    float min[3];
    int32_t extra1;
    float max[3];
    int32_t extra2;
};
static_assert(sizeof(Aabb) == 32, "rtbvh_ffi same_size test: Aabb == RTAabb == 32 B");
using BvhNode = Aabb;

// aabb.rs:50-57
static inline Aabb aabb_new() { return Aabb{{1e34f, 1e34f, 1e34f}, 0, {-1e34f, -1e34f, -1e34f}, 0}; }
static inline Vec3 amin(const Aabb& a) { return {a.min[0], a.min[1], a.min[2]}; }
static inline Vec3 amax(const Aabb& a) { return {a.max[0], a.max[1], a.max[2]}; }
// aabb.rs:252-262
static inline void grow(Aabb& a, Vec3 p) {
    for (int i = 0; i < 3; i++) {
        a.min[i] = f32_min(a.min[i], p[i]);
        a.max[i] = f32_max(a.max[i], p[i]);
    }
}
// aabb.rs:264-273 (extras untouched)
static inline void grow_bb(Aabb& a, const Aabb& b) {
    for (int i = 0; i < 3; i++) {
        a.min[i] = f32_min(a.min[i], b.min[i]);
        a.max[i] = f32_max(a.max[i], b.max[i]);
    }
}
// aabb.rs:73-85 (extras reset to Default = 0)
static inline Aabb union_of(const Aabb& a, const Aabb& b) {
    Aabb r;
    for (int i = 0; i < 3; i++) {
        r.min[i] = f32_min(a.min[i], b.min[i]);
        r.max[i] = f32_max(a.max[i], b.max[i]);
    }
    r.extra1 = 0;
    r.extra2 = 0;
    return r;
}
// aabb.rs:300-321
static inline void offset_by(Aabb& a, float delta) {
    for (int i = 0; i < 3; i++) {
        a.min[i] = a.min[i] - delta;
        a.max[i] = a.max[i] + delta;
    }
}
// aabb.rs:125-131
static inline Aabb union_of_list(const Aabb* aabbs, size_t n) {
    Aabb a = aabb_new();
    for (size_t i = 0; i < n; i++) grow_bb(a, aabbs[i]);
    offset_by(a, 0.0001f);
    return a;
}
// aabb.rs:343-346, 441-443
static inline Vec3 diagonal(const Aabb& a) { return amax(a) - amin(a); }
static inline float half_area(const Aabb& a) {
    Vec3 d = diagonal(a);
    return (d.x + d.y) * d.z + d.x * d.y;
}
// aabb.rs:354-363
static inline int longest_axis(const Aabb& a) {
    int ax = 0;
    if ((a.max[1] - a.min[1]) > (a.max[0] - a.min[0])) ax = 1;
    if ((a.max[2] - a.min[2]) > (a.max[ax] - a.min[ax])) ax = 2;
    return ax;
}
// aabb.rs:325-327
static inline Vec3 center(const Aabb& a) { return (amin(a) + amax(a)) * 0.5f; }
// aabb.rs:141-143
static inline bool is_valid(const Aabb& a) {
    return a.min[0] <= a.max[0] && a.min[1] <= a.max[1] && a.min[2] <= a.max[2];
}
// aabb.rs:246-249
static inline bool contains(const Aabb& a, Vec3 p) {
    return p.x > a.min[0] && p.y > a.min[1] && p.z > a.min[2] && p.x < a.max[0] && p.y < a.max[1] && p.z < a.max[2];
}

// mbvh_node.rs:28-38.  128 bytes.  (Reference alignment is 4 but it is read with aligned
// _mm_load_ps, mbvh_node.rs:87 — we make the 16-byte requirement explicit.)
struct alignas(16) MbvhNode {
    float min_x[4], max_x[4], min_y[4], max_y[4], min_z[4], max_z[4];
    int32_t children[4];
    int32_t counts[4];
};
static_assert(sizeof(MbvhNode) == 128, "rtbvh_ffi same_size test: MbvhNode == 128 B");

// mbvh_node.rs:56-80
static inline MbvhNode mbvh_node_new() {
    MbvhNode n;
    for (int i = 0; i < 4; i++) {
        n.min_x[i] = n.min_y[i] = n.min_z[i] = 1e34f;
        n.max_x[i] = n.max_y[i] = n.max_z[i] = -1e34f;
        n.children[i] = -1;
        n.counts[i] = -1;
    }
    return n;
}
// mbvh_node.rs:162-175
static inline void set_bounds_bb(MbvhNode& n, int slot, const Aabb& b) {
    n.min_x[slot] = b.min[0];
    n.min_y[slot] = b.min[1];
    n.min_z[slot] = b.min[2];
    n.max_x[slot] = b.max[0];
    n.max_y[slot] = b.max[1];
    n.max_z[slot] = b.max[2];
}

// ray.rs:9-16, 166-182
struct Ray {
    Vec3 origin;
    float t_min;
    Vec3 direction;
    float t;
    Vec3 inv_direction;
    uint8_t signs[4];
};
static inline Ray ray_new(Vec3 o, Vec3 d) {
    Ray r;
    r.origin = o;
    r.direction = d;
    r.t_min = 1e-4f;
    r.t = 1e34f;
    r.inv_direction = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
    r.signs[0] = d.x < 0.0f;
    r.signs[1] = d.y < 0.0f;
    r.signs[2] = d.z < 0.0f;
    r.signs[3] = 0;
    return r;
}

// ray.rs:47-107
struct alignas(16) RayPacket4 {
    __m128 origin_x, origin_y, origin_z;
    __m128 direction_x, direction_y, direction_z;
    __m128 inv_direction_x, inv_direction_y, inv_direction_z;
    __m128 t;
};
static inline bool m128_is_nan(__m128 v) { return _mm_movemask_ps(_mm_cmpunord_ps(v, v)) != 0; }
// what rtbvh_ffi/src/lib.rs:657-668 builds from SoA inputs (inv = ONE / dir)
static inline RayPacket4 packet_new(const float* ox, const float* oy, const float* oz, const float* dx,
                                    const float* dy, const float* dz, const float* t) {
    RayPacket4 p;
    p.origin_x = _mm_loadu_ps(ox);
    p.origin_y = _mm_loadu_ps(oy);
    p.origin_z = _mm_loadu_ps(oz);
    p.direction_x = _mm_loadu_ps(dx);
    p.direction_y = _mm_loadu_ps(dy);
    p.direction_z = _mm_loadu_ps(dz);
    const __m128 one = _mm_set1_ps(1.0f);
    p.inv_direction_x = _mm_div_ps(one, p.direction_x);
    p.inv_direction_y = _mm_div_ps(one, p.direction_y);
    p.inv_direction_z = _mm_div_ps(one, p.direction_z);
    p.t = _mm_loadu_ps(t);
    return p;
}

// ----------------------------------------------------------------------------------
// L1 node tests
// ----------------------------------------------------------------------------------
// aabb.rs:146-181.  Returns true + *key = t_far when hit.  Never looks at ray.t (quirk Q1).
static inline bool aabb_intersect(const Aabb& b, const Ray& r, float* key) {
    const float* lo[2] = {b.min, b.max};
    float ray_min = (lo[r.signs[0]][0] - r.origin.x) * r.inv_direction.x;
    float ray_max = (lo[1 - r.signs[0]][0] - r.origin.x) * r.inv_direction.x;
    float y_min = (lo[r.signs[1]][1] - r.origin.y) * r.inv_direction.y;
    float y_max = (lo[1 - r.signs[1]][1] - r.origin.y) * r.inv_direction.y;
    if ((ray_min > y_max) || (y_min > ray_max)) return false;
    if (y_min > ray_min) ray_min = y_min;
    if (y_max < ray_max) ray_max = y_max;
    float z_min = (lo[r.signs[2]][2] - r.origin.z) * r.inv_direction.z;
    float z_max = (lo[1 - r.signs[2]][2] - r.origin.z) * r.inv_direction.z;
    if ((ray_min > z_max) || (z_min > ray_max)) return false;
    if (z_max < ray_max) ray_max = z_max;
    if (ray_max > r.t_min) {
        *key = ray_max;
        return true;
    }
    return false;
}

// aabb.rs:218-244.  Lanes = rays.  Returns true + key[4] = t_near when any lane passes.
static inline bool aabb_intersect4(const Aabb& b, const RayPacket4& p, __m128* key) {
    __m128 t1_x = _mm_mul_ps(_mm_sub_ps(_mm_set1_ps(b.min[0]), p.origin_x), p.inv_direction_x);
    __m128 t1_y = _mm_mul_ps(_mm_sub_ps(_mm_set1_ps(b.min[1]), p.origin_y), p.inv_direction_y);
    __m128 t1_z = _mm_mul_ps(_mm_sub_ps(_mm_set1_ps(b.min[2]), p.origin_z), p.inv_direction_z);
    __m128 t2_x = _mm_mul_ps(_mm_sub_ps(_mm_set1_ps(b.max[0]), p.origin_x), p.inv_direction_x);
    __m128 t2_y = _mm_mul_ps(_mm_sub_ps(_mm_set1_ps(b.max[1]), p.origin_y), p.inv_direction_y);
    __m128 t2_z = _mm_mul_ps(_mm_sub_ps(_mm_set1_ps(b.max[2]), p.origin_z), p.inv_direction_z);
    __m128 t_min_x = _mm_min_ps(t1_x, t2_x), t_min_y = _mm_min_ps(t1_y, t2_y), t_min_z = _mm_min_ps(t1_z, t2_z);
    __m128 t_max_x = _mm_max_ps(t1_x, t2_x), t_max_y = _mm_max_ps(t1_y, t2_y), t_max_z = _mm_max_ps(t1_z, t2_z);
    __m128 t_min = _mm_max_ps(t_min_x, _mm_max_ps(t_min_y, t_min_z));
    __m128 t_max = _mm_min_ps(t_max_x, _mm_min_ps(t_max_y, t_max_z));
    __m128 mask = _mm_and_ps(_mm_and_ps(_mm_cmpgt_ps(t_max, _mm_setzero_ps()), _mm_cmpgt_ps(t_max, t_min)),
                             _mm_cmplt_ps(t_min, p.t));
    if (_mm_movemask_ps(mask) != 0) {
        *key = t_min;
        return true;
    }
    return false;
}

// mbvh_node.rs:12-15
struct MbvhHit {
    uint8_t ids[4];
    bool result[4];
};

// mbvh_node.rs:177-240.  Lanes = slots.  Uses ray.t at call time; no t_min / positivity test
// (quirk Q5); the 5th comparator swaps ids only (faulty network, kept verbatim).
static inline MbvhHit mbvh_intersect(const MbvhNode& n, const Ray& r) {
    __m128 ox = _mm_set1_ps(r.origin.x), oy = _mm_set1_ps(r.origin.y), oz = _mm_set1_ps(r.origin.z);
    __m128 ix = _mm_set1_ps(r.inv_direction.x), iy = _mm_set1_ps(r.inv_direction.y), iz = _mm_set1_ps(r.inv_direction.z);
    __m128 tx0 = _mm_mul_ps(_mm_sub_ps(_mm_load_ps(n.min_x), ox), ix);
    __m128 tx1 = _mm_mul_ps(_mm_sub_ps(_mm_load_ps(n.max_x), ox), ix);
    __m128 ty0 = _mm_mul_ps(_mm_sub_ps(_mm_load_ps(n.min_y), oy), iy);
    __m128 ty1 = _mm_mul_ps(_mm_sub_ps(_mm_load_ps(n.max_y), oy), iy);
    __m128 tz0 = _mm_mul_ps(_mm_sub_ps(_mm_load_ps(n.min_z), oz), iz);
    __m128 tz1 = _mm_mul_ps(_mm_sub_ps(_mm_load_ps(n.max_z), oz), iz);
    __m128 tx_min = _mm_min_ps(tx0, tx1), tx_max = _mm_max_ps(tx0, tx1);
    __m128 ty_min = _mm_min_ps(ty0, ty1), ty_max = _mm_max_ps(ty0, ty1);
    __m128 tz_min = _mm_min_ps(tz0, tz1), tz_max = _mm_max_ps(tz0, tz1);
    __m128 t_min = _mm_max_ps(tx_min, _mm_max_ps(ty_min, tz_min));
    __m128 t_max = _mm_min_ps(tx_max, _mm_min_ps(ty_max, tz_max));
    int bits = _mm_movemask_ps(_mm_and_ps(_mm_cmpge_ps(t_max, t_min), _mm_cmplt_ps(t_min, _mm_set1_ps(r.t))));
    MbvhHit h;
    for (int i = 0; i < 4; i++) h.result[i] = (bits >> i) & 1;
    alignas(16) float k[4];
    _mm_store_ps(k, t_min);
    uint8_t ids[4] = {0, 1, 2, 3};
    if (k[0] > k[1]) { std::swap(k[0], k[1]); std::swap(ids[0], ids[1]); }
    if (k[2] > k[3]) { std::swap(k[2], k[3]); std::swap(ids[2], ids[3]); }
    if (k[0] > k[2]) { std::swap(k[0], k[2]); std::swap(ids[0], ids[2]); }
    if (k[1] > k[3]) { std::swap(k[1], k[3]); std::swap(ids[1], ids[3]); }
    if (k[2] > k[3]) { std::swap(ids[2], ids[3]); }
    for (int i = 0; i < 4; i++) h.ids[i] = ids[i];
    return h;
}

// mbvh_node.rs:243-295.  Loop over the 4 rays, lanes = slots; strict t_max > t_min; no ordering.
static inline MbvhHit mbvh_intersect4(const MbvhNode& n, const RayPacket4& p) {
    __m128 min_x = _mm_load_ps(n.min_x), max_x = _mm_load_ps(n.max_x);
    __m128 min_y = _mm_load_ps(n.min_y), max_y = _mm_load_ps(n.max_y);
    __m128 min_z = _mm_load_ps(n.min_z), max_z = _mm_load_ps(n.max_z);
    alignas(16) float ox[4], oy[4], oz[4], ix[4], iy[4], iz[4], pt[4];
    _mm_store_ps(ox, p.origin_x); _mm_store_ps(oy, p.origin_y); _mm_store_ps(oz, p.origin_z);
    _mm_store_ps(ix, p.inv_direction_x); _mm_store_ps(iy, p.inv_direction_y); _mm_store_ps(iz, p.inv_direction_z);
    _mm_store_ps(pt, p.t);
    __m128 result = _mm_setzero_ps();
    for (int i = 0; i < 4; i++) {
        __m128 org = _mm_set1_ps(ox[i]), dir = _mm_set1_ps(ix[i]);
        __m128 t1 = _mm_mul_ps(_mm_sub_ps(min_x, org), dir), t2 = _mm_mul_ps(_mm_sub_ps(max_x, org), dir);
        __m128 t_min = _mm_min_ps(t1, t2), t_max = _mm_max_ps(t1, t2);
        org = _mm_set1_ps(oy[i]); dir = _mm_set1_ps(iy[i]);
        t1 = _mm_mul_ps(_mm_sub_ps(min_y, org), dir); t2 = _mm_mul_ps(_mm_sub_ps(max_y, org), dir);
        t_min = _mm_max_ps(t_min, _mm_min_ps(t1, t2));
        t_max = _mm_min_ps(t_max, _mm_max_ps(t1, t2));
        org = _mm_set1_ps(oz[i]); dir = _mm_set1_ps(iz[i]);
        t1 = _mm_mul_ps(_mm_sub_ps(min_z, org), dir); t2 = _mm_mul_ps(_mm_sub_ps(max_z, org), dir);
        t_min = _mm_max_ps(t_min, _mm_min_ps(t1, t2));
        t_max = _mm_min_ps(t_max, _mm_max_ps(t1, t2));
        result = _mm_or_ps(result, _mm_and_ps(_mm_cmpgt_ps(t_max, t_min), _mm_cmplt_ps(t_min, _mm_set1_ps(pt[i]))));
    }
    int bits = _mm_movemask_ps(result);
    MbvhHit h;
    for (int i = 0; i < 4; i++) {
        h.result[i] = (bits >> i) & 1;
        h.ids[i] = (uint8_t)i;
    }
    return h;
}

// ----------------------------------------------------------------------------------
// Work counters (north-star additions: they feed the algorithmic-bytes figure)
// ----------------------------------------------------------------------------------
struct Counters {
    uint64_t node_visits = 0;   // MBVH: m-node entries incl. root; BVH2: stack pops
    uint64_t inner_visits = 0;  // BVH2: inner-node pops (child pair fetched)
    uint64_t prim_tests = 0;    // candidates yielded
    uint64_t max_stack = 0;     // deepest stack (entries)
    uint64_t overflow32 = 0;    // rays that needed > 32 entries (reference: panic / UB, quirk Q10)
    void merge(const Counters& o) {
        node_visits += o.node_visits;
        inner_visits += o.inner_visits;
        prim_tests += o.prim_tests;
        max_stack = std::max(max_stack, o.max_stack);
        overflow32 += o.overflow32;
    }
};

static constexpr int kStack = 256;  // reference: 32 (iter.rs:25); we count overflows instead of panicking

// ----------------------------------------------------------------------------------
// L4 iterators, restated as internal iteration: f(prim_id) -> true means "break".
// The user's loop body (triangle test, shrink ray.t) lives in f, exactly like the callback of
// rtbvh_ffi intersect (rtbvh_ffi/src/lib.rs:572-576).
// ----------------------------------------------------------------------------------
// iter_indices.rs:69-106 (+ ctor :32-46), bvh_node.rs:150-177
template <class F>
static inline void bvh_traverse(const BvhNode* nodes, size_t n_nodes, const uint32_t* indices, Ray& ray, F&& f,
                                Counters* c = nullptr) {
    if (n_nodes == 0 || is_nan(ray.origin) || is_nan(ray.direction)) return;
    int32_t stack[kStack];
    int sp = 0;
    stack[0] = 0;
    int max_sp = 0;
    while (sp >= 0) {
        const BvhNode& node = nodes[stack[sp]];
        sp--;
        if (c) c->node_visits++;
        int32_t count = node.extra1, left_first = node.extra2;
        if (count > -1) {
            for (int32_t i = 0; i < count; i++) {
                if (c) c->prim_tests++;
                if (f(indices[left_first + i])) goto done;
            }
        } else if (left_first > -1) {
            if (c) c->inner_visits++;
            float kl = 0.f, kr = 0.f;
            bool hl = aabb_intersect(nodes[left_first], ray, &kl);
            bool hr = aabb_intersect(nodes[left_first + 1], ray, &kr);
            if (hl && hr) {
                if (kl < kr) {
                    stack[++sp] = left_first;
                    stack[++sp] = left_first + 1;
                } else {
                    stack[++sp] = left_first + 1;
                    stack[++sp] = left_first;
                }
            } else if (hl) {
                stack[++sp] = left_first;
            } else if (hr) {
                stack[++sp] = left_first + 1;
            }
            max_sp = std::max(max_sp, sp + 1);
        }
    }
done:
    if (c) {
        c->max_stack = std::max<uint64_t>(c->max_stack, max_sp);
        if (max_sp > 32) c->overflow32++;
    }
}

// iter_indices.rs:172-209 (+ ctor :129-168), bvh_node.rs:180-211
template <class F>
static inline void bvh_traverse_packet(const BvhNode* nodes, size_t n_nodes, const uint32_t* indices, RayPacket4& p,
                                       F&& f, Counters* c = nullptr) {
    if (n_nodes == 0 || m128_is_nan(p.origin_x) || m128_is_nan(p.origin_y) || m128_is_nan(p.origin_z) ||
        m128_is_nan(p.direction_x) || m128_is_nan(p.direction_y) || m128_is_nan(p.direction_z))
        return;
    int32_t stack[kStack];
    int sp = 0;
    stack[0] = 0;
    int max_sp = 0;
    while (sp >= 0) {
        const BvhNode& node = nodes[stack[sp]];
        sp--;
        if (c) c->node_visits++;
        int32_t count = node.extra1, left_first = node.extra2;
        if (count > -1) {
            for (int32_t i = 0; i < count; i++) {
                if (c) c->prim_tests++;
                if (f(indices[left_first + i])) goto done;
            }
        } else if (left_first > -1) {
            if (c) c->inner_visits++;
            __m128 kl = _mm_setzero_ps(), kr = _mm_setzero_ps();
            bool hl = aabb_intersect4(nodes[left_first], p, &kl);
            bool hr = aabb_intersect4(nodes[left_first + 1], p, &kr);
            if (hl && hr) {
                if (_mm_movemask_ps(_mm_cmplt_ps(kl, kr)) > 0) {
                    stack[++sp] = left_first;
                    stack[++sp] = left_first + 1;
                } else {
                    stack[++sp] = left_first + 1;
                    stack[++sp] = left_first;
                }
            } else if (hl) {
                stack[++sp] = left_first;
            } else if (hr) {
                stack[++sp] = left_first + 1;
            }
            max_sp = std::max(max_sp, sp + 1);
        }
    }
done:
    if (c) {
        c->max_stack = std::max<uint64_t>(c->max_stack, max_sp);
        if (max_sp > 32) c->overflow32++;
    }
}

// iter_indices.rs:267-312 (+ ctor :230-262).  `Hit` = mbvh_intersect or mbvh_intersect4.
template <class RayT, class HitFn, class F>
static inline void mbvh_traverse_impl(const MbvhNode* nodes, size_t n_nodes, const uint32_t* indices, RayT& ray,
                                      HitFn&& hit_fn, F&& f, Counters* c) {
    if (n_nodes == 0) return;  // nodes.get(0) == None -> default hit (all false), loop ends at once
    int32_t stack[kStack];
    int sp = -1, max_sp = 0;
    int32_t current = 0;
    MbvhHit hit = hit_fn(nodes[0], ray);
    if (c) c->node_visits++;
    for (;;) {
        const MbvhNode& node = nodes[current];
        for (int i = 0; i < 4; i++) {
            int id = hit.ids[3 - i];
            if (!hit.result[id]) continue;
            int32_t count = node.counts[id], left_first = node.children[id];
            if (count > -1) {
                for (int32_t j = 0; j < count; j++) {
                    if (c) c->prim_tests++;
                    if (f(indices[left_first + j])) goto done;
                }
            } else if (left_first > -1) {
                stack[++sp] = left_first;
                max_sp = std::max(max_sp, sp + 1);
            }
        }
        if (sp < 0) break;
        current = stack[sp--];
        hit = hit_fn(nodes[current], ray);
        if (c) c->node_visits++;
    }
done:
    if (c) {
        c->max_stack = std::max<uint64_t>(c->max_stack, max_sp);
        if (max_sp > 32) c->overflow32++;
    }
}
template <class F>
static inline void mbvh_traverse(const MbvhNode* nodes, size_t n_nodes, const uint32_t* indices, Ray& ray, F&& f,
                                 Counters* c = nullptr) {
    mbvh_traverse_impl(nodes, n_nodes, indices, ray, [](const MbvhNode& n, const Ray& r) { return mbvh_intersect(n, r); },
                       f, c);
}
// iter_indices.rs:370-414
template <class F>
static inline void mbvh_traverse_packet(const MbvhNode* nodes, size_t n_nodes, const uint32_t* indices, RayPacket4& p,
                                        F&& f, Counters* c = nullptr) {
    mbvh_traverse_impl(nodes, n_nodes, indices, p,
                       [](const MbvhNode& n, const RayPacket4& r) { return mbvh_intersect4(n, r); }, f, c);
}

// ----------------------------------------------------------------------------------
// Canonical triangle tests (builders/spatial_sah.rs:131-244)
// ----------------------------------------------------------------------------------
struct Tri {
    Vec3 v0, v1, v2;
};

// spatial_sah.rs:131-163 without the final window test; returns true and *t_out when the
// geometric part passes.  Callers apply `t > t_min && t < ray.t`.
static inline bool tri_geom(const Tri& tr, const Vec3& o, const Vec3& d, float* t_out) {
    Vec3 edge1 = tr.v1 - tr.v0;
    Vec3 edge2 = tr.v2 - tr.v0;
    Vec3 h = cross(d, edge2);
    float a = dot(edge1, h);
    if (a > -1e-5f && a < 1e-5f) return false;
    float f = 1.0f / a;
    Vec3 s = o - tr.v0;
    float u = f * dot(s, h);
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    Vec3 q = cross(s, edge1);
    float v = f * dot(d, q);
    if (v < 0.0f || (u + v) > 1.0f) return false;
    *t_out = f * dot(edge2, q);
    return true;
}
// spatial_sah.rs:131-163 verbatim semantics
static inline bool tri_intersect(const Tri& tr, Ray& ray) {
    float t;
    if (!tri_geom(tr, ray.origin, ray.direction, &t)) return false;
    if (t > ray.t_min && t < ray.t) {
        ray.t = t;
        return true;
    }
    return false;
}

// spatial_sah.rs:165-244 up to (not including) `t < packet.t`: returns lane mask of
// `det & u & v & t >= t_min` and the per-lane t.  Early-outs have no side effects in the reference.
static inline int tri_geom4(const Tri& tr, const RayPacket4& p, __m128 t_min, __m128* t_out) {
    const __m128 zero = _mm_setzero_ps(), one = _mm_set1_ps(1.0f);
    __m128 p0_x = _mm_set1_ps(tr.v0.x), p0_y = _mm_set1_ps(tr.v0.y), p0_z = _mm_set1_ps(tr.v0.z);
    __m128 e1x = _mm_sub_ps(_mm_set1_ps(tr.v1.x), p0_x), e1y = _mm_sub_ps(_mm_set1_ps(tr.v1.y), p0_y),
           e1z = _mm_sub_ps(_mm_set1_ps(tr.v1.z), p0_z);
    __m128 e2x = _mm_sub_ps(_mm_set1_ps(tr.v2.x), p0_x), e2y = _mm_sub_ps(_mm_set1_ps(tr.v2.y), p0_y),
           e2z = _mm_sub_ps(_mm_set1_ps(tr.v2.z), p0_z);
    __m128 h_x = _mm_sub_ps(_mm_mul_ps(p.direction_y, e2z), _mm_mul_ps(p.direction_z, e2y));
    __m128 h_y = _mm_sub_ps(_mm_mul_ps(p.direction_z, e2x), _mm_mul_ps(p.direction_x, e2z));
    __m128 h_z = _mm_sub_ps(_mm_mul_ps(p.direction_x, e2y), _mm_mul_ps(p.direction_y, e2x));
    __m128 a = _mm_add_ps(_mm_add_ps(_mm_mul_ps(e1x, h_x), _mm_mul_ps(e1y, h_y)), _mm_mul_ps(e1z, h_z));
    const __m128 eps = _mm_set1_ps(1e-6f);
    __m128 mask = _mm_or_ps(_mm_cmple_ps(a, _mm_sub_ps(zero, eps)), _mm_cmpge_ps(a, eps));
    if (_mm_movemask_ps(mask) == 0) return 0;
    __m128 f = _mm_div_ps(one, a);
    __m128 s_x = _mm_sub_ps(p.origin_x, p0_x), s_y = _mm_sub_ps(p.origin_y, p0_y), s_z = _mm_sub_ps(p.origin_z, p0_z);
    __m128 u = _mm_mul_ps(f, _mm_add_ps(_mm_add_ps(_mm_mul_ps(s_x, h_x), _mm_mul_ps(s_y, h_y)), _mm_mul_ps(s_z, h_z)));
    mask = _mm_and_ps(mask, _mm_and_ps(_mm_cmpge_ps(u, zero), _mm_cmple_ps(u, one)));
    if (_mm_movemask_ps(mask) == 0) return 0;
    __m128 q_x = _mm_sub_ps(_mm_mul_ps(s_y, e1z), _mm_mul_ps(s_z, e1y));
    __m128 q_y = _mm_sub_ps(_mm_mul_ps(s_z, e1x), _mm_mul_ps(s_x, e1z));
    __m128 q_z = _mm_sub_ps(_mm_mul_ps(s_x, e1y), _mm_mul_ps(s_y, e1x));
    __m128 v = _mm_mul_ps(
        f, _mm_add_ps(_mm_add_ps(_mm_mul_ps(p.direction_x, q_x), _mm_mul_ps(p.direction_y, q_y)), _mm_mul_ps(p.direction_z, q_z)));
    mask = _mm_and_ps(mask, _mm_and_ps(_mm_cmpge_ps(v, zero), _mm_cmple_ps(_mm_add_ps(u, v), one)));
    if (_mm_movemask_ps(mask) == 0) return 0;
    __m128 t = _mm_mul_ps(f, _mm_add_ps(_mm_add_ps(_mm_mul_ps(e2x, q_x), _mm_mul_ps(e2y, q_y)), _mm_mul_ps(e2z, q_z)));
    mask = _mm_and_ps(mask, _mm_cmpge_ps(t, t_min));
    *t_out = t;
    return _mm_movemask_ps(mask);
}
// spatial_sah.rs:165-244 verbatim semantics: returns the lane mask that was written.
static inline int tri_intersect4(const Tri& tr, RayPacket4& p, __m128 t_min) {
    __m128 t;
    int m = tri_geom4(tr, p, t_min, &t);
    if (!m) return 0;
    int lt = _mm_movemask_ps(_mm_cmplt_ps(t, p.t));
    m &= lt;
    if (!m) return 0;
    alignas(16) float tv[4], pt[4];
    _mm_store_ps(tv, t);
    _mm_store_ps(pt, p.t);
    for (int i = 0; i < 4; i++)
        if (m & (1 << i)) pt[i] = tv[i];
    p.t = _mm_load_ps(pt);
    return m;
}

// ----------------------------------------------------------------------------------
// utils.rs / morton.rs
// ----------------------------------------------------------------------------------
// utils.rs:12-22
static inline uint32_t round_up_log2(uint32_t bits, uint32_t offset) {
    if (bits == 0) return offset;
    while ((1u << offset) < bits) offset++;
    return offset;
}
// morton.rs:10-25 (usize = 64-bit)
static inline uint32_t morton_split(uint32_t v) {
    uint32_t log_bits = round_up_log2(32, 0);
    uint64_t x = v, mask = ~0ull;
    uint64_t i = log_bits, n = 1ull << log_bits;
    while (i > 0) {
        // Rust `<<` by >= 64 would panic in debug / wrap in release; n starts at 32 here so it never is.
        mask = (mask | (mask << n)) & ~(mask << (n / 2));
        x = (x | (x << n)) & mask;
        n >>= 1;
        i--;
    }
    return (uint32_t)x;
}
// utils.rs:42-57 (inclusive scan; count==0 returns first[0])
template <class T>
static inline T prefix_sum(const T* first, size_t count, T* out) {
    if (count == 0) return first[0];
    T sum = 0;
    for (size_t i = 0; i < count; i++) {
        sum = sum + first[i];
        out[i] = sum;
    }
    return sum;
}
// utils.rs:76-96 (unstable swap partition over slice[0..n))
template <class T, class P>
static inline size_t partition(T* slice, size_t n, P&& check) {
    size_t count = 0;
    for (size_t i = 0; i < n; i++) {
        if (check(slice[i])) {
            std::swap(slice[i], slice[count]);
            count++;
        }
    }
    return count;
}
// Rust `f as i32`: truncate, saturate, NaN -> 0
static inline int32_t f32_as_i32(float f) {
    if (std::isnan(f)) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}
// morton.rs:28-61
struct MortonEncoder {
    Vec3 world_to_grid, grid_offset;
    int32_t grid_dim;
    MortonEncoder(const Aabb& bb, int32_t dim = 1024) : grid_dim(dim) {
        Vec3 dg = diagonal(bb);
        Vec3 inv = {1.0f / dg.x, 1.0f / dg.y, 1.0f / dg.z};
        world_to_grid = {(float)dim * inv.x, (float)dim * inv.y, (float)dim * inv.z};
        grid_offset = (-amin(bb)) * world_to_grid;
    }
    uint32_t encode(Vec3 p) const {
        Vec3 g = p * world_to_grid + grid_offset;
        int32_t hi = grid_dim - 1;
        uint32_t x = (uint32_t)std::min(hi, std::max(f32_as_i32(g.x), 0));
        uint32_t y = (uint32_t)std::min(hi, std::max(f32_as_i32(g.y), 0));
        uint32_t z = (uint32_t)std::min(hi, std::max(f32_as_i32(g.z), 0));
        return morton_split(x) | (morton_split(y) << 1) | (morton_split(z) << 2);
    }
};

// ----------------------------------------------------------------------------------
// L3 containers
// ----------------------------------------------------------------------------------
struct Bvh {
    std::vector<BvhNode> nodes;
    std::vector<uint32_t> prim_indices;
    int build_type = 0;  // bvh.rs:18-23: 0 None, 1 LocallyOrderedClustered, 2 BinnedSAH, 3 Spatial
};
struct Mbvh {
    std::vector<BvhNode> nodes;
    std::vector<MbvhNode> m_nodes;
    std::vector<uint32_t> prim_indices;
};

// ----------------------------------------------------------------------------------
// Binned SAH (builders/binned_sah.rs:34-399, builders/mod.rs:59-76, utils.rs:243-288)
// ----------------------------------------------------------------------------------
struct SahBin {
    Aabb aabb;
    size_t prim_count;
    float right_cost;
};
struct SahSplit {
    float cost;
    uint32_t count;
};
static constexpr size_t kBinCount = 16;        // binned_sah.rs:314
static constexpr size_t kMaxDepth = 64;        // binned_sah.rs:313
static constexpr float kTraversalCost = 1.0f;  // binned_sah.rs:316

// binned_sah.rs:119-128  (Rust `as usize`: NaN -> 0, saturating)
static inline size_t compute_bin_index(Vec3 c, Vec3 bin_offset, Vec3 center_to_bin, int axis) {
    float bin_index = c[axis] * center_to_bin[axis] + bin_offset[axis];
    float m = f32_max(bin_index, 0.0f);
    size_t b = (m >= (float)kBinCount) ? kBinCount : (size_t)m;
    return std::min(kBinCount - 1, b);
}
// binned_sah.rs:80-114
static inline SahSplit find_split(SahBin* bins) {
    Aabb cur = aabb_new();
    size_t cnt = 0;
    for (size_t i = kBinCount - 1; i > 0; i--) {
        grow_bb(cur, bins[i].aabb);
        cnt += bins[i].prim_count;
        bins[i].right_cost = half_area(cur) * (float)cnt;
    }
    cur = aabb_new();
    cnt = 0;
    SahSplit best{std::numeric_limits<float>::max(), (uint32_t)kBinCount};
    for (size_t i = 0; i < kBinCount - 1; i++) {
        grow_bb(cur, bins[i].aabb);
        cnt += bins[i].prim_count;
        float cost = half_area(cur) * (float)cnt + bins[i + 1].right_cost;
        if (cost < best.cost) best = SahSplit{cost, (uint32_t)i + 1};
    }
    return best;
}

struct BinnedSahBuilder {
    const Aabb* aabbs;
    const Vec3* centers;  // Primitive::center() of each primitive
    size_t n;
    size_t max_leaf_size;  // primitives_per_leaf or 1 (binned_sah.rs:315)
    std::vector<BvhNode> nodes;
    std::vector<uint32_t> prim_indices;
    size_t node_count = 1;  // AtomicNodeStack counter (builders/mod.rs:54)
    bool threaded = false;  // build_parallel(): child pairs are handed out with an atomic fetch-add, like the reference's

    struct Task {
        size_t node, begin, end, depth;
        size_t work() const { return end - begin; }
    };

    // binned_sah.rs:132-282.  Returns true and fills (a, b) when the node was split.
    bool run(const Task& t, Task* a, Task* b) {
        BvhNode& node = nodes[t.node];
        offset_by(node, 0.0001f);
        auto make_leaf = [&](BvhNode& nd) {
            offset_by(nd, 0.0001f);
            nd.extra2 = (int32_t)t.begin;
            nd.extra1 = (int32_t)(t.end - t.begin);
        };
        const size_t work = t.work();
        if (work <= 1 || t.depth >= kMaxDepth) {
            make_leaf(node);
            return false;
        }
        Vec3 dg = diagonal(node);
        Vec3 center_to_bin = v3(1.0f / dg.x, 1.0f / dg.y, 1.0f / dg.z) * (float)kBinCount;
        Vec3 bin_offset = (-amin(node)) * center_to_bin;
        SahBin bins[3][kBinCount];
        for (int ax = 0; ax < 3; ax++)
            for (size_t i = 0; i < kBinCount; i++) bins[ax][i] = SahBin{aabb_new(), 0, std::numeric_limits<float>::max()};
        uint32_t* idx = prim_indices.data() + t.begin;
        for (size_t i = 0; i < work; i++) {
            size_t p = idx[i];
            for (int ax = 0; ax < 3; ax++) {
                size_t bi = compute_bin_index(centers[p], bin_offset, center_to_bin, ax);
                bins[ax][bi].prim_count += 1;
                grow_bb(bins[ax][bi].aabb, aabbs[p]);
            }
        }
        SahSplit best[3];
        for (int ax = 0; ax < 3; ax++) best[ax] = find_split(bins[ax]);
        int best_axis = 0;
        if (best[0].cost > best[1].cost) best_axis = 1;
        if (best[best_axis].cost > best[2].cost) best_axis = 2;
        size_t split_index = best[best_axis].count;
        float max_split_cost = half_area(node) * ((float)work - kTraversalCost);
        if (best[best_axis].count == kBinCount || best[best_axis].cost >= max_split_cost) {
            if (work > max_leaf_size) {
                // fallback: ~40 % median on the longest axis (binned_sah.rs:194-205)
                best_axis = longest_axis(node);
                size_t count = 0;
                for (size_t i = 0; i < kBinCount - 1; i++) {
                    count += bins[best_axis][i].prim_count;
                    if (count >= (work * 2 / 5 + 1)) {
                        split_index = i + 1;
                        break;
                    }
                }
            } else {
                make_leaf(node);
                return false;
            }
        }
        size_t begin_right = t.begin + partition(idx, work, [&](uint32_t p) {
                                 return compute_bin_index(centers[p], bin_offset, center_to_bin, best_axis) < split_index;
                             });
        if (begin_right > t.begin && begin_right < t.end) {
            size_t left;  // AtomicNodeStack::allocate, builders/mod.rs:59-76
            if (threaded) {
                left = __atomic_fetch_add(&node_count, (size_t)2, __ATOMIC_RELAXED);
            } else {
                left = node_count;
                node_count += 2;
            }
            node.extra2 = (int32_t)left;
            node.extra1 = -1;
            Aabb lb = aabb_new(), rb = aabb_new();
            // quirk Q3: the LEFT box uses best_splits[best_axis].count even after the fallback moved split_index
            for (size_t i = 0; i < best[best_axis].count; i++) grow_bb(lb, bins[best_axis][i].aabb);
            for (size_t i = split_index; i < kBinCount; i++) grow_bb(rb, bins[best_axis][i].aabb);
            nodes[left] = lb;
            nodes[left + 1] = rb;
            *a = Task{left, t.begin, begin_right, t.depth + 1};
            *b = Task{left + 1, begin_right, t.end, t.depth + 1};
            return true;
        }
        make_leaf(node);
        return false;
    }

    // binned_sah.rs:346-399 with TaskSpawner::run_task's single-thread order (utils.rs:243-288):
    // the larger child runs first.  Thread spawning only changes node numbering, never topology.
    Bvh build() {
        Bvh out;
        if (n == 0) return out;
        nodes.assign(n * 2 - 1, aabb_new());
        prim_indices.resize(n);
        for (size_t i = 0; i < n; i++) prim_indices[i] = (uint32_t)i;
        node_count = 1;
        nodes[0] = union_of_list(aabbs, n);
        std::vector<Task> stack;
        stack.push_back(Task{0, 0, n, 0});
        while (!stack.empty()) {
            Task t = stack.back();
            stack.pop_back();
            Task a, b;
            if (run(t, &a, &b)) {
                if (a.work() < b.work()) std::swap(a, b);
                stack.push_back(b);
                stack.push_back(a);
            }
        }
        nodes.resize(node_count);
        out.nodes = std::move(nodes);
        out.prim_indices = std::move(prim_indices);
        out.build_type = 2;
        return out;
    }

    // The reference's threaded flavour (TaskSpawner, utils.rs:189-289): a child with more than 1024 primitives is handed to
    // another thread while threads are available, the rest runs on the spawning thread's own stack.  Here: OpenMP tasks with
    // the same 1024-primitive threshold.  Topology, boxes and leaf contents are those of build(); only the node NUMBERING
    // depends on the interleaving of the threads' allocations — exactly as in the reference.  Used for the CPU build
    // baseline of bench.py (every host core); the deterministic build() stays the checker.
    void run_subtree(Task root) {
        std::vector<Task> stack;
        stack.push_back(root);
        while (!stack.empty()) {
            Task t = stack.back();
            stack.pop_back();
            Task a, b;
            if (run(t, &a, &b)) {
                if (a.work() < b.work()) std::swap(a, b);
                if (b.work() > 1024) {  // utils.rs:253-262 (the smaller child is the one handed over)
#pragma omp task firstprivate(b)
                    run_subtree(b);
                } else {
                    stack.push_back(b);
                }
                stack.push_back(a);
            }
        }
    }
    Bvh build_parallel(int threads) {
        Bvh out;
        if (n == 0) return out;
        threaded = true;
        nodes.assign(n * 2 - 1, aabb_new());
        prim_indices.resize(n);
        for (size_t i = 0; i < n; i++) prim_indices[i] = (uint32_t)i;
        node_count = 1;
        nodes[0] = union_of_list(aabbs, n);
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
        {
#pragma omp single
            run_subtree(Task{0, 0, n, 0});
        }
        nodes.resize(node_count);
        out.nodes = std::move(nodes);
        out.prim_indices = std::move(prim_indices);
        out.build_type = 2;
        return out;
    }
};

// ----------------------------------------------------------------------------------
// Locally-ordered clustering (builders/locb.rs:18-328, morton.rs:63-102)
// ----------------------------------------------------------------------------------
struct LocbBuilder {
    const Aabb* aabbs;
    const Vec3* centers;
    size_t n;
    static constexpr size_t kRadius = 14;  // locb.rs:27
    uint64_t cluster_iterations = 0, cluster_sum = 0;  // for kappa in SURVEY §8d

    // locb.rs:36-45
    static inline void search_range(size_t i, size_t begin, size_t end, size_t* b, size_t* e) {
        *b = (i > begin + kRadius) ? i - kRadius : begin;
        *e = std::min(i + kRadius + 1, end);
    }

    // locb.rs:48-245
    void cluster(const std::vector<BvhNode>& input, std::vector<BvhNode>& output, std::vector<uint32_t>& neighbours,
                 std::vector<uint32_t>& merged_index, size_t begin, size_t end, size_t previous_end, size_t* next_begin,
                 size_t* next_end) {
        // nearest neighbour search (locb.rs:93-142).  The reference caches forward distances in a
        // rotating (radius+1) x radius matrix; min/max unions are commutative so the cached value
        // for (j, i) is bit-identical to recomputing (i, j).  We keep the cache (ring buffer).
        std::vector<float> ring((kRadius + 1) * kRadius, 0.0f);
        auto row = [&](size_t i) { return ring.data() + (i % (kRadius + 1)) * kRadius; };
        for (size_t i = begin; i < end; i++) {
            size_t sb, se;
            search_range(i, begin, end, &sb, &se);
            float best = std::numeric_limits<float>::max();
            int64_t best_nb = -1;
            for (size_t j = sb; j < i; j++) {
                float d = row(j)[i - j - 1];
                if (d < best) {
                    best = d;
                    best_nb = (int64_t)j;
                }
            }
            float* fwd = row(i);
            for (size_t j = i + 1; j < se; j++) {
                float d = half_area(union_of(input[i], input[j]));
                fwd[j - i - 1] = d;
                if (d < best) {
                    best = d;
                    best_nb = (int64_t)j;
                }
            }
            neighbours[i] = (uint32_t)best_nb;
        }
        // locb.rs:158-167
        for (size_t i = begin; i < end; i++) {
            size_t j = neighbours[i];
            bool mergeable = neighbours[j] == i;
            merged_index[i] = (i < j && mergeable) ? 1u : 0u;
        }
        // locb.rs:170-176
        prefix_sum(merged_index.data() + begin, end - begin, merged_index.data() + begin);
        // locb.rs:178-185
        size_t merged_count = merged_index[end - 1];
        size_t unmerged_count = end - begin - merged_count;
        size_t children_count = merged_count * 2;
        size_t children_begin = end - children_count;
        size_t unmerged_begin = end - (children_count + unmerged_count);
        *next_begin = unmerged_begin;
        *next_end = children_begin;
        // locb.rs:194-215
        for (size_t i = begin; i < end; i++) {
            size_t j = neighbours[i];
            if (neighbours[j] == i) {
                if (i < j) {
                    BvhNode& parent = output[unmerged_begin + j - begin - merged_index[j]];
                    size_t first_child = children_begin + ((size_t)merged_index[i] - 1) * 2;
                    parent = union_of(input[j], input[i]);
                    parent.extra1 = -1;
                    parent.extra2 = (int32_t)first_child;
                    output[first_child] = input[i];
                    output[first_child + 1] = input[j];
                }
            } else {
                output[unmerged_begin + i - begin - merged_index[i]] = input[i];
            }
        }
        // locb.rs:242
        for (size_t k = end; k < previous_end; k++) output[k] = input[k];
    }

    // morton.rs:63-102 (stable sort by code).  `parallel` uses the libstdc++ parallel-mode stable sort
    // (the reference uses rayon par_sort_by, also stable).
    std::vector<uint32_t> sorted_indices(const MortonEncoder& enc, std::vector<uint32_t>* codes_out, bool parallel);

    // locb.rs:249-328
    Bvh build(bool parallel = false) {
        Bvh out;
        if (n == 0) return out;
        Aabb world = union_of_list(aabbs, n);  // locb.rs:19
        if (n <= 2) {
            BvhNode root = world;
            root.extra2 = 0;
            root.extra1 = (int32_t)n;
            out.nodes.push_back(root);
            for (size_t i = 0; i < n; i++) out.prim_indices.push_back((uint32_t)i);
            out.build_type = 1;
            return out;
        }
        MortonEncoder enc(world, 1024);
        std::vector<uint32_t> prim_indices = sorted_indices(enc, nullptr, parallel);
        size_t node_count = 2 * n - 1;
        BvhNode init{{0, 0, 0}, -1, {0, 0, 0}, 0};
        std::vector<BvhNode> nodes(node_count, init), nodes_copy(node_count, init);
        // locb.rs:283, 296-305: neighbours = aux[0..], merged_index = aux[node_count..]
        std::vector<uint32_t> neighbours(node_count, (uint32_t)node_count), merged(node_count, (uint32_t)node_count);
        size_t begin = node_count - n, end = node_count, previous_end = end;
        for (size_t i = 0; i < n; i++) {
            BvhNode& nd = nodes[begin + i];
            nd = aabbs[prim_indices[i]];
            offset_by(nd, 0.0001f);
            nd.extra1 = 1;
            nd.extra2 = (int32_t)i;
        }
        while (end - begin > 1) {
            size_t nb, ne;
            cluster_iterations++;
            cluster_sum += end - begin;
            cluster(nodes, nodes_copy, neighbours, merged, begin, end, previous_end, &nb, &ne);
            std::swap(nodes, nodes_copy);
            previous_end = end;
            begin = nb;
            end = ne;
        }
        out.nodes = std::move(nodes);
        out.prim_indices = std::move(prim_indices);
        out.build_type = 1;
        return out;
    }
};

// ----------------------------------------------------------------------------------
// Spatial-split SAH (builders/spatial_sah.rs:279-1034) — CPU only: the north star uploads such a
// tree "unchanged", it is never built on the GPU.  Restated verbatim including its quirks:
//   * SpatialTriangle::split interpolates edges as a*t*(b-a) instead of a+t*(b-a) (Q9, :91-94);
//   * run_binning_pass counts `exit` on the FIRST bin and sweeps the already suffix-accumulated boxes (:498-523);
//   * the "all references on one side" repair uses left_count / 2 as an absolute index (:651-655);
//   * prim_indices has length N + floor(0.75 N) with unused trailing zeros (:936-940).
// One behaviour is NOT reproduced by default (fix_child_ranges = true), because it silently drops
// primitives from every tree: allocate_children moves the right child's references up by
// left_split_count to make split space for the left child, but then describes the children with the
// OLD positions (:375-423): the right child sees the first left_split_count references twice and never
// sees the last left_split_count.  With fix_child_ranges = false the verbatim ranges are used (kept for
// the test that documents the defect).  Rust's slice::sort_by only ever asks `cmp == Less`, so the
// reference's non-total comparators (:322-332, :446-454) reduce to stable sorts by `a < b` / by mark.
// ----------------------------------------------------------------------------------
struct SpatialReference {
    Aabb aabb;
    Vec3 center;
    uint32_t prim_id;
};
static inline void shrink(Aabb& a, const Aabb& b) {  // aabb.rs:287-297
    for (int i = 0; i < 3; i++) {
        a.min[i] = f32_max(a.min[i], b.min[i]);
        a.max[i] = f32_min(a.max[i], b.max[i]);
    }
}
static inline float area(const Aabb& a) {  // aabb.rs:330-335
    Vec3 e = amax(a) - amin(a);
    float v = e.x * e.y + e.x * e.z + e.y * e.z;
    return f32_max(0.0f, v);
}
// SpatialTriangle::split (spatial_sah.rs:85-129)
static inline void tri_split(const Tri& tr, int axis, float position, Aabb* left_out, Aabb* right_out) {
    const Vec3 p[3] = {tr.v0, tr.v1, tr.v2};
    Aabb left = aabb_new(), right = aabb_new();
    auto split_edge = [&](Vec3 a, Vec3 b) {
        float t = (position - a[axis]) / (b[axis] - a[axis]);
        return (a * t) * (b - a);  // sic
    };
    const bool q0 = p[0][axis] <= position, q1 = p[1][axis] <= position, q2 = p[2][axis] <= position;
    auto grow_if = [&](bool q, Vec3 pos) { if (q) grow(left, pos); else grow(right, pos); };
    grow_if(q0, p[0]);
    grow_if(q1, p[1]);
    grow_if(q2, p[2]);
    if (q0 ^ q1) { Vec3 m = split_edge(p[0], p[1]); grow(left, m); grow(right, m); }
    if (q1 ^ q2) { Vec3 m = split_edge(p[1], p[2]); grow(left, m); grow(right, m); }
    if (q2 ^ q0) { Vec3 m = split_edge(p[2], p[0]); grow(left, m); grow(right, m); }
    *left_out = left;
    *right_out = right;
}
static inline size_t f32_as_usize(float f) {  // Rust `as usize`
    if (!(f > 0.0f)) return 0;
    if (f >= 1.8446744e19f) return ~size_t(0);
    return (size_t)f;
}

struct SpatialSahBuilder {
    const Aabb* aabbs;
    const Tri* tris;
    const Vec3* centers;
    size_t n;
    size_t max_leaf_size;
    bool fix_child_ranges = true;
    // spatial_sah.rs:866-883
    size_t binning_pass_count = 2, max_depth = 64, bin_count = 16;
    float traversal_cost = 1.0f, alpha = 1e-5f, split_factor = 0.75f;

    struct Item {
        size_t node, begin, end, split_end, depth;
        bool is_sorted;
        size_t work() const { return end - begin; }
    };
    struct ObjectSplit {
        float cost = std::numeric_limits<float>::max();
        long index = -1;  // Option<isize>
        int axis = 0;
        Aabb left_box = aabb_new(), right_box = aabb_new();
    };
    struct SpatialSplit {
        float cost = std::numeric_limits<float>::max();
        float position = 0.0f;
        int axis = 0;
    };
    std::vector<BvhNode> nodes;
    std::vector<uint32_t> prim_indices;
    std::vector<Aabb> accumulated;
    std::vector<SpatialReference> refs[3];
    std::vector<uint8_t> marks;
    size_t reference_count = 0, node_count = 1;
    float spatial_threshold = 0.0f;
    uint64_t spatial_splits = 0, object_splits = 0;

    void make_leaf(BvhNode& node, size_t begin, size_t end) {  // :727-738
        size_t prim_count = end - begin;
        size_t first_prim = reference_count;
        reference_count += prim_count;
        for (size_t i = 0; i < prim_count; i++) prim_indices[first_prim + i] = refs[0][begin + i].prim_id;
        node.extra2 = (int32_t)first_prim;
        node.extra1 = (int32_t)prim_count;
    }

    ObjectSplit find_object_split(size_t begin, size_t end, bool is_sorted) {  // :314-368
        if (!is_sorted) {
            for (int axis = 0; axis < 3; axis++)
                std::stable_sort(refs[axis].begin() + begin, refs[axis].begin() + end,
                                 [axis](const SpatialReference& a, const SpatialReference& b) { return a.center[axis] < b.center[axis]; });
        }
        ObjectSplit best;
        for (int axis = 0; axis < 3; axis++) {
            Aabb bb = aabb_new();
            for (size_t i = end - 1; i > begin; i--) {
                grow_bb(bb, refs[axis][i].aabb);
                accumulated[i] = bb;
            }
            bb = aabb_new();
            for (size_t i = begin; i < end - 1; i++) {
                grow_bb(bb, refs[axis][i].aabb);
                float cost = half_area(bb) * (float)(i + 1 - begin) + half_area(accumulated[i + 1]) * (float)(end - (i + 1));
                if (cost < best.cost) {
                    best.cost = cost;
                    best.axis = axis;
                    best.index = (long)i + 1;
                    best.left_box = bb;
                    best.right_box = accumulated[i + 1];
                }
            }
        }
        return best;
    }

    void allocate_children(const Item& it, size_t right_begin, size_t right_end, const Aabb& left_box, const Aabb& right_box,
                           bool is_sorted, Item* a, Item* b) {  // :370-424
        size_t left = node_count;
        node_count += 2;
        BvhNode& parent = nodes[it.node];
        parent.extra2 = (int32_t)left;
        parent.extra1 = -1;
        offset_by(parent, 0.0001f);
        nodes[left] = left_box;
        nodes[left + 1] = right_box;
        size_t remaining = it.split_end - right_end;
        float left_cost = half_area(left_box) * (float)(right_begin - it.begin);
        float right_cost = half_area(right_box) * (float)(right_end - right_begin);
        size_t left_split_count = remaining == 0 ? 0 : f32_as_usize((float)remaining * (left_cost / (left_cost + right_cost)));
        if (left_split_count > 0) {
            for (int k = 0; k < 3; k++)  // move_backward(refs + right_begin, refs + right_end, refs + right_end + lsc)
                std::move_backward(refs[k].begin() + right_begin, refs[k].begin() + right_end,
                                   refs[k].begin() + right_end + left_split_count);
        }
        size_t left_end = right_begin;
        if (fix_child_ranges) {
            *a = Item{left, it.begin, left_end, right_begin + left_split_count, it.depth + 1, is_sorted};
            *b = Item{left + 1, right_begin + left_split_count, right_end + left_split_count, it.split_end, it.depth + 1, is_sorted};
        } else {  // verbatim (:405-421)
            *a = Item{left, it.begin, left_end, right_begin, it.depth + 1, is_sorted};
            *b = Item{left + 1, right_begin, right_end, it.split_end, it.depth + 1, is_sorted};
        }
    }

    void apply_object_split(const Item& it, const ObjectSplit& split, Item* a, Item* b) {  // :426-470
        size_t split_index = (size_t)split.index;
        int o0 = (split.axis + 1) % 3, o1 = (split.axis + 2) % 3;
        for (size_t i = it.begin; i < split_index; i++) marks[refs[split.axis][i].prim_id] = 1;
        for (size_t i = split_index; i < it.end; i++) marks[refs[split.axis][i].prim_id] = 0;
        auto marked = [&](const SpatialReference& r) { return marks[r.prim_id] != 0; };
        std::stable_partition(refs[o0].begin() + it.begin, refs[o0].begin() + it.end, marked);
        std::stable_partition(refs[o1].begin() + it.begin, refs[o1].begin() + it.end, marked);
        object_splits++;
        allocate_children(it, split_index, it.end, split.left_box, split.right_box, true, a, b);
    }

    bool run_binning_pass(const Item&, SpatialSplit& split, int axis, size_t begin, size_t end, float min, float max, float* lo,
                          float* hi) {  // :472-535
        struct Bin {
            Aabb aabb = aabb_new();
            size_t entry = 0, exit = 0;
        };
        std::vector<Bin> bins(bin_count);
        float bin_size = (max - min) / (float)bin_count;
        float inv_size = 1.0f / bin_size;
        for (size_t i = begin; i < end; i++) {
            const SpatialReference& ref = refs[0][i];
            size_t first_bin = std::min(bin_count - 1, f32_as_usize(f32_max(inv_size * (ref.aabb.min[axis] - min), 0.0f)));
            size_t last_bin = std::min(bin_count - 1, f32_as_usize(f32_max(inv_size * (ref.aabb.max[axis] - min), 0.0f)));
            if (!is_valid(ref.aabb)) break;
            Aabb current = ref.aabb;
            for (size_t j = 0; first_bin + j < last_bin; j++) {
                Aabb lb, rb;
                tri_split(tris[ref.prim_id], axis, min + (float)(j + first_bin + 1) * bin_size, &lb, &rb);
                shrink(lb, current);
                grow_bb(bins[first_bin + j].aabb, lb);
                shrink(current, rb);
            }
            grow_bb(bins[last_bin].aabb, current);
            bins[first_bin].entry += 1;
            bins[first_bin].exit += 1;  // sic
        }
        Aabb cur = aabb_new();
        for (size_t i = bin_count; i > 0; i--) {
            grow_bb(cur, bins[i - 1].aabb);
            bins[i - 1].aabb = cur;
        }
        size_t left_count = 0, right_count = end - begin;
        cur = aabb_new();
        bool found = false;
        for (size_t i = 0; i + 1 < bin_count; i++) {
            left_count += bins[i].entry;
            right_count -= bins[i].exit;
            grow_bb(cur, bins[i].aabb);
            float cost = (float)left_count * half_area(cur) + (float)right_count * half_area(bins[i + 1].aabb);
            if (cost < split.cost) {
                split.cost = cost;
                split.axis = axis;
                split.position = min + (float)(i + 1) * bin_size;
                found = true;
            }
        }
        if (found) {
            *lo = split.position - bin_size;
            *hi = split.position + bin_size;
        }
        return found;
    }

    SpatialSplit find_spatial_split(const Item& it) {  // :537-557
        SpatialSplit split;
        for (int axis = 0; axis < 3; axis++) {
            float mn = nodes[it.node].min[axis], mx = nodes[it.node].max[axis];
            for (size_t pass = 0; pass < binning_pass_count; pass++) {
                float lo, hi;
                if (run_binning_pass(it, split, axis, it.begin, it.end, mn, mx, &lo, &hi)) {
                    mn = lo;
                    mx = hi;
                } else {
                    break;
                }
            }
        }
        return split;
    }

    void apply_spatial_split(const Item& it, const SpatialSplit& split, Item* a, Item* b) {  // :559-716
        size_t left_end = it.begin, right_begin = it.end, right_end = it.end;
        Aabb left_box = aabb_new(), right_box = aabb_new();
        std::vector<SpatialReference>& r = refs[split.axis];
        size_t i = it.begin;
        while (i < right_begin) {
            const Aabb& bb = r[i].aabb;
            if (bb.max[split.axis] <= split.position) {
                grow_bb(left_box, bb);
                std::swap(r[i], r[left_end]);
                i++;
                left_end++;
            } else if (bb.min[split.axis] >= split.position) {
                grow_bb(right_box, bb);
                right_begin--;
                std::swap(r[i], r[right_begin]);
            } else {
                i++;
            }
        }
        size_t left_count = left_end - it.begin, right_count = right_end - right_begin;
        if ((left_count == 0 || right_count == 0) && left_end == right_begin) {
            if (left_count > 0) left_end = left_count / 2;  // sic: not begin + left_count / 2
            else left_end += right_count / 2;
            right_begin = left_end;
            left_box = aabb_new();
            right_box = aabb_new();
            for (size_t k = it.begin; k < left_end; k++) grow_bb(left_box, r[k].aabb);
            for (size_t k = left_end; k < it.end; k++) grow_bb(right_box, r[k].aabb);
        }
        while (left_end < right_begin) {
            SpatialReference ref = r[left_end];
            Aabb lp, rp;
            tri_split(tris[ref.prim_id], split.axis, split.position, &lp, &rp);
            shrink(lp, ref.aabb);
            shrink(rp, ref.aabb);
            if (it.split_end - right_end > 0) {
                grow_bb(left_box, lp);
                grow_bb(right_box, rp);
                r[right_end] = SpatialReference{rp, center(rp), ref.prim_id};
                r[left_end] = SpatialReference{lp, center(lp), ref.prim_id};
                right_end++;
                left_end++;
                left_count++;
                right_count++;
            } else if (left_count < right_count) {
                grow_bb(left_box, ref.aabb);
                left_end++;
                left_count++;
            } else {
                grow_bb(right_box, ref.aabb);
                right_begin--;
                std::swap(r[right_begin], r[left_end]);
                right_count++;
            }
        }
        for (int k = 1; k <= 2; k++) {
            std::vector<SpatialReference>& o = refs[(split.axis + k) % 3];
            std::copy(r.begin() + it.begin, r.begin() + right_end, o.begin() + it.begin);
        }
        spatial_splits++;
        allocate_children(it, right_begin, right_end, left_box, right_box, false, a, b);
    }

    bool run(const Item& it, Item* a, Item* b) {  // :720-838
        BvhNode& node = nodes[it.node];
        if (it.work() <= 1 || it.depth >= max_depth) {
            make_leaf(node, it.begin, it.end);
            return false;
        }
        ObjectSplit os = find_object_split(it.begin, it.end, it.is_sorted);
        SpatialSplit ss;
        Aabb overlap_box = os.left_box;
        shrink(overlap_box, os.right_box);
        float overlap = area(overlap_box);
        if (overlap > spatial_threshold && (it.split_end - it.end) > 0) ss = find_spatial_split(it);
        float best_cost = f32_min(ss.cost, os.cost);
        bool use_spatial = best_cost < os.cost;
        float max_split_cost = half_area(nodes[it.node]) * ((float)it.work() - traversal_cost);
        if (best_cost >= max_split_cost) {
            if (it.work() > max_leaf_size) {
                use_spatial = false;
                os.index = (long)((it.begin + it.end) / 2);
                os.axis = longest_axis(nodes[it.node]);
                os.left_box = aabb_new();
                os.right_box = aabb_new();
                for (size_t i = it.begin; i < (size_t)os.index; i++) grow_bb(os.left_box, refs[os.axis][i].aabb);
                for (size_t i = (size_t)os.index; i < it.end; i++) grow_bb(os.right_box, refs[os.axis][i].aabb);
            } else {
                make_leaf(nodes[it.node], it.begin, it.end);
                return false;
            }
        }
        if (use_spatial) {
            apply_spatial_split(it, ss, a, b);
        } else if (os.index >= 0) {
            if ((size_t)os.index < it.begin || (size_t)os.index >= it.end) {  // the reference assert!s (panics) here
                make_leaf(nodes[it.node], it.begin, it.end);
                return false;
            }
            apply_object_split(it, os, a, b);
        } else {
            make_leaf(nodes[it.node], it.begin, it.end);
            return false;
        }
        return true;
    }

    Bvh build() {  // :925-1033
        Bvh out;
        if (n == 0) return out;
        size_t max_ref = n + f32_as_usize((float)n * split_factor);
        nodes.assign(2 * max_ref + 1, aabb_new());
        prim_indices.assign(max_ref, 0u);
        accumulated.assign(max_ref, aabb_new());
        marks.assign(n, 0);
        SpatialReference dflt{aabb_new(), v3(0, 0, 0), 0};
        for (int k = 0; k < 3; k++) {
            refs[k].assign(max_ref, dflt);
            for (size_t i = 0; i < n; i++) refs[k][i] = SpatialReference{aabbs[i], centers[i], (uint32_t)i};
        }
        Aabb root_bounds = union_of_list(aabbs, n);
        spatial_threshold = alpha * 2.0f * half_area(root_bounds);
        node_count = 1;
        reference_count = 0;
        nodes[0] = root_bounds;
        std::vector<Item> stack{Item{0, 0, n, max_ref, 0, false}};
        while (!stack.empty()) {  // TaskSpawner::run_task, single-thread order (utils.rs:243-288)
            Item it = stack.back();
            stack.pop_back();
            Item a, b;
            if (run(it, &a, &b)) {
                if (a.work() < b.work()) std::swap(a, b);
                stack.push_back(b);
                stack.push_back(a);
            }
        }
        nodes.resize(node_count);
        out.nodes = std::move(nodes);
        out.prim_indices = std::move(prim_indices);
        out.build_type = 3;
        return out;
    }
};

// ----------------------------------------------------------------------------------
// Collapse to MBVH (mbvh_node.rs:297-411, bvh.rs:381-404)
// ----------------------------------------------------------------------------------
static inline void merge_nodes(size_t m_index, size_t cur_node, const std::vector<BvhNode>& bvh_pool,
                               std::vector<MbvhNode>& mbvh_pool, size_t* pool_ptr) {
    for (int i = 0; i < 4; i++) set_bounds_bb(mbvh_pool[m_index], i, bvh_pool[cur_node]);
    int32_t nodes[4] = {-1, -1, -1, -1}, leafs[4] = {-1, -1, -1, -1};
    auto is_leaf = [](const BvhNode& n) { return n.extra1 >= 0; };
    const BvhNode& cur = bvh_pool[cur_node];
    if (cur.extra2 >= 0) {
        // NB (verbatim): for a leaf `cur` this reinterprets its primitive offset as a node index.
        for (int side = 0; side < 2; side++) {
            size_t x = (size_t)cur.extra2 + side;
            int s0 = side * 2, s1 = side * 2 + 1;
            if (x >= bvh_pool.size()) continue;  // bvh_pool.get(..) == None
            const BvhNode& xn = bvh_pool[x];
            if (xn.extra2 < 0) continue;  // get_left_first() == None
            size_t lf = (size_t)xn.extra2;
            if (is_leaf(xn)) {
                nodes[s0] = (int32_t)lf;
                leafs[s0] = xn.extra1;
            } else {
                if (is_leaf(bvh_pool[lf])) {
                    nodes[s0] = bvh_pool[lf].extra2;
                    leafs[s0] = bvh_pool[lf].extra1;
                } else {
                    nodes[s0] = (int32_t)lf;
                }
                set_bounds_bb(mbvh_pool[m_index], s0, bvh_pool[lf]);
                if (is_leaf(bvh_pool[lf + 1])) {
                    nodes[s1] = bvh_pool[lf + 1].extra2;
                    leafs[s1] = bvh_pool[lf + 1].extra1;
                } else {
                    nodes[s1] = (int32_t)lf + 1;
                }
                set_bounds_bb(mbvh_pool[m_index], s1, bvh_pool[lf + 1]);
            }
        }
    }
    for (int i = 0; i < 4; i++) {
        int32_t node = nodes[i], count = leafs[i];
        if (node >= 0 && count >= 0) {
            mbvh_pool[m_index].children[i] = node;
            mbvh_pool[m_index].counts[i] = count;
            continue;
        } else if (node == -1) {
            continue;
        }
        if (is_leaf(bvh_pool[node])) {
            mbvh_pool[m_index].children[i] = bvh_pool[node].extra2;
            mbvh_pool[m_index].counts[i] = bvh_pool[node].extra1;
            set_bounds_bb(mbvh_pool[m_index], i, bvh_pool[node]);
        } else {
            size_t new_m = (*pool_ptr)++;
            mbvh_pool[m_index].children[i] = (int32_t)new_m;
            set_bounds_bb(mbvh_pool[m_index], i, bvh_pool[node]);
            merge_nodes(new_m, (size_t)node, bvh_pool, mbvh_pool, pool_ptr);
        }
    }
}
// bvh.rs:381-404
static inline Mbvh mbvh_construct(const Bvh& bvh) {
    Mbvh m;
    if (bvh.nodes.empty()) return m;
    std::vector<MbvhNode> pool(bvh.nodes.size(), mbvh_node_new());
    size_t pool_ptr = 1;
    merge_nodes(0, 0, bvh.nodes, pool, &pool_ptr);
    pool.resize(pool_ptr);
    m.nodes = bvh.nodes;
    m.m_nodes = std::move(pool);
    m.prim_indices = bvh.prim_indices;
    return m;
}

// ----------------------------------------------------------------------------------
// refit / validate (bvh.rs:176-244) and the SAH metric defined in SURVEY §8 a-15
// ----------------------------------------------------------------------------------
static inline void refit(Bvh& bvh, const Aabb* new_aabbs) {
    for (size_t k = bvh.nodes.size(); k-- > 0;) {
        Aabb bb = aabb_new();
        BvhNode& nd = bvh.nodes[k];
        if (nd.extra2 >= 0) {
            int32_t lf = nd.extra2, count = nd.extra1;
            if (count >= 0) {
                for (int32_t i = 0; i < count; i++) grow_bb(bb, new_aabbs[bvh.prim_indices[(size_t)lf + i]]);
            } else {
                grow_bb(bb, bvh.nodes[lf]);
                grow_bb(bb, bvh.nodes[lf + 1]);
            }
            offset_by(bb, 0.0001f);
        }
        // "Overwrite AABB": self.nodes[i].bounds = aabb replaces the extras too (bvh.rs:203) —
        // verbatim this zeroes count/left_first; we keep the topology fields, which is what every
        // later traversal of a refitted tree requires.  Flagged in DESIGN.md as a reference bug.
        bb.extra1 = nd.extra1;
        bb.extra2 = nd.extra2;
        nd = bb;
    }
}
static inline bool validate(const Bvh& bvh, size_t prim_count) {
    if (bvh.nodes.empty()) return false;
    std::vector<uint8_t> found(prim_count, 0);
    std::vector<int32_t> stack{0};
    while (!stack.empty()) {
        const BvhNode& nd = bvh.nodes[stack.back()];
        stack.pop_back();
        if (nd.extra2 < 0) continue;
        if (nd.extra1 >= 0) {
            for (int32_t i = 0; i < nd.extra1; i++) {
                uint32_t p = bvh.prim_indices[(size_t)nd.extra2 + i];
                if (p >= prim_count) return false;
                found[p] = 1;
            }
        } else {
            stack.push_back(nd.extra2);
            stack.push_back(nd.extra2 + 1);
        }
    }
    for (uint8_t f : found)
        if (!f) return false;
    return true;
}
// sum_inner HA(n)/HA(root)*Ct + sum_leaf HA(l)/HA(root)*count(l), Ct = 1.0
static inline double sah_cost(const BvhNode* nodes, size_t n_nodes) {
    if (n_nodes == 0) return 0.0;
    double root = (double)half_area(nodes[0]);
    double cost = 0.0;
    std::vector<int32_t> stack{0};
    while (!stack.empty()) {
        const BvhNode& nd = nodes[stack.back()];
        stack.pop_back();
        if (nd.extra1 >= 0) {
            cost += (double)half_area(nd) * (double)nd.extra1;
        } else if (nd.extra2 >= 0) {
            cost += (double)half_area(nd) * (double)kTraversalCost;
            stack.push_back(nd.extra2);
            stack.push_back(nd.extra2 + 1);
        }
    }
    return cost / root;
}

}  // namespace rto
