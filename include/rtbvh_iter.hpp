// rtbvh_iter.hpp — host-side mirror of the reference's index iterators, used ONLY by the legacy
// per-candidate callback entry points of include/rtbvh.h (intersect, intersect_packet,
// intersect_mbvh, intersect_mbvh_packet).  A host function pointer cannot be invoked from a kernel,
// so these four entry points necessarily run where the callback lives; they are the compatibility
// shim of the drop-in boundary, not the measured path (that is rtbvh_gpu_intersect* in capi.cu).
//
// Mirrors: src/iter_indices.rs (BvhIndexIterator :18-107, BvhPacketIndexIterator :109-210,
// MbvhIndexIterator :212-313, MbvhPacketIndexIterator :315-415), src/aabb.rs:146-244,
// src/bvh_node.rs:150-211, src/mbvh_node.rs:177-295, src/ray.rs:166-182.
// Iterator protocol: `bool next(uint32_t* prim)` == Rust's `next() -> Option<(u32, &mut Ray)>`;
// the caller mutates ray.t / packet.t between calls exactly like the reference's loop body.
#pragma once
#include <cmath>
#include <cstdint>

#include "rtbvh.h"

namespace rtbvh_host {

struct Ray {  // src/ray.rs:9-16
    float origin[3];
    float t_min;
    float direction[3];
    float t;
    float inv_direction[3];
    uint8_t signs[4];
    static Ray make(const float* o, const float* d) {  // Ray::new, src/ray.rs:166-182
        Ray r;
        for (int k = 0; k < 3; k++) {
            r.origin[k] = o[k];
            r.direction[k] = d[k];
            r.inv_direction[k] = 1.0f / d[k];
            r.signs[k] = d[k] < 0.0f;
        }
        r.signs[3] = 0;
        r.t_min = 1e-4f;
        r.t = 1e34f;
        return r;
    }
    bool has_nan() const {
        for (int k = 0; k < 3; k++)
            if (std::isnan(origin[k]) || std::isnan(direction[k])) return true;
        return false;
    }
};

struct RayPacket4 {  // src/ray.rs:47-61
    float origin[3][4], direction[3][4], inv_direction[3][4];
    float t[4];
    bool has_nan() const {
        for (int k = 0; k < 3; k++)
            for (int l = 0; l < 4; l++)
                if (std::isnan(origin[k][l]) || std::isnan(direction[k][l])) return true;
        return false;
    }
};

// _mm_min_ps / _mm_max_ps operand rule (glam Vec4 on x86-64)
static inline float vmin(float a, float b) { return a < b ? a : b; }
static inline float vmax(float a, float b) { return a > b ? a : b; }

static inline bool aabb_intersect(const RTAabb& b, const Ray& r, float* key) {  // src/aabb.rs:146-181
    const float* p[2] = {b.min, b.max};
    float ray_min = (p[r.signs[0]][0] - r.origin[0]) * r.inv_direction[0];
    float ray_max = (p[1 - r.signs[0]][0] - r.origin[0]) * r.inv_direction[0];
    const float y_min = (p[r.signs[1]][1] - r.origin[1]) * r.inv_direction[1];
    const float y_max = (p[1 - r.signs[1]][1] - r.origin[1]) * r.inv_direction[1];
    if ((ray_min > y_max) || (y_min > ray_max)) return false;
    if (y_min > ray_min) ray_min = y_min;
    if (y_max < ray_max) ray_max = y_max;
    const float z_min = (p[r.signs[2]][2] - r.origin[2]) * r.inv_direction[2];
    const float z_max = (p[1 - r.signs[2]][2] - r.origin[2]) * r.inv_direction[2];
    if ((ray_min > z_max) || (z_min > ray_max)) return false;
    if (z_max < ray_max) ray_max = z_max;
    *key = ray_max;
    return ray_max > r.t_min;
}

static inline bool aabb_intersect4(const RTAabb& b, const RayPacket4& p, float key[4]) {  // src/aabb.rs:218-244
    bool any = false;
    for (int l = 0; l < 4; l++) {
        float lo[3], hi[3];
        for (int k = 0; k < 3; k++) {
            const float t1 = (b.min[k] - p.origin[k][l]) * p.inv_direction[k][l];
            const float t2 = (b.max[k] - p.origin[k][l]) * p.inv_direction[k][l];
            lo[k] = vmin(t1, t2);
            hi[k] = vmax(t1, t2);
        }
        const float t_min = vmax(lo[0], vmax(lo[1], lo[2]));
        const float t_max = vmin(hi[0], vmin(hi[1], hi[2]));
        key[l] = t_min;
        any |= (t_max > 0.0f) && (t_max > t_min) && (t_min < p.t[l]);
    }
    return any;
}

struct MbvhHit {  // src/mbvh_node.rs:12-15
    uint8_t ids[4] = {0, 0, 0, 0};
    bool result[4] = {false, false, false, false};
};

static inline MbvhHit mbvh_intersect(const RTMbvhNode& n, const Ray& r) {  // src/mbvh_node.rs:177-240
    MbvhHit h;
    float key[4];
    for (int s = 0; s < 4; s++) {
        const float tx0 = (n.min_x[s] - r.origin[0]) * r.inv_direction[0], tx1 = (n.max_x[s] - r.origin[0]) * r.inv_direction[0];
        const float ty0 = (n.min_y[s] - r.origin[1]) * r.inv_direction[1], ty1 = (n.max_y[s] - r.origin[1]) * r.inv_direction[1];
        const float tz0 = (n.min_z[s] - r.origin[2]) * r.inv_direction[2], tz1 = (n.max_z[s] - r.origin[2]) * r.inv_direction[2];
        const float t_min = vmax(vmin(tx0, tx1), vmax(vmin(ty0, ty1), vmin(tz0, tz1)));
        const float t_max = vmin(vmax(tx0, tx1), vmin(vmax(ty0, ty1), vmax(tz0, tz1)));
        h.result[s] = (t_max >= t_min) && (t_min < r.t);
        key[s] = t_min;
        h.ids[s] = (uint8_t)s;
    }
    auto cswap = [&](int i, int j) {
        if (key[i] > key[j]) {
            const float k = key[i]; key[i] = key[j]; key[j] = k;
            const uint8_t d = h.ids[i]; h.ids[i] = h.ids[j]; h.ids[j] = d;
        }
    };
    cswap(0, 1); cswap(2, 3); cswap(0, 2); cswap(1, 3);
    if (key[2] > key[3]) { const uint8_t d = h.ids[2]; h.ids[2] = h.ids[3]; h.ids[3] = d; }
    return h;
}

static inline MbvhHit mbvh_intersect4(const RTMbvhNode& n, const RayPacket4& p) {  // src/mbvh_node.rs:243-295
    MbvhHit h;
    for (int s = 0; s < 4; s++) {
        bool res = false;
        for (int i = 0; i < 4; i++) {
            float t1 = (n.min_x[s] - p.origin[0][i]) * p.inv_direction[0][i], t2 = (n.max_x[s] - p.origin[0][i]) * p.inv_direction[0][i];
            float t_min = vmin(t1, t2), t_max = vmax(t1, t2);
            t1 = (n.min_y[s] - p.origin[1][i]) * p.inv_direction[1][i]; t2 = (n.max_y[s] - p.origin[1][i]) * p.inv_direction[1][i];
            t_min = vmax(t_min, vmin(t1, t2)); t_max = vmin(t_max, vmax(t1, t2));
            t1 = (n.min_z[s] - p.origin[2][i]) * p.inv_direction[2][i]; t2 = (n.max_z[s] - p.origin[2][i]) * p.inv_direction[2][i];
            t_min = vmax(t_min, vmin(t1, t2)); t_max = vmin(t_max, vmax(t1, t2));
            res |= (t_max > t_min) && (t_min < p.t[i]);
        }
        h.result[s] = res;
        h.ids[s] = (uint8_t)s;
    }
    return h;
}

constexpr int kHostStack = 256;  // reference: 32 entries, panic / UB beyond (src/iter.rs:25)

// BvhIndexIterator / BvhPacketIndexIterator (src/iter_indices.rs:18-210)
template <class RayT>
class BvhIndexIteratorT {
  public:
    BvhIndexIteratorT(RayT* ray, const RTBvhNode* nodes, size_t node_count, const uint32_t* indices)
        : ray_(ray), nodes_(nodes), indices_(indices) {
        stack_ptr_ = (node_count == 0 || ray->has_nan()) ? -1 : 0;
        stack_[0] = 0;
    }
    bool next(uint32_t* prim) {
        for (;;) {
            if (stack_ptr_ < 0) return false;
            const RTAabb& node = nodes_[stack_[stack_ptr_]].aabb;
            stack_ptr_--;
            const int32_t count = node.count, left_first = node.left_first;
            if (count > -1) {
                if (i_ < count) {
                    *prim = indices_[left_first + i_];
                    i_++;
                    stack_ptr_++;  // the leaf stays on the stack until exhausted
                    return true;
                }
                i_ = 0;
            } else if (left_first > -1) {
                push_children(nodes_[left_first].aabb, nodes_[left_first + 1].aabb, left_first);
            }
        }
    }

  private:
    void push(int32_t v) {
        if (stack_ptr_ + 1 < kHostStack) stack_[++stack_ptr_] = v;
    }
    void push_children(const RTAabb& l, const RTAabb& r, int32_t left_first);
    RayT* ray_;
    const RTBvhNode* nodes_;
    const uint32_t* indices_;
    int32_t i_ = 0;
    int32_t stack_[kHostStack];
    int32_t stack_ptr_;
};
template <>
inline void BvhIndexIteratorT<Ray>::push_children(const RTAabb& l, const RTAabb& r, int32_t left_first) {
    float kl = 0.f, kr = 0.f;  // BvhNode::sort_nodes, src/bvh_node.rs:150-177
    const bool hl = aabb_intersect(l, *ray_, &kl), hr = aabb_intersect(r, *ray_, &kr);
    if (hl && hr) {
        if (kl < kr) { push(left_first); push(left_first + 1); } else { push(left_first + 1); push(left_first); }
    } else if (hl) {
        push(left_first);
    } else if (hr) {
        push(left_first + 1);
    }
}
template <>
inline void BvhIndexIteratorT<RayPacket4>::push_children(const RTAabb& l, const RTAabb& r, int32_t left_first) {
    float kl[4], kr[4];  // BvhNode::sort_nodes4, src/bvh_node.rs:180-211
    const bool hl = aabb_intersect4(l, *ray_, kl), hr = aabb_intersect4(r, *ray_, kr);
    if (hl && hr) {
        bool any_lt = false;
        for (int k = 0; k < 4; k++) any_lt |= kl[k] < kr[k];
        if (any_lt) { push(left_first); push(left_first + 1); } else { push(left_first + 1); push(left_first); }
    } else if (hl) {
        push(left_first);
    } else if (hr) {
        push(left_first + 1);
    }
}
using BvhIndexIterator = BvhIndexIteratorT<Ray>;
using BvhPacketIndexIterator = BvhIndexIteratorT<RayPacket4>;

// MbvhIndexIterator / MbvhPacketIndexIterator (src/iter_indices.rs:212-415)
template <class RayT>
class MbvhIndexIteratorT {
  public:
    MbvhIndexIteratorT(RayT* ray, const RTMbvhNode* nodes, size_t node_count, const uint32_t* indices)
        : ray_(ray), nodes_(nodes), indices_(indices), empty_(node_count == 0) {
        if (!empty_) hit_ = test(nodes_[0]);
    }
    bool next(uint32_t* prim) {
        if (empty_) return false;
        for (;;) {
            const RTMbvhNode& node = nodes_[current_];
            if (i_ >= 4) {
                if (stack_ptr_ < 0) return false;
                current_ = stack_[stack_ptr_--];
                hit_ = test(nodes_[current_]);
                i_ = 0;
            } else {
                const int id = hit_.ids[3 - i_];
                if (hit_.result[id]) {
                    const int32_t count = node.counts[id], left_first = node.children[id];
                    if (count > -1) {
                        if (j_ < count) {
                            *prim = indices_[left_first + j_];
                            j_++;
                            return true;
                        }
                        j_ = 0;
                    } else if (left_first > -1) {
                        if (stack_ptr_ + 1 < kHostStack) stack_[++stack_ptr_] = left_first;
                    }
                }
                i_++;
            }
        }
    }

  private:
    MbvhHit test(const RTMbvhNode& n) const;
    RayT* ray_;
    const RTMbvhNode* nodes_;
    const uint32_t* indices_;
    bool empty_;
    MbvhHit hit_;
    int32_t current_ = 0, i_ = 0, j_ = 0;
    int32_t stack_[kHostStack];
    int32_t stack_ptr_ = -1;
};
template <>
inline MbvhHit MbvhIndexIteratorT<Ray>::test(const RTMbvhNode& n) const { return mbvh_intersect(n, *ray_); }
template <>
inline MbvhHit MbvhIndexIteratorT<RayPacket4>::test(const RTMbvhNode& n) const { return mbvh_intersect4(n, *ray_); }
using MbvhIndexIterator = MbvhIndexIteratorT<Ray>;
using MbvhPacketIndexIterator = MbvhIndexIteratorT<RayPacket4>;

}  // namespace rtbvh_host
