// traverse.cu — closest-hit / any-hit traversal of Bvh and Mbvh for single rays and RayPacket4.
//
// Replaces, as ONE batched kernel per (tree, ray kind, query), the loop every caller of the reference
// writes around its iterators:
//     src/iter_indices.rs:69-106   BvhIndexIterator::next          + src/aabb.rs:146-181, src/bvh_node.rs:150-177
//     src/iter_indices.rs:172-209  BvhPacketIndexIterator::next    + src/aabb.rs:218-244, src/bvh_node.rs:180-211
//     src/iter_indices.rs:267-312  MbvhIndexIterator::next         + src/mbvh_node.rs:177-240
//     src/iter_indices.rs:370-414  MbvhPacketIndexIterator::next   + src/mbvh_node.rs:243-295
//     src/builders/spatial_sah.rs:131-163 / :165-244               SpatialTriangle::intersect / intersect4
//     examples/benchmark.rs:25-31, :55-61; rtbvh_ffi/src/lib.rs:572-576 (callback `true` => break => any hit)
//
// Parity contract: identical visitation rules and predicates, fp32 with one
// rounding per operation (the *_rn intrinsics never fuse).  Inside one node the reference's slot
// results are frozen at node entry (`hit` is computed once, iter_indices.rs:281-285), so the leaf
// slots of a node may be tested in any order without changing ray.t after the node; inner slots are
// pushed in the reference's order (ids[3], ids[2], ids[1], ids[0] of its 5-comparator network).
//
// Layout / mapping (B200): one thread per ray (packets: 4 adjacent lanes = one RayPacket4, decisions
// by quad vote).  Node = 8 x LDG.128 (Mbvh, 128 B = one L2 line) or a 64-byte adjacent child pair
// (Bvh), fetched with 256-bit loads (4 resp. 2 LDG.256 per visit: one L1 wavefront per instruction
// and lane).  Triangles are 64-byte leaf-ordered records (LDG.256 + LDG.128, no index indirection).  The
// traversal stack lives in shared memory, [entry][thread] so it is bank-conflict free.
#include <cub/cub.cuh>

#include "traverse.cuh"

namespace rtb {

namespace {

#ifndef RTB_SMEM_STACK
#define RTB_SMEM_STACK 32
#endif
constexpr int kSmemStack = RTB_SMEM_STACK;        // entries kept in shared memory (32 = the reference's whole stack, src/iter.rs:25)
constexpr int kSpillStack = 128 - kSmemStack;     // further entries spill to thread-local memory (only touched by deep rays); beyond 128 -> overflow flag
#ifndef RTB_BLOCK
#define RTB_BLOCK 128
#endif
constexpr int kBlock = RTB_BLOCK;
#ifndef RTB_PBLOCK
#define RTB_PBLOCK RTB_BLOCK
#endif
constexpr int kPBlock = RTB_PBLOCK;  // threads per block of the persistent single-ray kernel
#ifndef RTB_TOPK
#define RTB_TOPK 0
#endif
constexpr int kTopK = RTB_TOPK;      // Mbvh nodes staged in shared memory by the persistent single-ray kernel (0: none)
constexpr int kTopRow = 9;           // float4 per staged node: 128 B + 16 B padding (rows 144 B apart: LDS.128 of lanes on
                                     // different rows spreads over all banks)
#ifndef RTB_REFILL
#define RTB_REFILL 8
#endif
#ifndef RTB_MINBLOCKS
#define RTB_MINBLOCKS 1
#endif
constexpr int kRefillIdle = RTB_REFILL;  // persistent kernels: refill a warp once this many lanes are idle
#ifndef RTB_RAYCHUNK
#define RTB_RAYCHUNK 64
#endif
constexpr unsigned kRayChunk = RTB_RAYCHUNK;  // rays a warp reserves per global atomic (measured: 256 -> 1665, 128 -> 1816, 64 -> 1899, 32 -> 1911, 16 -> 1884 Mrays/s; 64 is best end to end)

struct RayRegs {
    float ox, oy, oz;
    float dx, dy, dz;
    float ix, iy, iz;
    float t_min;   // single rays: ray.t_min; packets: the t_min argument of intersect4
    float t;       // ray.t / packet.t[lane]
    uint32_t prim;
    bool exact;    // slab products may be NaN: use the SSE min/max operand rule
    bool nan;      // origin or direction has a NaN component
};

__device__ __forceinline__ void finish_ray_setup(RayRegs& r) {
    // Ray::new: inv_direction = Vec3::ONE / direction (src/ray.rs:179)
    r.ix = fdiv(1.0f, r.dx);
    r.iy = fdiv(1.0f, r.dy);
    r.iz = fdiv(1.0f, r.dz);
    r.prim = kNoHit;
    const bool fin = isfinite(r.ox) && isfinite(r.oy) && isfinite(r.oz) && isfinite(r.dx) && isfinite(r.dy) &&
                     isfinite(r.dz) && isfinite(r.ix) && isfinite(r.iy) && isfinite(r.iz);
    r.exact = !fin;
    r.nan = isnan(r.ox) || isnan(r.oy) || isnan(r.oz) || isnan(r.dx) || isnan(r.dy) || isnan(r.dz);
}

// ---- SpatialTriangle::intersect (single) / intersect4 lane (packet) -----------------------------
// Returns true when the candidate was accepted with t < ray.t (ray.t shrunk).  Ties (t == ray.t from
// an earlier accepted hit) only lower the reported id (north-star rule).
template <bool PACKET>
__device__ __forceinline__ bool tri_candidate(const TriRec* __restrict__ tris, int pos, RayRegs& r) {
    const F8 ab = ld256_tri(&tris[pos].a);
    const float4 A = ab.lo, E1 = ab.hi;
    const float4 E2 = __ldg(&tris[pos].c);
    // h = direction x edge2
    const float hx = fsub(fmul(r.dy, E2.z), fmul(E2.y, r.dz));
    const float hy = fsub(fmul(r.dz, E2.x), fmul(E2.z, r.dx));
    const float hz = fsub(fmul(r.dx, E2.y), fmul(E2.x, r.dy));
    const float a = fadd(fadd(fmul(E1.x, hx), fmul(E1.y, hy)), fmul(E1.z, hz));
    if (PACKET) {
        if (!(a <= -1e-6f || a >= 1e-6f)) return false;  // spatial_sah.rs:191-196
    } else {
        if (a > -1e-5f && a < 1e-5f) return false;  // spatial_sah.rs:140-142
    }
    const float f = fdiv(1.0f, a);
    const float sx = fsub(r.ox, A.x), sy = fsub(r.oy, A.y), sz = fsub(r.oz, A.z);
    const float u = fmul(f, fadd(fadd(fmul(sx, hx), fmul(sy, hy)), fmul(sz, hz)));
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    // q = s x edge1
    const float qx = fsub(fmul(sy, E1.z), fmul(E1.y, sz));
    const float qy = fsub(fmul(sz, E1.x), fmul(E1.z, sx));
    const float qz = fsub(fmul(sx, E1.y), fmul(E1.x, sy));
    const float v = fmul(f, fadd(fadd(fmul(r.dx, qx), fmul(r.dy, qy)), fmul(r.dz, qz)));
    if (PACKET) {
        if (!(v >= 0.0f && fadd(u, v) <= 1.0f)) return false;  // spatial_sah.rs:218-222
    } else {
        if (v < 0.0f || fadd(u, v) > 1.0f) return false;  // spatial_sah.rs:150-152
    }
    const float t = fmul(f, fadd(fadd(fmul(E2.x, qx), fmul(E2.y, qy)), fmul(E2.z, qz)));
    const bool above = PACKET ? (t >= r.t_min) : (t > r.t_min);
    if (!above) return false;
    const uint32_t id = __float_as_uint(A.w);
    if (t < r.t) {
        r.t = t;
        r.prim = id;
        return true;
    }
    if (r.prim != kNoHit && t == r.t && id < r.prim) r.prim = id;
    return false;
}

// ---- Aabb::intersect (src/aabb.rs:146-181) -------------------------------------------------------
__device__ __forceinline__ bool aabb_single(const float4 lo, const float4 hi, const RayRegs& r, float& key) {
    const bool sx = r.dx < 0.0f, sy = r.dy < 0.0f, sz = r.dz < 0.0f;
    float ray_min = fmul(fsub(sx ? hi.x : lo.x, r.ox), r.ix);
    float ray_max = fmul(fsub(sx ? lo.x : hi.x, r.ox), r.ix);
    const float y_min = fmul(fsub(sy ? hi.y : lo.y, r.oy), r.iy);
    const float y_max = fmul(fsub(sy ? lo.y : hi.y, r.oy), r.iy);
    if ((ray_min > y_max) || (y_min > ray_max)) return false;
    if (y_min > ray_min) ray_min = y_min;
    if (y_max < ray_max) ray_max = y_max;
    const float z_min = fmul(fsub(sz ? hi.z : lo.z, r.oz), r.iz);
    const float z_max = fmul(fsub(sz ? lo.z : hi.z, r.oz), r.iz);
    if ((ray_min > z_max) || (z_min > ray_max)) return false;
    if (z_max < ray_max) ray_max = z_max;
    key = ray_max;
    return ray_max > r.t_min;
}

// ---- Aabb::intersect4, one lane (src/aabb.rs:218-244) -------------------------------------------
template <bool EXACT>
__device__ __forceinline__ bool aabb_lane(const float4 lo, const float4 hi, const RayRegs& r, float& key) {
    const float t1x = fmul(fsub(lo.x, r.ox), r.ix), t1y = fmul(fsub(lo.y, r.oy), r.iy), t1z = fmul(fsub(lo.z, r.oz), r.iz);
    const float t2x = fmul(fsub(hi.x, r.ox), r.ix), t2y = fmul(fsub(hi.y, r.oy), r.iy), t2z = fmul(fsub(hi.z, r.oz), r.iz);
    const float tmin = vmax<EXACT>(vmin<EXACT>(t1x, t2x), vmax<EXACT>(vmin<EXACT>(t1y, t2y), vmin<EXACT>(t1z, t2z)));
    const float tmax = vmin<EXACT>(vmax<EXACT>(t1x, t2x), vmin<EXACT>(vmax<EXACT>(t1y, t2y), vmax<EXACT>(t1z, t2z)));
    key = tmin;
    return tmax > 0.0f && tmax > tmin && tmin < r.t;
}

// ---- MbvhNode::intersect slab part (src/mbvh_node.rs:177-206) ------------------------------------
// key[s] = t_min of slot s; returns the 4-bit `result` mask (t_max >= t_min && t_min < ray.t).
template <bool EXACT>
__device__ __forceinline__ uint32_t mbvh_slabs(const float4 mnx, const float4 mxx, const float4 mny, const float4 mxy,
                                               const float4 mnz, const float4 mxz, const RayRegs& r, float key[4]) {
    const float a_mnx[4] = {mnx.x, mnx.y, mnx.z, mnx.w}, a_mxx[4] = {mxx.x, mxx.y, mxx.z, mxx.w};
    const float a_mny[4] = {mny.x, mny.y, mny.z, mny.w}, a_mxy[4] = {mxy.x, mxy.y, mxy.z, mxy.w};
    const float a_mnz[4] = {mnz.x, mnz.y, mnz.z, mnz.w}, a_mxz[4] = {mxz.x, mxz.y, mxz.z, mxz.w};
    uint32_t mask = 0;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const float tx0 = fmul(fsub(a_mnx[s], r.ox), r.ix), tx1 = fmul(fsub(a_mxx[s], r.ox), r.ix);
        const float ty0 = fmul(fsub(a_mny[s], r.oy), r.iy), ty1 = fmul(fsub(a_mxy[s], r.oy), r.iy);
        const float tz0 = fmul(fsub(a_mnz[s], r.oz), r.iz), tz1 = fmul(fsub(a_mxz[s], r.oz), r.iz);
        const float tmn = vmax<EXACT>(vmin<EXACT>(tx0, tx1), vmax<EXACT>(vmin<EXACT>(ty0, ty1), vmin<EXACT>(tz0, tz1)));
        const float tmx = vmin<EXACT>(vmax<EXACT>(tx0, tx1), vmin<EXACT>(vmax<EXACT>(ty0, ty1), vmax<EXACT>(tz0, tz1)));
        key[s] = tmn;
        if (tmx >= tmn && tmn < r.t) mask |= 1u << s;
    }
    return mask;
}

// ---- MbvhNode::intersect4, one lane = one ray against the 4 slots (src/mbvh_node.rs:243-281) -----
template <bool EXACT>
__device__ __forceinline__ uint32_t mbvh_slabs_lane(const float4 mnx, const float4 mxx, const float4 mny,
                                                    const float4 mxy, const float4 mnz, const float4 mxz,
                                                    const RayRegs& r) {
    const float a_mnx[4] = {mnx.x, mnx.y, mnx.z, mnx.w}, a_mxx[4] = {mxx.x, mxx.y, mxx.z, mxx.w};
    const float a_mny[4] = {mny.x, mny.y, mny.z, mny.w}, a_mxy[4] = {mxy.x, mxy.y, mxy.z, mxy.w};
    const float a_mnz[4] = {mnz.x, mnz.y, mnz.z, mnz.w}, a_mxz[4] = {mxz.x, mxz.y, mxz.z, mxz.w};
    uint32_t mask = 0;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        float t1 = fmul(fsub(a_mnx[s], r.ox), r.ix), t2 = fmul(fsub(a_mxx[s], r.ox), r.ix);
        float tmin = vmin<EXACT>(t1, t2), tmax = vmax<EXACT>(t1, t2);
        t1 = fmul(fsub(a_mny[s], r.oy), r.iy);
        t2 = fmul(fsub(a_mxy[s], r.oy), r.iy);
        tmin = vmax<EXACT>(tmin, vmin<EXACT>(t1, t2));
        tmax = vmin<EXACT>(tmax, vmax<EXACT>(t1, t2));
        t1 = fmul(fsub(a_mnz[s], r.oz), r.iz);
        t2 = fmul(fsub(a_mxz[s], r.oz), r.iz);
        tmin = vmax<EXACT>(tmin, vmin<EXACT>(t1, t2));
        tmax = vmin<EXACT>(tmax, vmax<EXACT>(t1, t2));
        if (tmax > tmin && tmin < r.t) mask |= 1u << s;
    }
    return mask;
}

// L1 prefetch of the 128-byte line at p (no destination register)
__device__ __forceinline__ void prefetch_l1(const void* p) {
#if RTB_PREFETCH
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}

__device__ __forceinline__ int sel4(const int4 v, int s) { return s == 0 ? v.x : (s == 1 ? v.y : (s == 2 ? v.z : v.w)); }

#define RTB_CSWAP(i, j)                 \
    if (key[i] > key[j]) {              \
        float tk = key[i];              \
        key[i] = key[j];                \
        key[j] = tk;                    \
        int tp = pay[i];                \
        pay[i] = pay[j];                \
        pay[j] = tp;                    \
    }

struct Stack {
    int* base;     // &smem[threadIdx.x]; entry e < kSmemStack at base[e * stride]
    int* spill;    // thread-local array; entry e >= kSmemStack at spill[e - kSmemStack]
    int sp;
    uint32_t* overflow;
    int stride;    // threads per block
    __device__ __forceinline__ void push(int v) {
        if (sp < kSmemStack) {
            base[sp * stride] = v;
            sp++;
        } else if (sp < kSmemStack + kSpillStack) {
            spill[sp - kSmemStack] = v;
            sp++;
        } else {
            *overflow = 1u;
        }
    }
    __device__ __forceinline__ int pop() {
        --sp;
        return sp < kSmemStack ? base[sp * stride] : spill[sp - kSmemStack];
    }
    __device__ __forceinline__ void reset() { sp = 0; }
};

// ================================================================================================
// Mbvh, single rays  (MbvhIndexIterator).
// ================================================================================================
struct MNode {  // one 128-byte MbvhNode in registers
    float4 mnx, mxx, mny, mxy, mnz, mxz;
    int4 ch, cn;
};
__device__ __forceinline__ int4 as_int4(const float4 f) {
    return make_int4(__float_as_int(f.x), __float_as_int(f.y), __float_as_int(f.z), __float_as_int(f.w));
}
__device__ __forceinline__ MNode mnode_load_global(const float4* __restrict__ nodes, int cur) {
    const float4* n = nodes + (size_t)cur * 8;
    const F8 q0 = ld256_node(n), q1 = ld256_node(n + 2), q2 = ld256_node(n + 4), q3 = ld256_node(n + 6);
    return MNode{q0.lo, q0.hi, q1.lo, q1.hi, q2.lo, q2.hi, as_int4(q3.lo), as_int4(q3.hi)};
}

// Node entry: MbvhNode::intersect with the ray.t of this moment, then the inner slots are pushed in the
// order ids[3], ids[2], ids[1], ids[0] (iter_indices.rs:287, :304-309).  The pushes do not depend on what
// the leaf slots of this node do to ray.t (the slot results are frozen at node entry), so they are done
// BEFORE the leaf slots: the next node is then known early and can be fetched while triangles are tested.
// Returns the next node (-1: stack empty); `leaves` = hit leaf slots still to be tested.
__device__ __forceinline__ int mbvh_visit_push(const MNode& nd, const RayRegs& r, Stack& st, uint32_t& leaves) {
    float key[4];
    const uint32_t mask = r.exact ? mbvh_slabs<true>(nd.mnx, nd.mxx, nd.mny, nd.mxy, nd.mnz, nd.mxz, r, key)
                                  : mbvh_slabs<false>(nd.mnx, nd.mxx, nd.mny, nd.mxy, nd.mnz, nd.mxz, r, key);
    const int4 ch = nd.ch, cn = nd.cn;
    const uint32_t leafbits = (cn.x > -1 ? 1u : 0u) | (cn.y > -1 ? 2u : 0u) | (cn.z > -1 ? 4u : 0u) | (cn.w > -1 ? 8u : 0u);
    const uint32_t childbits = (ch.x > -1 ? 1u : 0u) | (ch.y > -1 ? 2u : 0u) | (ch.z > -1 ? 4u : 0u) | (ch.w > -1 ? 8u : 0u);
    leaves = mask & leafbits;
    const uint32_t inner = mask & ~leafbits & childbits;
    if (inner) {
        int pay[4] = {(inner & 1u) ? ch.x : -1, (inner & 2u) ? ch.y : -1, (inner & 4u) ? ch.z : -1, (inner & 8u) ? ch.w : -1};
        // the reference's 5-comparator network; the last comparator swaps ids only (mbvh_node.rs:219-237)
        RTB_CSWAP(0, 1)
        RTB_CSWAP(2, 3)
        RTB_CSWAP(0, 2)
        RTB_CSWAP(1, 3)
        if (key[2] > key[3]) {
            int tp = pay[2];
            pay[2] = pay[3];
            pay[3] = tp;
        }
        // Push order is pay[3], pay[2], pay[1], pay[0]; the entry pushed last is popped right away, so it is
        // handed over in a register (no shared-memory round trip on the critical path) and only the others
        // are stored.  `next` = first valid of pay[0], pay[1], pay[2], pay[3].
        const int next = pay[0] >= 0 ? pay[0] : (pay[1] >= 0 ? pay[1] : (pay[2] >= 0 ? pay[2] : pay[3]));
        const int skip = pay[0] >= 0 ? 0 : (pay[1] >= 0 ? 1 : (pay[2] >= 0 ? 2 : 3));
        if (st.sp + 3 <= kSmemStack) {  // fast path: compact predicated stores
            int* b = st.base + st.sp * st.stride;
            int c = 0;
            if (skip < 3 && pay[3] >= 0) { b[c * st.stride] = pay[3]; c++; }
            if (skip < 2 && pay[2] >= 0) { b[c * st.stride] = pay[2]; c++; }
            if (skip < 1 && pay[1] >= 0) { b[c * st.stride] = pay[1]; c++; }
            st.sp += c;
        } else {
            if (skip < 3 && pay[3] >= 0) st.push(pay[3]);
            if (skip < 2 && pay[2] >= 0) st.push(pay[2]);
            if (skip < 1 && pay[1] >= 0) st.push(pay[1]);
        }
        return next;
    }
    return st.sp > 0 ? st.pop() : -1;
}
// leaf slots: yield every primitive (iter_indices.rs:292-303).  Returns true when an any-hit query is done.
template <bool ANY>
__device__ __forceinline__ bool mbvh_visit_leaves(const int4 ch, const int4 cn, uint32_t leaves, const DeviceTree& tree,
                                                  RayRegs& r) {
    while (leaves) {
        const int s = __ffs(leaves) - 1;
        leaves &= leaves - 1;
        const int first = sel4(ch, s), count = sel4(cn, s);
        for (int j = 0; j < count; j++) {
            const bool hit = tri_candidate<false>(tree.tris, first + j, r);
            if (ANY && hit) return true;
        }
    }
    return false;
}
// One call = one node visit; returns true when the ray is done.
template <bool ANY>
__device__ __forceinline__ bool mbvh_single_step(const DeviceTree& tree, RayRegs& r, Stack& st, int& cur) {
    const MNode nd = mnode_load_global(tree.nodes, cur);
    uint32_t leaves;
    const int next = mbvh_visit_push(nd, r, st, leaves);
    if (next >= 0) prefetch_l1(tree.nodes + (size_t)next * 8);
    if (mbvh_visit_leaves<ANY>(nd.ch, nd.cn, leaves, tree, r)) return true;
    if (next < 0) return true;
    cur = next;
    return false;
}

// ================================================================================================
// Bvh, single rays  (BvhIndexIterator).  One call = one popped node (the root is popped without a box
// test, iter_indices.rs:32-46); returns true when the stack is empty.
// ================================================================================================
template <bool ANY>
__device__ __forceinline__ bool bvh_single_step(const DeviceTree& tree, RayRegs& r, Stack& st, int& cur) {
    const float4* __restrict__ nodes = tree.nodes;
    const F8 nd = ld256(nodes + (size_t)cur * 2);
    const int count = __float_as_int(nd.lo.w), left_first = __float_as_int(nd.hi.w);
    int next = -1;
    if (count > -1) {
        for (int i = 0; i < count; i++) {
            const bool hit = tri_candidate<false>(tree.tris, left_first + i, r);
            if (ANY && hit) return true;
        }
    } else if (left_first > -1) {
        const float4* c = nodes + (size_t)left_first * 2;
        const F8 lc = ld256(c), rc = ld256(c + 2);
        const float4 l0 = lc.lo, l1 = lc.hi, r0 = rc.lo, r1 = rc.hi;
        float kl = 0.f, kr = 0.f;
        const bool hl = aabb_single(l0, l1, r, kl);
        const bool hr = aabb_single(r0, r1, r, kr);
        // BvhNode::sort_nodes (bvh_node.rs:150-177); the entry it pushes last is popped next, so it stays in a register
        if (hl && hr) {
            if (kl < kr) {
                st.push(left_first);
                next = left_first + 1;
            } else {
                st.push(left_first + 1);
                next = left_first;
            }
        } else if (hl) {
            next = left_first;
        } else if (hr) {
            next = left_first + 1;
        }
    }
    if (next < 0) {
        if (st.sp == 0) return true;
        next = st.pop();
    }
    cur = next;
    return false;
}

template <int TREE, bool ANY>
__device__ __forceinline__ bool single_step(const DeviceTree& tree, RayRegs& r, Stack& st, int& cur) {
    if (TREE == RT_TREE_MBVH) return mbvh_single_step<ANY>(tree, r, st, cur);
    return bvh_single_step<ANY>(tree, r, st, cur);
}

// ================================================================================================
// Packets: 4 adjacent lanes = one RayPacket4; every traversal decision is a quad vote, so the four
// lanes execute the same control flow, exactly like the SSE lanes of the reference.
// ================================================================================================
__device__ __forceinline__ uint32_t quad_mask() { return 0xFu << (threadIdx.x & 28u); }

// any-hit retirement of a lane: the callback writes t = -1e34 (see rtbvh_gpu.h)
template <bool ANY>
__device__ __forceinline__ bool packet_candidate(const TriRec* __restrict__ tris, int pos, RayRegs& r, bool& retired,
                                                 uint32_t qm) {
    const bool hit = tri_candidate<true>(tris, pos, r);
    if (ANY) {
        if (hit) {
            r.t = -1e34f;
            retired = true;
        }
        return __all_sync(qm, retired) != 0;
    }
    return false;
}

// One call = one node visit of the packet; returns true when the packet is done (quad-uniform).
template <bool ANY>
__device__ __forceinline__ bool mbvh_packet_step(const DeviceTree& tree, RayRegs& r, Stack& st, int& cur, bool& retired,
                                                 uint32_t qm) {
    const MNode nd = mnode_load_global(tree.nodes, cur);
    const uint32_t mine = r.exact ? mbvh_slabs_lane<true>(nd.mnx, nd.mxx, nd.mny, nd.mxy, nd.mnz, nd.mxz, r)
                                  : mbvh_slabs_lane<false>(nd.mnx, nd.mxx, nd.mny, nd.mxy, nd.mnz, nd.mxz, r);
    const uint32_t mask = __reduce_or_sync(qm, mine);  // result |= ... over the 4 rays (mbvh_node.rs:277-279)
    // no ordering: ids = [0,1,2,3], slots visited 3, 2, 1, 0 (iter_indices.rs:390).  As for single rays the slot
    // results are frozen at node entry, so the inner slots are pushed first (same stack order) and the leaf
    // slots are tested afterwards.
    const int4 ch = nd.ch, cn = nd.cn;
#pragma unroll
    for (int s = 3; s >= 0; s--) {
        const int first = sel4(ch, s), count = sel4(cn, s);
        if (((mask >> s) & 1u) && count <= -1 && first > -1) st.push(first);
    }
#pragma unroll 1
    for (int s = 3; s >= 0; s--) {
        if (!((mask >> s) & 1u)) continue;
        const int first = sel4(ch, s), count = sel4(cn, s);
        for (int j = 0; j < count; j++)
            if (packet_candidate<ANY>(tree.tris, first + j, r, retired, qm)) return true;
    }
    if (st.sp == 0) return true;
    cur = st.pop();
    return false;
}

// One call = one popped node (the root is popped without a box test); returns true when the stack is empty.
template <bool ANY>
__device__ __forceinline__ bool bvh_packet_step(const DeviceTree& tree, RayRegs& r, Stack& st, int& cur, bool& retired,
                                                uint32_t qm) {
    const float4* __restrict__ nodes = tree.nodes;
    const F8 nd = ld256(nodes + (size_t)cur * 2);
    const int count = __float_as_int(nd.lo.w), left_first = __float_as_int(nd.hi.w);
    int next = -1;
    if (count > -1) {
        for (int i = 0; i < count; i++)
            if (packet_candidate<ANY>(tree.tris, left_first + i, r, retired, qm)) return true;
    } else if (left_first > -1) {
        const float4* c = nodes + (size_t)left_first * 2;
        const F8 lc = ld256(c), rc = ld256(c + 2);
        const float4 l0 = lc.lo, l1 = lc.hi, r0 = rc.lo, r1 = rc.hi;
        float kl, kr;
        const bool ml = r.exact ? aabb_lane<true>(l0, l1, r, kl) : aabb_lane<false>(l0, l1, r, kl);
        const bool mr = r.exact ? aabb_lane<true>(r0, r1, r, kr) : aabb_lane<false>(r0, r1, r, kr);
        const bool hl = __any_sync(qm, ml) != 0, hr = __any_sync(qm, mr) != 0;
        if (hl && hr) {  // BvhNode::sort_nodes4: any lane with t_near_left < t_near_right (bvh_node.rs:180-211)
            if (__any_sync(qm, kl < kr)) {
                st.push(left_first);
                next = left_first + 1;
            } else {
                st.push(left_first + 1);
                next = left_first;
            }
        } else if (hl) {
            next = left_first;
        } else if (hr) {
            next = left_first + 1;
        }
    }
    if (next < 0) {
        if (st.sp == 0) return true;
        next = st.pop();
    }
    cur = next;
    return false;
}

template <int TREE, bool ANY>
__device__ __forceinline__ bool packet_step(const DeviceTree& tree, RayRegs& r, Stack& st, int& cur, bool& retired, uint32_t qm) {
    if (TREE == RT_TREE_MBVH) return mbvh_packet_step<ANY>(tree, r, st, cur, retired, qm);
    return bvh_packet_step<ANY>(tree, r, st, cur, retired, qm);
}


// ================================================================================================
// Phase-split stepping (RTBVH_TRACE_MODE=phased; the structure of the packet kernels).  ncu of the one-step-per-iteration kernel
// (profiles/r2g_*): 14.7 of 32 lanes active per issued instruction, because every iteration ran
// "node visit, then the triangles of its hit leaf slots" and ~20 % of the lanes had triangles —
// the other ~80 % idled through one or more ~80-instruction triangle tests per visit.
// Here a lane is in one of two phases — N: it holds a node to visit; T: it holds a leaf range with
// triangles left — and the WARP picks per iteration the phase most of its lanes are in; the lanes
// of the other phase wait one iteration.  A lane never visits a node while it still has triangles
// pending: the reference shrinks ray.t with the candidates of a node before it pops the next node
// (iter_indices.rs:292-309), and with its non-conservative boxes the set of visited
// nodes — hence the result — depends on that, so the order per ray stays the reference's; only the
// interleaving BETWEEN rays changes.
// Hit leaf slots of a node beyond the first stay with the lane as a 4-bit mask + the node's index
// (Lane::pmask / pnode); their ranges are re-read from the node when the current range is used up.
// ================================================================================================
#ifndef RTB_TRI_NUM
#define RTB_TRI_NUM 3  // tri phase is chosen when lanes_T * RTB_TRI_NUM >= lanes_N * RTB_TRI_DEN
#endif
#ifndef RTB_TRI_FLAT
#define RTB_TRI_FLAT 1  // T phase: the triangle test without early exits (0: the branching test of the other kernels)
#endif
#ifndef RTB_TRI_DEN
#define RTB_TRI_DEN 1
#endif
struct Lane {
    int cur;               // node to visit next (-1: none in hand: pop)
    int tri_pos, tri_end;  // leaf range in the leaf-ordered triangle records; T phase while tri_pos < tri_end
    int pnode;             // Mbvh: the node whose further hit leaf slots (pmask) are still to be tested
    uint32_t pmask;
};
// Next pending leaf slot of node L.pnode: its (first, count) words are read again from the node (two scalar loads that hit L1:
// the lane fetched this line a few instructions ago) instead of keeping eight registers alive or parking ranges on the stack.
__device__ __forceinline__ void lane_next_leaf(const DeviceTree& tree, const float4* __restrict__ top_s, Lane& L) {
    const int s = __ffs(L.pmask) - 1;
    L.pmask &= L.pmask - 1;
    int first, count;
    if (kTopK > 0 && (L.pnode & kTopFlag)) {
        const int* w = reinterpret_cast<const int*>(top_s + (L.pnode & ~kTopFlag) * kTopRow + 6);
        first = w[s];
        count = w[4 + s];
    } else {
        const int* w = reinterpret_cast<const int*>(tree.nodes + (size_t)L.pnode * 8 + 6);
        first = __ldg(w + s);
        count = __ldg(w + 4 + s);
    }
    L.tri_pos = first;
    L.tri_end = first + count;
}
// Called when the lane's current range is used up: next pending leaf slot, else the next node; true = ray finished.
__device__ __forceinline__ bool lane_advance(const DeviceTree& tree, const float4* __restrict__ top_s, Stack& st, Lane& L) {
    while (L.pmask != 0) {
        lane_next_leaf(tree, top_s, L);
        if (L.tri_pos < L.tri_end) return false;
    }
    if (L.cur < 0) {
        if (st.sp == 0) return true;
        L.cur = st.pop();
    }
    return false;
}

// SpatialTriangle::intersect without early exits (same operations, same predicates, evaluated in full): in a T-phase
// step nearly all lanes hold a triangle, so some lane needs every instruction anyway and branches only add
// divergence bookkeeping.  Returns true when the candidate was accepted with t < ray.t.
__device__ __forceinline__ bool tri_candidate_flat(const TriRec* __restrict__ tris, int pos, RayRegs& r) {
    const F8 ab = ld256_tri(&tris[pos].a);
    const float4 A = ab.lo, E1 = ab.hi;
    const float4 E2 = __ldg(&tris[pos].c);
    const float hx = fsub(fmul(r.dy, E2.z), fmul(E2.y, r.dz));
    const float hy = fsub(fmul(r.dz, E2.x), fmul(E2.z, r.dx));
    const float hz = fsub(fmul(r.dx, E2.y), fmul(E2.x, r.dy));
    const float a = fadd(fadd(fmul(E1.x, hx), fmul(E1.y, hy)), fmul(E1.z, hz));
    const bool p_par = (a > -1e-5f && a < 1e-5f);  // spatial_sah.rs:140-142
    const float f = fdiv(1.0f, a);
    const float sx = fsub(r.ox, A.x), sy = fsub(r.oy, A.y), sz = fsub(r.oz, A.z);
    const float u = fmul(f, fadd(fadd(fmul(sx, hx), fmul(sy, hy)), fmul(sz, hz)));
    const bool p_u = (u >= 0.0f && u <= 1.0f);
    const float qx = fsub(fmul(sy, E1.z), fmul(E1.y, sz));
    const float qy = fsub(fmul(sz, E1.x), fmul(E1.z, sx));
    const float qz = fsub(fmul(sx, E1.y), fmul(E1.x, sy));
    const float v = fmul(f, fadd(fadd(fmul(r.dx, qx), fmul(r.dy, qy)), fmul(r.dz, qz)));
    const bool p_v = !(v < 0.0f || fadd(u, v) > 1.0f);  // spatial_sah.rs:150-152
    const float t = fmul(f, fadd(fadd(fmul(E2.x, qx), fmul(E2.y, qy)), fmul(E2.z, qz)));
    const bool ok = !p_par && p_u && p_v && (t > r.t_min);
    const uint32_t id = __float_as_uint(A.w);
    const bool closer = ok && t < r.t;
    const bool tie = ok && r.prim != kNoHit && t == r.t && id < r.prim;
    if (closer) r.t = t;
    if (closer || tie) r.prim = id;
    return closer;
}

// N phase, Mbvh: MbvhNode::intersect at node entry; inner slots pushed in the reference's order with the entry that
// would be popped next kept in L.cur; hit leaf slots become leaf ranges.
__device__ __forceinline__ MNode mnode_load_shared(const float4* __restrict__ top_s, int slot) {
    const float4* n = top_s + slot * kTopRow;
    return MNode{n[0], n[1], n[2], n[3], n[4], n[5], as_int4(n[6]), as_int4(n[7])};
}
template <bool EXACT>
__device__ __forceinline__ void mbvh_node_phase(const DeviceTree& tree, const float4* __restrict__ top_s, const RayRegs& r,
                                                Stack& st, Lane& L) {
    MNode nd;
    if (kTopK > 0 && (L.cur & kTopFlag))  // a staged node: its child fields carry the flag for staged children
        nd = mnode_load_shared(top_s, L.cur & ~kTopFlag);
    else
        nd = mnode_load_global(tree.nodes, L.cur);
    float key[4];
    const uint32_t mask = mbvh_slabs<EXACT>(nd.mnx, nd.mxx, nd.mny, nd.mxy, nd.mnz, nd.mxz, r, key);
    const int4 ch = nd.ch, cn = nd.cn;
    const uint32_t leafbits = (cn.x > -1 ? 1u : 0u) | (cn.y > -1 ? 2u : 0u) | (cn.z > -1 ? 4u : 0u) | (cn.w > -1 ? 8u : 0u);
    const uint32_t childbits = (ch.x > -1 ? 1u : 0u) | (ch.y > -1 ? 2u : 0u) | (ch.z > -1 ? 4u : 0u) | (ch.w > -1 ? 8u : 0u);
    const uint32_t leaves = mask & leafbits;
    const uint32_t inner = mask & ~leafbits & childbits;
    int next = -1;
    if (inner) {
        int pay[4] = {(inner & 1u) ? ch.x : -1, (inner & 2u) ? ch.y : -1, (inner & 4u) ? ch.z : -1, (inner & 8u) ? ch.w : -1};
        RTB_CSWAP(0, 1)
        RTB_CSWAP(2, 3)
        RTB_CSWAP(0, 2)
        RTB_CSWAP(1, 3)
        if (key[2] > key[3]) {  // the last comparator swaps ids only (mbvh_node.rs:219-237)
            int tp = pay[2];
            pay[2] = pay[3];
            pay[3] = tp;
        }
        next = pay[0] >= 0 ? pay[0] : (pay[1] >= 0 ? pay[1] : (pay[2] >= 0 ? pay[2] : pay[3]));
        const int skip = pay[0] >= 0 ? 0 : (pay[1] >= 0 ? 1 : (pay[2] >= 0 ? 2 : 3));
        if (st.sp + 3 <= kSmemStack) {
            int* b = st.base + st.sp * st.stride;
            int c = 0;
            if (skip < 3 && pay[3] >= 0) { b[c * st.stride] = pay[3]; c++; }
            if (skip < 2 && pay[2] >= 0) { b[c * st.stride] = pay[2]; c++; }
            if (skip < 1 && pay[1] >= 0) { b[c * st.stride] = pay[1]; c++; }
            st.sp += c;
        } else {
            if (skip < 3 && pay[3] >= 0) st.push(pay[3]);
            if (skip < 2 && pay[2] >= 0) st.push(pay[2]);
            if (skip < 1 && pay[1] >= 0) st.push(pay[1]);
        }
    }
    const int node = L.cur;
    L.cur = next;
    if (leaves) {  // usually one slot: it becomes the lane's current range (the lane has none: it is in the N phase)
        const int s0 = __ffs(leaves) - 1;
        const int c0 = sel4(cn, s0), f0 = sel4(ch, s0);
        L.tri_pos = f0;
        L.tri_end = f0 + c0;
        L.pmask = leaves & (leaves - 1);  // further hit leaf slots of the same node: fetched when this range is used up
        L.pnode = node;
    }
}
// N phase, Bvh: one popped node (the root is popped without a box test, iter_indices.rs:32-46).
__device__ __forceinline__ void bvh_node_phase(const DeviceTree& tree, const RayRegs& r, Stack& st, Lane& L) {
    const float4* __restrict__ nodes = tree.nodes;
    const F8 nd = ld256(nodes + (size_t)L.cur * 2);
    const int count = __float_as_int(nd.lo.w), left_first = __float_as_int(nd.hi.w);
    int next = -1;
    if (count > -1) {
        L.tri_pos = left_first;
        L.tri_end = left_first + count;
    } else if (left_first > -1) {
        const float4* c = nodes + (size_t)left_first * 2;
        const F8 lc = ld256(c), rc = ld256(c + 2);
        float kl = 0.f, kr = 0.f;
        const bool hl = aabb_single(lc.lo, lc.hi, r, kl);
        const bool hr = aabb_single(rc.lo, rc.hi, r, kr);
        if (hl && hr) {  // BvhNode::sort_nodes (bvh_node.rs:150-177)
            if (kl < kr) {
                st.push(left_first);
                next = left_first + 1;
            } else {
                st.push(left_first + 1);
                next = left_first;
            }
        } else if (hl) {
            next = left_first;
        } else if (hr) {
            next = left_first + 1;
        }
    }
    L.cur = next;
}

// ================================================================================================
// kernels
// ================================================================================================
#ifndef RTB_STREAM
#define RTB_STREAM 1  // rays are read once and records written once: streaming (evict-first) accesses keep the tree in L2
#endif
__device__ __forceinline__ void load_ray(const RTRay* __restrict__ rays, size_t i, RayRegs& r) {
#if RTB_STREAM
    const float4 a = __ldcs(reinterpret_cast<const float4*>(rays) + i * 2), b = __ldcs(reinterpret_cast<const float4*>(rays) + i * 2 + 1);
#else
    const F8 ab = ld256(reinterpret_cast<const float4*>(rays) + i * 2);
    const float4 a = ab.lo, b = ab.hi;
#endif
    r.ox = a.x; r.oy = a.y; r.oz = a.z; r.t_min = a.w;
    r.dx = b.x; r.dy = b.y; r.dz = b.z; r.t = b.w;
    finish_ray_setup(r);
}
// Gated launches read rays the copy engine has just written: L2-only loads (each ray is read once anyway).
__device__ __forceinline__ void load_ray_cg(const RTRay* __restrict__ rays, size_t i, RayRegs& r) {
    const float4* p = reinterpret_cast<const float4*>(rays) + i * 2;
    float4 a, b;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p) : "memory");
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 1) : "memory");
    r.ox = a.x; r.oy = a.y; r.oz = a.z; r.t_min = a.w;
    r.dx = b.x; r.dy = b.y; r.dz = b.z; r.t = b.w;
    finish_ray_setup(r);
}
// Split input: origins (the kernel's ray pointer) and directions as tightly packed float3 arrays, common t_min / t.
__device__ __forceinline__ void load_ray_od(const float* __restrict__ origins, const PeerDests& pd, size_t i, RayRegs& r) {
    const float* o = origins + i * 3;
    const float* d = pd.directions + i * 3;
    r.ox = __ldcg(o); r.oy = __ldcg(o + 1); r.oz = __ldcg(o + 2);
    r.dx = __ldcg(d); r.dy = __ldcg(d + 1); r.dz = __ldcg(d + 2);
    r.t_min = pd.t_min;
    r.t = pd.t_max;
    finish_ray_setup(r);
}
template <class T>
__device__ __forceinline__ void st_stream(T* p, T v) {
#if RTB_STREAM
    __stcs(p, v);
#else
    *p = v;
#endif
}
// Result store.  With peer destinations (multi-GPU gather fused into the kernel) the record is also written,
// the moment its ray finishes, into the gather buffer of every peer GPU (P2P stores over NVLink / NVSwitch to
// cudaIpc-mapped memory): the transfer overlaps the traversal ray by ray and no collective follows the kernel.
template <bool ANY>
__device__ __forceinline__ void store_result(const RayRegs& r, size_t i, RTHit* __restrict__ hits,
                                             uint8_t* __restrict__ occluded, const PeerDests& pd) {
    if (ANY) {
        const uint8_t v = r.prim != kNoHit ? 1 : 0;
        if (occluded) st_stream(occluded + i, v);
#pragma unroll
        for (int k = 0; k < 8; k++)  // unrolled: pd stays in the constant bank (no local copy for dynamic indexing)
            if (k < pd.count) static_cast<uint8_t*>(pd.p[k])[pd.offset + i] = v;
    } else {
        const float2 v = make_float2(r.t, __uint_as_float(r.prim));
        if (hits) st_stream(reinterpret_cast<float2*>(hits) + i, v);
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (k < pd.count) static_cast<float2*>(pd.p[k])[pd.offset + i] = v;
    }
}
template <bool ANY>
__device__ __forceinline__ void store_result(const RayRegs& r, size_t i, RTHit* __restrict__ hits,
                                             uint8_t* __restrict__ occluded) {
    if (ANY)
        st_stream(occluded + i, (uint8_t)(r.prim != kNoHit ? 1 : 0));
    else
        st_stream(reinterpret_cast<float2*>(hits) + i, make_float2(r.t, __uint_as_float(r.prim)));
}

// Position in reservation order -> ray index: 8x8 pixel tiles inside the tiled part of the batch (see the refill code).
__device__ __forceinline__ size_t tiled_index(const PeerDests& pd, unsigned long long pos) {
    if (kRayChunk == 64 && pd.tile_w != 0 && pos < pd.tile_n) {
        const unsigned long long band = 8ull * pd.tile_w, b = pos / band, rr = pos % band;
        const unsigned tx = (unsigned)(rr >> 6), j = (unsigned)(rr & 63u);
        return (size_t)(b * band + (unsigned long long)(j >> 3) * pd.tile_w + tx * 8u + (j & 7u));
    }
    return (size_t)pos;
}
// Fused multi-GPU gather, chunk-wise.  Every chunk of kRayChunk rays is reserved — and therefore traced — by ONE warp, so
// the bookkeeping is warp-local: a four-entry table per warp in shared memory (chunk id, rays still running); finished rays
// store their record into the LOCAL result buffer, and when the last ray of a chunk has finished the warp copies the
// chunk's records to every destination — lane l moves records 2l, 2l+1 (one 16-byte load from L2, one 16-byte store per
// destination; any hit: two bytes), 64 contiguous bytes per tile row — instead of one 8-byte P2P store per ray and
// destination scattered over the refills.  No global atomics, no device-scope fence: writer and reader are the same warp.
// A chunk that finds the table full (a warp with more than four chunks in flight: very long rays) falls back to per-ray stores.
constexpr int kPushSlots = 4;
template <bool ANY>
__device__ __forceinline__ void push_chunk(unsigned c, size_t n, const RTHit* __restrict__ hits,
                                           const uint8_t* __restrict__ occluded, const PeerDests& pd, unsigned lane) {
    __threadfence_block();  // the records were stored by lanes of this warp
    const unsigned long long lo = (unsigned long long)c * kRayChunk;
    const unsigned size = (unsigned)((n - lo) < (unsigned long long)kRayChunk ? (n - lo) : (unsigned long long)kRayChunk);
    const unsigned j = 2u * lane;
    if (j >= size) return;
    const size_t idx = tiled_index(pd, lo + j);  // even, and idx + 1 is the next position's index (same tile row)
    const bool two = j + 1 < size;
    if (ANY) {
        const uint8_t* src = occluded + idx;
        unsigned short v2 = 0;
        if (two) {
            asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v2) : "l"(src) : "memory");
        } else {
            asm volatile("ld.global.cg.u8 %0, [%1];" : "=h"(v2) : "l"(src) : "memory");
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (k < pd.count) {
                uint8_t* d = static_cast<uint8_t*>(pd.p[k]) + pd.offset + idx;
                if (two)
                    *reinterpret_cast<unsigned short*>(d) = v2;
                else
                    *d = (uint8_t)v2;
            }
        }
    } else {
        const float2* src = reinterpret_cast<const float2*>(hits) + idx;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (two) {
            asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src) : "memory");
        } else {
            asm volatile("ld.global.cg.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(src) : "memory");
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (k < pd.count) {
                float2* d = static_cast<float2*>(pd.p[k]) + pd.offset + idx;
                if (two)
                    *reinterpret_cast<float4*>(d) = v;
                else
                    *d = make_float2(v.x, v.y);
            }
        }
    }
}
// A warp registers the chunk it has just reserved: false when its table is full (that chunk's rays then scatter one by one).
__device__ __forceinline__ void push_register(unsigned* tab, unsigned c, unsigned size, unsigned lane) {
    __syncwarp();
    int e = -1;
#pragma unroll
    for (int k = 0; k < kPushSlots; k++)
        if (e < 0 && tab[kPushSlots + k] == 0) e = k;  // warp-uniform (broadcast reads)
    __syncwarp();
    if (e >= 0 && lane == 0) {
        tab[e] = c;
        tab[kPushSlots + e] = size;
    }
    __syncwarp();
}
// Result flush of the finished lanes of a warp (all lanes call it).  PUSH: local store, the chunk table counts the rays
// down, completed chunks are copied to the destinations; otherwise the plain (or per-ray scattered) store.
template <bool ANY, bool PUSH>
__device__ __forceinline__ void flush_finished(bool& fin, const RayRegs& r, size_t my, unsigned slot, size_t n,
                                               RTHit* __restrict__ hits, uint8_t* __restrict__ occluded, const PeerDests& pd,
                                               unsigned* tab, unsigned lane) {
    unsigned pending = __ballot_sync(0xFFFFFFFFu, fin);
    if (pending == 0) return;
    if (!PUSH) {
        if (fin) store_result<ANY>(r, my, hits, occluded, pd);
        fin = false;
        return;
    }
    const unsigned cid = slot / kRayChunk;
    if (fin) store_result<ANY>(r, my, hits, occluded);
    while (pending) {
        const unsigned c = __shfl_sync(0xFFFFFFFFu, cid, __ffs(pending) - 1);
        const unsigned grp = __ballot_sync(0xFFFFFFFFu, fin && cid == c);
        pending &= ~grp;
        int e = -1;
#pragma unroll
        for (int k = 0; k < kPushSlots; k++)
            if (tab[k] == c && tab[kPushSlots + k] != 0) e = k;
        __syncwarp();
        if (e < 0) {  // unregistered chunk: one store per ray and destination
            if (fin && cid == c) {
                PeerDests q = pd;
                store_result<ANY>(r, my, nullptr, nullptr, q);
            }
            continue;
        }
        const unsigned left = tab[kPushSlots + e] - __popc(grp);
        __syncwarp();
        if (lane == 0) tab[kPushSlots + e] = left;
        __syncwarp();
        if (left == 0) push_chunk<ANY>(c, n, hits, occluded, pd, lane);
    }
    fin = false;
}

// Static assignment: thread i traces ray i.  Kept for A/B runs (RTBVH_TRACE_MODE=static) and for
// small batches; a warp lives as long as its slowest ray.
template <int TREE, bool ANY>
__global__ void __launch_bounds__(kBlock) trace_single_kernel(const DeviceTree tree, const RTRay* __restrict__ rays,
                                                              size_t n, RTHit* __restrict__ hits,
                                                              uint8_t* __restrict__ occluded,
                                                              const uint32_t* __restrict__ perm,
                                                              uint32_t* __restrict__ overflow) {
    __shared__ int smem[kSmemStack * kBlock];
    int deep[kSpillStack];
    const size_t slot = (size_t)blockIdx.x * kBlock + threadIdx.x;
    if (slot >= n) return;
    const size_t i = perm ? (size_t)perm[slot] : slot;  // sorted launch order -> original ray index
    RayRegs r;
    load_ray(rays, i, r);
    Stack st{smem + threadIdx.x, deep, 0, overflow, kBlock};
    if (tree.node_count != 0 && !r.nan) {
        int cur = 0;
        while (!single_step<TREE, ANY>(tree, r, st, cur)) {
        }
    }
    store_result<ANY>(r, i, hits, occluded);
}

// Persistent warps with dynamic ray refill (the "persistent-thread, warp-cooperative" kernel of the
// north star): the grid is sized to the machine (SMs x resident blocks), every warp reserves rays in
// chunks of kRayChunk from a global counter and, whenever kRefillIdle or more lanes have finished,
// hands the idle lanes the next rays (ballot + prefix popcount).  Lanes therefore sit at different
// depths of different rays, but all execute the same node-visit step, which keeps the SIMD lanes
// busy when ray lengths differ (one missing ray no longer pins 31 idle lanes).
template <int TREE, bool ANY, bool PHASED, bool PUSH>
__global__ void __launch_bounds__(kPBlock, RTB_MINBLOCKS) trace_single_persistent_kernel(const DeviceTree tree,
                                                                         const RTRay* __restrict__ rays, size_t n,
                                                                         RTHit* __restrict__ hits,
                                                                         uint8_t* __restrict__ occluded,
                                                                         const uint32_t* __restrict__ perm,
                                                                         unsigned long long* __restrict__ counter,
                                                                         uint32_t* __restrict__ overflow,
                                                                         const PeerDests pd) {
    // dynamic shared memory: the traversal stacks ([entry][thread]) and, behind them, the staged top of the tree
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    int* smem = reinterpret_cast<int*>(dyn_smem);
    __shared__ unsigned push_tab[PUSH ? (kPBlock / 32) * 2 * kPushSlots : 1];  // per warp: chunk ids, rays still running
    unsigned* tab = push_tab + (PUSH ? (threadIdx.x >> 5) * 2 * kPushSlots : 0);
    if (PUSH) {
        if ((threadIdx.x & 31u) < 2 * kPushSlots) tab[threadIdx.x & 31u] = 0;
        __syncwarp();
    }
    const float4* top_s = nullptr;
    if constexpr (kTopK > 0 && PHASED && TREE == RT_TREE_MBVH) {
        float4* t = reinterpret_cast<float4*>(dyn_smem + (size_t)kSmemStack * kPBlock * sizeof(int));
        const uint32_t cnt = min(tree.top_count, (uint32_t)kTopK);
        for (uint32_t i = threadIdx.x; i < cnt * 8; i += kPBlock) t[(i >> 3) * kTopRow + (i & 7)] = __ldg(tree.top + i);
        __syncthreads();
        top_s = t;
    }
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    int deep[kSpillStack];
    Stack st{smem + threadIdx.x, deep, 0, overflow, kPBlock};
    RayRegs r;
    int cur = 0;
    Lane L{-1, 0, 0, 0, 0u};
    size_t my = 0;
    unsigned slot = 0;  // PUSH: the ray's position in reservation order (its chunk = slot / kRayChunk)
    bool active = false;
    bool fin = false;  // the lane holds the record of a finished ray that is not stored yet (stored at the next refill, by
                       // all finished lanes of the warp in the same instructions, instead of lane by lane as rays end)
    unsigned long long res_next = 0, res_end = 0;  // this warp's reserved index range (warp-uniform)
    bool exhausted = false;                        // the global counter ran past n (warp-uniform)
    for (;;) {
        unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
        if (idle == 0xFFFFFFFFu || (!exhausted && __popc(idle) >= kRefillIdle)) {
            while (idle != 0 && !exhausted) {
                if (res_next >= res_end) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(counter, (unsigned long long)kRayChunk);
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    if (base >= n) {
                        exhausted = true;
                        break;
                    }
                    res_next = base;
                    res_end = base + kRayChunk < n ? base + kRayChunk : n;
                    if (PUSH) push_register(tab, (unsigned)(base / kRayChunk), (unsigned)(res_end - res_next), lane);
                    if (pd.ready) {  // host-buffer pipeline: wait until the copy engine has delivered this range
                        if (lane == 0) {
                            unsigned long long have;
                            unsigned backoff = 250;  // thousands of warps may wait on this one L2 line: poll politely
                            for (;;) {
                                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(have) : "l"(pd.ready) : "memory");
                                if (have >= res_end) break;
                                __nanosleep(backoff);
                                if (backoff < 4000) backoff <<= 1;
                            }
                        }
                        __syncwarp();
                    }
                }
                const unsigned long long avail = res_end - res_next;
                const unsigned want = __popc(idle);
                const unsigned take = avail < want ? (unsigned)avail : want;
                const unsigned rank = __popc(idle & lt_mask);
                flush_finished<ANY, PUSH>(fin, r, my, slot, n, hits, occluded, pd, tab, lane);
                if (!active && rank < take) {
                    // chunk of 64 consecutive positions -> one 8x8 pixel tile of the same 8-row band (neighbouring lanes then
                    // walk neighbouring parts of the tree: fewer distinct sectors per load instruction)
                    if (PUSH) slot = (unsigned)(res_next + rank);
                    my = tiled_index(pd, res_next + rank);
                    if (perm) my = (size_t)perm[my];
                    if (pd.directions)
                        load_ray_od(reinterpret_cast<const float*>(rays), pd, my, r);
                    else if (pd.ready)
                        load_ray_cg(rays, my, r);
                    else
                        load_ray(rays, my, r);
                    st.reset();
                    cur = 0;
                    L = Lane{(kTopK > 0 && PHASED && TREE == RT_TREE_MBVH && tree.top_count != 0) ? kTopFlag : 0, 0, 0, 0, 0u};
                    if (tree.node_count != 0 && !r.nan)
                        active = true;
                    else
                        fin = true;
                }
                res_next += take;
                idle = __ballot_sync(0xFFFFFFFFu, !active);
            }
            if (idle == 0xFFFFFFFFu) break;  // nothing left to trace for this warp
        }
        if constexpr (!PHASED) {
            if (active) {
                if (single_step<TREE, ANY>(tree, r, st, cur)) {
                    fin = true;
                    active = false;
                }
            }
        } else {
            // phase vote: T = lanes holding a triangle range, N = the other active lanes
            const unsigned act_m = ~idle;  // idle was re-balloted by the refill; otherwise it is this iteration's ballot
            const bool in_t = active && L.tri_pos < L.tri_end;
            const unsigned t_m = __ballot_sync(0xFFFFFFFFu, in_t);
            const int n_t = __popc(t_m), n_n = __popc(act_m & ~t_m);
            const bool in_n = active && !in_t;
            bool done = false;
            if (n_t * RTB_TRI_NUM >= n_n * RTB_TRI_DEN && n_t > 0) {
                if (in_t) {
#if RTB_TRI_FLAT
                    const bool hit = tri_candidate_flat(tree.tris, L.tri_pos, r);
#else
                    const bool hit = tri_candidate<false>(tree.tris, L.tri_pos, r);
#endif
                    L.tri_pos++;
                    if (ANY && hit)
                        done = true;
                    else if (L.tri_pos >= L.tri_end)
                        done = lane_advance(tree, top_s, st, L);
                }
            } else {
                // slabs with the SSE operand rule for the whole warp as soon as one visiting lane needs it (a zero /
                // non-finite component): for every other ray both flavours give the same predicates and keys
                const bool any_exact = TREE == RT_TREE_MBVH && __any_sync(0xFFFFFFFFu, in_n && r.exact) != 0;
                if (in_n) {
                    if (TREE == RT_TREE_MBVH) {
                        if (any_exact)
                            mbvh_node_phase<true>(tree, top_s, r, st, L);
                        else
                            mbvh_node_phase<false>(tree, top_s, r, st, L);
                    } else {
                        bvh_node_phase(tree, r, st, L);
                    }
                    if (L.tri_pos >= L.tri_end) done = lane_advance(tree, top_s, st, L);
                }
            }
            if (done) {
                fin = true;
                active = false;
            }
        }
    }
    flush_finished<ANY, PUSH>(fin, r, my, slot, n, hits, occluded, pd, tab, lane);
}

// ---- lane-cooperative node fetch -------------------------------------------------------------------
// When every lane fetches its own 128-byte node, the L1 data pipe moves 16 bytes per wavefront (one
// wavefront per lane and 16-byte chunk: 256 per warp and visit) and becomes the limiter (ncu:
// l1tex__data_pipe_lsu_wavefronts ~70 % with long-scoreboard stalls).  Here the 8 lanes of a group
// fetch ONE node together — lane j copies chunk j with cp.async (LDGSTS), a full 128-byte line per
// wavefront — into a per-warp shared-memory tile, and every lane then reads its own node back with
// conflict-free LDS.128 (144-byte row stride).  ~3x fewer data-pipe wavefronts per visit, and the copy
// of the NEXT node is issued before the triangles of the current node are tested.
constexpr int kNodeRowWords = 36;  // 128-byte node + 16 bytes of padding: LDS.128 of 8 lanes hits 32 distinct banks

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Collective over the warp: lane L wants node `want` (or -1).  Group g = L / 8 serves its 8 owners in turn.
__device__ __forceinline__ void coop_fetch(const float4* __restrict__ nodes, int want, float* warp_tile, unsigned lane) {
    const unsigned sub = lane & 7u, base = lane & 24u;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int c = __shfl_sync(0xFFFFFFFFu, want, (int)(base + i));
        if (c >= 0) cp_async16(warp_tile + (base + i) * kNodeRowWords + sub * 4, nodes + (size_t)c * 8 + sub);
    }
}

template <bool ANY>
__global__ void __launch_bounds__(kBlock, 6) trace_mbvh_coop_kernel(const DeviceTree tree,
                                                                                const RTRay* __restrict__ rays, size_t n,
                                                                                RTHit* __restrict__ hits,
                                                                                uint8_t* __restrict__ occluded,
                                                                                const uint32_t* __restrict__ perm,
                                                                                unsigned long long* __restrict__ counter,
                                                                                uint32_t* __restrict__ overflow) {
    __shared__ int smem[kSmemStack * kBlock];
    __shared__ __align__(16) float tiles[(kBlock / 32) * 32 * kNodeRowWords];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    float* warp_tile = tiles + (threadIdx.x >> 5) * 32 * kNodeRowWords;
    const float4* my_row = reinterpret_cast<const float4*>(warp_tile + lane * kNodeRowWords);
    int deep[kSpillStack];
    Stack st{smem + threadIdx.x, deep, 0, overflow, kBlock};
    RayRegs r;
    int cur = 0, fetched = -1;  // fetched: the node whose copy is in (or on its way to) this lane's row
    size_t my = 0;
    bool active = false;
    unsigned long long res_next = 0, res_end = 0;
    bool exhausted = false;
    for (;;) {
        unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
        if (idle == 0xFFFFFFFFu || (!exhausted && __popc(idle) >= kRefillIdle)) {
            while (idle != 0 && !exhausted) {
                if (res_next >= res_end) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(counter, (unsigned long long)kRayChunk);
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    if (base >= n) {
                        exhausted = true;
                        break;
                    }
                    res_next = base;
                    res_end = base + kRayChunk < n ? base + kRayChunk : n;
                }
                const unsigned long long avail = res_end - res_next;
                const unsigned want = __popc(idle);
                const unsigned take = avail < want ? (unsigned)avail : want;
                const unsigned rank = __popc(idle & lt_mask);
                if (!active && rank < take) {
                    my = (size_t)(res_next + rank);
                    if (perm) my = (size_t)perm[my];
                    load_ray(rays, my, r);
                    st.reset();
                    cur = 0;
                    fetched = -1;
                    if (tree.node_count != 0 && !r.nan)
                        active = true;
                    else
                        store_result<ANY>(r, my, hits, occluded);
                }
                res_next += take;
                idle = __ballot_sync(0xFFFFFFFFu, !active);
            }
            if (idle == 0xFFFFFFFFu) break;
        }
        // rows that do not hold the node their lane is about to visit (fresh rays): fetch now
        const int want = (active && fetched != cur) ? cur : -1;
        if (__any_sync(0xFFFFFFFFu, want >= 0)) {
            if (ANY) {  // a ray that ended on a hit may have left a prefetch in flight into a row that is re-targeted now
                cp_async_wait_all();
                __syncwarp();
            }
            coop_fetch(tree.nodes, want, warp_tile, lane);
        }
        cp_async_wait_all();
        __syncwarp();
        MNode nd;
        if (active) {
            nd.mnx = my_row[0]; nd.mxx = my_row[1]; nd.mny = my_row[2]; nd.mxy = my_row[3];
            nd.mnz = my_row[4]; nd.mxz = my_row[5];
            nd.ch = as_int4(my_row[6]);
            nd.cn = as_int4(my_row[7]);
        }
        __syncwarp();  // every lane has its node in registers: the tile may be overwritten
        int next = -1;
        uint32_t leaves = 0;
        int4 ch = make_int4(-1, -1, -1, -1), cn = ch;
        if (active) {
            next = mbvh_visit_push(nd, r, st, leaves);
            ch = nd.ch;
            cn = nd.cn;
        }
        fetched = (active && next >= 0) ? next : -1;
        coop_fetch(tree.nodes, fetched, warp_tile, lane);  // in flight while the triangles below are tested
        if (active) {
            const bool done = mbvh_visit_leaves<ANY>(ch, cn, leaves, tree, r);
            if (done || next < 0) {
                store_result<ANY>(r, my, hits, occluded);
                active = false;
            } else {
                cur = next;
            }
        }
    }
    cp_async_wait_all();
}

__device__ __forceinline__ void load_packet_lane(const RTRayPacket4* __restrict__ packets, size_t p, int ql, float t_min,
                                                 RayRegs& r) {
    const float* pk = reinterpret_cast<const float*>(packets + p);
    r.ox = __ldg(pk + 0 + ql); r.oy = __ldg(pk + 4 + ql); r.oz = __ldg(pk + 8 + ql);
    r.dx = __ldg(pk + 12 + ql); r.dy = __ldg(pk + 16 + ql); r.dz = __ldg(pk + 20 + ql);
    r.t = __ldg(pk + 24 + ql);
    r.t_min = t_min;
    finish_ray_setup(r);
}
template <bool ANY>
__device__ __forceinline__ void store_packet_lane(const RayRegs& r, bool retired, size_t p, int ql,
                                                  RTHitPacket4* __restrict__ hits, uint8_t* __restrict__ occluded) {
    if (ANY) {
        occluded[p * 4 + ql] = retired ? 1 : 0;
    } else {
        hits[p].t[ql] = r.t;
        hits[p].prim[ql] = r.prim;
    }
}

// Static assignment: quad q of the grid traces packet q (A/B: RTBVH_TRACE_MODE=static).
template <int TREE, bool ANY>
__global__ void __launch_bounds__(kBlock) trace_packet_kernel(const DeviceTree tree,
                                                              const RTRayPacket4* __restrict__ packets, size_t n_packets,
                                                              float t_min, RTHitPacket4* __restrict__ hits,
                                                              uint8_t* __restrict__ occluded,
                                                              uint32_t* __restrict__ overflow) {
    __shared__ int smem[kSmemStack * kBlock];
    const size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x;  // ray index; packet = i / 4
    const size_t p = i >> 2;
    const int ql = (int)(i & 3);
    if (p >= n_packets) return;  // whole quads leave together
    RayRegs r;
    load_packet_lane(packets, p, ql, t_min, r);
    const uint32_t qm = quad_mask();
    int deep[kSpillStack];
    Stack st{smem + threadIdx.x, deep, 0, overflow, kBlock};
    bool retired = false;
    // BvhPacketIndexIterator rejects the packet when ANY lane has a NaN (iter_indices.rs:129-144);
    // MbvhPacketIndexIterator has no such check (the NaN lane just never passes a comparison).
    const bool reject = (TREE == RT_TREE_BVH) && (__any_sync(qm, r.nan) != 0);
    if (tree.node_count != 0 && !reject) {
        int cur = 0;
        while (!packet_step<TREE, ANY>(tree, r, st, cur, retired, qm)) {
        }
    }
    store_packet_lane<ANY>(r, retired, p, ql, hits, occluded);
}

// Persistent warps for packets: 8 quads per warp, idle quads are refilled with the next packets.
template <int TREE, bool ANY>
__global__ void __launch_bounds__(kBlock) trace_packet_persistent_kernel(const DeviceTree tree,
                                                                         const RTRayPacket4* __restrict__ packets,
                                                                         size_t n_packets, float t_min,
                                                                         RTHitPacket4* __restrict__ hits,
                                                                         uint8_t* __restrict__ occluded,
                                                                         unsigned long long* __restrict__ counter,
                                                                         uint32_t* __restrict__ overflow) {
    __shared__ int smem[kSmemStack * kBlock];
    const unsigned lane = threadIdx.x & 31u;
    const int ql = (int)(lane & 3u);
    const unsigned below = (1u << (lane & 28u)) - 1u;  // lanes of lower quads
    const uint32_t qm = quad_mask();
    int deep[kSpillStack];
    Stack st{smem + threadIdx.x, deep, 0, overflow, kBlock};
    RayRegs r;
    int cur = 0;
    size_t my = 0;
    bool active = false, retired = false;  // active is quad-uniform
    unsigned long long res_next = 0, res_end = 0;
    bool exhausted = false;
    constexpr unsigned kPacketChunk = kRayChunk / 4;
    for (;;) {
        unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
        if (idle == 0xFFFFFFFFu || (!exhausted && __popc(idle) >= kRefillIdle)) {
            while (idle != 0 && !exhausted) {
                if (res_next >= res_end) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(counter, (unsigned long long)kPacketChunk);
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    if (base >= n_packets) {
                        exhausted = true;
                        break;
                    }
                    res_next = base;
                    res_end = base + kPacketChunk < n_packets ? base + kPacketChunk : n_packets;
                }
                const unsigned long long avail = res_end - res_next;
                const unsigned want = (unsigned)__popc(idle) >> 2;  // idle quads
                const unsigned take = avail < want ? (unsigned)avail : want;
                const unsigned rank = (unsigned)__popc(idle & below) >> 2;  // rank of this quad among the idle quads
                if (!active && rank < take) {
                    my = (size_t)(res_next + rank);
                    load_packet_lane(packets, my, ql, t_min, r);
                    st.reset();
                    cur = 0;
                    retired = false;
                    const bool reject = (TREE == RT_TREE_BVH) && (__any_sync(qm, r.nan) != 0);
                    if (tree.node_count != 0 && !reject)
                        active = true;
                    else
                        store_packet_lane<ANY>(r, retired, my, ql, hits, occluded);
                }
                res_next += take;
                idle = __ballot_sync(0xFFFFFFFFu, !active);
            }
            if (idle == 0xFFFFFFFFu) break;
        }
        if (active) {
            if (packet_step<TREE, ANY>(tree, r, st, cur, retired, qm)) {
                store_packet_lane<ANY>(r, retired, my, ql, hits, occluded);
                active = false;
            }
        }
    }
}

// ================================================================================================
// Mbvh packets, ONE LANE PER PACKET (the default Mbvh packet kernel).
// With four lanes per packet every lane repeats the packet's control flow (stack, slot loop, leaf
// bookkeeping) and the packet pays the union of its four rays' node visits in every lane.  Here one
// lane owns the whole RayPacket4: the 16 ray x slot slab tests of a visit are 16 independent
// dependency chains in ONE thread (instruction-level parallelism instead of occupancy), the control
// flow is paid once per packet, and a node fetch (4 x LDG.256) is shared by four rays — a quarter of
// the L1 data-pipe wavefronts per ray of the single-ray kernel.  Persistent warps, refill and the
// warp-wide node / triangle phases are those of the phased single-ray kernel.
// Semantics are MbvhPacketIndexIterator's (iter_indices.rs:370-414) + intersect4 (mbvh_node.rs:243-295,
// spatial_sah.rs:165-244): a slot is entered when ANY ray passes `t_max > t_min && t_min < packet.t[i]`
// with the packet.t of node entry, no ordering (inner slots pushed 3, 2, 1, 0), leaf slots yield every
// primitive to the four-ray triangle test (eps 1e-6, t >= t_min).
// ================================================================================================
#ifndef RTB_LBLOCK
#define RTB_LBLOCK 128
#endif
#ifndef RTB_LMINBLOCKS
#define RTB_LMINBLOCKS 5  // 96 registers: measured 4 blocks (128 regs) 1 781-1 806, 5 blocks 2 008, 6 blocks (80 regs, spills) 1 795, 7 blocks 1 454 Mrays/s
#endif
constexpr int kLBlock = RTB_LBLOCK;
#ifndef RTB_LTRI_NUM
#define RTB_LTRI_NUM RTB_TRI_NUM  // lane-per-packet kernels: triangle phase when lanes_T * RTB_LTRI_NUM >= lanes_N
#endif
#ifndef RTB_LREFILL
#define RTB_LREFILL RTB_REFILL    // ... and refill once this many lanes are idle
#endif
#ifndef RTB_LKEEPDIR
#define RTB_LKEEPDIR 0  // 1: keep the four directions in registers instead of re-reading them in the triangle phase
#endif
struct Packet {
    float ox[4], oy[4], oz[4];
#if RTB_LKEEPDIR
    float dx[4], dy[4], dz[4];
#endif
    float ix[4], iy[4], iz[4];  // node visits need the inverse directions only; the triangle phase re-reads the directions
    float t[4];                 // from the packet record (3 x LDG.128 per tested triangle: 12 registers less per lane)
    uint32_t prim[4];
};
__device__ __forceinline__ void load_packet(const RTRayPacket4* __restrict__ packets, size_t p, Packet& k, bool& exact) {
    const float4* q = reinterpret_cast<const float4*>(packets + p);  // 112-byte records: 16-byte aligned
    const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3), e = __ldg(q + 4), f = __ldg(q + 5),
                 g = __ldg(q + 6);
    k.ox[0] = a.x; k.ox[1] = a.y; k.ox[2] = a.z; k.ox[3] = a.w;
    k.oy[0] = b.x; k.oy[1] = b.y; k.oy[2] = b.z; k.oy[3] = b.w;
    k.oz[0] = c.x; k.oz[1] = c.y; k.oz[2] = c.z; k.oz[3] = c.w;
    const float dx[4] = {d.x, d.y, d.z, d.w}, dy[4] = {e.x, e.y, e.z, e.w}, dz[4] = {f.x, f.y, f.z, f.w};
    k.t[0] = g.x; k.t[1] = g.y; k.t[2] = g.z; k.t[3] = g.w;
    bool fin = true;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        k.ix[i] = fdiv(1.0f, dx[i]);  // RayPacket4::new: inv_direction = 1 / direction (src/ray.rs:64-148)
        k.iy[i] = fdiv(1.0f, dy[i]);
        k.iz[i] = fdiv(1.0f, dz[i]);
#if RTB_LKEEPDIR
        k.dx[i] = dx[i]; k.dy[i] = dy[i]; k.dz[i] = dz[i];
#endif
        k.prim[i] = kNoHit;
        fin = fin && isfinite(k.ox[i]) && isfinite(k.oy[i]) && isfinite(k.oz[i]) && isfinite(dx[i]) && isfinite(dy[i]) &&
              isfinite(dz[i]) && isfinite(k.ix[i]) && isfinite(k.iy[i]) && isfinite(k.iz[i]);
    }
    exact = !fin;
}
// MbvhNode::intersect4 (mbvh_node.rs:243-281): `result |= ...` over the four rays; returns the 4-bit slot mask
template <bool EXACT>
__device__ __forceinline__ uint32_t mbvh_slabs_packet(const MNode& nd, const Packet& k) {
    const float a_mnx[4] = {nd.mnx.x, nd.mnx.y, nd.mnx.z, nd.mnx.w}, a_mxx[4] = {nd.mxx.x, nd.mxx.y, nd.mxx.z, nd.mxx.w};
    const float a_mny[4] = {nd.mny.x, nd.mny.y, nd.mny.z, nd.mny.w}, a_mxy[4] = {nd.mxy.x, nd.mxy.y, nd.mxy.z, nd.mxy.w};
    const float a_mnz[4] = {nd.mnz.x, nd.mnz.y, nd.mnz.z, nd.mnz.w}, a_mxz[4] = {nd.mxz.x, nd.mxz.y, nd.mxz.z, nd.mxz.w};
    uint32_t mask = 0;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        bool any = false;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float t1 = fmul(fsub(a_mnx[s], k.ox[i]), k.ix[i]), t2 = fmul(fsub(a_mxx[s], k.ox[i]), k.ix[i]);
            float tmin = vmin<EXACT>(t1, t2), tmax = vmax<EXACT>(t1, t2);
            t1 = fmul(fsub(a_mny[s], k.oy[i]), k.iy[i]);
            t2 = fmul(fsub(a_mxy[s], k.oy[i]), k.iy[i]);
            tmin = vmax<EXACT>(tmin, vmin<EXACT>(t1, t2));
            tmax = vmin<EXACT>(tmax, vmax<EXACT>(t1, t2));
            t1 = fmul(fsub(a_mnz[s], k.oz[i]), k.iz[i]);
            t2 = fmul(fsub(a_mxz[s], k.oz[i]), k.iz[i]);
            tmin = vmax<EXACT>(tmin, vmin<EXACT>(t1, t2));
            tmax = vmin<EXACT>(tmax, vmax<EXACT>(t1, t2));
            any = any || (tmax > tmin && tmin < k.t[i]);
        }
        if (any) mask |= 1u << s;
    }
    return mask;
}
// SpatialTriangle::intersect4 (spatial_sah.rs:165-244) for the four rays of the lane's packet, without early exits.
// Returns the 4-bit mask of rays that accepted the candidate with t < packet.t[i].
__device__ __forceinline__ uint32_t tri_candidate_packet(const TriRec* __restrict__ tris, int pos, float t_min,
                                                         const RTRayPacket4* __restrict__ rec, Packet& k) {
    const F8 ab = ld256_tri(&tris[pos].a);
    const float4 A = ab.lo, E1 = ab.hi;
    const float4 E2 = __ldg(&tris[pos].c);
    const uint32_t id = __float_as_uint(A.w);
#if RTB_LKEEPDIR
    const float* dx = k.dx; const float* dy = k.dy; const float* dz = k.dz;
#else
    const float4* q = reinterpret_cast<const float4*>(rec);
    const float4 d = __ldg(q + 3), e = __ldg(q + 4), f4 = __ldg(q + 5);
    const float dx[4] = {d.x, d.y, d.z, d.w}, dy[4] = {e.x, e.y, e.z, e.w}, dz[4] = {f4.x, f4.y, f4.z, f4.w};
#endif
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float hx = fsub(fmul(dy[i], E2.z), fmul(E2.y, dz[i]));
        const float hy = fsub(fmul(dz[i], E2.x), fmul(E2.z, dx[i]));
        const float hz = fsub(fmul(dx[i], E2.y), fmul(E2.x, dy[i]));
        const float a = fadd(fadd(fmul(E1.x, hx), fmul(E1.y, hy)), fmul(E1.z, hz));
        const bool p_a = (a <= -1e-6f || a >= 1e-6f);  // spatial_sah.rs:191-196
        const float f = fdiv(1.0f, a);
        const float sx = fsub(k.ox[i], A.x), sy = fsub(k.oy[i], A.y), sz = fsub(k.oz[i], A.z);
        const float u = fmul(f, fadd(fadd(fmul(sx, hx), fmul(sy, hy)), fmul(sz, hz)));
        const bool p_u = (u >= 0.0f && u <= 1.0f);
        const float qx = fsub(fmul(sy, E1.z), fmul(E1.y, sz));
        const float qy = fsub(fmul(sz, E1.x), fmul(E1.z, sx));
        const float qz = fsub(fmul(sx, E1.y), fmul(E1.x, sy));
        const float v = fmul(f, fadd(fadd(fmul(dx[i], qx), fmul(dy[i], qy)), fmul(dz[i], qz)));
        const bool p_v = (v >= 0.0f && fadd(u, v) <= 1.0f);  // spatial_sah.rs:218-222
        const float t = fmul(f, fadd(fadd(fmul(E2.x, qx), fmul(E2.y, qy)), fmul(E2.z, qz)));
        const bool ok = p_a && p_u && p_v && (t >= t_min);
        const bool closer = ok && t < k.t[i];
        const bool tie = ok && k.prim[i] != kNoHit && t == k.t[i] && id < k.prim[i];
        if (closer) k.t[i] = t;
        if (closer || tie) k.prim[i] = id;
        if (closer) acc |= 1u << i;
    }
    return acc;
}

template <bool ANY>
__global__ void __launch_bounds__(kLBlock, RTB_LMINBLOCKS) trace_mbvh_packet_lane_kernel(
    const DeviceTree tree, const RTRayPacket4* __restrict__ packets, size_t n_packets, float t_min,
    RTHitPacket4* __restrict__ hits, uint8_t* __restrict__ occluded, unsigned long long* __restrict__ counter,
    uint32_t* __restrict__ overflow) {
    __shared__ int smem[kSmemStack * kLBlock];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    int deep[kSpillStack];
    Stack st{smem + threadIdx.x, deep, 0, overflow, kLBlock};
    Packet k;
    bool exact = false;
    uint32_t retired = 0;  // any hit: rays of the packet that are done (their t is -1e34, see rtbvh_gpu.h)
    Lane L{-1, 0, 0, 0, 0u};
    size_t my = 0;
    bool active = false, fin = false;
    unsigned long long res_next = 0, res_end = 0;
    bool exhausted = false;
    constexpr unsigned kPacketChunk = 32;  // packets a warp reserves per global atomic: one full refill
    auto store = [&]() {
        if (ANY) {
            const uint32_t v = (retired & 1u) | ((retired & 2u) << 7) | ((retired & 4u) << 14) | ((retired & 8u) << 21);
            reinterpret_cast<uint32_t*>(occluded)[my] = v;
        } else {
            float4* o = reinterpret_cast<float4*>(hits + my);
            o[0] = make_float4(k.t[0], k.t[1], k.t[2], k.t[3]);
            o[1] = make_float4(__uint_as_float(k.prim[0]), __uint_as_float(k.prim[1]), __uint_as_float(k.prim[2]),
                               __uint_as_float(k.prim[3]));
        }
    };
    for (;;) {
        unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
        if (idle == 0xFFFFFFFFu || (!exhausted && __popc(idle) >= RTB_LREFILL)) {
            while (idle != 0 && !exhausted) {
                if (res_next >= res_end) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(counter, (unsigned long long)kPacketChunk);
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    if (base >= n_packets) {
                        exhausted = true;
                        break;
                    }
                    res_next = base;
                    res_end = base + kPacketChunk < n_packets ? base + kPacketChunk : n_packets;
                }
                const unsigned long long avail = res_end - res_next;
                const unsigned want = __popc(idle);
                const unsigned take = avail < want ? (unsigned)avail : want;
                const unsigned rank = __popc(idle & lt_mask);
                if (__any_sync(0xFFFFFFFFu, fin)) {
                    if (fin) store();
                    fin = false;
                }
                if (!active && rank < take) {
                    my = (size_t)(res_next + rank);
                    load_packet(packets, my, k, exact);
                    st.reset();
                    retired = 0;
                    L = Lane{0, 0, 0, 0, 0u};
                    if (tree.node_count != 0)  // MbvhPacketIndexIterator has no NaN check (iter_indices.rs:327-352)
                        active = true;
                    else
                        fin = true;
                }
                res_next += take;
                idle = __ballot_sync(0xFFFFFFFFu, !active);
            }
            if (idle == 0xFFFFFFFFu) break;
        }
        const unsigned act_m = ~idle;
        const bool in_t = active && L.tri_pos < L.tri_end;
        const unsigned t_m = __ballot_sync(0xFFFFFFFFu, in_t);
        const int n_t = __popc(t_m), n_n = __popc(act_m & ~t_m);
        const bool in_n = active && !in_t;
        bool done = false;
        if (n_t * RTB_LTRI_NUM >= n_n && n_t > 0) {
            if (in_t) {
                const uint32_t acc = tri_candidate_packet(tree.tris, L.tri_pos, t_min, packets + my, k);
                L.tri_pos++;
                if (ANY && acc) {
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        if (acc & (1u << i)) k.t[i] = -1e34f;
                    retired |= acc;
                    if (retired == 0xFu) done = true;
                }
                if (!done && L.tri_pos >= L.tri_end) done = lane_advance(tree, nullptr, st, L);
            }
        } else {
            const bool any_exact = __any_sync(0xFFFFFFFFu, in_n && exact) != 0;
            if (in_n) {
                const int node = L.cur;
                const MNode nd = mnode_load_global(tree.nodes, node);
                const uint32_t mask = any_exact ? mbvh_slabs_packet<true>(nd, k) : mbvh_slabs_packet<false>(nd, k);
                const int4 ch = nd.ch, cn = nd.cn;
                const uint32_t leafbits = (cn.x > -1 ? 1u : 0u) | (cn.y > -1 ? 2u : 0u) | (cn.z > -1 ? 4u : 0u) | (cn.w > -1 ? 8u : 0u);
                const uint32_t childbits = (ch.x > -1 ? 1u : 0u) | (ch.y > -1 ? 2u : 0u) | (ch.z > -1 ? 4u : 0u) | (ch.w > -1 ? 8u : 0u);
                const uint32_t leaves = mask & leafbits;
                const uint32_t inner = mask & ~leafbits & childbits;
                // inner slots are pushed 3, 2, 1, 0 (iter_indices.rs:390): the lowest hit slot is popped next and stays in a register
                int next = -1;
                if (inner) {
                    const int lowest = __ffs(inner) - 1;
                    if ((inner & 8u) && lowest != 3) st.push(ch.w);
                    if ((inner & 4u) && lowest != 2) st.push(ch.z);
                    if ((inner & 2u) && lowest != 1) st.push(ch.y);
                    next = sel4(ch, lowest);
                }
                L.cur = next;
                if (leaves) {
                    const int s0 = __ffs(leaves) - 1;
                    const int c0 = sel4(cn, s0), f0 = sel4(ch, s0);
                    L.tri_pos = f0;
                    L.tri_end = f0 + c0;
                    L.pmask = leaves & (leaves - 1);
                    L.pnode = node;
                }
                if (L.tri_pos >= L.tri_end) done = lane_advance(tree, nullptr, st, L);
            }
        }
        if (done) {
            fin = true;
            active = false;
        }
    }
    if (fin) store();
}

// ================================================================================================
// Bvh packets, one lane per packet: BvhPacketIndexIterator (iter_indices.rs:172-209) + Aabb::intersect4
// (aabb.rs:218-244) + BvhNode::sort_nodes4 (bvh_node.rs:180-211) with the structure of the Mbvh kernel above.
// A popped node is a leaf (its primitives go through the four-ray triangle test) or an inner node whose two
// children are tested against the four rays: a child is entered when ANY ray passes
// `t_max > 0 && t_max > t_min && t_min < packet.t[i]`; when both are entered the left child is pushed (visited
// second) iff ANY ray has t_near_left < t_near_right — compared on all four rays, masks or not, like the SSE code.
// The root is popped without a box test; a packet with a NaN lane is rejected as a whole (iter_indices.rs:129-144).
// ================================================================================================
template <bool EXACT>
__device__ __forceinline__ void bvh_children_packet(const F8& lc, const F8& rc, const Packet& k, bool& hl, bool& hr, bool& left_nearer) {
    hl = hr = left_nearer = false;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float key[2];
        bool hit[2];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const float4 lo = c == 0 ? lc.lo : rc.lo, hi = c == 0 ? lc.hi : rc.hi;
            const float t1x = fmul(fsub(lo.x, k.ox[i]), k.ix[i]), t1y = fmul(fsub(lo.y, k.oy[i]), k.iy[i]), t1z = fmul(fsub(lo.z, k.oz[i]), k.iz[i]);
            const float t2x = fmul(fsub(hi.x, k.ox[i]), k.ix[i]), t2y = fmul(fsub(hi.y, k.oy[i]), k.iy[i]), t2z = fmul(fsub(hi.z, k.oz[i]), k.iz[i]);
            const float tmin = vmax<EXACT>(vmin<EXACT>(t1x, t2x), vmax<EXACT>(vmin<EXACT>(t1y, t2y), vmin<EXACT>(t1z, t2z)));
            const float tmax = vmin<EXACT>(vmax<EXACT>(t1x, t2x), vmin<EXACT>(vmax<EXACT>(t1y, t2y), vmax<EXACT>(t1z, t2z)));
            key[c] = tmin;
            hit[c] = tmax > 0.0f && tmax > tmin && tmin < k.t[i];
        }
        hl = hl || hit[0];
        hr = hr || hit[1];
        left_nearer = left_nearer || (key[0] < key[1]);
    }
}

template <bool ANY>
__global__ void __launch_bounds__(kLBlock, RTB_LMINBLOCKS) trace_bvh_packet_lane_kernel(
    const DeviceTree tree, const RTRayPacket4* __restrict__ packets, size_t n_packets, float t_min,
    RTHitPacket4* __restrict__ hits, uint8_t* __restrict__ occluded, unsigned long long* __restrict__ counter,
    uint32_t* __restrict__ overflow) {
    __shared__ int smem[kSmemStack * kLBlock];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    int deep[kSpillStack];
    Stack st{smem + threadIdx.x, deep, 0, overflow, kLBlock};
    Packet k;
    bool exact = false;
    uint32_t retired = 0;
    int cur = -1, tri_pos = 0, tri_end = 0;
    size_t my = 0;
    bool active = false, fin = false;
    unsigned long long res_next = 0, res_end = 0;
    bool exhausted = false;
    constexpr unsigned kPacketChunk = 32;
    auto store = [&]() {
        if (ANY) {
            const uint32_t v = (retired & 1u) | ((retired & 2u) << 7) | ((retired & 4u) << 14) | ((retired & 8u) << 21);
            reinterpret_cast<uint32_t*>(occluded)[my] = v;
        } else {
            float4* o = reinterpret_cast<float4*>(hits + my);
            o[0] = make_float4(k.t[0], k.t[1], k.t[2], k.t[3]);
            o[1] = make_float4(__uint_as_float(k.prim[0]), __uint_as_float(k.prim[1]), __uint_as_float(k.prim[2]),
                               __uint_as_float(k.prim[3]));
        }
    };
    for (;;) {
        unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
        if (idle == 0xFFFFFFFFu || (!exhausted && __popc(idle) >= RTB_LREFILL)) {
            while (idle != 0 && !exhausted) {
                if (res_next >= res_end) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(counter, (unsigned long long)kPacketChunk);
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    if (base >= n_packets) {
                        exhausted = true;
                        break;
                    }
                    res_next = base;
                    res_end = base + kPacketChunk < n_packets ? base + kPacketChunk : n_packets;
                }
                const unsigned long long avail = res_end - res_next;
                const unsigned want = __popc(idle);
                const unsigned take = avail < want ? (unsigned)avail : want;
                const unsigned rank = __popc(idle & lt_mask);
                if (__any_sync(0xFFFFFFFFu, fin)) {
                    if (fin) store();
                    fin = false;
                }
                if (!active && rank < take) {
                    my = (size_t)(res_next + rank);
                    load_packet(packets, my, k, exact);
                    st.reset();
                    retired = 0;
                    cur = 0;
                    tri_pos = tri_end = 0;
                    bool nan = false;  // BvhPacketIndexIterator::new rejects a packet with any NaN origin / direction lane
                    {
                        const float4* q = reinterpret_cast<const float4*>(packets + my);
#pragma unroll
                        for (int j = 0; j < 6; j++) {
                            const float4 v = __ldg(q + j);
                            nan = nan || isnan(v.x) || isnan(v.y) || isnan(v.z) || isnan(v.w);
                        }
                    }
                    if (tree.node_count != 0 && !nan)
                        active = true;
                    else
                        fin = true;
                }
                res_next += take;
                idle = __ballot_sync(0xFFFFFFFFu, !active);
            }
            if (idle == 0xFFFFFFFFu) break;
        }
        const unsigned act_m = ~idle;
        const bool in_t = active && tri_pos < tri_end;
        const unsigned t_m = __ballot_sync(0xFFFFFFFFu, in_t);
        const int n_t = __popc(t_m), n_n = __popc(act_m & ~t_m);
        const bool in_n = active && !in_t;
        bool done = false;
        if (n_t * RTB_LTRI_NUM >= n_n && n_t > 0) {
            if (in_t) {
                const uint32_t acc = tri_candidate_packet(tree.tris, tri_pos, t_min, packets + my, k);
                tri_pos++;
                if (ANY && acc) {
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        if (acc & (1u << i)) k.t[i] = -1e34f;
                    retired |= acc;
                    if (retired == 0xFu) done = true;
                }
            }
        } else {
            const bool any_exact = __any_sync(0xFFFFFFFFu, in_n && exact) != 0;
            if (in_n) {
                const float4* __restrict__ nodes = tree.nodes;
                const F8 nd = ld256(nodes + (size_t)cur * 2);
                const int count = __float_as_int(nd.lo.w), left_first = __float_as_int(nd.hi.w);
                int next = -1;
                if (count > -1) {
                    tri_pos = left_first;
                    tri_end = left_first + count;
                } else if (left_first > -1) {
                    const float4* c = nodes + (size_t)left_first * 2;
                    const F8 lc = ld256(c), rc = ld256(c + 2);
                    bool hl, hr, ln;
                    if (any_exact)
                        bvh_children_packet<true>(lc, rc, k, hl, hr, ln);
                    else
                        bvh_children_packet<false>(lc, rc, k, hl, hr, ln);
                    if (hl && hr) {
                        if (ln) {
                            st.push(left_first);
                            next = left_first + 1;
                        } else {
                            st.push(left_first + 1);
                            next = left_first;
                        }
                    } else if (hl) {
                        next = left_first;
                    } else if (hr) {
                        next = left_first + 1;
                    }
                }
                cur = next;
            }
        }
        // a lane without triangles left needs a node: the register hand-over, else the stack; none left: the packet is done
        if (active && !done && tri_pos >= tri_end && cur < 0) {
            if (st.sp == 0)
                done = true;
            else
                cur = st.pop();
        }
        if (done) {
            fin = true;
            active = false;
        }
    }
    if (fin) store();
}

// ---- scene upload helpers ---------------------------------------------------------------------
__global__ void gather_tris_kernel(const float* __restrict__ verts, uint32_t stride_f, const uint32_t* __restrict__ indices,
                                   uint32_t index_count, uint32_t tri_count, TriRec* __restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= index_count) return;
    const uint32_t id = indices[k];
    TriRec rec;
    if (id < tri_count) {
        const float* v = verts + (size_t)id * 3 * stride_f;
        const float v0x = v[0], v0y = v[1], v0z = v[2];
        const float v1x = v[stride_f], v1y = v[stride_f + 1], v1z = v[stride_f + 2];
        const float v2x = v[2 * stride_f], v2y = v[2 * stride_f + 1], v2z = v[2 * stride_f + 2];
        rec.a = make_float4(v0x, v0y, v0z, __uint_as_float(id));
        rec.b = make_float4(fsub(v1x, v0x), fsub(v1y, v0y), fsub(v1z, v0z), 0.f);
        rec.c = make_float4(fsub(v2x, v0x), fsub(v2y, v0y), fsub(v2z, v0z), 0.f);
        rec.d = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {  // out-of-range id (never produced by the builders): degenerate triangle, never hit
        rec.a = make_float4(0.f, 0.f, 0.f, __uint_as_float(id));
        rec.b = make_float4(0.f, 0.f, 0.f, 0.f);
        rec.c = make_float4(0.f, 0.f, 0.f, 0.f);
        rec.d = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    out[k] = rec;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float hash_unit(uint64_t seed, uint64_t index, uint32_t lane) {
    return (float)(splitmix64(seed ^ (index * 16ull + lane)) >> 40) * (1.0f / 16777216.0f);
}

// CameraView3D::generate_ray (shared/src/lib.rs:157-165); normalize = v * (1 / sqrt(v.v))
__global__ void camera_rays_kernel(float3 pos, float3 p1, float3 right, float3 up, uint32_t width, uint32_t height,
                                   uint32_t row0, uint32_t rows, uint64_t seed, uint64_t frame, RTRay* __restrict__ out) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (size_t)width * rows) return;
    const uint32_t x = (uint32_t)(k % width), y = row0 + (uint32_t)(k / width);
    float fx = (float)x, fy = (float)y;
    if (seed != 0) {
        const uint64_t pix = frame * ((uint64_t)width * height) + (uint64_t)y * width + x;
        fx = fadd(fx, hash_unit(seed, pix, 0));
        fy = fadd(fy, hash_unit(seed, pix, 1));
    }
    const float u = fmul(fx, fdiv(1.0f, (float)width)), v = fmul(fy, fdiv(1.0f, (float)height));
    const float px = fadd(fadd(p1.x, fmul(u, right.x)), fmul(v, up.x));
    const float py = fadd(fadd(p1.y, fmul(u, right.y)), fmul(v, up.y));
    const float pz = fadd(fadd(p1.z, fmul(u, right.z)), fmul(v, up.z));
    const float dx = fsub(px, pos.x), dy = fsub(py, pos.y), dz = fsub(pz, pos.z);
    const float inv = fdiv(1.0f, __fsqrt_rn(fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz))));
    float4* o = reinterpret_cast<float4*>(out + k);
    o[0] = make_float4(pos.x, pos.y, pos.z, 1e-4f);
    o[1] = make_float4(fmul(dx, inv), fmul(dy, inv), fmul(dz, inv), 1e34f);
}

// Breadth-first copy of the top of an Mbvh (one warp): slot 0 = root; a child that also got a slot is addressed as
// (slot | kTopFlag) in its parent's copy, every other field is the node's own.  Slot order inside a level is arbitrary.
__global__ void build_top_table_kernel(const float4* __restrict__ nodes, uint32_t node_count, uint32_t cap,
                                       float4* __restrict__ top, uint32_t* __restrict__ top_count) {
    extern __shared__ int q[];  // node index per slot
    __shared__ int tail;
    const int lane = threadIdx.x;
    if (node_count == 0 || cap == 0) {
        if (lane == 0) *top_count = 0;
        return;
    }
    if (lane == 0) {
        q[0] = 0;
        tail = 1;
    }
    __syncwarp();
    int head = 0;
    for (;;) {
        const int end = tail;
        if (head >= end) break;
        __syncwarp();
        for (int i = head + lane; i < end; i += 32) {
            const float4* n = nodes + (size_t)q[i] * 8;
            int4 ch = as_int4(n[6]);
            const int4 cn = as_int4(n[7]);
            int* c = &ch.x;
            const int* k = &cn.x;
            for (int s = 0; s < 4; s++) {
                if (k[s] <= -1 && c[s] > -1) {  // inner child
                    const int slot = atomicAdd(&tail, 1);
                    if (slot < (int)cap) {
                        q[slot] = c[s];
                        c[s] = slot | kTopFlag;
                    }
                }
            }
            float4* o = top + (size_t)i * 8;
            for (int j = 0; j < 6; j++) o[j] = n[j];
            o[6] = make_float4(__int_as_float(ch.x), __int_as_float(ch.y), __int_as_float(ch.z), __int_as_float(ch.w));
            o[7] = n[7];
        }
        head = end;
        __syncwarp();
        if (lane == 0 && tail > (int)cap) tail = (int)cap;
        __syncwarp();
    }
    if (lane == 0) *top_count = (uint32_t)tail;
}

}  // namespace

int top_table_capacity() { return kTopK; }

cudaError_t launch_build_top_table(const float4* d_nodes, uint32_t node_count, float4* d_top, uint32_t* d_top_count,
                                   cudaStream_t stream) {
    if (kTopK == 0) return cudaSuccess;
    build_top_table_kernel<<<1, 32, (size_t)kTopK * sizeof(int), stream>>>(d_nodes, node_count, (uint32_t)kTopK, d_top, d_top_count);
    return cudaGetLastError();
}

// ---- launchers ----------------------------------------------------------------------------------
// ---- optional ray sorting ----------------------------------------------------------------------------
// Incoherent batches (shadow / bounce rays) are traced in Morton order of (origin, direction): neighbouring
// lanes then walk neighbouring parts of the tree (fewer distinct lines per load, better L1/L2 hit rates).
// Results are scattered back through the permutation, so every ray's result is unchanged, bit for bit.
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void ray_keys_kernel(const RTRay* __restrict__ rays, size_t n, float3 lo, float3 inv_ext,
                                unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const F8 ab = ld256(reinterpret_cast<const float4*>(rays) + i * 2);
    const float ox = (ab.lo.x - lo.x) * inv_ext.x, oy = (ab.lo.y - lo.y) * inv_ext.y, oz = (ab.lo.z - lo.z) * inv_ext.z;
    const float len = sqrtf(ab.hi.x * ab.hi.x + ab.hi.y * ab.hi.y + ab.hi.z * ab.hi.z);
    const float il = len > 0.f ? 0.5f / len : 0.f;
    const float dx = ab.hi.x * il + 0.5f, dy = ab.hi.y * il + 0.5f, dz = ab.hi.z * il + 0.5f;
    auto q = [](float v) { return (uint32_t)fminf(fmaxf(v * 1024.0f, 0.0f), 1023.0f); };  // NaN -> 0
    const uint32_t ko = spread10(q(ox)) | (spread10(q(oy)) << 1) | (spread10(q(oz)) << 2);
    const uint32_t kd = spread10(q(dx)) | (spread10(q(dy)) << 1) | (spread10(q(dz)) << 2);
    keys[i] = ((unsigned long long)ko << 30) | kd;
    idx[i] = (uint32_t)i;
}

// grid of a persistent kernel: resident blocks per SM x number of SMs (queried once per kernel)
template <class K>
static unsigned persistent_grid(K kernel, int block = kBlock, size_t dyn_smem = 0) {
    int dev = 0, sms = kSmCount, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (dyn_smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, dyn_smem);
    return (unsigned)(sms * (per_sm > 0 ? per_sm : 1));
}
// dynamic shared memory of the persistent single-ray kernel: stacks, plus the staged top of the tree where it is used
template <int TREE, bool PHASED>
constexpr size_t persistent_smem() {
    return (size_t)kSmemStack * kPBlock * sizeof(int) +
           ((kTopK > 0 && PHASED && TREE == RT_TREE_MBVH) ? (size_t)kTopK * kTopRow * sizeof(float4) : 0);
}

template <int TREE, bool ANY>
static cudaError_t launch_single_t(const DeviceTree& tree, const RTRay* d_rays, size_t n, RTHit* d_hits,
                                   uint8_t* d_occluded, const uint32_t* d_perm, unsigned long long* d_counter,
                                   uint32_t* d_overflow, int mode, const PeerDests& pd, cudaStream_t stream) {
    if ((pd.count > 0 || pd.ready || pd.directions) && mode != kTracePersistent && mode != kTracePhased) return cudaErrorNotSupported;  // fused gather / input gate / split input live in the default kernel
    const size_t blocks_needed = ceil_div(n, kBlock);
    const size_t pblocks_needed = ceil_div(n, kPBlock);
    const bool persistent = mode != kTraceStatic;
    if (TREE == RT_TREE_MBVH && mode == kTraceCoop) {
        static const unsigned machine = persistent_grid(trace_mbvh_coop_kernel<ANY>);
        const unsigned grid = (unsigned)(blocks_needed < machine ? blocks_needed : machine);
        cudaError_t e = cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return e;
        trace_mbvh_coop_kernel<ANY><<<grid, kBlock, 0, stream>>>(tree, d_rays, n, d_hits, d_occluded, d_perm, d_counter, d_overflow);
        return cudaGetLastError();
    }
    if (mode == kTracePhased) {
        constexpr size_t smem = persistent_smem<TREE, true>();
        static const unsigned machine = persistent_grid(trace_single_persistent_kernel<TREE, ANY, true, false>, kPBlock, smem);
        const unsigned grid = (unsigned)(pblocks_needed < machine ? pblocks_needed : machine);
        cudaError_t e = cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return e;
        trace_single_persistent_kernel<TREE, ANY, true, false><<<grid, kPBlock, smem, stream>>>(tree, d_rays, n, d_hits, d_occluded, d_perm,
                                                                                    d_counter, d_overflow, pd);
    } else if (persistent) {
        constexpr size_t smem = persistent_smem<TREE, false>();
        cudaError_t e = cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return e;
        const bool push = pd.push != 0 && pd.count > 0 && d_perm == nullptr && (pd.offset & 1) == 0 &&
                          n < 0xFFFFFFFFull && (ANY ? d_occluded != nullptr : d_hits != nullptr);
        if (push) {
            static const unsigned machine = persistent_grid(trace_single_persistent_kernel<TREE, ANY, false, true>, kPBlock, smem);
            const unsigned grid = (unsigned)(pblocks_needed < machine ? pblocks_needed : machine);
            trace_single_persistent_kernel<TREE, ANY, false, true><<<grid, kPBlock, smem, stream>>>(tree, d_rays, n, d_hits, d_occluded,
                                                                                        d_perm, d_counter, d_overflow, pd);
        } else {
            static const unsigned machine = persistent_grid(trace_single_persistent_kernel<TREE, ANY, false, false>, kPBlock, smem);
            const unsigned grid = (unsigned)(pblocks_needed < machine ? pblocks_needed : machine);
            trace_single_persistent_kernel<TREE, ANY, false, false><<<grid, kPBlock, smem, stream>>>(tree, d_rays, n, d_hits, d_occluded,
                                                                                         d_perm, d_counter, d_overflow, pd);
        }
    } else {
        trace_single_kernel<TREE, ANY><<<(unsigned)blocks_needed, kBlock, 0, stream>>>(tree, d_rays, n, d_hits, d_occluded,
                                                                                     d_perm, d_overflow);
    }
    return cudaGetLastError();
}

cudaError_t launch_trace_single(const DeviceTree& tree, int tree_kind, bool any, const RTRay* d_rays, size_t n,
                                RTHit* d_hits, uint8_t* d_occluded, unsigned long long* d_counter,
                                uint32_t* d_overflow, int mode, const float* sort_bounds, const PeerDests* peers,
                                cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    PeerDests pd{};
    if (peers) pd = *peers;
    if ((pd.ready || pd.directions) && (sort_bounds != nullptr || (mode != kTracePersistent && mode != kTracePhased))) return cudaErrorNotSupported;  // gate / split input: default kernel, caller's order
    uint32_t* d_perm = nullptr;
    void* scratch = nullptr;
    if (sort_bounds != nullptr && n >= 4096 && n <= (size_t)0x7FFFFFFF) {  // CUB's num_items is an int; larger batches are traced unsorted
        // scratch from the stream-ordered pool: keys in/out (8 B), indices in/out (4 B), CUB temp
        size_t temp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                        (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 60, stream);
        const size_t kb = (n * 8 + 255) & ~size_t(255), ib = (n * 4 + 255) & ~size_t(255);
        cudaError_t e = cudaMallocAsync(&scratch, 2 * kb + 2 * ib + temp_bytes, stream);
        if (e != cudaSuccess) return e;
        char* p = (char*)scratch;
        unsigned long long* k_in = (unsigned long long*)p;
        unsigned long long* k_out = (unsigned long long*)(p + kb);
        uint32_t* i_in = (uint32_t*)(p + 2 * kb);
        uint32_t* i_out = (uint32_t*)(p + 2 * kb + ib);
        void* temp = p + 2 * kb + 2 * ib;
        const float3 lo = make_float3(sort_bounds[0], sort_bounds[1], sort_bounds[2]);
        const float3 ie = make_float3(1.0f / fmaxf(sort_bounds[3] - sort_bounds[0], 1e-30f),
                                      1.0f / fmaxf(sort_bounds[4] - sort_bounds[1], 1e-30f),
                                      1.0f / fmaxf(sort_bounds[5] - sort_bounds[2], 1e-30f));
        ray_keys_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(d_rays, n, lo, ie, k_in, i_in);
        e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k_in, k_out, i_in, i_out, (int)n, 0, 60, stream);
        if (e != cudaSuccess) {
            cudaFreeAsync(scratch, stream);
            return e;
        }
        d_perm = i_out;
    }
    cudaError_t e;
    if (tree_kind == RT_TREE_MBVH)
        e = any ? launch_single_t<RT_TREE_MBVH, true>(tree, d_rays, n, d_hits, d_occluded, d_perm, d_counter, d_overflow, mode, pd, stream)
                : launch_single_t<RT_TREE_MBVH, false>(tree, d_rays, n, d_hits, d_occluded, d_perm, d_counter, d_overflow, mode, pd, stream);
    else
        e = any ? launch_single_t<RT_TREE_BVH, true>(tree, d_rays, n, d_hits, d_occluded, d_perm, d_counter, d_overflow, mode, pd, stream)
                : launch_single_t<RT_TREE_BVH, false>(tree, d_rays, n, d_hits, d_occluded, d_perm, d_counter, d_overflow, mode, pd, stream);
    if (scratch) cudaFreeAsync(scratch, stream);
    return e;
}

template <int TREE, bool ANY>
static cudaError_t launch_packets_t(const DeviceTree& tree, const RTRayPacket4* d_packets, size_t n_packets, float t_min,
                                    RTHitPacket4* d_hits, uint8_t* d_occluded, unsigned long long* d_counter,
                                    uint32_t* d_overflow, int mode, cudaStream_t stream) {
    const size_t blocks_needed = ceil_div(n_packets * 4, kBlock);
    // the lane kernels move packets and results with 16-byte accesses (any hit: one 32-bit word per packet); the C types only
    // promise 4-byte alignment, so oddly placed caller buffers take the quad kernel (scalar accesses)
    const bool aligned = ((uintptr_t)d_packets & 15u) == 0 && ((uintptr_t)d_hits & 15u) == 0 && ((uintptr_t)d_occluded & 3u) == 0;
    if (mode == kTraceLane && !aligned) mode = kTraceStatic;
    if (mode == kTraceLane) {  // one lane per packet
        const size_t need = ceil_div(n_packets, kLBlock);
        cudaError_t e = cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return e;
        if (TREE == RT_TREE_MBVH) {
            static const unsigned machine = persistent_grid(trace_mbvh_packet_lane_kernel<ANY>, kLBlock);
            const unsigned grid = (unsigned)(need < machine ? need : machine);
            trace_mbvh_packet_lane_kernel<ANY><<<grid, kLBlock, 0, stream>>>(tree, d_packets, n_packets, t_min, d_hits, d_occluded,
                                                                            d_counter, d_overflow);
        } else {
            static const unsigned machine = persistent_grid(trace_bvh_packet_lane_kernel<ANY>, kLBlock);
            const unsigned grid = (unsigned)(need < machine ? need : machine);
            trace_bvh_packet_lane_kernel<ANY><<<grid, kLBlock, 0, stream>>>(tree, d_packets, n_packets, t_min, d_hits, d_occluded,
                                                                           d_counter, d_overflow);
        }
    } else if (mode != kTraceStatic) {
        static const unsigned machine = persistent_grid(trace_packet_persistent_kernel<TREE, ANY>);
        const unsigned grid = (unsigned)(blocks_needed < machine ? blocks_needed : machine);
        cudaError_t e = cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return e;
        trace_packet_persistent_kernel<TREE, ANY><<<grid, kBlock, 0, stream>>>(tree, d_packets, n_packets, t_min, d_hits,
                                                                              d_occluded, d_counter, d_overflow);
    } else {
        trace_packet_kernel<TREE, ANY><<<(unsigned)blocks_needed, kBlock, 0, stream>>>(tree, d_packets, n_packets, t_min, d_hits,
                                                                                     d_occluded, d_overflow);
    }
    return cudaGetLastError();
}

cudaError_t launch_trace_packets(const DeviceTree& tree, int tree_kind, bool any, const RTRayPacket4* d_packets,
                                 size_t n_packets, float t_min, RTHitPacket4* d_hits, uint8_t* d_occluded,
                                 unsigned long long* d_counter, uint32_t* d_overflow, int mode, cudaStream_t stream) {
    if (n_packets == 0) return cudaSuccess;
    if (tree_kind == RT_TREE_MBVH)
        return any ? launch_packets_t<RT_TREE_MBVH, true>(tree, d_packets, n_packets, t_min, d_hits, d_occluded, d_counter, d_overflow, mode, stream)
                   : launch_packets_t<RT_TREE_MBVH, false>(tree, d_packets, n_packets, t_min, d_hits, d_occluded, d_counter, d_overflow, mode, stream);
    return any ? launch_packets_t<RT_TREE_BVH, true>(tree, d_packets, n_packets, t_min, d_hits, d_occluded, d_counter, d_overflow, mode, stream)
               : launch_packets_t<RT_TREE_BVH, false>(tree, d_packets, n_packets, t_min, d_hits, d_occluded, d_counter, d_overflow, mode, stream);
}

cudaError_t launch_gather_tris(const float* d_verts, uint32_t stride_floats, const uint32_t* d_indices,
                               uint32_t index_count, uint32_t tri_count, TriRec* d_out, cudaStream_t stream) {
    if (index_count == 0) return cudaSuccess;
    gather_tris_kernel<<<(unsigned)ceil_div(index_count, 256), 256, 0, stream>>>(d_verts, stride_floats, d_indices,
                                                                                  index_count, tri_count, d_out);
    return cudaGetLastError();
}

cudaError_t launch_camera_rays(const float pos[3], const float p1[3], const float right[3], const float up[3],
                               uint32_t width, uint32_t height, uint32_t row0, uint32_t rows, uint64_t seed,
                               uint64_t frame, RTRay* d_rays, cudaStream_t stream) {
    const size_t n = (size_t)width * rows;
    if (n == 0) return cudaSuccess;
    camera_rays_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(
        make_float3(pos[0], pos[1], pos[2]), make_float3(p1[0], p1[1], p1[2]), make_float3(right[0], right[1], right[2]),
        make_float3(up[0], up[1], up[2]), width, height, row0, rows, seed, frame, d_rays);
    return cudaGetLastError();
}

}  // namespace rtb
