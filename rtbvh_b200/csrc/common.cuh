// common.cuh — device-side types and exact-arithmetic helpers shared by the sm_100a kernels.
//
// All floating-point work on the hot path is written with the round-to-nearest intrinsics
// (__fadd_rn / __fsub_rn / __fmul_rn / __fdiv_rn): nvcc never contracts those into FMA, which is
// what keeps t bit-identical to the reference (Rust never fuses).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rtbvh_gpu.h"

namespace rtb {

constexpr int kSmCount = 148;  // B200: 2 dies x 74 SMs
constexpr uint32_t kNoHit = 0xFFFFFFFFu;

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// glam Vec4::min / max on x86-64 are _mm_min_ps / _mm_max_ps: `a < b ? a : b` (second operand when
// either is NaN).  When no operand can be NaN this equals FMNMX (fminf/fmaxf) up to the sign of zero,
// which no predicate on the path can observe.  EXACT = true is used only for rays whose slab
// products can produce NaN (a zero / non-finite direction or origin component).
template <bool EXACT>
__device__ __forceinline__ float vmin(float a, float b) {
    if (EXACT) return a < b ? a : b;
    return fminf(a, b);
}
template <bool EXACT>
__device__ __forceinline__ float vmax(float a, float b) {
    if (EXACT) return a > b ? a : b;
    return fmaxf(a, b);
}

// Triangle record in leaf order: 64 bytes = 2 x LDG.256 (sm_100 has 256-bit global loads).
//   a = (v0.xyz, bitcast prim id), b = (v1 - v0, 0), c = (v2 - v0, 0), d = unused padding
// edge1 / edge2 are the same single fp32 subtractions the reference performs per test
// (src/builders/spatial_sah.rs:136-137), done once at upload.
struct alignas(32) TriRec {
    float4 a, b, c, d;
};
static_assert(sizeof(TriRec) == 64, "");

// 256-bit read-only global load (PTX ISA 8.8, sm_100+: SASS LDG.E.ENL2.256.CONSTANT).  A lane that
// fetches its own node pays one L1 wavefront per load instruction, so halving the instruction count
// halves the L1 data-pipe work of a node visit.  RTB_LD256=0 builds the 2 x LDG.128 variant for A/B.
#ifndef RTB_PREFETCH
#define RTB_PREFETCH 1
#endif
#ifndef RTB_LD256
#define RTB_LD256 1
#endif
struct alignas(32) F8 {
    float4 lo, hi;
};
__device__ __forceinline__ F8 ld256(const void* p) {
    F8 r;
#if RTB_LD256
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
                 : "l"(p));
#else
    r.lo = __ldg(reinterpret_cast<const float4*>(p));
    r.hi = __ldg(reinterpret_cast<const float4*>(p) + 1);
#endif
    return r;
}

// Experiment knobs (A/B only; defaults are the plain loads above): RTB_TRI_NOALLOC = 1 loads triangle records without
// allocating in L1 (they are touched once or twice, the nodes many times), RTB_NODE_EVICT_LAST = 1 marks node lines evict-last in L1.
#ifndef RTB_TRI_NOALLOC
#define RTB_TRI_NOALLOC 0
#endif
#ifndef RTB_NODE_EVICT_LAST
#define RTB_NODE_EVICT_LAST 0
#endif
__device__ __forceinline__ F8 ld256_tri(const void* p) {
#if RTB_TRI_NOALLOC
    F8 r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w) : "l"(p));
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w) : "l"(p));
    return r;
#else
    return ld256(p);
#endif
}
__device__ __forceinline__ F8 ld256_node(const void* p) {
#if RTB_NODE_EVICT_LAST
    F8 r;
    asm("ld.global.nc.L1::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
        : "l"(p));
    return r;
#else
    return ld256(p);
#endif
}

struct DeviceTree {
    const float4* nodes;      // Bvh: 2 float4 per node; Mbvh: 8 float4 per node
    uint32_t node_count;
    const TriRec* tris;       // index_count records (leaf order: tris[k] = triangle indices[k])
    uint32_t index_count;
    // Mbvh only, optional: copies of the top `top_count` nodes in breadth-first order whose child fields address other
    // copies as (slot | kTopFlag); the persistent kernel stages them in shared memory (traverse.cu, RTB_TOPK)
    const float4* top;
    uint32_t top_count;
};
constexpr int kTopFlag = 1 << 30;

inline __host__ __device__ size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }

}  // namespace rtb
