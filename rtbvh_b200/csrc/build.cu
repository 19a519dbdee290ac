// build.cu — GPU builders behind create_bvh / create_mbvh / refit.
//
// Replaces (reference file:line):
//   Builder::construct_binned_sah            src/bvh.rs:87-111  -> BinnedSahBuilder::build  src/builders/binned_sah.rs:346-399,
//                                            BinnedSahBuildTask::run :132-282, find_split :80-114, compute_bin_index :119-128
//   Builder::construct_locally_ordered_clustered  src/bvh.rs:113-137 -> LocallyOrderedClusteringBuilder src/builders/locb.rs:18-328,
//                                            MortonEncoder src/morton.rs:28-102
//   Mbvh::construct / MbvhNode::merge_nodes  src/bvh.rs:381-404, src/mbvh_node.rs:297-411
//   Bvh::refit                               src/bvh.rs:176-205
//
// Design (B200-first, not a port of the task-per-thread CPU builders):
//   * binned SAH runs breadth-first, one LEVEL of the tree per pass over the primitives:
//       bin     : one thread per primitive position, 3 x 16 bins per active node accumulated with integer
//                 atomics on order-preserving float keys (block-private shared-memory bins when a block
//                 lies inside one node, i.e. on the top levels where contention would be highest);
//       split   : one warp per node: suffix/prefix scans of the 16 bins by shuffles, SAH argmin, the
//                 reference's leaf / 40 %-median-fallback rules, child boxes;
//       scatter : stable partition of every node's index range by one segmented scan (CUB
//                 ExclusiveSumByKey) + one scatter into the ping-pong index buffer.
//     min/max/count are exact and order independent, every float expression is written with the *_rn
//     intrinsics in the reference's order, so the GPU tree has the SAME topology and boxes as the
//     reference's (node numbering is level order instead of the reference's scheduling-dependent order,
//     and primitives inside a leaf are in stable instead of swap-partition order).  Quirks Q3/Q4 of
//     SURVEY.md are reproduced on purpose: the contract is "same hits as the reference-built tree".
//   * LOCB: world box reduce -> 30-bit Morton codes -> CUB radix sort (stable) -> per iteration
//     nearest-neighbour search (radius 14) / mutual-pair flags / inclusive scan / write, all data
//     parallel; the result is bit-identical to the reference algorithm, node for node.
//   * collapse and refit are pointer-chasing kernels over the finished binary tree.
#include <cub/cub.cuh>
#include <mutex>

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <atomic>

#include "build.cuh"

namespace rtb {

#define RTB_CUDA(call)                                   \
    do {                                                 \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) return fail(#call, e__); \
    } while (0)

thread_local BuildStats g_build_stats;

namespace {

constexpr int kBins = 16;            // binned_sah.rs:314
constexpr int kMaxDepth = 64;        // binned_sah.rs:313
constexpr int kBinWords = 7;         // min xyz, max xyz (ordered keys), count
constexpr int kTaskBinWords = 3 * kBins * kBinWords;  // 336 words = 1344 B per active node
constexpr float kPad = 0.0001f;
constexpr int kRadius = 14;          // locb.rs:27
#ifndef RTB_SMALL
#define RTB_SMALL 32
#endif
constexpr uint32_t kSmall = RTB_SMALL;  // subtrees with <= kSmall (<= 32) primitives are finished by one warp (sah_small_kernel)
// Level tasks with <= kWarpTask primitives are binned, split AND partitioned by one warp each (sah_warp_task_kernel: bins
// and the task's index range in shared memory, no global atomics, no global partition pass); larger ("span-class") tasks go
// through the span-based bin kernel + sah_split_kernel + the segmented-scan partition.  pos_task holds, per index
// position, the span-class task it belongs to or -1 (finished, or inside a warp-class task: nothing for the span passes).
#ifndef RTB_WARP_TASK
#define RTB_WARP_TASK 512
#endif
constexpr uint32_t kWarpTask = RTB_WARP_TASK;
__host__ __device__ inline int32_t pt_encode(uint32_t t, uint32_t n) { return n <= kWarpTask ? -1 : (int32_t)t; }
__device__ __forceinline__ int32_t pt_task(int32_t pt) { return pt; }

// ---- order-preserving float <-> uint keys (so min/max can be integer atomics) --------------------
__host__ __device__ inline uint32_t fkey(float f) {
#ifdef __CUDA_ARCH__
    const uint32_t b = __float_as_uint(f);
#else
    uint32_t b;
    memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ inline float fkey_inv(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

struct Box {
    float mn[3], mx[3];
};
__device__ __forceinline__ Box box_empty() { return Box{{1e34f, 1e34f, 1e34f}, {-1e34f, -1e34f, -1e34f}}; }  // aabb.rs:50-57
__device__ __forceinline__ Box box_union(const Box& a, const Box& b) {  // grow_bb: f32::min / f32::max
    Box r;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        r.mn[k] = fminf(a.mn[k], b.mn[k]);
        r.mx[k] = fmaxf(a.mx[k], b.mx[k]);
    }
    return r;
}
__device__ __forceinline__ void box_pad(Box& b, float d) {  // offset_by, aabb.rs:313-321
#pragma unroll
    for (int k = 0; k < 3; k++) {
        b.mn[k] = fsub(b.mn[k], d);
        b.mx[k] = fadd(b.mx[k], d);
    }
}
__device__ __forceinline__ float box_half_area(const Box& b) {  // aabb.rs:343-346
    const float dx = fsub(b.mx[0], b.mn[0]), dy = fsub(b.mx[1], b.mn[1]), dz = fsub(b.mx[2], b.mn[2]);
    return fadd(fmul(fadd(dx, dy), dz), fmul(dx, dy));
}
__device__ __forceinline__ int box_longest_axis(const Box& b) {  // aabb.rs:354-363
    const float e0 = fsub(b.mx[0], b.mn[0]), e1 = fsub(b.mx[1], b.mn[1]), e2 = fsub(b.mx[2], b.mn[2]);
    int a = 0;
    if (e1 > e0) a = 1;
    if (e2 > (a == 0 ? e0 : e1)) a = 2;
    return a;
}
__device__ __forceinline__ Box load_box(const float4* __restrict__ bb, size_t i) {
    const float4 lo = bb[i * 2], hi = bb[i * 2 + 1];
    return Box{{lo.x, lo.y, lo.z}, {hi.x, hi.y, hi.z}};
}
__device__ __forceinline__ void store_node(float4* nodes, size_t i, const Box& b, int count, int left_first) {
    nodes[i * 2] = make_float4(b.mn[0], b.mn[1], b.mn[2], __int_as_float(count));
    nodes[i * 2 + 1] = make_float4(b.mx[0], b.mx[1], b.mx[2], __int_as_float(left_first));
}

// compute_bin_index, binned_sah.rs:119-128 (`as usize` saturates, NaN -> 0)
__device__ __forceinline__ int bin_index(float c, float k, float off) {
    const float v = fmaxf(fadd(fmul(c, k), off), 0.0f);
    return v >= (float)kBins ? kBins - 1 : (int)v;
}

// ---- primitives ----------------------------------------------------------------------------------
// aabbs == null: the centers act as point primitives (rtbvh_ffi/src/lib.rs:396-422): aabb = grow(empty, center)
__global__ void point_boxes_kernel(const float* __restrict__ cen, uint32_t stride, uint32_t n, float4* __restrict__ bb) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = cen[(size_t)i * stride], y = cen[(size_t)i * stride + 1], z = cen[(size_t)i * stride + 2];
    bb[(size_t)i * 2] = make_float4(fminf(1e34f, x), fminf(1e34f, y), fminf(1e34f, z), 0.f);
    bb[(size_t)i * 2 + 1] = make_float4(fmaxf(-1e34f, x), fmaxf(-1e34f, y), fmaxf(-1e34f, z), 0.f);
}

// Triangle::aabb / Triangle::center of the bench primitive (shared/src/lib.rs:27-39)
__global__ void tri_prims_kernel(const float* __restrict__ verts, uint32_t stride, uint32_t n, float4* __restrict__ bb,
                                 float* __restrict__ cen) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* v = verts + (size_t)i * 3 * stride;
    float mn[3], mx[3], c[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float a = v[k], b = v[stride + k], d = v[2 * stride + k];
        mn[k] = fminf(fminf(fminf(1e34f, a), b), d);
        mx[k] = fmaxf(fmaxf(fmaxf(-1e34f, a), b), d);
        c[k] = fmul(fadd(fadd(a, b), d), 1.0f / 3.0f);
    }
    bb[(size_t)i * 2] = make_float4(mn[0], mn[1], mn[2], 0.f);
    bb[(size_t)i * 2 + 1] = make_float4(mx[0], mx[1], mx[2], 0.f);
    cen[(size_t)i * 3] = c[0];
    cen[(size_t)i * 3 + 1] = c[1];
    cen[(size_t)i * 3 + 2] = c[2];
}

__global__ void tri_boxes_kernel(const float* __restrict__ verts, uint32_t stride, uint32_t n, float4* __restrict__ bb) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* v = verts + (size_t)i * 3 * stride;
    float mn[3], mx[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float a = v[k], b = v[stride + k], d = v[2 * stride + k];
        mn[k] = fminf(fminf(fminf(1e34f, a), b), d);
        mx[k] = fmaxf(fmaxf(fmaxf(-1e34f, a), b), d);
    }
    bb[(size_t)i * 2] = make_float4(mn[0], mn[1], mn[2], 0.f);
    bb[(size_t)i * 2 + 1] = make_float4(mx[0], mx[1], mx[2], 0.f);
}

// Aabb::union_of_list without the pad (aabb.rs:125-129): warp shuffle reduce, block reduce through shared-memory key atomics,
// then 6 global key atomics per BLOCK (per warp they were 57 k atomics on six addresses: 44 us of a 1 Mi-triangle build)
__global__ void world_reduce_kernel(const float4* __restrict__ bb, uint32_t n, uint32_t* __restrict__ keys) {
    __shared__ uint32_t s_keys[6];
    if (threadIdx.x < 3) s_keys[threadIdx.x] = fkey(1e34f);
    else if (threadIdx.x < 6) s_keys[threadIdx.x] = fkey(-1e34f);
    __syncthreads();
    Box b = box_empty();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        b = box_union(b, load_box(bb, i));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            b.mn[k] = fminf(b.mn[k], __shfl_xor_sync(0xFFFFFFFFu, b.mn[k], off));
            b.mx[k] = fmaxf(b.mx[k], __shfl_xor_sync(0xFFFFFFFFu, b.mx[k], off));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            atomicMin(&s_keys[k], fkey(b.mn[k]));
            atomicMax(&s_keys[3 + k], fkey(b.mx[k]));
        }
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicMin(&keys[threadIdx.x], s_keys[threadIdx.x]);
    else if (threadIdx.x < 6) atomicMax(&keys[threadIdx.x], s_keys[threadIdx.x]);
}
__global__ void world_init_kernel(uint32_t* keys) {
    if (threadIdx.x < 3) keys[threadIdx.x] = fkey(1e34f);
    else if (threadIdx.x < 6) keys[threadIdx.x] = fkey(-1e34f);
}
// world = union_of_list(aabbs) incl. its 1e-4 pad, written as node `dst` (count/left_first given)
__global__ void world_to_node_kernel(const uint32_t* __restrict__ keys, float4* nodes, uint32_t dst, int count, int left_first,
                                     float* __restrict__ world_out) {
    Box b;
    for (int k = 0; k < 3; k++) {
        b.mn[k] = fkey_inv(keys[k]);
        b.mx[k] = fkey_inv(keys[3 + k]);
    }
    box_pad(b, kPad);
    if (nodes) store_node(nodes, dst, b, count, left_first);
    if (world_out)
        for (int k = 0; k < 3; k++) {
            world_out[k] = b.mn[k];
            world_out[3 + k] = b.mx[k];
        }
}
__global__ void iota_kernel(uint32_t* a, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}
__global__ void fill_i32_kernel(int32_t* a, uint32_t n, int32_t v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

// =================================================================================================
// Binned SAH, breadth first
// =================================================================================================
struct Task {
    uint32_t node, begin, end;
};
struct TaskAux {
    float k[3], off[3];  // center_to_bin, bin_offset (binned_sah.rs:146-147)
    uint32_t slot;       // span-class tasks (> kWarpTask primitives): which 3 x 16 bin block in HBM is theirs this level
    uint32_t pad;
};
struct alignas(32) PartTask {  // everything the partition needs to know about a task, one 32-byte sector (written by emit)
    uint32_t begin, nleft;         // nleft == 0: the task did not split (a split always has 0 < nleft < n)
    int32_t child_left, child_right;  // encoded next-level task of each side, -1 if that child needs no further level pass
    uint32_t shift, split_index;   // predicate: ((binidx >> shift) & 15) < split_index; split_index == 0: no split
    uint32_t pad[2];
};
struct Decision {
    uint32_t split;        // 1: node becomes inner
    uint32_t axis;         // final best_axis
    uint32_t split_index;  // bins < split_index go left
    uint32_t nleft;
    float lmn[3], lmx[3], rmn[3], rmx[3];
};

// What the level loop knows about level d (input of its kernels), kept on the device: the host enqueues levels without
// waiting for their results.  state[d + 1] is written by level d's emit kernel.
struct LevelState {
    uint32_t A;            // tasks of this level
    uint32_t node_count;   // level nodes allocated so far
    uint32_t S;            // small subtrees collected so far
    uint32_t small_slots;  // node slots reserved for them
};

// Entry of BinnedSahBuildTask::run (binned_sah.rs:133-147): pad the node box, compute the binning transform.  Every
// task has >= 2 primitives and depth < 64 (children that are leaves on entry are finalised by emit_kernel).
__device__ __forceinline__ Box task_enter(float4* nodes, uint32_t node, Box b, TaskAux* aux_slot, uint32_t bin_slot) {
    box_pad(b, kPad);
    store_node(nodes, node, b, 0, 0);
    TaskAux a;
    a.slot = bin_slot;
    a.pad = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        a.k[k] = fmul(fdiv(1.0f, fsub(b.mx[k], b.mn[k])), (float)kBins);
        a.off[k] = fmul(-b.mn[k], a.k[k]);
    }
    *aux_slot = a;
    return b;
}
// Level 0: the root task (node 0 already holds union_of_list of the primitives).
__global__ void sah_root_task_kernel(float4* nodes, uint32_t n, Task* tasks, TaskAux* aux, LevelState* state,
                                     uint32_t* span_tasks) {
    tasks[0] = Task{0u, 0u, n};
    task_enter(nodes, 0, load_box(nodes, 0), &aux[0], 0u);
    state[0] = LevelState{1u, 1u, 0u, 0u};
    span_tasks[0] = n > kWarpTask ? 1u : 0u;
}

__global__ void sah_bins_init_kernel(uint32_t* __restrict__ bins, size_t words) {
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= words) return;
    const uint32_t f = (uint32_t)(w % kBinWords);
    bins[w] = f < 3 ? fkey(1e34f) : (f < 6 ? fkey(-1e34f) : 0u);
}

__device__ __forceinline__ void bin_accumulate(uint32_t* b, const Box& box) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
        atomicMin(&b[k], fkey(box.mn[k]));
        atomicMax(&b[3 + k], fkey(box.mx[k]));
    }
    atomicAdd(&b[6], 1u);
}

// Fill bins with primitives (binned_sah.rs:157-172).  Every block walks a contiguous span of index positions in
// chunks of kBinBlock; while consecutive chunks lie inside ONE task (always, on the top levels) the bins are
// accumulated in shared memory and flushed once per (block, task): the root level then issues grid x 336 global
// atomics instead of (n / 256) x 336.  Chunks that straddle tasks fall back to direct global atomics (deep
// levels: small tasks, low contention).
constexpr int kBinBlock = 256;
// Runs at least this long get lane-private bin copies (42 KB to init and fold per run; 384 measured slower than 1024).
constexpr uint32_t kLanePrivateRun = 1024;
// Fill bins with primitives (binned_sah.rs:157-172) for the span-class tasks (> kWarpTask primitives) of the level.
// Every block owns a contiguous span of index positions and walks it RUN by run (a run = the part of one task's range
// inside the span; the task table gives its end, so no search).  A run is accumulated in shared memory and folded into
// the task's bin block in HBM with 336 atomics, so a primitive never costs a global atomic.  With ONE copy of the
// 3 x 16 bins the 21 shared-memory atomics of a primitive collide inside the warp (32 lanes on 16 bins, ~5 replays each:
// measured 77 ns per primitive and level at 10 M triangles); long runs therefore use a private copy per LANE — entry e
// of lane l at sb[e * 32 + l]: one instruction touches 32 different banks and never the same address twice — which
// brought the root level of a 10 M build from 776 to 122 us.  Short runs keep the single copy (cheap to init and fold).
template <int COPIES>
__device__ __forceinline__ void bin_run(uint32_t* sb, uint32_t pos, uint32_t run_end, const TaskAux& a,
                                        const uint32_t* __restrict__ idx, const float4* __restrict__ bb,
                                        const float* __restrict__ cen, uint32_t cstride, uint32_t* __restrict__ g,
                                        uint16_t* __restrict__ binidx) {
    const unsigned lane = threadIdx.x & (unsigned)(COPIES - 1);  // COPIES is a power of two <= 32
    for (int w = threadIdx.x; w < kTaskBinWords * COPIES; w += kBinBlock) {
        const int f = (w / COPIES) % kBinWords;
        sb[w] = f < 3 ? fkey(1e34f) : (f < 6 ? fkey(-1e34f) : 0u);
    }
    __syncthreads();
    for (uint32_t i = pos + threadIdx.x; i < run_end; i += kBinBlock) {
        const uint32_t p = idx[i];
        const Box box = load_box(bb, p);
        uint32_t packed = 0;
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
            const int b = bin_index(cen[(size_t)p * cstride + ax], a.k[ax], a.off[ax]);
            uint32_t* e = &sb[((ax * kBins + b) * kBinWords) * COPIES + lane];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                atomicMin(e + k * COPIES, fkey(box.mn[k]));
                atomicMax(e + (3 + k) * COPIES, fkey(box.mx[k]));
            }
            atomicAdd(e + 6 * COPIES, 1u);
            packed |= (uint32_t)b << (4 * ax);
        }
        binidx[i] = (uint16_t)packed;  // the partition pass reads this instead of gathering the centroid again
    }
    __syncthreads();
    for (int w = threadIdx.x; w < kTaskBinWords; w += kBinBlock) {
        const int f = w % kBinWords;
        uint32_t v = f < 3 ? fkey(1e34f) : (f < 6 ? fkey(-1e34f) : 0u);
        for (unsigned k = 0; k < (unsigned)COPIES; k++) {
            const uint32_t c = sb[w * COPIES + ((threadIdx.x + k) & (COPIES - 1))];  // rotated: conflict-free across the warp
            v = f < 3 ? min(v, c) : (f < 6 ? max(v, c) : v + c);
        }
        if (f < 3) {
            if (v != fkey(1e34f)) atomicMin(&g[w], v);
        } else if (f < 6) {
            if (v != fkey(-1e34f)) atomicMax(&g[w], v);
        } else if (v) {
            atomicAdd(&g[w], v);
        }
    }
    __syncthreads();
}
// PRIV = false: the small-build variant without the lane-private path (1.3 KB instead of 42 KB of shared memory per block,
// 8 instead of 5 blocks per SM): below a few M primitives a block sees too few primitives per run for the copies to pay
// (measured at 1 Mi triangles: 0.55 ms per build for this variant, 0.71-0.77 ms with lane-private runs).
#ifndef RTB_BIN_COPIES
#define RTB_BIN_COPIES 1  // small-build variant: copies of the bins per block (lane & (COPIES - 1) picks one): fewer replays
#endif
constexpr int kBinCopies = RTB_BIN_COPIES;
template <bool PRIV>
__global__ void __launch_bounds__(kBinBlock) sah_bin_kernel(const uint32_t* __restrict__ idx, const int32_t* __restrict__ pos_task,
                                                            uint32_t n, uint32_t span, const Task* __restrict__ tasks,
                                                            const TaskAux* __restrict__ aux, const float4* __restrict__ bb,
                                                            const float* __restrict__ cen, uint32_t cstride,
                                                            uint32_t* __restrict__ bins,
                                                            uint16_t* __restrict__ binidx /* 3 x 4-bit bin per position */,
                                                            const LevelState* __restrict__ state,
                                                            const uint32_t* __restrict__ span_tasks /* of this level */) {
    __shared__ uint32_t sb[kTaskBinWords * (PRIV ? 32 : kBinCopies)];
    __shared__ uint32_t s_next;
    const uint64_t begin64 = (uint64_t)blockIdx.x * span;
    if (begin64 >= n || state->A == 0 || *span_tasks == 0) return;  // deep levels: only warp-class tasks are left
    const uint32_t end = (uint32_t)min((uint64_t)n, begin64 + span);
    uint32_t pos = (uint32_t)begin64;
    while (pos < end) {  // every decision below is block-uniform
        const int32_t pt = pos_task[pos];
        if (pt >= 0) {  // span-class task: bin the run
            const uint32_t run_end = min(end, tasks[pt].end);
            const TaskAux a = aux[pt];
            uint32_t* g = bins + (size_t)a.slot * kTaskBinWords;
            if (PRIV && run_end - pos >= kLanePrivateRun)
                bin_run<PRIV ? 32 : 1>(sb, pos, run_end, a, idx, bb, cen, cstride, g, binidx);
            else if (!PRIV && kBinCopies > 1 && run_end - pos >= 64u * kBinCopies)
                bin_run<kBinCopies>(sb, pos, run_end, a, idx, bb, cen, cstride, g, binidx);
            else
                bin_run<1>(sb, pos, run_end, a, idx, bb, cen, cstride, g, binidx);
            pos = run_end;
        } else {  // finished / warp-class positions carry no range: find the next span-class one in this chunk
            const uint32_t c1 = min(pos + kBinBlock, end);
            if (threadIdx.x == 0) s_next = c1;
            __syncthreads();
            const uint32_t i = pos + threadIdx.x;
            if (i < c1 && pos_task[i] >= 0) atomicMin(&s_next, i);
            __syncthreads();
            pos = s_next;
            __syncthreads();
        }
    }
}

__device__ __forceinline__ Box shfl_box_down(const Box& b, int off) {
    Box r;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        r.mn[k] = __shfl_down_sync(0xFFFFFFFFu, b.mn[k], off);
        r.mx[k] = __shfl_down_sync(0xFFFFFFFFu, b.mx[k], off);
    }
    return r;
}
__device__ __forceinline__ Box shfl_box_up(const Box& b, int off) {
    Box r;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        r.mn[k] = __shfl_up_sync(0xFFFFFFFFu, b.mn[k], off);
        r.mx[k] = __shfl_up_sync(0xFFFFFFFFu, b.mx[k], off);
    }
    return r;
}
__device__ __forceinline__ Box warp_union(Box b) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            b.mn[k] = fminf(b.mn[k], __shfl_xor_sync(0xFFFFFFFFu, b.mn[k], off));
            b.mx[k] = fmaxf(b.mx[k], __shfl_xor_sync(0xFFFFFFFFu, b.mx[k], off));
        }
    }
    return b;
}

// One warp per task: find_split on the three axes, axis choice, leaf / fallback rules, child boxes
// (binned_sah.rs:80-114, :174-247).  Lane b < 16 owns bin b.
struct SplitOut {  // warp-uniform result of sah_split_task
    bool split;
    uint32_t axis, split_index, nleft;
};
__device__ __forceinline__ SplitOut sah_split_task(const Task task, uint32_t t, int lane, const uint32_t* taskbins,
                                                   const float4* __restrict__ nodes, uint32_t max_leaf, uint32_t depth,
                                                   Decision* __restrict__ dec, uint4* __restrict__ counts) {
    const uint32_t n = task.end - task.begin;
    Box bin[3];
    uint32_t cnt[3];
    float best_cost[3];
    uint32_t best_count[3];
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
        bin[ax] = box_empty();
        cnt[ax] = 0;
        if (lane < kBins) {
            const uint32_t* g = taskbins + (ax * kBins + lane) * kBinWords;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                bin[ax].mn[k] = fkey_inv(g[k]);
                bin[ax].mx[k] = fkey_inv(g[3 + k]);
            }
            cnt[ax] = g[6];
        }
        // suffix: R[i] = union of bins[i..16), cntR[i]; right_cost[i] = half_area(R[i]) * cntR[i]
        Box R = bin[ax];
        uint32_t cR = cnt[ax];
#pragma unroll
        for (int off = 1; off < kBins; off <<= 1) {
            const Box o = shfl_box_down(R, off);
            const uint32_t oc = __shfl_down_sync(0xFFFFFFFFu, cR, off);
            if (lane + off < kBins) {
                R = box_union(R, o);
                cR += oc;
            }
        }
        const float right_cost = fmul(box_half_area(R), (float)cR);
        // prefix: L[i] = union of bins[0..i], cntL[i]
        Box L = bin[ax];
        uint32_t cL = cnt[ax];
#pragma unroll
        for (int off = 1; off < kBins; off <<= 1) {
            const Box o = shfl_box_up(L, off);
            const uint32_t oc = __shfl_up_sync(0xFFFFFFFFu, cL, off);
            if (lane >= off && lane < kBins) {
                L = box_union(L, o);
                cL += oc;
            }
        }
        const float rc_next = __shfl_down_sync(0xFFFFFFFFu, right_cost, 1);
        float cost = fadd(fmul(box_half_area(L), (float)cL), rc_next);
        // first strict minimum below f32::MAX, scanning i = 0..14 (binned_sah.rs:101-111)
        int bi = lane;
        if (!(lane < kBins - 1 && cost < FLT_MAX)) {
            cost = FLT_MAX;
            bi = kBins;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float oc = __shfl_xor_sync(0xFFFFFFFFu, cost, off);
            const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, off);
            if (oc < cost || (oc == cost && oi < bi)) {
                cost = oc;
                bi = oi;
            }
        }
        best_cost[ax] = cost;
        best_count[ax] = bi == kBins ? (uint32_t)kBins : (uint32_t)bi + 1u;
    }
    int best_axis = 0;
    if (best_cost[0] > best_cost[1]) best_axis = 1;
    if ((best_axis == 0 ? best_cost[0] : best_cost[1]) > best_cost[2]) best_axis = 2;
    auto pick_u = [&](const uint32_t* a, int ax) { return ax == 0 ? a[0] : (ax == 1 ? a[1] : a[2]); };
    auto pick_f = [&](const float* a, int ax) { return ax == 0 ? a[0] : (ax == 1 ? a[1] : a[2]); };
    uint32_t split_index = pick_u(best_count, best_axis);
    const Box nb = load_box(nodes, task.node);  // already padded by sah_prepare_kernel
    const float max_split_cost = fmul(box_half_area(nb), fsub((float)n, 1.0f));  // traversal_cost = 1.0
    bool do_split = true;
    if (pick_u(best_count, best_axis) == (uint32_t)kBins || pick_f(best_cost, best_axis) >= max_split_cost) {
        if (n > max_leaf) {
            // fallback: ~40 % median on the longest axis (binned_sah.rs:189-205)
            best_axis = box_longest_axis(nb);
            uint32_t cum = pick_u(cnt, best_axis);
#pragma unroll
            for (int off = 1; off < kBins; off <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, cum, off);
                if (lane >= off) cum += o;
            }
            const uint32_t need = (uint32_t)(((uint64_t)n * 2ull) / 5ull + 1ull);
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, lane < kBins - 1 && cum >= need);
            if (m) split_index = (uint32_t)__ffs(m);  // i + 1 of the first such bin
        } else {
            do_split = false;
        }
    }
    const uint32_t my_cnt = pick_u(cnt, best_axis);
    uint32_t nleft = (lane < (int)split_index) ? my_cnt : 0u;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) nleft += __shfl_xor_sync(0xFFFFFFFFu, nleft, off);
    if (nleft == 0 || nleft == n) do_split = false;  // one side empty -> leaf (binned_sah.rs:222, :277-281)
    const Box mine = best_axis == 0 ? bin[0] : (best_axis == 1 ? bin[1] : bin[2]);
    // quirk Q3: the left box always uses the SAH split count of the final axis (binned_sah.rs:232-235)
    const uint32_t left_count_q3 = pick_u(best_count, best_axis);
    const Box lb = warp_union(lane < (int)left_count_q3 && lane < kBins ? mine : box_empty());
    const Box rb = warp_union(lane >= (int)split_index && lane < kBins ? mine : box_empty());
    if (lane == 0) {
        Decision d;
        d.split = do_split ? 1u : 0u;
        d.axis = (uint32_t)best_axis;
        d.split_index = split_index;
        d.nleft = nleft;
        for (int k = 0; k < 3; k++) {
            d.lmn[k] = lb.mn[k];
            d.lmx[k] = lb.mx[k];
            d.rmn[k] = rb.mn[k];
            d.rmx[k] = rb.mx[k];
        }
        dec[t] = d;
        // children: leaf on entry (n <= 1 or depth cap) / small subtree (one warp finishes it) / next-level task
        uint4 c = make_uint4(d.split, 0u, 0u, 0u);
        if (do_split) {
            const bool cap = depth + 1 >= (uint32_t)kMaxDepth;
            const uint32_t nc[2] = {nleft, n - nleft};
            for (int k = 0; k < 2; k++) {
                if (nc[k] <= 1 || cap) continue;
                if (nc[k] <= kSmall) {
                    c.z += 1u;
                    c.w += 2u * nc[k] - 2u;  // node slots reserved for the subtree below this child
                } else {
                    c.y += 1u;
                }
            }
        }
        counts[t] = c;
    }
    return SplitOut{do_split, (uint32_t)best_axis, split_index, nleft};
}
// Launched for A_ub >= A warps (the host only knows an upper bound of the level's task count): warps beyond A zero
// their counts entry so that the fixed-size scan that follows is exact.
__global__ void __launch_bounds__(128) sah_split_kernel(const Task* __restrict__ tasks, uint32_t A_ub,
                                                        const TaskAux* __restrict__ aux, uint32_t* __restrict__ bins,
                                                        const float4* __restrict__ nodes,
                                                        uint32_t max_leaf, uint32_t depth, Decision* __restrict__ dec,
                                                        uint4* __restrict__ counts, const LevelState* __restrict__ state) {
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= A_ub) return;
    if (t >= state->A) {
        if (lane == 0) counts[t] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    const Task task = tasks[t];
    if (task.end - task.begin <= kWarpTask) return;  // sah_warp_task_kernel's
    uint32_t* tb = bins + (size_t)aux[t].slot * kTaskBinWords;
    sah_split_task(task, t, lane, tb, nodes, max_leaf, depth, dec, counts);
    // leave the bins of slot t clean for whichever task gets this index on a later level
    __syncwarp();
    for (int k = lane; k < kTaskBinWords; k += 32) {
        const int f = k % kBinWords;
        tb[k] = f < 3 ? fkey(1e34f) : (f < 6 ? fkey(-1e34f) : 0u);
    }
}

// Levels without span-class tasks skip sah_split_kernel, which also zeroes the counts of the slots beyond A.
__global__ void zero_counts_tail_kernel(uint4* __restrict__ counts, uint32_t A_ub, const LevelState* __restrict__ state) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < A_ub && t >= state->A) counts[t] = make_uint4(0u, 0u, 0u, 0u);
}
// Warp-class tasks (<= kWarpTask primitives): one warp does the whole node step of its task — it stages the task's index
// range in shared memory, fills the 3 x 16 bins there (no global atomics, no bins in HBM), evaluates the split and, if
// the node splits, writes the range back stably partitioned IN PLACE (it owns the range; ballot prefix sums give the
// destinations).  Such a task never shows up in the global partition pass, and once a level holds no span-class task
// any more the level loop consists of this kernel, the scan of the per-task counts and the emit kernel only.
constexpr int kWarpTaskWarps = 4;
// MERGED (the default, RTBVH_SAH_MERGE=0 turns it off): the same launch also plays sah_split_kernel's part for the
// span-class tasks of the level (after the bin pass) and zeroes the counts beyond A, so a level needs neither
// sah_split_kernel nor zero_counts_tail_kernel: one launch less per level, and the handful of latency-bound split warps
// (9 us for ONE task under ncu, profiles/r3b_build_ncu.md) run next to the warp-class tasks instead of before them.
template <bool MERGED>
__global__ void __launch_bounds__(kWarpTaskWarps * 32) sah_warp_task_kernel(const Task* __restrict__ tasks,
                                                                            uint32_t* __restrict__ idx,
                                                                            const TaskAux* __restrict__ aux,
                                                                            const float4* __restrict__ bb,
                                                                            const float* __restrict__ cen, uint32_t cstride,
                                                                            const float4* __restrict__ nodes, uint32_t max_leaf,
                                                                            uint32_t depth, Decision* __restrict__ dec,
                                                                            uint4* __restrict__ counts,
                                                                            const LevelState* __restrict__ state, uint32_t A_ub,
                                                                            uint32_t* __restrict__ bins) {
    __shared__ uint32_t sb[kWarpTaskWarps][kTaskBinWords];
    __shared__ uint32_t s_idx[kWarpTaskWarps][kWarpTask];   // the task's index range as read
    __shared__ uint32_t s_out[kWarpTaskWarps][kWarpTask];   // ... and partitioned
    __shared__ uint16_t s_pk[kWarpTaskWarps][kWarpTask];    // 3 x 4-bit bin per primitive
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x * kWarpTaskWarps + w;
    if (MERGED) {
        if (t >= A_ub) return;
        if (t >= state->A) {  // the scan over A_ub entries that follows must see zeros here
            if (lane == 0) counts[t] = make_uint4(0u, 0u, 0u, 0u);
            return;
        }
    } else if (t >= state->A) {
        return;
    }
    const Task task = tasks[t];
    const uint32_t n = task.end - task.begin;
    if (n > kWarpTask) {
        if (MERGED) {  // span-class task: its bins are in HBM (sah_bin_kernel ran before this launch)
            uint32_t* tb = bins + (size_t)aux[t].slot * kTaskBinWords;
            sah_split_task(task, t, lane, tb, nodes, max_leaf, depth, dec, counts);
            __syncwarp();
            for (int k = lane; k < kTaskBinWords; k += 32) {  // leave the bin block clean for its next owner
                const int f = k % kBinWords;
                tb[k] = f < 3 ? fkey(1e34f) : (f < 6 ? fkey(-1e34f) : 0u);
            }
        }
        return;
    }
    for (int k = lane; k < kTaskBinWords; k += 32) {
        const int f = k % kBinWords;
        sb[w][k] = f < 3 ? fkey(1e34f) : (f < 6 ? fkey(-1e34f) : 0u);
    }
    __syncwarp();
    const TaskAux a = aux[t];
    // (measured: four bin copies per warp to thin out the atomic replays made this kernel 1.4x SLOWER at 1 Mi triangles —
    // with ~150 primitives per task it is bound by the gather latency and the init/fold of the copies, not by replays)
    for (uint32_t j = lane; j < n; j += 32) {
        const uint32_t p = idx[task.begin + j];
        const Box box = load_box(bb, p);
        uint32_t packed = 0;
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
            const int b = bin_index(cen[(size_t)p * cstride + ax], a.k[ax], a.off[ax]);
            bin_accumulate(&sb[w][(ax * kBins + b) * kBinWords], box);
            packed |= (uint32_t)b << (4 * ax);
        }
        s_idx[w][j] = p;
        s_pk[w][j] = (uint16_t)packed;
    }
    __syncwarp();
    const SplitOut so = sah_split_task(task, t, lane, sb[w], nodes, max_leaf, depth, dec, counts);
    if (!so.split) return;
    // stable partition (binned_sah.rs:213-219 predicate): lefts to [0, nleft), rights behind them, original order kept
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t lbase = 0, rbase = so.nleft;
    for (uint32_t j0 = 0; j0 < n; j0 += 32) {
        const uint32_t j = j0 + lane;
        const bool valid = j < n;
        const bool left = valid && (((uint32_t)s_pk[w][valid ? j : 0] >> (4u * so.axis)) & 15u) < so.split_index;
        const uint32_t lm = __ballot_sync(0xFFFFFFFFu, left), vm = __ballot_sync(0xFFFFFFFFu, valid);
        const uint32_t rm = vm & ~lm;
        if (valid) s_out[w][left ? lbase + (uint32_t)__popc(lm & lt) : rbase + (uint32_t)__popc(rm & lt)] = s_idx[w][j];
        lbase += (uint32_t)__popc(lm);
        rbase += (uint32_t)__popc(rm);
    }
    __syncwarp();
    for (uint32_t j = lane; j < n; j += 32) idx[task.begin + j] = s_out[w][j];
}

// make_leaf (binned_sah.rs:134-138): second pad, left_first = begin, count = n
__device__ __forceinline__ void make_leaf(float4* nodes, uint32_t node, Box b, uint32_t begin, uint32_t n) {
    box_pad(b, kPad);
    store_node(nodes, node, b, (int)n, (int)begin);
}

__global__ void root_leaf_kernel(float4* nodes, uint32_t n) {
    Box b = load_box(nodes, 0);
    box_pad(b, kPad);
    make_leaf(nodes, 0, b, 0, n);
}

struct SmallTask {
    uint32_t node, begin, end, depth, node_base;
};

// Allocate the child pairs (level order), write inner nodes / leaves, emit next-level tasks (entered right here: box
// pad + binning transform, so the next level starts with its bin pass) and small subtrees.
// rank[t] = exclusive scan of counts: x pairs, y next-level tasks, z small subtrees, w node slots of small subtrees.
// The thread of the last task publishes the next level's state.
__device__ __forceinline__ void sah_emit_task(const uint32_t t, const LevelState cur, const uint4 r /* rank[t] */,
                                              const uint4 c /* counts[t] */, const Task* __restrict__ tasks,
                                              const Decision* __restrict__ dec, uint32_t depth, float4* nodes,
                                              Task* __restrict__ next_tasks, TaskAux* __restrict__ next_aux,
                                              SmallTask* __restrict__ small_tasks, PartTask* __restrict__ ptask,
                                              uint32_t* __restrict__ span_tasks, LevelState* __restrict__ state) {
    if (t == cur.A - 1) {
        state[depth + 1] = LevelState{r.y + c.y, cur.node_count + 2u * (r.x + c.x), cur.S + r.z + c.z, cur.small_slots + r.w + c.w};
    }
    const uint32_t node_base = cur.node_count, small_base = cur.S, small_node_base = cur.small_slots;
    const Task task = tasks[t];
    const Decision d = dec[t];
    const uint32_t n = task.end - task.begin;
    Box nb = load_box(nodes, task.node);
    if (!d.split) {
        make_leaf(nodes, task.node, nb, task.begin, n);
        ptask[t] = PartTask{task.begin, 0u, -1, -1, 0u, 0u, {0u, 0u}};
        return;
    }
    const uint32_t left = node_base + 2u * r.x;
    store_node(nodes, task.node, nb, -1, (int)left);
    const bool cap = depth + 1 >= (uint32_t)kMaxDepth;
    uint32_t next = r.y, small = small_base + r.z, small_nodes = small_node_base + r.w;
    const uint32_t mid = task.begin + d.nleft;
    Box cb[2] = {Box{{d.lmn[0], d.lmn[1], d.lmn[2]}, {d.lmx[0], d.lmx[1], d.lmx[2]}},
                 Box{{d.rmn[0], d.rmn[1], d.rmn[2]}, {d.rmx[0], d.rmx[1], d.rmx[2]}}};
    const uint32_t cbeg[2] = {task.begin, mid}, cend[2] = {mid, task.end};
    int32_t child[2] = {-1, -1};
    for (int k = 0; k < 2; k++) {
        const uint32_t nc = cend[k] - cbeg[k];
        if (nc <= 1 || cap) {  // leaf on entry: pad (run) + pad (make_leaf), binned_sah.rs:133-143
            box_pad(cb[k], kPad);
            make_leaf(nodes, left + k, cb[k], cbeg[k], nc);
        } else if (nc <= kSmall) {
            store_node(nodes, left + k, cb[k], 0, 0);
            small_tasks[small] = SmallTask{left + k, cbeg[k], cend[k], depth + 1, small_nodes};
            small++;
            small_nodes += 2u * nc - 2u;
        } else {
            // which bin block a span-class task gets is irrelevant to the result (scratch), so a counter will do
            const uint32_t slot = nc > kWarpTask ? atomicAdd(&span_tasks[depth + 1], 1u) : 0u;
            task_enter(nodes, left + k, cb[k], &next_aux[next], slot);
            next_tasks[next] = Task{left + k, cbeg[k], cend[k]};
            child[k] = pt_encode(next++, nc);
        }
    }
    ptask[t] = PartTask{task.begin, d.nleft, child[0], child[1], 4u * d.axis, d.split_index, {0u, 0u}};
}
__global__ void sah_emit_kernel(const Task* __restrict__ tasks, uint32_t A_ub, const Decision* __restrict__ dec,
                                const uint4* __restrict__ rank, const uint4* __restrict__ counts, uint32_t depth, float4* nodes,
                                Task* __restrict__ next_tasks, TaskAux* __restrict__ next_aux,
                                SmallTask* __restrict__ small_tasks,
                                PartTask* __restrict__ ptask /* per task: what the partition pass needs */,
                                uint32_t* __restrict__ span_tasks /* [depth + 1]: bin blocks handed out for the next level */,
                                LevelState* __restrict__ state /* [depth] in, [depth + 1] out */) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const LevelState cur = state[depth];
    if (cur.A == 0) {
        if (t == 0) state[depth + 1] = cur;
        return;
    }
    if (t >= cur.A) return;
    sah_emit_task(t, cur, rank[t], counts[t], tasks, dec, depth, nodes, next_tasks, next_aux, small_tasks, ptask, span_tasks, state);
}
struct Uint4Sum {
    __device__ __forceinline__ uint4 operator()(const uint4& a, const uint4& b) const {
        return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
};
// Levels with at most kScanEmitTasks tasks (the top ten): ONE block scans the per-task counts and emits, instead of
// cub::DeviceScan (two launches) + sah_emit_kernel — the scan of <= 1024 entries is a block scan (the default;
// RTBVH_SAH_SCANEMIT=0 turns it off).  Same ranks, same emit code.
constexpr int kScanEmitTasks = 1024;
__global__ void __launch_bounds__(kScanEmitTasks, 1) sah_scan_emit_kernel(const Task* __restrict__ tasks, uint32_t A_ub,
                                                                       const Decision* __restrict__ dec,
                                                                       const uint4* __restrict__ counts, uint32_t depth,
                                                                       float4* nodes, Task* __restrict__ next_tasks,
                                                                       TaskAux* __restrict__ next_aux,
                                                                       SmallTask* __restrict__ small_tasks,
                                                                       PartTask* __restrict__ ptask,
                                                                       uint32_t* __restrict__ span_tasks,
                                                                       LevelState* __restrict__ state) {
    __shared__ uint4 s_warp[kScanEmitTasks / 32];
    const uint32_t t = threadIdx.x;
    const int lane = (int)(t & 31u), w = (int)(t >> 5);
    const LevelState cur = state[depth];  // block-uniform
    if (cur.A == 0) {
        if (t == 0) state[depth + 1] = cur;
        return;
    }
    const uint4 c = (t < cur.A && t < A_ub) ? counts[t] : make_uint4(0u, 0u, 0u, 0u);
    // exclusive block scan: shuffles inside the warps, the 32 warp totals scanned by warp 0
    const Uint4Sum add;
    uint4 inc = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint4 o = make_uint4(__shfl_up_sync(0xFFFFFFFFu, inc.x, off), __shfl_up_sync(0xFFFFFFFFu, inc.y, off),
                                   __shfl_up_sync(0xFFFFFFFFu, inc.z, off), __shfl_up_sync(0xFFFFFFFFu, inc.w, off));
        if (lane >= off) inc = add(inc, o);
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        const uint4 mine = s_warp[lane];
        uint4 ws = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint4 o = make_uint4(__shfl_up_sync(0xFFFFFFFFu, ws.x, off), __shfl_up_sync(0xFFFFFFFFu, ws.y, off),
                                       __shfl_up_sync(0xFFFFFFFFu, ws.z, off), __shfl_up_sync(0xFFFFFFFFu, ws.w, off));
            if (lane >= off) ws = add(ws, o);
        }
        s_warp[lane] = make_uint4(ws.x - mine.x, ws.y - mine.y, ws.z - mine.z, ws.w - mine.w);  // exclusive warp offsets
    }
    __syncthreads();
    const uint4 base = s_warp[w];
    const uint4 r = make_uint4(base.x + inc.x - c.x, base.y + inc.y - c.y, base.z + inc.z - c.z, base.w + inc.w - c.w);
    if (t >= cur.A) return;
    sah_emit_task(t, cur, r, c, tasks, dec, depth, nodes, next_tasks, next_aux, small_tasks, ptask, span_tasks, state);
}
bool level_flag(const char* name) {  // the level-loop variants above default to on; NAME=0 selects the older path
    const char* e = std::getenv(name);
    return !(e && e[0] == '0');
}

// ---- small subtrees: one warp runs BinnedSahBuildTask::run for every node below a <= 32-primitive task -----
// Same decisions as the level kernels, computed without bins in memory: for split position s on an axis the
// left box / count is the union / number of the primitives whose bin is < s — identical to the prefix of the
// reference's bins because min/max/count do not depend on the order of accumulation.
struct SmallEntry {
    uint32_t node, b, e, depth;
    Box box;
};
constexpr int kSmallWarps = 4;
#ifndef RTB_SMALL_MINBLOCKS
#define RTB_SMALL_MINBLOCKS 8  // 64 registers: the kernel is latency-bound, resident warps matter more than registers
#endif
__device__ __forceinline__ Box shfl_box(const Box& b, int src) {
    Box r;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        r.mn[k] = __shfl_sync(0xFFFFFFFFu, b.mn[k], src);
        r.mx[k] = __shfl_sync(0xFFFFFFFFu, b.mx[k], src);
    }
    return r;
}
// One candidate split (axis ax, bins < sidx go left) over the primitives [b, e) of the warp's shared-memory tile:
// lo = {min xyz, packed bin ids}, hi = {max xyz, primitive id}.
__device__ __forceinline__ float small_candidate(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t b,
                                                 uint32_t e, int ax, uint32_t sidx, Box& L, Box& R) {
    L = box_empty();
    R = box_empty();
    uint32_t cl = 0, cr = 0;
    for (uint32_t j = b; j < e; j++) {
        const float4 l4 = lo[j], h4 = hi[j];
        const Box pb{{l4.x, l4.y, l4.z}, {h4.x, h4.y, h4.z}};
        const uint32_t bj = (__float_as_uint(l4.w) >> (8 * ax)) & 0xFFu;
        if (bj < sidx) {
            L = box_union(L, pb);
            cl++;
        } else {
            R = box_union(R, pb);
            cr++;
        }
    }
    return fadd(fmul(box_half_area(L), (float)cl), fmul(box_half_area(R), (float)cr));
}
__device__ __forceinline__ unsigned inmask2(uint32_t b) { return 3u << b; }  // the two lanes of a 2-primitive node
// find_split over the 15 candidates held by one aligned 16-lane group (lane & 15 = split - 1; lane 15 of the group holds
// no candidate): first strict minimum below f32::MAX in order s = 1..15 (binned_sah.rs:95-111).  All lanes of the group
// return the group's (cost, count).
__device__ __forceinline__ void small_argmin16(float cost, int lane, float& best_cost, uint32_t& best_count) {
    uint32_t cnt = (uint32_t)(lane & 15) + 1u;
    if ((lane & 15) == 15 || !(cost < FLT_MAX)) {
        cost = FLT_MAX;
        cnt = (uint32_t)kBins;
    }
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) {
        const float oc = __shfl_xor_sync(0xFFFFFFFFu, cost, off);
        const uint32_t oi = __shfl_xor_sync(0xFFFFFFFFu, cnt, off);
        if (oc < cost || (oc == cost && oi < cnt)) {
            cost = oc;
            cnt = oi;
        }
    }
    best_cost = cost;
    best_count = cnt;
}
__global__ void __launch_bounds__(kSmallWarps * 32, RTB_SMALL_MINBLOCKS) sah_small_kernel(const SmallTask* __restrict__ tasks, uint32_t S,
                                                                      uint32_t* __restrict__ idx, const float4* __restrict__ bb,
                                                                      const float* __restrict__ cen, uint32_t cstride,
                                                                      float4* nodes, uint32_t max_leaf,
                                                                      uint32_t* __restrict__ used_nodes) {
    __shared__ float4 s_lo[kSmallWarps][32];  // min xyz | packed bin ids of the node being split
    __shared__ float4 s_hi[kSmallWarps][32];  // max xyz | primitive id
    __shared__ float s_cen[kSmallWarps][32][3];
    __shared__ SmallEntry s_stack[kSmallWarps][34];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x * kSmallWarps + w;
    if (t >= S) return;
    const SmallTask task = tasks[t];
    const uint32_t n = task.end - task.begin;
    if ((uint32_t)lane < n) {
        const uint32_t p = idx[task.begin + lane];
        const float4 l4 = bb[(size_t)p * 2], h4 = bb[(size_t)p * 2 + 1];
        s_lo[w][lane] = make_float4(l4.x, l4.y, l4.z, 0.f);
        s_hi[w][lane] = make_float4(h4.x, h4.y, h4.z, __uint_as_float(p));
        for (int k = 0; k < 3; k++) s_cen[w][lane][k] = cen[(size_t)p * cstride + k];
    }
    int sp = 0;
    if (lane == 0) s_stack[w][0] = SmallEntry{task.node, 0u, n, task.depth, load_box(nodes, task.node)};
    sp = 1;
    uint32_t next_free = task.node_base;
    __syncwarp();
    while (sp > 0) {
        const SmallEntry e = s_stack[w][--sp];
        __syncwarp();
        const uint32_t nn = e.e - e.b;
        Box nb = e.box;
        box_pad(nb, kPad);  // entry pad (binned_sah.rs:133)
        if (nn <= 1 || e.depth >= (uint32_t)kMaxDepth) {
            if (lane == 0) make_leaf(nodes, e.node, nb, task.begin + e.b, nn);
            continue;
        }
        float k3[3], off3[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            k3[k] = fmul(fdiv(1.0f, fsub(nb.mx[k], nb.mn[k])), (float)kBins);
            off3[k] = fmul(-nb.mn[k], k3[k]);
        }
        const bool in = (uint32_t)lane >= e.b && (uint32_t)lane < e.e;
        uint32_t mybins = 0;
        if (in) {
#pragma unroll
            for (int k = 0; k < 3; k++) mybins |= (uint32_t)bin_index(s_cen[w][lane][k], k3[k], off3[k]) << (8 * k);
            s_lo[w][lane].w = __uint_as_float(mybins);
        }
        __syncwarp();
        if (nn == 2) {
            // Two primitives (about a third of all nodes): the 45 candidates collapse to one.  On an axis where the two
            // bins differ every split position between them has the same cost half_area(A) + half_area(B) (first one:
            // min bin + 1); where they are equal one side is always empty and the cost is NaN (never chosen).  The cost is the
            // same on every usable axis, so find_split's strict comparisons keep the first usable axis.  Anything else
            // (no usable axis, cost not below the leaf cost -> fallback / leaf rules) takes the general path below.
            const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, mybins, (int)e.b), b1 = __shfl_sync(0xFFFFFFFFu, mybins, (int)e.b + 1);
            int ax = -1;
#pragma unroll
            for (int k = 2; k >= 0; k--)
                if (((b0 >> (8 * k)) & 0xFFu) != ((b1 >> (8 * k)) & 0xFFu)) ax = k;
            if (ax >= 0) {
                const float4 l0 = s_lo[w][e.b], h0 = s_hi[w][e.b], l1 = s_lo[w][e.b + 1], h1 = s_hi[w][e.b + 1];
                const Box A = box_union(box_empty(), Box{{l0.x, l0.y, l0.z}, {h0.x, h0.y, h0.z}});
                const Box B = box_union(box_empty(), Box{{l1.x, l1.y, l1.z}, {h1.x, h1.y, h1.z}});
                const float cost = fadd(fmul(box_half_area(A), 1.0f), fmul(box_half_area(B), 1.0f));
                const float max_cost2 = fmul(box_half_area(nb), fsub(2.0f, 1.0f));
                if (cost < FLT_MAX && cost < max_cost2) {
                    const bool first_left = ((b0 >> (8 * ax)) & 0xFFu) < ((b1 >> (8 * ax)) & 0xFFu);
                    __syncwarp();
                    if (!first_left && in) {  // stable partition of two: the right-going primitive sits first -> swap
                        const uint32_t other = (uint32_t)lane == e.b ? e.b + 1 : e.b;
                        const float4 olo = s_lo[w][other], ohi = s_hi[w][other];
                        float oc[3];
                        for (int k = 0; k < 3; k++) oc[k] = s_cen[w][other][k];
                        __syncwarp(inmask2(e.b));
                        s_lo[w][lane] = olo;
                        s_hi[w][lane] = ohi;
                        for (int k = 0; k < 3; k++) s_cen[w][lane][k] = oc[k];
                    }
                    const uint32_t left = next_free;
                    next_free += 2;
                    if (lane == 0) {
                        store_node(nodes, e.node, nb, -1, (int)left);
                        Box bl = first_left ? A : B, br = first_left ? B : A;
                        box_pad(bl, kPad);  // leaf on entry: entry pad + make_leaf pad (binned_sah.rs:133-143)
                        box_pad(br, kPad);
                        make_leaf(nodes, left, bl, task.begin + e.b, 1u);
                        make_leaf(nodes, left + 1, br, task.begin + e.b + 1, 1u);
                    }
                    __syncwarp();
                    continue;
                }
            }
        }
        // 45 candidates (3 axes x split positions 1..15) in two rounds: lanes 0-15 / 16-31 hold axes 0 / 1, then lanes
        // 0-15 hold axis 2; lane 15 of a group holds nothing.  Every lane keeps the child boxes of its candidates.
        Box L1, R1, L2, R2;
        float c1 = FLT_MAX, c2 = FLT_MAX;
        L1 = R1 = L2 = R2 = box_empty();
        float best_cost[3];
        uint32_t best_count[3];
        // Up to 10 primitives (most nodes): only split positions right behind an occupied bin can differ from their
        // predecessor — position s and the next occupied-bin boundary below it give the same partition, hence the same
        // cost, and find_split keeps the first of equals — so lane (axis, j) evaluates the ONE candidate s = bin_j + 1:
        // 3 nn <= 30 lanes, one round instead of two, and the per-axis "first strict minimum" is two integer warp
        // reductions on order-preserving keys (minimum cost, then the smallest s among its holders).
        const bool compact = nn <= 10;
        int cax = 3;          // compact mode: this lane's axis (3: no candidate)
        uint32_t cs = kBins;  // ... and split position
        if (compact) {
            if ((uint32_t)lane < 3u * nn) {
                cax = lane / (int)nn;
                const uint32_t j = e.b + (uint32_t)lane % nn;
                cs = ((__float_as_uint(s_lo[w][j].w) >> (8 * cax)) & 0xFFu) + 1u;
            }
            const bool cvalid = cax < 3 && cs < (uint32_t)kBins;
            if (cvalid) c1 = small_candidate(s_lo[w], s_hi[w], e.b, e.e, cax, cs, L1, R1);
            const uint32_t key = (cvalid && c1 < FLT_MAX) ? fkey(c1) : 0xFFFFFFFFu;
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const uint32_t m = __reduce_min_sync(0xFFFFFFFFu, cax == a ? key : 0xFFFFFFFFu);
                const uint32_t sa = __reduce_min_sync(0xFFFFFFFFu, (cax == a && key == m && m != 0xFFFFFFFFu) ? cs : (uint32_t)kBins);
                best_cost[a] = m == 0xFFFFFFFFu ? FLT_MAX : fkey_inv(m);
                best_count[a] = m == 0xFFFFFFFFu ? (uint32_t)kBins : sa;
            }
        } else {
            if ((lane & 15) != 15) c1 = small_candidate(s_lo[w], s_hi[w], e.b, e.e, lane >> 4, (uint32_t)(lane & 15) + 1u, L1, R1);
            if (lane < 15) c2 = small_candidate(s_lo[w], s_hi[w], e.b, e.e, 2, (uint32_t)lane + 1u, L2, R2);
            float bc1, bc2;
            uint32_t bn1, bn2;
            small_argmin16(c1, lane, bc1, bn1);
            small_argmin16(c2, lane, bc2, bn2);
            best_cost[0] = __shfl_sync(0xFFFFFFFFu, bc1, 0);
            best_count[0] = __shfl_sync(0xFFFFFFFFu, bn1, 0);
            best_cost[1] = __shfl_sync(0xFFFFFFFFu, bc1, 16);
            best_count[1] = __shfl_sync(0xFFFFFFFFu, bn1, 16);
            best_cost[2] = __shfl_sync(0xFFFFFFFFu, bc2, 0);
            best_count[2] = __shfl_sync(0xFFFFFFFFu, bn2, 0);
        }
        int best_axis = 0;
        if (best_cost[0] > best_cost[1]) best_axis = 1;
        if ((best_axis == 0 ? best_cost[0] : best_cost[1]) > best_cost[2]) best_axis = 2;
        uint32_t split_index = best_axis == 0 ? best_count[0] : (best_axis == 1 ? best_count[1] : best_count[2]);
        const float axis_cost = best_axis == 0 ? best_cost[0] : (best_axis == 1 ? best_cost[1] : best_cost[2]);
        const float max_split_cost = fmul(box_half_area(nb), fsub((float)nn, 1.0f));
        bool do_split = true, fallback = false;
        if (split_index == (uint32_t)kBins || axis_cost >= max_split_cost) {
            if (nn > max_leaf) {  // fallback (binned_sah.rs:189-205)
                fallback = true;
                best_axis = box_longest_axis(nb);
                // per-bin counts on that axis: lane b < 16 counts the primitives of bin b
                uint32_t cnt = 0;
                if (lane < kBins)
                    for (uint32_t j = e.b; j < e.e; j++)
                        cnt += (((__float_as_uint(s_lo[w][j].w) >> (8 * best_axis)) & 0xFFu) == (uint32_t)lane) ? 1u : 0u;
                uint32_t cum = cnt;
#pragma unroll
                for (int o = 1; o < kBins; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, cum, o);
                    if (lane >= o) cum += v;
                }
                const uint32_t need = (uint32_t)(((uint64_t)nn * 2ull) / 5ull + 1ull);
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, lane < kBins - 1 && cum >= need);
                if (m) split_index = (uint32_t)__ffs(m);
            } else {
                do_split = false;
            }
        }
        const uint32_t mybin = (mybins >> (8 * best_axis)) & 0xFFu;
        const bool goes_left = in && mybin < split_index;
        const uint32_t lmask = __ballot_sync(0xFFFFFFFFu, goes_left);
        const uint32_t inmask = __ballot_sync(0xFFFFFFFFu, in);
        const uint32_t nleft = (uint32_t)__popc(lmask);
        if (nleft == 0 || nleft == nn) do_split = false;
        if (!do_split) {
            if (lane == 0) make_leaf(nodes, e.node, nb, task.begin + e.b, nn);
            continue;
        }
        // child boxes.  Regular split: the lane that evaluated the winning candidate already holds them.  Fallback:
        // quirk Q3 — the left box uses the SAH split count of the final axis, the right box the fallback index.
        Box lb, rb;
        if (!fallback && compact) {
            const uint32_t holders = __ballot_sync(0xFFFFFFFFu, cax == best_axis && cs == split_index);
            const int src = __ffs((int)holders) - 1;  // any holder has the same boxes (same partition, same order of unions)
            lb = shfl_box(L1, src);
            rb = shfl_box(R1, src);
        } else if (!fallback) {
            const int src = best_axis < 2 ? 16 * best_axis + (int)split_index - 1 : (int)split_index - 1;
            lb = shfl_box(best_axis < 2 ? L1 : L2, src);
            rb = shfl_box(best_axis < 2 ? R1 : R2, src);
        } else {
            const uint32_t q3 = best_axis == 0 ? best_count[0] : (best_axis == 1 ? best_count[1] : best_count[2]);
            const float4 l4 = s_lo[w][lane], h4 = s_hi[w][lane];
            const Box mine = in ? Box{{l4.x, l4.y, l4.z}, {h4.x, h4.y, h4.z}} : box_empty();
            lb = warp_union((in && mybin < q3) ? mine : box_empty());
            rb = warp_union((in && mybin >= split_index) ? mine : box_empty());
        }
        // stable partition of [b, e) inside the warp
        const uint32_t lt = (1u << lane) - 1u;
        uint32_t dest = (uint32_t)lane;
        if (in) dest = goes_left ? e.b + (uint32_t)__popc(lmask & lt) : e.b + nleft + (uint32_t)__popc((inmask & ~lmask) & lt);
        float4 mylo = make_float4(0.f, 0.f, 0.f, 0.f), myhi = mylo;
        float myc[3] = {0.f, 0.f, 0.f};
        if (in) {
            mylo = s_lo[w][lane];
            myhi = s_hi[w][lane];
            for (int k = 0; k < 3; k++) myc[k] = s_cen[w][lane][k];
        }
        __syncwarp();
        if (in) {
            s_lo[w][dest] = mylo;
            s_hi[w][dest] = myhi;
            for (int k = 0; k < 3; k++) s_cen[w][dest][k] = myc[k];
        }
        const uint32_t left = next_free;
        next_free += 2;
        // children that are leaves on entry (one primitive, or the depth cap) are finalised here: entry pad +
        // make_leaf pad (binned_sah.rs:133-143); only real subtrees go onto the stack
        const bool cap = e.depth + 1 >= (uint32_t)kMaxDepth;
        const bool r_leaf = (nn - nleft) <= 1 || cap, l_leaf = nleft <= 1 || cap;
        if (lane == 0) {
            store_node(nodes, e.node, nb, -1, (int)left);
            int k = sp;
            if (r_leaf) {
                Box b2 = rb;
                box_pad(b2, kPad);
                make_leaf(nodes, left + 1, b2, task.begin + e.b + nleft, nn - nleft);
            } else {
                s_stack[w][k++] = SmallEntry{left + 1, e.b + nleft, e.e, e.depth + 1, rb};
            }
            if (l_leaf) {
                Box b2 = lb;
                box_pad(b2, kPad);
                make_leaf(nodes, left, b2, task.begin + e.b, nleft);
            } else {
                s_stack[w][k++] = SmallEntry{left, e.b, e.b + nleft, e.depth + 1, lb};
            }
        }
        sp += (r_leaf ? 0 : 1) + (l_leaf ? 0 : 1);
        __syncwarp();
    }
    if ((uint32_t)lane < n) idx[task.begin + lane] = __float_as_uint(s_hi[w][lane].w);
    if (lane == 0) used_nodes[t] = next_free - task.node_base;
}

// ---- EXPERIMENTAL (RTBVH_SAH_SMALL=ls, default off): the same subtrees, LEVEL-SYNCHRONOUS inside the warp.  Measured
// (profiles/r3e_ls_ab.json, AB_LS=1 scripts/partition_ab.py): trees byte-identical to sah_small_kernel's on every scene, but
// 1.7x SLOWER per build at 1 Mi triangles — every pass walks as far as the largest open node of its depth, so small
// siblings wait for big ones.  Kept as a verified-correct base for a variant with per-node walk lengths.
// sah_small_kernel is issue-bound at ~530 warp instructions per inner node (profiles/r3b_build_ncu.md) because the warp
// works on one node at a time and most nodes hold 2-8 primitives.  Here lane j holds tile position j, every position
// belongs to exactly one OPEN node of the current subtree depth (a contiguous range [b, e)), and all open nodes of a depth
// go through the same instructions:
//   * candidates: lane j evaluates, per axis, the one split position s = bin_j + 1 of its own node (the rule of the
//     compact path of sah_small_kernel, valid for any node size: a position that is not right behind an occupied bin
//     repeats its predecessor's partition and can never be find_split's first strict minimum) by walking its node's range
//     in the shared-memory tile;
//   * per-node minima, counts and child boxes: warp reductions over the node's lanes (__match_any_sync on b);
//   * stable partition of all open nodes with one ballot;
//   * nodes are collected in a shared-memory table and written at the end with sah_small_kernel's numbering (child pairs in
//     DFS pre-order of the splitting nodes: rank = # splitting nodes with a smaller b, or the same b and a smaller depth).
// Decisions, boxes, index order and node numbering are meant to be identical to sah_small_kernel's, bit for bit.
constexpr int kLsWarps = 4;
constexpr int kLsNodes = 64;  // <= 31 inner + 32 leaf nodes per subtree
__device__ __forceinline__ uint32_t seg_min_key(unsigned peers, bool in, float v) { return __reduce_min_sync(peers, in ? fkey(v) : fkey(1e34f)); }
__device__ __forceinline__ uint32_t seg_max_key(unsigned peers, bool in, float v) { return __reduce_max_sync(peers, in ? fkey(v) : fkey(-1e34f)); }
__global__ void __launch_bounds__(kLsWarps * 32, 4) sah_small_ls_kernel(const SmallTask* __restrict__ tasks, uint32_t S,
                                                                        uint32_t* __restrict__ idx, const float4* __restrict__ bb,
                                                                        const float* __restrict__ cen, uint32_t cstride,
                                                                        float4* nodes, uint32_t max_leaf,
                                                                        uint32_t* __restrict__ used_nodes) {
    __shared__ float4 s_lo[kLsWarps][32];       // min xyz | packed bin ids (this level)
    __shared__ float4 s_hi[kLsWarps][32];       // max xyz | primitive id
    __shared__ float s_cen[kLsWarps][32][3];
    __shared__ float4 t_lo[kLsWarps][kLsNodes];  // node table: min xyz | count (-1: inner)
    __shared__ float4 t_hi[kLsWarps][kLsNodes];  //             max xyz | left_first of a leaf
    __shared__ uint32_t t_meta[kLsWarps][kLsNodes];  // parent | side << 8 | b << 16 | level << 24
    __shared__ uint32_t t_pair[kLsWarps][kLsNodes];  // inner nodes: global index of the child pair
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x * kLsWarps + w;
    if (t >= S) return;
    const SmallTask task = tasks[t];
    const uint32_t n = task.end - task.begin;
    if ((uint32_t)lane < n) {
        const uint32_t p = idx[task.begin + lane];
        const float4 l4 = bb[(size_t)p * 2], h4 = bb[(size_t)p * 2 + 1];
        s_lo[w][lane] = make_float4(l4.x, l4.y, l4.z, 0.f);
        s_hi[w][lane] = make_float4(h4.x, h4.y, h4.z, __uint_as_float(p));
        for (int k = 0; k < 3; k++) s_cen[w][lane][k] = cen[(size_t)p * cstride + k];
    }
    // per-lane view of the open node this tile position belongs to (uniform over the node's lanes)
    bool open = (uint32_t)lane < n;
    uint32_t sb = 0, se = n, entry = 0;
    Box nbox = load_box(nodes, task.node);  // as left by sah_emit_kernel: the un-padded child box
    uint32_t n_entries = 1, level = 0;
    if (lane == 0) t_meta[w][0] = 0u;  // the root: parent / side unused, b = 0, level = 0
    __syncwarp();
    const unsigned lt = (1u << lane) - 1u;
    while (__any_sync(0xFFFFFFFFu, open)) {
        const uint32_t depth = task.depth + level;  // of every open node of this pass
        const uint32_t nn = se - sb;
        Box nb = nbox;
        box_pad(nb, kPad);  // entry pad (binned_sah.rs:133)
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, open ? sb : 0x100u + (uint32_t)lane);
        // an open node always has >= 2 primitives and depth < kMaxDepth: leaves on entry are finalised by their parent,
        // and a small task is created with >= 2 primitives below the depth cap
        uint32_t bins3 = 0;
        if (open) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float kk = fmul(fdiv(1.0f, fsub(nb.mx[k], nb.mn[k])), (float)kBins);
                const float off = fmul(-nb.mn[k], kk);
                bins3 |= (uint32_t)bin_index(s_cen[w][lane][k], kk, off) << (8 * k);
            }
            s_lo[w][lane].w = __uint_as_float(bins3);
        }
        __syncwarp();
        // ---- candidates: s = bin + 1 per axis, cost over the node's range --------------------------------------------
        uint32_t cs[3], cl[3], cr[3];
        Box L[3], R[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            cs[a] = ((bins3 >> (8 * a)) & 0xFFu) + 1u;
            cl[a] = cr[a] = 0u;
            L[a] = R[a] = box_empty();
        }
        const uint32_t maxlen = __reduce_max_sync(0xFFFFFFFFu, open ? nn : 0u);
        for (uint32_t d = 0; d < maxlen; d++) {
            const uint32_t j = sb + d;
            if (open && j < se) {
                const float4 l4 = s_lo[w][j], h4 = s_hi[w][j];
                const Box pb{{l4.x, l4.y, l4.z}, {h4.x, h4.y, h4.z}};
                const uint32_t bj = __float_as_uint(l4.w);
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    if (((bj >> (8 * a)) & 0xFFu) < cs[a]) {
                        L[a] = box_union(L[a], pb);
                        cl[a]++;
                    } else {
                        R[a] = box_union(R[a], pb);
                        cr[a]++;
                    }
                }
            }
        }
        // ---- find_split per axis: first strict minimum below f32::MAX in order s = 1..15 (binned_sah.rs:95-111) ------
        float best_cost[3];
        uint32_t best_count[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const bool valid = open && cs[a] < (uint32_t)kBins;
            const float c = fadd(fmul(box_half_area(L[a]), (float)cl[a]), fmul(box_half_area(R[a]), (float)cr[a]));
            const uint32_t key = (valid && c < FLT_MAX) ? fkey(c) : 0xFFFFFFFFu;
            const uint32_t m = __reduce_min_sync(peers, key);
            const uint32_t sa = __reduce_min_sync(peers, (key == m && m != 0xFFFFFFFFu) ? cs[a] : (uint32_t)kBins);
            best_cost[a] = m == 0xFFFFFFFFu ? FLT_MAX : fkey_inv(m);
            best_count[a] = m == 0xFFFFFFFFu ? (uint32_t)kBins : sa;
        }
        int best_axis = 0;
        if (best_cost[0] > best_cost[1]) best_axis = 1;
        if ((best_axis == 0 ? best_cost[0] : best_cost[1]) > best_cost[2]) best_axis = 2;
        uint32_t split_index = best_axis == 0 ? best_count[0] : (best_axis == 1 ? best_count[1] : best_count[2]);
        const float axis_cost = best_axis == 0 ? best_cost[0] : (best_axis == 1 ? best_cost[1] : best_cost[2]);
        const float max_split_cost = fmul(box_half_area(nb), fsub((float)nn, 1.0f));
        bool do_split = open, fallback = false;
        if (open && (split_index == (uint32_t)kBins || axis_cost >= max_split_cost)) {
            if (nn > max_leaf) {
                fallback = true;  // ~40 % median on the longest axis (binned_sah.rs:189-205)
                best_axis = box_longest_axis(nb);
            } else {
                do_split = false;
            }
        }
        {   // fallback index: first bin i < 15 whose cumulative count reaches floor(2n/5) + 1.  cl[axis] of lane j is the
            // number of primitives of the node with bin <= bin_j; empty bins repeat their predecessor's count, so the
            // first bin that qualifies is an occupied one.  All lanes take part in the reduction (segments are disjoint).
            const uint32_t need = (uint32_t)(((uint64_t)nn * 2ull) / 5ull + 1ull);
            const uint32_t mycl = best_axis == 0 ? cl[0] : (best_axis == 1 ? cl[1] : cl[2]);
            const uint32_t mys = best_axis == 0 ? cs[0] : (best_axis == 1 ? cs[1] : cs[2]);
            const uint32_t cand = (fallback && mys < (uint32_t)kBins && mycl >= need) ? mys : (uint32_t)kBins;
            const uint32_t fs = __reduce_min_sync(peers, cand);
            if (fallback && fs < (uint32_t)kBins) split_index = fs;
        }
        const uint32_t mybin = (bins3 >> (8 * best_axis)) & 0xFFu;
        const bool goes_left = do_split && mybin < split_index;
        const uint32_t lmask = __ballot_sync(0xFFFFFFFFu, goes_left) & peers;
        const uint32_t nleft = (uint32_t)__popc(lmask);
        if (nleft == 0 || nleft == nn) do_split = false;  // one side empty -> leaf (binned_sah.rs:222, :277-281)
        // ---- child boxes: unions over the node's lanes.  Regular split: bins < split_index | >= split_index (what the
        // winning candidate accumulated); fallback, quirk Q3: the left box uses the SAH split count of the final axis.
        const uint32_t q_left = fallback ? (best_axis == 0 ? best_count[0] : (best_axis == 1 ? best_count[1] : best_count[2])) : split_index;
        const bool in_l = do_split && mybin < q_left, in_r = do_split && mybin >= split_index;
        const float4 mylo = s_lo[w][lane], myhi = s_hi[w][lane];
        Box lb, rb;
        lb.mn[0] = fkey_inv(seg_min_key(peers, in_l, mylo.x));
        lb.mn[1] = fkey_inv(seg_min_key(peers, in_l, mylo.y));
        lb.mn[2] = fkey_inv(seg_min_key(peers, in_l, mylo.z));
        lb.mx[0] = fkey_inv(seg_max_key(peers, in_l, myhi.x));
        lb.mx[1] = fkey_inv(seg_max_key(peers, in_l, myhi.y));
        lb.mx[2] = fkey_inv(seg_max_key(peers, in_l, myhi.z));
        rb.mn[0] = fkey_inv(seg_min_key(peers, in_r, mylo.x));
        rb.mn[1] = fkey_inv(seg_min_key(peers, in_r, mylo.y));
        rb.mn[2] = fkey_inv(seg_min_key(peers, in_r, mylo.z));
        rb.mx[0] = fkey_inv(seg_max_key(peers, in_r, myhi.x));
        rb.mx[1] = fkey_inv(seg_max_key(peers, in_r, myhi.y));
        rb.mx[2] = fkey_inv(seg_max_key(peers, in_r, myhi.z));
        // ---- node table: this node, and two entries per splitting node (numbered by the order of the nodes' b) --------
        const bool leader = open && (uint32_t)lane == sb;
        const uint32_t split_leaders = __ballot_sync(0xFFFFFFFFu, leader && do_split);
        const uint32_t child0 = n_entries + 2u * (uint32_t)__popc(split_leaders & ((1u << sb) - 1u));  // left child entry
        const bool cap = depth + 1 >= (uint32_t)kMaxDepth;
        const uint32_t nright = nn - nleft;
        const bool l_leaf = nleft <= 1 || cap, r_leaf = nright <= 1 || cap;
        if (leader) {
            if (!do_split) {  // make_leaf(nb): second pad, left_first = begin, count = n
                Box b2 = nb;
                box_pad(b2, kPad);
                t_lo[w][entry] = make_float4(b2.mn[0], b2.mn[1], b2.mn[2], __int_as_float((int)nn));
                t_hi[w][entry] = make_float4(b2.mx[0], b2.mx[1], b2.mx[2], __int_as_float((int)(task.begin + sb)));
            } else {
                t_lo[w][entry] = make_float4(nb.mn[0], nb.mn[1], nb.mn[2], __int_as_float(-1));
                t_hi[w][entry] = make_float4(nb.mx[0], nb.mx[1], nb.mx[2], __int_as_float(-1));
                t_meta[w][child0] = entry | (0u << 8) | (sb << 16) | ((level + 1u) << 24);
                t_meta[w][child0 + 1] = entry | (1u << 8) | ((sb + nleft) << 16) | ((level + 1u) << 24);
                if (l_leaf) {  // leaf on entry: entry pad + make_leaf pad (binned_sah.rs:133-143)
                    Box b2 = lb;
                    box_pad(b2, kPad);
                    box_pad(b2, kPad);
                    t_lo[w][child0] = make_float4(b2.mn[0], b2.mn[1], b2.mn[2], __int_as_float((int)nleft));
                    t_hi[w][child0] = make_float4(b2.mx[0], b2.mx[1], b2.mx[2], __int_as_float((int)(task.begin + sb)));
                }
                if (r_leaf) {
                    Box b2 = rb;
                    box_pad(b2, kPad);
                    box_pad(b2, kPad);
                    t_lo[w][child0 + 1] = make_float4(b2.mn[0], b2.mn[1], b2.mn[2], __int_as_float((int)nright));
                    t_hi[w][child0 + 1] = make_float4(b2.mx[0], b2.mx[1], b2.mx[2], __int_as_float((int)(task.begin + sb + nleft)));
                }
            }
        }
        n_entries += 2u * (uint32_t)__popc(split_leaders);
        // ---- stable partition of every splitting node, all at once ---------------------------------------------------
        const uint32_t inmask = __ballot_sync(0xFFFFFFFFu, open && do_split) & peers;
        uint32_t dest = (uint32_t)lane;
        if (open && do_split)
            dest = goes_left ? sb + (uint32_t)__popc(lmask & lt) : sb + nleft + (uint32_t)__popc((inmask & ~lmask) & lt);
        float myc[3];
#pragma unroll
        for (int k = 0; k < 3; k++) myc[k] = s_cen[w][lane][k];
        __syncwarp();
        if ((uint32_t)lane < n) {
            s_lo[w][dest] = mylo;
            s_hi[w][dest] = myhi;
#pragma unroll
            for (int k = 0; k < 3; k++) s_cen[w][dest][k] = myc[k];
        }
        __syncwarp();
        // ---- next pass: tile position `lane` now belongs to the left or the right child of the node it was in ----------
        if (open && do_split) {
            const bool is_left = (uint32_t)lane < sb + nleft;
            open = is_left ? !l_leaf : !r_leaf;
            entry = is_left ? child0 : child0 + 1u;
            nbox = is_left ? lb : rb;
            const uint32_t mid = sb + nleft;
            if (is_left) se = mid; else sb = mid;
        } else {
            open = false;
        }
        level++;
    }
    __syncwarp();
    // ---- numbering and write-out --------------------------------------------------------------------------------------
    // child pair of a splitting node = node_base + 2 * (# splitting nodes before it in DFS pre-order); in pre-order node A
    // precedes node B iff A.b < B.b, or A.b == B.b and A is the shallower one (then A is an ancestor of B).
    uint32_t inner_count = 0;
    for (uint32_t i = (uint32_t)lane; i < n_entries; i += 32) {
        if (__float_as_int(t_lo[w][i].w) >= 0) continue;  // a leaf
        const uint32_t mi = t_meta[w][i], bi = (mi >> 16) & 0xFFu, li = mi >> 24;
        uint32_t rank = 0;
        for (uint32_t j = 0; j < n_entries; j++) {
            if (j == i || __float_as_int(t_lo[w][j].w) >= 0) continue;
            const uint32_t mj = t_meta[w][j], bj = (mj >> 16) & 0xFFu, lj = mj >> 24;
            if (bj < bi || (bj == bi && lj < li)) rank++;
        }
        t_pair[w][i] = task.node_base + 2u * rank;
    }
    for (uint32_t j = 0; j < n_entries; j++) inner_count += __float_as_int(t_lo[w][j].w) < 0 ? 1u : 0u;
    __syncwarp();
    for (uint32_t i = (uint32_t)lane; i < n_entries; i += 32) {
        const uint32_t mi = t_meta[w][i];
        const uint32_t gid = i == 0 ? task.node : t_pair[w][mi & 0xFFu] + ((mi >> 8) & 1u);
        float4 lo = t_lo[w][i], hi = t_hi[w][i];
        if (__float_as_int(lo.w) < 0) hi.w = __int_as_float((int)t_pair[w][i]);  // inner: left_first = its child pair
        nodes[(size_t)gid * 2] = lo;
        nodes[(size_t)gid * 2 + 1] = hi;
    }
    if ((uint32_t)lane < n) idx[task.begin + lane] = __float_as_uint(s_hi[w][lane].w);
    if (lane == 0) used_nodes[t] = 2u * inner_count;
}
bool small_ls_mode() {
    static const bool v = [] {
        const char* e = std::getenv("RTBVH_SAH_SMALL");
        return e && std::string(e) == "ls";
    }();
    return v;
}

// ---- compaction of the node array when small subtrees used fewer slots than reserved ---------------
// waste_prefix[t] = exclusive scan of (reserved - used) over the small tasks (ordered by node_base).
__device__ __forceinline__ uint32_t small_remap(uint32_t node, uint32_t level_nodes, const SmallTask* __restrict__ tasks,
                                                const uint32_t* __restrict__ waste_prefix, uint32_t S) {
    if (node < level_nodes) return node;
    uint32_t lo = 0, hi = S;  // last task with node_base <= node
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (tasks[mid].node_base <= node) lo = mid; else hi = mid;
    }
    return node - waste_prefix[lo];
}
__global__ void small_rebase_kernel(SmallTask* tasks, uint32_t S, uint32_t level_nodes) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < S) tasks[t].node_base += level_nodes;
}
__global__ void small_waste_kernel(const SmallTask* __restrict__ tasks, const uint32_t* __restrict__ used, uint32_t S,
                                   uint32_t* __restrict__ waste) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S) return;
    waste[t] = 2u * (tasks[t].end - tasks[t].begin) - 2u - used[t];
}
__global__ void compact_nodes_kernel(const float4* __restrict__ in, float4* __restrict__ out, uint32_t level_nodes,
                                     const SmallTask* __restrict__ tasks, const uint32_t* __restrict__ used,
                                     const uint32_t* __restrict__ waste_prefix, uint32_t S, uint32_t total_slots) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total_slots) return;
    uint32_t dst = i;
    if (i >= level_nodes) {
        uint32_t lo = 0, hi = S;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (tasks[mid].node_base <= i) lo = mid; else hi = mid;
        }
        if (i - tasks[lo].node_base >= used[lo]) return;  // reserved but unused slot
        dst = i - waste_prefix[lo];
    }
    float4 a = in[(size_t)i * 2], b = in[(size_t)i * 2 + 1];
    const int count = __float_as_int(a.w), left = __float_as_int(b.w);
    if (count < 0 && left >= 0) b.w = __int_as_float((int)small_remap((uint32_t)left, level_nodes, tasks, waste_prefix, S));
    out[(size_t)dst * 2] = a;
    out[(size_t)dst * 2 + 1] = b;
}

// ---- stable partition of every task's index range: ONE segmented scan (cub::DeviceScan::InclusiveScanByKey over the
// task keys) whose input iterator evaluates the partition predicate on the fly (binned_sah.rs:213-219) and whose output
// iterator scatters the index to its place in the ping-pong buffer: no flag array, no rank array, no extra kernels.
struct FlagSum {
    uint32_t sum, flag;  // inclusive count of left-going positions of the segment; this position's own predicate
};
struct FlagSumOp {
    __host__ __device__ FlagSum operator()(const FlagSum& a, const FlagSum& b) const { return FlagSum{a.sum + b.sum, b.flag}; }
};
struct PartitionParams {
    const uint32_t* idx;
    const int32_t* pos_task;
    const PartTask* ptask;
    const uint16_t* binidx;
    uint32_t* idx_out;
    int32_t* pos_task_out;
};
struct PartitionFlagIn {
    const PartitionParams* p;
    __device__ FlagSum operator()(uint32_t i) const {
        const int32_t t = pt_task(p->pos_task[i]);
        uint32_t f = 0;
        if (t >= 0) {
            const uint2 k = *reinterpret_cast<const uint2*>(&p->ptask[t].shift);  // {shift, split_index}
            if (k.y != 0u) f = (((uint32_t)p->binidx[i] >> k.x) & 15u) < k.y ? 1u : 0u;
        }
        return FlagSum{f, f};
    }
};
struct PartitionScatterOut {
    const PartitionParams* p;
    uint32_t base;
    struct Ref {
        const PartitionParams* p;
        uint32_t i;
        __device__ const Ref& operator=(const FlagSum& v) const {
            const int32_t t = pt_task(p->pos_task[i]);
            uint4 q = make_uint4(0u, 0u, 0u, 0u);
            if (t >= 0) q = *reinterpret_cast<const uint4*>(&p->ptask[t]);  // {begin, nleft, child_left, child_right}
            if (t < 0 || q.y == 0u) {  // finished position, or a task that became a leaf this level
                p->idx_out[i] = p->idx[i];
                p->pos_task_out[i] = -1;
            } else {
                const uint32_t r = v.sum - v.flag;
                const bool left = v.flag != 0;
                const uint32_t dest = left ? q.x + r : q.x + q.y + (i - q.x - r);
                p->idx_out[dest] = p->idx[i];
                p->pos_task_out[dest] = (int32_t)(left ? q.z : q.w);
            }
            return *this;
        }
    };
    using iterator_category = std::random_access_iterator_tag;
    using value_type = FlagSum;
    using difference_type = ptrdiff_t;
    using pointer = void;
    using reference = Ref;
    __host__ __device__ PartitionScatterOut operator+(difference_type k) const { return PartitionScatterOut{p, base + (uint32_t)k}; }
    __host__ __device__ PartitionScatterOut& operator+=(difference_type k) {
        base += (uint32_t)k;
        return *this;
    }
    __host__ __device__ Ref operator[](difference_type k) const { return Ref{p, base + (uint32_t)k}; }
    __host__ __device__ Ref operator*() const { return Ref{p, base}; }
};
using PartitionFlagIter = cub::TransformInputIterator<FlagSum, PartitionFlagIn, cub::CountingInputIterator<uint32_t>>;

// ---- the same stable partition without the chained scan (the default; RTBVH_SAH_PARTITION=cub selects the CUB pass
// above).  ncu of a 1 Mi build (profiles/r3b_build_ncu.md): DeviceScanByKey needs 35 us (+ 4 us init) per level for 10 MB
// of traffic — the latency of its decoupled look-back chain, not bandwidth.  Here every block owns a fixed chunk of
// kPartChunk index positions and two short kernels replace the scan:
//   count   : tail_left[c] = left-going positions of chunk c that belong to the task of the chunk's LAST position (the
//             only task that can continue into chunk c + 1);
//   scatter : the carry of the task running through the chunk's FIRST position is the sum of tail_left over the chunks
//             since that task's begin (they lie in the task entirely, except the first one, whose tail is exactly the
//             task's part); inside the chunk a block-wide segmented scan of the predicate gives every position its rank,
//             and the destination rule is PartitionScatterOut's, verbatim.
// Same inputs, same outputs (idx_out, pos_task_out), bit for bit: a stable partition has one result.
constexpr int kPartBlock = 256, kPartItems = 8;
constexpr uint32_t kPartChunk = kPartBlock * kPartItems;
constexpr uint32_t kPartMaxChunks = 8192;  // the carry is a block reduction over the task's chunks: bound it (n <= 16 Mi)

struct SegSum {
    uint32_t sum, head;  // head: a segment (task) starts at or before this position, inside the scanned range
};
struct SegSumOp {
    __device__ __forceinline__ SegSum operator()(const SegSum& a, const SegSum& b) const {
        return SegSum{b.head ? b.sum : a.sum + b.sum, a.head | b.head};
    }
};
__device__ __forceinline__ uint32_t part_flag(const PartitionParams* p, uint32_t i, int32_t t) {
    if (t < 0) return 0u;
    const uint2 k = *reinterpret_cast<const uint2*>(&p->ptask[t].shift);  // {shift, split_index}
    return (k.y != 0u && (((uint32_t)p->binidx[i] >> k.x) & 15u) < k.y) ? 1u : 0u;
}
__global__ void __launch_bounds__(kPartBlock) part_count_kernel(const PartitionParams* __restrict__ p, uint32_t n,
                                                                uint32_t* __restrict__ tail_left) {
    using Reduce = cub::BlockReduce<uint32_t, kPartBlock>;
    __shared__ typename Reduce::TempStorage tmp;
    const uint32_t base = blockIdx.x * kPartChunk;
    const uint32_t end = min(n, base + kPartChunk);
    const int32_t t_last = pt_task(p->pos_task[end - 1]);  // block-uniform
    uint32_t cnt = 0;
    if (t_last >= 0) {
        const uint2 k = *reinterpret_cast<const uint2*>(&p->ptask[t_last].shift);
        if (k.y != 0u) {
            const uint32_t first = max(base, p->ptask[t_last].begin);  // the task's positions inside this chunk: [first, end)
            for (uint32_t i = first + threadIdx.x; i < end; i += kPartBlock)
                cnt += (((uint32_t)p->binidx[i] >> k.x) & 15u) < k.y ? 1u : 0u;
        }
    }
    const uint32_t total = Reduce(tmp).Sum(cnt);
    if (threadIdx.x == 0) tail_left[blockIdx.x] = total;
}
__global__ void __launch_bounds__(kPartBlock) part_scatter_kernel(const PartitionParams* __restrict__ p, uint32_t n,
                                                                  const uint32_t* __restrict__ tail_left) {
    using Reduce = cub::BlockReduce<uint32_t, kPartBlock>;
    using Scan = cub::BlockScan<SegSum, kPartBlock>;
    __shared__ union {
        typename Reduce::TempStorage reduce;
        typename Scan::TempStorage scan;
    } tmp;
    __shared__ uint32_t s_carry;
    const uint32_t base = blockIdx.x * kPartChunk;
    const uint32_t end = min(n, base + kPartChunk);
    // carry of the task that runs through the first position of the chunk
    const int32_t t0 = pt_task(p->pos_task[base]);  // block-uniform
    uint32_t part = 0;
    if (t0 >= 0) {
        const uint32_t tb = p->ptask[t0].begin;
        if (tb < base)
            for (uint32_t c = tb / kPartChunk + threadIdx.x; c < blockIdx.x; c += kPartBlock) part += tail_left[c];
    }
    const uint32_t carry = Reduce(tmp.reduce).Sum(part);
    if (threadIdx.x == 0) s_carry = carry;
    __syncthreads();  // also separates the two uses of the shared temp storage
    // segmented inclusive scan of the predicate over the chunk (blocked arrangement: kPartItems consecutive positions per thread)
    const uint32_t i0 = base + threadIdx.x * kPartItems;
    int32_t task[kPartItems];
    uint32_t flag[kPartItems];
    SegSum v[kPartItems];
    int32_t prev = (i0 > base && i0 < end) ? pt_task(p->pos_task[i0 - 1]) : 0;
#pragma unroll
    for (int k = 0; k < kPartItems; k++) {
        const uint32_t i = i0 + k;
        if (i < end) {
            task[k] = pt_task(p->pos_task[i]);
            flag[k] = part_flag(p, i, task[k]);
            v[k] = SegSum{flag[k], (i > base && task[k] != prev) ? 1u : 0u};
            prev = task[k];
        } else {  // padding behind the end of the array: its own segments, never read
            task[k] = -1;
            flag[k] = 0u;
            v[k] = SegSum{0u, 1u};
        }
    }
    Scan(tmp.scan).InclusiveScan(v, v, SegSumOp());
    const uint32_t first_carry = s_carry;
#pragma unroll
    for (int k = 0; k < kPartItems; k++) {
        const uint32_t i = i0 + k;
        if (i >= end) continue;
        const int32_t t = task[k];
        uint4 q = make_uint4(0u, 0u, 0u, 0u);
        if (t >= 0) q = *reinterpret_cast<const uint4*>(&p->ptask[t]);  // {begin, nleft, child_left, child_right}
        if (t < 0 || q.y == 0u) {  // finished position, or a task that became a leaf this level
            p->idx_out[i] = p->idx[i];
            p->pos_task_out[i] = -1;
        } else {
            const uint32_t incl = v[k].sum + (v[k].head ? 0u : first_carry);  // no segment start since the chunk's first position
            const uint32_t r = incl - flag[k];
            const bool left = flag[k] != 0u;
            const uint32_t dest = left ? q.x + r : q.x + q.y + (i - q.x - r);
            p->idx_out[dest] = p->idx[i];
            p->pos_task_out[dest] = (int32_t)(left ? q.z : q.w);
        }
    }
}
// Default for builds of up to kPartMaxChunks chunks; RTBVH_SAH_PARTITION=cub selects the scan-by-key pass (A/B:
// scripts/partition_ab.py — byte-identical trees on every scene, 1 Mi soup 2.44 -> 2.25 ms, 3 Mi 5.81 -> 5.45 ms).
bool partition_block_mode() {
    static const bool v = [] {
        const char* e = std::getenv("RTBVH_SAH_PARTITION");
        return !(e && std::string(e) == "cub");
    }();
    return v;
}

// =================================================================================================
// LOCB
// =================================================================================================
__device__ __forceinline__ uint32_t part1by2(uint32_t v) {  // == morton_split for 10-bit inputs (morton.rs:10-25)
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
// MortonEncoder::new + encode (morton.rs:36-61); world = [min xyz, max xyz] incl. pad
__global__ void morton_kernel(const float* __restrict__ cen, uint32_t cstride, uint32_t n, const float* __restrict__ world,
                              uint32_t* __restrict__ codes, uint32_t* __restrict__ idx) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t g[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float w2g = fmul(1024.0f, fdiv(1.0f, fsub(world[3 + k], world[k])));
        const float off = fmul(-world[k], w2g);
        const float p = fadd(fmul(cen[(size_t)i * cstride + k], w2g), off);
        const int v = __float2int_rz(p);  // Rust `as i32`: truncate, saturate, NaN -> 0
        g[k] = (uint32_t)min(1023, max(v, 0));
    }
    codes[i] = part1by2(g[0]) | (part1by2(g[1]) << 1) | (part1by2(g[2]) << 2);
    idx[i] = i;
}
// leaves in Morton order (locb.rs:289-294)
__global__ void locb_leaves_kernel(const float4* __restrict__ bb, const uint32_t* __restrict__ idx, uint32_t n,
                                   float4* __restrict__ nodes, uint32_t begin) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Box b = load_box(bb, idx[i]);
    box_pad(b, kPad);
    store_node(nodes, begin + i, b, 1, (int)i);
}
// nearest neighbour within the search window (locb.rs:93-142): ascending j, strict <
__global__ void locb_nn_kernel(const float4* __restrict__ in, uint32_t begin, uint32_t end, uint32_t* __restrict__ nb) {
    const uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    const uint32_t sb = (i > begin + kRadius) ? i - kRadius : begin;  // search_range, locb.rs:36-45
    const uint32_t se = min(i + kRadius + 1, end);
    const Box me = load_box(in, i);
    float best = FLT_MAX;
    uint32_t best_j = 0xFFFFFFFFu;
    for (uint32_t j = sb; j < se; j++) {
        if (j == i) continue;
        const float d = box_half_area(box_union(me, load_box(in, j)));
        if (d < best) {
            best = d;
            best_j = j;
        }
    }
    nb[i] = best_j;
}
// locb.rs:158-167
__global__ void locb_flag_kernel(const uint32_t* __restrict__ nb, uint32_t begin, uint32_t end, uint32_t* __restrict__ merged) {
    const uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    const uint32_t j = nb[i];
    merged[i] = (j < end && j >= begin && i < j && nb[j] == i) ? 1u : 0u;
}
// layout math + merge / copy (locb.rs:178-242).  counts[0] = merged_count (device), filled by the scan.
__global__ void locb_write_kernel(const float4* __restrict__ in, float4* __restrict__ out, const uint32_t* __restrict__ nb,
                                  const uint32_t* __restrict__ P /* inclusive scan of merged */, uint32_t begin, uint32_t end,
                                  uint32_t previous_end) {
    const uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t merged_count = P[end - 1];
    const uint32_t children_begin = end - 2u * merged_count;
    const uint32_t unmerged_begin = children_begin - (end - begin - merged_count);
    if (i < end) {
        const uint32_t j = nb[i];
        const bool mutual = j >= begin && j < end && nb[j] == i;
        if (mutual) {
            if (i < j) {
                const uint32_t parent = unmerged_begin + (j - begin) - P[j];
                const uint32_t first_child = children_begin + (P[i] - 1u) * 2u;
                const Box u = box_union(load_box(in, j), load_box(in, i));
                store_node(out, parent, u, -1, (int)first_child);
                out[(size_t)first_child * 2] = in[(size_t)i * 2];
                out[(size_t)first_child * 2 + 1] = in[(size_t)i * 2 + 1];
                out[(size_t)first_child * 2 + 2] = in[(size_t)j * 2];
                out[(size_t)first_child * 2 + 3] = in[(size_t)j * 2 + 1];
            }
        } else {
            const uint32_t dst = unmerged_begin + (i - begin) - P[i];
            out[(size_t)dst * 2] = in[(size_t)i * 2];
            out[(size_t)dst * 2 + 1] = in[(size_t)i * 2 + 1];
        }
    } else if (i < previous_end) {  // carry the pairs placed by the previous iteration (locb.rs:242)
        out[(size_t)i * 2] = in[(size_t)i * 2];
        out[(size_t)i * 2 + 1] = in[(size_t)i * 2 + 1];
    }
}

// The tail of the clustering: once a range holds <= kLocbTail clusters an iteration is a handful of microseconds of work
// behind five launches and a host round trip (58 iterations at 1 Mi triangles, ~35 of them this small).  One block runs all
// remaining iterations back to back — the same four phases, separated by block barriers, on the same global arrays — and
// reports how many it ran (the parity tells the host which node buffer holds the tree).
constexpr uint32_t kLocbTail = 2048;
constexpr int kLocbTailThreads = 1024;
// (everything the block re-reads after one of its own barriers goes through L2: __ldcg)
__device__ __forceinline__ Box tail_box(const float4* nodes, size_t i) {
    const float4 lo = __ldcg(&nodes[i * 2]), hi = __ldcg(&nodes[i * 2 + 1]);
    return Box{{lo.x, lo.y, lo.z}, {hi.x, hi.y, hi.z}};
}
__global__ void __launch_bounds__(kLocbTailThreads) locb_tail_kernel(float4* bufA, float4* bufB, uint32_t* nb, uint32_t* P,
                                                                     uint32_t begin, uint32_t end, uint32_t previous_end,
                                                                     uint32_t* out_state /* [0] iterations, [1] error */) {
    typedef cub::BlockScan<uint32_t, kLocbTailThreads> BlockScan;
    __shared__ typename BlockScan::TempStorage scan_tmp;
    float4* cur = bufA;
    float4* other = bufB;
    uint32_t iters = 0, err = 0;
    const uint32_t tid = threadIdx.x;
    while (end - begin > 1) {
        // nearest neighbour within the search window (locb.rs:93-142): ascending j, strict <
        for (uint32_t i = begin + tid; i < end; i += kLocbTailThreads) {
            const uint32_t sb = (i > begin + kRadius) ? i - kRadius : begin;
            const uint32_t se = min(i + kRadius + 1, end);
            const Box me = tail_box(cur, i);
            float best = FLT_MAX;
            uint32_t best_j = 0xFFFFFFFFu;
            for (uint32_t j = sb; j < se; j++) {
                if (j == i) continue;
                const float d = box_half_area(box_union(me, tail_box(cur, j)));
                if (d < best) {
                    best = d;
                    best_j = j;
                }
            }
            nb[i] = best_j;
        }
        __syncthreads();
        // mutual pairs (locb.rs:158-167) and their inclusive scan
        uint32_t carry = 0;
        for (uint32_t base = begin; base < end; base += kLocbTailThreads) {
            const uint32_t i = base + tid;
            uint32_t f = 0;
            if (i < end) {
                const uint32_t j = __ldcg(&nb[i]);
                f = (j < end && j >= begin && i < j && __ldcg(&nb[j]) == i) ? 1u : 0u;
            }
            uint32_t incl, agg;
            BlockScan(scan_tmp).InclusiveSum(f, incl, agg);
            if (i < end) P[i] = carry + incl;
            carry += agg;
            __syncthreads();
        }
        const uint32_t merged_count = carry;
        if (merged_count == 0) {  // no mutual pair (degenerate input): the multi-kernel loop reports the same
            err = 1;
            break;
        }
        // layout math + merge / copy (locb.rs:178-242)
        const uint32_t children_begin = end - 2u * merged_count;
        const uint32_t unmerged_begin = children_begin - (end - begin - merged_count);
        for (uint32_t i = begin + tid; i < previous_end; i += kLocbTailThreads) {
            if (i < end) {
                const uint32_t j = __ldcg(&nb[i]);
                const bool mutual = j >= begin && j < end && __ldcg(&nb[j]) == i;
                if (mutual) {
                    if (i < j) {
                        const uint32_t parent = unmerged_begin + (j - begin) - __ldcg(&P[j]);
                        const uint32_t first_child = children_begin + (__ldcg(&P[i]) - 1u) * 2u;
                        const Box u = box_union(tail_box(cur, j), tail_box(cur, i));
                        store_node(other, parent, u, -1, (int)first_child);
                        other[(size_t)first_child * 2] = __ldcg(&cur[(size_t)i * 2]);
                        other[(size_t)first_child * 2 + 1] = __ldcg(&cur[(size_t)i * 2 + 1]);
                        other[(size_t)first_child * 2 + 2] = __ldcg(&cur[(size_t)j * 2]);
                        other[(size_t)first_child * 2 + 3] = __ldcg(&cur[(size_t)j * 2 + 1]);
                    }
                } else {
                    const uint32_t dst = unmerged_begin + (i - begin) - __ldcg(&P[i]);
                    other[(size_t)dst * 2] = __ldcg(&cur[(size_t)i * 2]);
                    other[(size_t)dst * 2 + 1] = __ldcg(&cur[(size_t)i * 2 + 1]);
                }
            } else {  // carry the pairs placed by the previous iteration (locb.rs:242)
                other[(size_t)i * 2] = __ldcg(&cur[(size_t)i * 2]);
                other[(size_t)i * 2 + 1] = __ldcg(&cur[(size_t)i * 2 + 1]);
            }
        }
        __syncthreads();
        float4* t = cur;
        cur = other;
        other = t;
        previous_end = end;
        begin = unmerged_begin;
        end = children_begin;
        iters++;
    }
    if (tid == 0) {
        out_state[0] = iters;
        out_state[1] = err;
    }
}

// =================================================================================================
// Collapse (merge_nodes) and refit
// =================================================================================================
__device__ __forceinline__ int node_count_of(const float4* nodes, int i) { return __float_as_int(nodes[(size_t)i * 2].w); }
__device__ __forceinline__ int node_left_of(const float4* nodes, int i) { return __float_as_int(nodes[(size_t)i * 2 + 1].w); }

__global__ void parents_kernel(const float4* __restrict__ nodes, uint32_t n_nodes, int32_t* __restrict__ parent) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    if (i == 0) parent[0] = -1;
    const int c = node_count_of(nodes, i), l = node_left_of(nodes, i);
    if (c < 0 && l >= 0 && (uint32_t)l + 1 < n_nodes) {
        parent[l] = (int32_t)i;
        parent[l + 1] = (int32_t)i;
    }
}
// depth by walking up (reads only); m-roots = inner nodes at even depth (every second level of
// merge_nodes' recursion).
__global__ void mroot_flag_kernel(const float4* __restrict__ nodes, uint32_t n_nodes, const int32_t* __restrict__ parent,
                                  uint8_t* __restrict__ is_mroot) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const bool inner = node_count_of(nodes, i) < 0 && node_left_of(nodes, i) >= 0;
    uint32_t depth = 0;
    for (int p = parent[i]; p >= 0; p = parent[p]) depth++;
    is_mroot[i] = (inner && (depth & 1u) == 0u) ? 1 : 0;
}
// sub[i] = number of m-roots in the subtree of i, bottom-up: every leaf climbs, the second child to
// arrive at a node finishes it (one arrival counter per node, no contended counters).
__global__ void mroot_sub_kernel(const float4* __restrict__ nodes, uint32_t n_nodes, const int32_t* __restrict__ parent,
                                 const uint8_t* __restrict__ is_mroot, uint32_t* sub, uint32_t* __restrict__ arrived) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    if (!(node_count_of(nodes, i) >= 0 || node_left_of(nodes, i) < 0)) return;  // start from leaves (and invalid nodes)
    sub[i] = 0;
    int cur = (int)i;
    for (;;) {
        const int p = parent[cur];
        if (p < 0) break;
        __threadfence();
        if (atomicAdd(&arrived[p], 1u) == 0u) break;
        const int l = node_left_of(nodes, p);
        sub[p] = __ldcg(&sub[l]) + __ldcg(&sub[l + 1]) + (uint32_t)is_mroot[p];
        cur = p;
    }
}
// pre-order (depth-first, slot order) index of every m-root == the pool_ptr numbering of merge_nodes
__global__ void mroot_index_kernel(const float4* __restrict__ nodes, uint32_t n_nodes, const int32_t* __restrict__ parent,
                                   const uint8_t* __restrict__ is_mroot, const uint32_t* __restrict__ sub,
                                   uint32_t* __restrict__ mindex) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes || !is_mroot[i]) return;
    uint32_t pre = 0;
    int cur = (int)i;
    while (parent[cur] >= 0) {
        const int mid = parent[cur];     // the binary child of the m-parent that `cur` hangs under
        const int mp = parent[mid];      // the m-parent (even depth)
        const int L = node_left_of(nodes, mp);
        // m-children of mp in slot order: children of L, then children of L+1 (inner ones only)
        pre += 1;                        // mp itself precedes its subtree
        for (int side = 0; side < 2; side++) {
            const int x = L + side;
            if (node_count_of(nodes, x) >= 0) continue;  // leaf child: no grandchildren
            const int g0 = node_left_of(nodes, x);
            for (int k = 0; k < 2; k++) {
                const int g = g0 + k;
                if (g == cur) goto counted;
                pre += sub[g];           // 0 for leaves
            }
        }
    counted:
        cur = mp;
    }
    mindex[i] = pre;
}
// merge_nodes body for one m-root (mbvh_node.rs:297-411)
__global__ void collapse_emit_kernel(const float4* __restrict__ nodes, uint32_t n_nodes, const uint8_t* __restrict__ is_mroot,
                                     const uint32_t* __restrict__ mindex, float4* __restrict__ mnodes) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes || !is_mroot[i]) return;
    const Box pb = load_box(nodes, i);
    Box sb[4] = {pb, pb, pb, pb};
    int child[4] = {-1, -1, -1, -1}, count[4] = {-1, -1, -1, -1};
    const int L = node_left_of(nodes, i);
    for (int side = 0; side < 2; side++) {
        const int x = L + side;
        const int s0 = side * 2;
        const int xc = node_count_of(nodes, x), xl = node_left_of(nodes, x);
        if (xl < 0) continue;
        if (xc >= 0) {  // direct leaf child: one slot, keeps the parent's box (quirk Q6)
            child[s0] = xl;
            count[s0] = xc;
        } else {
            for (int k = 0; k < 2; k++) {
                const int g = xl + k;
                const int gc = node_count_of(nodes, g), gl = node_left_of(nodes, g);
                sb[s0 + k] = load_box(nodes, g);
                if (gc >= 0) {
                    child[s0 + k] = gl;
                    count[s0 + k] = gc;
                } else {
                    child[s0 + k] = (int)mindex[g];
                }
            }
        }
    }
    float4* m = mnodes + (size_t)mindex[i] * 8;
    m[0] = make_float4(sb[0].mn[0], sb[1].mn[0], sb[2].mn[0], sb[3].mn[0]);
    m[1] = make_float4(sb[0].mx[0], sb[1].mx[0], sb[2].mx[0], sb[3].mx[0]);
    m[2] = make_float4(sb[0].mn[1], sb[1].mn[1], sb[2].mn[1], sb[3].mn[1]);
    m[3] = make_float4(sb[0].mx[1], sb[1].mx[1], sb[2].mx[1], sb[3].mx[1]);
    m[4] = make_float4(sb[0].mn[2], sb[1].mn[2], sb[2].mn[2], sb[3].mn[2]);
    m[5] = make_float4(sb[0].mx[2], sb[1].mx[2], sb[2].mx[2], sb[3].mx[2]);
    m[6] = make_float4(__int_as_float(child[0]), __int_as_float(child[1]), __int_as_float(child[2]), __int_as_float(child[3]));
    m[7] = make_float4(__int_as_float(count[0]), __int_as_float(count[1]), __int_as_float(count[2]), __int_as_float(count[3]));
}

// Bvh::refit (bvh.rs:176-205), bottom-up with one arrival counter per inner node; the topology
// fields are kept (the reference's `self.nodes[i].bounds = aabb` would zero them — see DESIGN.md).
__global__ void refit_kernel(float4* nodes, uint32_t n_nodes, const int32_t* __restrict__ parent,
                             const uint32_t* __restrict__ indices, const float4* __restrict__ new_bb,
                             uint32_t* __restrict__ arrived) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const int c = node_count_of(nodes, i), l = node_left_of(nodes, i);
    if (c < 0 || l < 0) return;  // start from valid leaves only
    Box b = box_empty();
    for (int k = 0; k < c; k++) b = box_union(b, load_box(new_bb, indices[l + k]));
    box_pad(b, kPad);
    store_node(nodes, i, b, c, l);
    int cur = (int)i;
    for (;;) {
        const int p = parent[cur];
        if (p < 0) break;
        __threadfence();
        if (atomicAdd(&arrived[p], 1u) == 0u) break;  // first child to arrive: the sibling will finish the parent
        const int pl = node_left_of(nodes, p);
        // both children are complete (the sibling's stores precede its fence + atomic); bypass L1
        const float4 a0 = __ldcg(nodes + (size_t)pl * 2), a1 = __ldcg(nodes + (size_t)pl * 2 + 1);
        const float4 b0 = __ldcg(nodes + (size_t)pl * 2 + 2), b1 = __ldcg(nodes + (size_t)pl * 2 + 3);
        Box u = box_union(Box{{a0.x, a0.y, a0.z}, {a1.x, a1.y, a1.z}}, Box{{b0.x, b0.y, b0.z}, {b1.x, b1.y, b1.z}});
        u = box_union(box_empty(), u);
        box_pad(u, kPad);
        store_node(nodes, p, u, -1, pl);
        cur = p;
    }
}

// ---- builder workspace ------------------------------------------------------------------------------
// Every scratch and result buffer of a build is carved out of one device arena per host thread (bump pointer, reset at
// the start of each builder entry point, grown — never shrunk — when a build needs more).  Measured on B200: the same
// buffers taken from the stream-ordered pool (cudaMallocAsync) cost 15-3500 ms of host time per 10-30 M triangle build
// (profiles/r2t_build_trace.txt) against 20-90 ms of kernels; from the arena they cost nothing after the first build.
// rtbvh_gpu_trim_workspace() gives the memory back.
static thread_local double g_alloc_host_ms = 0;  // host time spent inside cudaMalloc for the arena (RTBVH_BUILD_TRACE)
static bool build_trace() {
    static const bool v = std::getenv("RTBVH_BUILD_TRACE") != nullptr;
    return v;
}
struct Arena {
    struct Chunk {
        char* base;
        size_t cap, used;
    };
    std::vector<Chunk> chunks;
    int device = -1;
    ~Arena() { release(); }
    void release() {
        for (auto& c : chunks) cudaFree(c.base);
        chunks.clear();
    }
    // Start of a builder entry point: everything handed out before is dead.  Several chunks mean the last build had to
    // grow: merge them into one allocation of the combined size so that the steady state is a single chunk.
    cudaError_t reset() {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev != device) {
            release();
            device = dev;
        }
        if (chunks.size() > 1) {
            size_t total = 0;
            for (auto& c : chunks) total += c.cap;
            release();
            const cudaError_t e = grow(total);
            if (e != cudaSuccess) return e;
        }
        for (auto& c : chunks) c.used = 0;
        return cudaSuccess;
    }
    // Position of the bump pointer, to rewind to later (what was handed out before the mark stays valid).
    struct Mark {
        size_t chunks = 0, used = 0;
    };
    Mark mark() const { return Mark{chunks.size(), chunks.empty() ? 0 : chunks.back().used}; }
    void rewind(const Mark& m) {
        if (m.chunks == 0 || m.chunks > chunks.size()) return;
        for (size_t i = m.chunks; i < chunks.size(); i++) chunks[i].used = 0;
        chunks[m.chunks - 1].used = m.used;
    }
    cudaError_t grow(size_t cap) {
        Chunk c{nullptr, cap, 0};
        const auto t0 = std::chrono::steady_clock::now();
        const cudaError_t e = cudaMalloc((void**)&c.base, cap);
        g_alloc_host_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (e == cudaSuccess) chunks.push_back(c);
        return e;
    }
    cudaError_t take(size_t bytes, void** out) {
        bytes = (bytes + 255) & ~size_t(255);
        // (after a rewind the bump pointer may sit in an earlier chunk: allocation always continues in the last one)
        if (chunks.empty() || chunks.back().cap - chunks.back().used < bytes) {
            const size_t last = chunks.empty() ? 0 : chunks.back().cap;
            const cudaError_t e = grow(std::max({bytes, last, size_t(64) << 20}));
            if (e != cudaSuccess) return e;
        }
        Chunk& c = chunks.back();
        *out = c.base + c.used;
        c.used += bytes;
        return cudaSuccess;
    }
};
static Arena& arena() {
    static thread_local Arena a;
    return a;
}
struct DevBuf {  // a view into the arena: nothing to free
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    cudaError_t alloc(size_t b) {
        if (b <= bytes && p) return cudaSuccess;
        const cudaError_t e = arena().take(b ? b : 16, &p);
        bytes = e == cudaSuccess ? b : 0;
        return e;
    }
    template <class T>
    T* as() { return (T*)p; }
};
// A cudaMalloc'ed copy of an arena buffer, owned by the caller (the resident scene path).
static cudaError_t detach(const DevBuf& b, size_t bytes, void** out) {
    *out = nullptr;
    cudaError_t e = dev_block_alloc(out, bytes ? bytes : 16);
    if (e != cudaSuccess) return e;
    return bytes ? cudaMemcpyAsync(*out, b.p, bytes, cudaMemcpyDeviceToDevice, 0) : cudaSuccess;
}
inline unsigned blocks(size_t n, int b) { return (unsigned)((n + b - 1) / b); }

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    Timer() {
        cudaEventCreate(&a);
        cudaEventCreate(&b);
    }
    ~Timer() {
        cudaEventDestroy(a);
        cudaEventDestroy(b);
    }
    void start() { cudaEventRecord(a, 0); }
    float stop() {
        cudaEventRecord(b, 0);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        return ms;
    }
};

}  // namespace

// =================================================================================================
// Device-resident builders
// =================================================================================================
struct DeviceBvh {
    DevBuf nodes;    // float4[2 * node_count]
    DevBuf indices;  // uint32[n]
    uint32_t node_count = 0, index_count = 0;
};

struct Uint4Add {
    __host__ __device__ uint4 operator()(const uint4& a, const uint4& b) const {
        return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
};

static ResultCode build_binned_sah_device(const float4* d_bb, const float* d_cen, uint32_t cstride, uint32_t n,
                                          uint32_t max_leaf, DeviceBvh* out) {
    const uint32_t max_nodes = 2 * n - 1;
    DevBuf nodesA;
    RTB_CUDA(nodesA.alloc((size_t)max_nodes * 32));
    float4* nodes = nodesA.as<float4>();
    DevBuf idxB[2], ptB[2], tasksB[2], auxB[2], binidx, span_tasks, dec, bins, counts, ranks4, ptask, world, temp, state, params, small_tasks, used,
        waste, waste_prefix;
    for (int k = 0; k < 2; k++) {
        RTB_CUDA(idxB[k].alloc((size_t)n * 4));
        RTB_CUDA(ptB[k].alloc((size_t)n * 4));
    }
    RTB_CUDA(world.alloc(6 * 4));
    // every small subtree holds >= 2 primitives, so there are at most n / 2 of them
    const uint32_t small_cap = n / 2 + 1;
    RTB_CUDA(small_tasks.alloc((size_t)small_cap * sizeof(SmallTask)));
    // root = union_of_list(aabbs) (binned_sah.rs:361)
    world_init_kernel<<<1, 32>>>(world.as<uint32_t>());
    world_reduce_kernel<<<std::min(blocks(n, 256), 148u * 8u), 256>>>(d_bb, n, world.as<uint32_t>());
    world_to_node_kernel<<<1, 1>>>(world.as<uint32_t>(), nodes, 0, 0, 0, nullptr);
    iota_kernel<<<blocks(n, 256), 256>>>(idxB[0].as<uint32_t>(), n);
    LevelState fin{0u, 1u, 0u, 0u};
    int parity = 0;  // idxB[parity] holds the current index order
    if (n <= 1) {  // the root is a leaf on entry: pad (run) + pad (make_leaf)
        // world_to_node wrote union_of_list's own pad; apply run's and make_leaf's pads
        root_leaf_kernel<<<1, 1>>>(nodes, n);
    } else {
        // A level task holds > kSmall primitives (smaller children go to the small-subtree list), so no level has more
        // than n / (kSmall + 1) tasks: every per-task array is sized once, nothing grows inside the loop.
        const uint32_t task_cap = n / (kSmall + 1) + 2;
        for (int k = 0; k < 2; k++) {
            RTB_CUDA(tasksB[k].alloc((size_t)task_cap * sizeof(Task)));
            RTB_CUDA(auxB[k].alloc((size_t)task_cap * sizeof(TaskAux)));
        }
        RTB_CUDA(dec.alloc((size_t)task_cap * sizeof(Decision)));
        // only span-class tasks (> kWarpTask primitives) own a bin block in HBM: at most n / (kWarpTask + 1) per level
        const uint32_t span_cap = n / (kWarpTask + 1) + 2;
        RTB_CUDA(bins.alloc((size_t)span_cap * kTaskBinWords * 4));
        RTB_CUDA(span_tasks.alloc((size_t)(kMaxDepth + 3) * 4));
        RTB_CUDA(cudaMemsetAsync(span_tasks.p, 0, (size_t)(kMaxDepth + 3) * 4, 0));
        RTB_CUDA(counts.alloc((size_t)task_cap * sizeof(uint4)));
        RTB_CUDA(ranks4.alloc((size_t)task_cap * sizeof(uint4)));
        RTB_CUDA(ptask.alloc((size_t)task_cap * sizeof(PartTask)));
        RTB_CUDA(binidx.alloc((size_t)n * 2));
        RTB_CUDA(state.alloc((size_t)(kMaxDepth + 3) * sizeof(LevelState)));
        RTB_CUDA(params.alloc(2 * sizeof(PartitionParams)));
        const uint32_t part_chunks = blocks(n, kPartChunk);
        const bool part_block = partition_block_mode() && part_chunks <= kPartMaxChunks;
        DevBuf part_tail;
        if (part_block) RTB_CUDA(part_tail.alloc((size_t)part_chunks * 4));
        PartitionParams hp[2];
        for (int k = 0; k < 2; k++)
            hp[k] = PartitionParams{idxB[k].as<uint32_t>(), ptB[k].as<int32_t>(), ptask.as<PartTask>(), binidx.as<uint16_t>(),
                                    idxB[k ^ 1].as<uint32_t>(), ptB[k ^ 1].as<int32_t>()};
        RTB_CUDA(cudaMemcpyAsync(params.p, hp, sizeof(hp), cudaMemcpyHostToDevice, 0));
        // CUB temp storage: sized for the largest scans we run
        size_t temp_bytes = 0, tb = 0;
        const PartitionParams* dp = params.as<PartitionParams>();
        cub::DeviceScan::InclusiveScanByKey(nullptr, tb, ptB[0].as<int32_t>(),
                                            PartitionFlagIter(cub::CountingInputIterator<uint32_t>(0), PartitionFlagIn{dp}),
                                            PartitionScatterOut{dp, 0}, FlagSumOp(), n);
        temp_bytes = tb;
        cub::DeviceScan::ExclusiveScan(nullptr, tb, (uint4*)nullptr, (uint4*)nullptr, Uint4Add(), make_uint4(0, 0, 0, 0), (int)task_cap);
        temp_bytes = std::max(temp_bytes, tb);
        cub::DeviceScan::ExclusiveSum(nullptr, tb, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)small_cap);
        temp_bytes = std::max(temp_bytes, tb);
        RTB_CUDA(temp.alloc(temp_bytes));
        // bins start clean and are left clean by every split (sah_split_kernel re-initialises what it has read)
        const size_t words = (size_t)span_cap * kTaskBinWords;
        sah_bins_init_kernel<<<blocks(words, 256), 256>>>(bins.as<uint32_t>(), words);
        fill_i32_kernel<<<blocks(n, 256), 256>>>(ptB[0].as<int32_t>(), n, pt_encode(0, n));  // every position belongs to task 0
        LevelState* d_state = state.as<LevelState>();
        sah_root_task_kernel<<<1, 1>>>(nodes, n, tasksB[0].as<Task>(), auxB[0].as<TaskAux>(), d_state, span_tasks.as<uint32_t>());
        // bin kernel: a machine-sized grid, every block owns a contiguous span (multiple of the chunk size)
        const uint32_t bin_chunks = blocks(n, kBinBlock);
        const bool bin_priv = n >= (4u << 20);
        const uint32_t bin_grid = std::min(bin_chunks, 148u * (bin_priv ? 5u : 8u));  // one wave of resident blocks
        const uint32_t bin_span = ((bin_chunks + bin_grid - 1) / bin_grid) * kBinBlock;
        // The level loop runs on the device's own bookkeeping (LevelState): the host enqueues level after level with
        // launch sizes from an upper bound of the task count and only looks at the state — one level behind, through a
        // pinned copy — once the tree could be finished (a task can only vanish after log2(n / kSmall) halvings unless the
        // input is degenerate; a level enqueued after the last one is a no-op that passes the index order through).
        static thread_local LevelState* h_state = nullptr;
        if (!h_state) RTB_CUDA(cudaHostAlloc(&h_state, (size_t)(kMaxDepth + 3) * sizeof(LevelState), cudaHostAllocDefault));
        struct EventPair {  // destroyed on every exit path of the level loop
            cudaEvent_t e[2] = {nullptr, nullptr};
            ~EventPair() {
                if (e[0]) cudaEventDestroy(e[0]);
                if (e[1]) cudaEventDestroy(e[1]);
            }
            cudaEvent_t& operator[](unsigned i) { return e[i]; }
        } ev;
        RTB_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
        RTB_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
        // From the first level that could be free of span-class tasks (all tasks <= kWarpTask needs log2(n / kWarpTask)
        // halvings) the host reads the state one level behind; the GPU still has a full level queued, so this costs no
        // bubble.  Once a level has no span-class task none will follow (children only get smaller): from then on the
        // span bin, the span split and the global partition pass — and with it the ping-pong of the index buffers — are
        // left out.  The loop ends when a level had no task at all.
        static thread_local uint32_t* h_span = nullptr;
        if (!h_span) RTB_CUDA(cudaHostAlloc(&h_span, (size_t)(kMaxDepth + 3) * 4, cudaHostAllocDefault));
        uint32_t first_check = 0;
        while ((uint64_t(kWarpTask) << (first_check + 1)) < n) first_check++;
        static const bool merge_split = level_flag("RTBVH_SAH_MERGE"), scan_emit = level_flag("RTBVH_SAH_SCANEMIT");
        uint32_t depth = 0;
        bool done = false, span_left = true;
        for (; !done && depth <= (uint32_t)kMaxDepth; depth++) {
            const int par = (int)(depth & 1u);  // tasks / aux ping-pong every level
            const uint32_t A_ub = depth >= 31 ? task_cap : std::min(task_cap, 1u << depth);
            const Task* t_cur = tasksB[par].as<Task>();
            const TaskAux* aux_cur = auxB[par].as<TaskAux>();
            const LevelState* st = d_state + depth;
            if (span_left) {
                if (bin_priv)
                    sah_bin_kernel<true><<<bin_grid, kBinBlock>>>(idxB[parity].as<uint32_t>(), ptB[parity].as<int32_t>(), n, bin_span, t_cur,
                                                                  aux_cur, d_bb, d_cen, cstride, bins.as<uint32_t>(), binidx.as<uint16_t>(),
                                                                  st, span_tasks.as<uint32_t>() + depth);
                else
                    sah_bin_kernel<false><<<bin_grid, kBinBlock>>>(idxB[parity].as<uint32_t>(), ptB[parity].as<int32_t>(), n, bin_span, t_cur,
                                                                   aux_cur, d_bb, d_cen, cstride, bins.as<uint32_t>(), binidx.as<uint16_t>(),
                                                                   st, span_tasks.as<uint32_t>() + depth);
                if (!merge_split)
                    sah_split_kernel<<<blocks((size_t)A_ub * 32, 128), 128>>>(t_cur, A_ub, aux_cur, bins.as<uint32_t>(), nodes, max_leaf,
                                                                              depth, dec.as<Decision>(), counts.as<uint4>(), st);
            } else if (!merge_split) {
                zero_counts_tail_kernel<<<blocks(A_ub, 256), 256>>>(counts.as<uint4>(), A_ub, st);
            }
            if (merge_split)
                sah_warp_task_kernel<true><<<blocks(A_ub, kWarpTaskWarps), kWarpTaskWarps * 32>>>(
                    t_cur, idxB[parity].as<uint32_t>(), aux_cur, d_bb, d_cen, cstride, nodes, max_leaf, depth, dec.as<Decision>(),
                    counts.as<uint4>(), st, A_ub, bins.as<uint32_t>());
            else
                sah_warp_task_kernel<false><<<blocks(A_ub, kWarpTaskWarps), kWarpTaskWarps * 32>>>(
                    t_cur, idxB[parity].as<uint32_t>(), aux_cur, d_bb, d_cen, cstride, nodes, max_leaf, depth, dec.as<Decision>(),
                    counts.as<uint4>(), st, A_ub, bins.as<uint32_t>());
            size_t tbytes = temp.bytes;
            if (scan_emit && A_ub <= (uint32_t)kScanEmitTasks) {
                sah_scan_emit_kernel<<<1, kScanEmitTasks>>>(t_cur, A_ub, dec.as<Decision>(), counts.as<uint4>(), depth, nodes,
                                                            tasksB[par ^ 1].as<Task>(), auxB[par ^ 1].as<TaskAux>(),
                                                            small_tasks.as<SmallTask>(), ptask.as<PartTask>(),
                                                            span_tasks.as<uint32_t>(), d_state);
            } else {
                RTB_CUDA(cub::DeviceScan::ExclusiveScan(temp.p, tbytes, counts.as<uint4>(), ranks4.as<uint4>(), Uint4Add(),
                                                        make_uint4(0, 0, 0, 0), (int)A_ub));
                sah_emit_kernel<<<blocks(A_ub, 128), 128>>>(t_cur, A_ub, dec.as<Decision>(), ranks4.as<uint4>(), counts.as<uint4>(),
                                                            depth, nodes, tasksB[par ^ 1].as<Task>(), auxB[par ^ 1].as<TaskAux>(),
                                                            small_tasks.as<SmallTask>(), ptask.as<PartTask>(),
                                                            span_tasks.as<uint32_t>(), d_state);
            }
            if (span_left) {
                if (part_block) {
                    part_count_kernel<<<part_chunks, kPartBlock>>>(dp + parity, n, part_tail.as<uint32_t>());
                    part_scatter_kernel<<<part_chunks, kPartBlock>>>(dp + parity, n, part_tail.as<uint32_t>());
                } else {
                    tbytes = temp.bytes;
                    RTB_CUDA(cub::DeviceScan::InclusiveScanByKey(temp.p, tbytes, ptB[parity].as<int32_t>(),
                                                                 PartitionFlagIter(cub::CountingInputIterator<uint32_t>(0), PartitionFlagIn{dp + parity}),
                                                                 PartitionScatterOut{dp + parity, 0}, FlagSumOp(), n));
                }
                parity ^= 1;
            }
            RTB_CUDA(cudaMemcpyAsync(&h_state[depth + 1], d_state + depth + 1, sizeof(LevelState), cudaMemcpyDeviceToHost, 0));
            RTB_CUDA(cudaMemcpyAsync(&h_span[depth + 1], span_tasks.as<uint32_t>() + depth + 1, 4, cudaMemcpyDeviceToHost, 0));
            RTB_CUDA(cudaEventRecord(ev[par], 0));
            if (depth >= 1 && depth - 1 >= first_check) {
                RTB_CUDA(cudaEventSynchronize(ev[par ^ 1]));  // level depth - 1 has finished: h_state / h_span [depth] are valid
                if (h_state[depth].A == 0) done = true;      // the level just enqueued is a no-op; nothing follows it
                if (h_span[depth] == 0) span_left = false;   // level `depth` had no span-class task: none from depth + 1 on
            }
        }
        RTB_CUDA(cudaEventSynchronize(ev[(depth - 1) & 1u]));
        RTB_CUDA(cudaGetLastError());
        fin = h_state[depth];
        if (fin.A != 0) return fail("binned SAH: level loop did not terminate");
        if (fin.S > small_cap) return fail("binned SAH: small-subtree list overflow");
    }
    uint32_t* idx_cur = idxB[parity].as<uint32_t>();
    const uint32_t S = fin.S, small_slots = fin.small_slots;
    const uint32_t level_nodes = fin.node_count;
    uint32_t total_nodes = level_nodes;
    if (S > 0) {
        // rebase the reserved ranges behind the level nodes, then let one warp finish each subtree
        RTB_CUDA(used.alloc((size_t)S * 4));
        small_rebase_kernel<<<blocks(S, 256), 256>>>(small_tasks.as<SmallTask>(), S, level_nodes);
        if (small_ls_mode())
            sah_small_ls_kernel<<<blocks(S, kLsWarps), kLsWarps * 32>>>(small_tasks.as<SmallTask>(), S, idx_cur, d_bb, d_cen, cstride,
                                                                        nodes, max_leaf, used.as<uint32_t>());
        else
            sah_small_kernel<<<blocks(S, kSmallWarps), kSmallWarps * 32>>>(small_tasks.as<SmallTask>(), S, idx_cur, d_bb, d_cen, cstride,
                                                                           nodes, max_leaf, used.as<uint32_t>());
        // compaction only if some subtree ended early (leaves with several primitives)
        RTB_CUDA(waste.alloc((size_t)S * 4));
        RTB_CUDA(waste_prefix.alloc((size_t)(S + 1) * 4));
        small_waste_kernel<<<blocks(S, 256), 256>>>(small_tasks.as<SmallTask>(), used.as<uint32_t>(), S, waste.as<uint32_t>());
        size_t tbytes = temp.bytes;
        RTB_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, tbytes, waste.as<uint32_t>(), waste_prefix.as<uint32_t>(), (int)S));
        uint32_t hw[2];
        RTB_CUDA(cudaMemcpy(&hw[0], waste_prefix.as<uint32_t>() + (S - 1), 4, cudaMemcpyDeviceToHost));
        RTB_CUDA(cudaMemcpy(&hw[1], waste.as<uint32_t>() + (S - 1), 4, cudaMemcpyDeviceToHost));
        const uint32_t wasted = hw[0] + hw[1];
        const uint32_t slots = level_nodes + small_slots;
        total_nodes = slots - wasted;
        if (wasted > 0) {
            RTB_CUDA(out->nodes.alloc((size_t)total_nodes * 32));
            compact_nodes_kernel<<<blocks(slots, 256), 256>>>(nodes, out->nodes.as<float4>(), level_nodes,
                                                             small_tasks.as<SmallTask>(), used.as<uint32_t>(),
                                                             waste_prefix.as<uint32_t>(), S, slots);
        }
    }
    if (!out->nodes.p) {  // no compaction: hand the build buffer over
        std::swap(out->nodes.p, nodesA.p);
        std::swap(out->nodes.bytes, nodesA.bytes);
    }
    // hand the index buffer over as well
    std::swap(out->indices.p, idxB[parity].p);
    std::swap(out->indices.bytes, idxB[parity].bytes);
    out->node_count = total_nodes;
    out->index_count = n;
    RTB_CUDA(cudaDeviceSynchronize());
    return Ok;
}

static ResultCode build_locb_device(const float4* d_bb, const float* d_cen, uint32_t cstride, uint32_t n, DeviceBvh* out,
                                    uint32_t* iterations) {
    DevBuf world_keys, world;
    RTB_CUDA(world_keys.alloc(6 * 4));
    RTB_CUDA(world.alloc(6 * 4));
    world_init_kernel<<<1, 32>>>(world_keys.as<uint32_t>());
    world_reduce_kernel<<<std::min(blocks(n, 256), 148u * 8u), 256>>>(d_bb, n, world_keys.as<uint32_t>());
    RTB_CUDA(out->indices.alloc((size_t)n * 4));
    if (n <= 2) {  // single root leaf holding everything (locb.rs:258-269)
        RTB_CUDA(out->nodes.alloc(32));
        world_to_node_kernel<<<1, 1>>>(world_keys.as<uint32_t>(), out->nodes.as<float4>(), 0, (int)n, 0, nullptr);
        iota_kernel<<<1, 32>>>(out->indices.as<uint32_t>(), n);
        out->node_count = 1;
        out->index_count = n;
        RTB_CUDA(cudaDeviceSynchronize());
        return Ok;
    }
    world_to_node_kernel<<<1, 1>>>(world_keys.as<uint32_t>(), nullptr, 0, 0, 0, world.as<float>());
    const uint32_t node_count = 2 * n - 1;
    DevBuf codes, codes2, idx2, nodesB, nb, merged, P, temp;
    RTB_CUDA(codes.alloc((size_t)n * 4));
    RTB_CUDA(codes2.alloc((size_t)n * 4));
    RTB_CUDA(idx2.alloc((size_t)n * 4));
    RTB_CUDA(out->nodes.alloc((size_t)node_count * 32));
    RTB_CUDA(nodesB.alloc((size_t)node_count * 32));
    RTB_CUDA(nb.alloc((size_t)node_count * 4));
    RTB_CUDA(merged.alloc((size_t)node_count * 4));
    RTB_CUDA(P.alloc((size_t)node_count * 4));
    morton_kernel<<<blocks(n, 256), 256>>>(d_cen, cstride, n, world.as<float>(), codes.as<uint32_t>(), idx2.as<uint32_t>());
    size_t tb = 0, tb2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, codes.as<uint32_t>(), codes2.as<uint32_t>(), idx2.as<uint32_t>(),
                                    out->indices.as<uint32_t>(), (int)n, 0, 30);
    cub::DeviceScan::InclusiveSum(nullptr, tb2, merged.as<uint32_t>(), P.as<uint32_t>(), (int)n);
    RTB_CUDA(temp.alloc(std::max(tb, tb2)));
    tb = temp.bytes;
    RTB_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, tb, codes.as<uint32_t>(), codes2.as<uint32_t>(), idx2.as<uint32_t>(),
                                             out->indices.as<uint32_t>(), (int)n, 0, 30));  // LSD radix sort: stable
    // all nodes start as {0-box, count -1, left_first 0} (locb.rs:273-281); every slot is overwritten below
    float4* cur = out->nodes.as<float4>();
    float4* other = nodesB.as<float4>();
    RTB_CUDA(cudaMemset(cur, 0, (size_t)node_count * 32));
    RTB_CUDA(cudaMemset(other, 0, (size_t)node_count * 32));
    uint32_t begin = node_count - n, end = node_count, previous_end = end;
    locb_leaves_kernel<<<blocks(n, 256), 256>>>((const float4*)d_bb, out->indices.as<uint32_t>(), n, cur, begin);
    uint32_t iters = 0;
    while (end - begin > kLocbTail) {
        const uint32_t c = end - begin;
        locb_nn_kernel<<<blocks(c, 128), 128>>>(cur, begin, end, nb.as<uint32_t>());
        locb_flag_kernel<<<blocks(c, 256), 256>>>(nb.as<uint32_t>(), begin, end, merged.as<uint32_t>());
        tb = temp.bytes;
        RTB_CUDA(cub::DeviceScan::InclusiveSum(temp.p, tb, merged.as<uint32_t>() + begin, P.as<uint32_t>() + begin, (int)c));
        locb_write_kernel<<<blocks(previous_end - begin, 256), 256>>>(cur, other, nb.as<uint32_t>(), P.as<uint32_t>(), begin, end,
                                                                      previous_end);
        uint32_t merged_count = 0;
        RTB_CUDA(cudaMemcpy(&merged_count, P.as<uint32_t>() + (end - 1), 4, cudaMemcpyDeviceToHost));
        if (merged_count == 0) return fail("LOCB: no mutual pair found (degenerate input)");
        const uint32_t children_begin = end - 2 * merged_count;
        const uint32_t unmerged_begin = children_begin - (c - merged_count);
        std::swap(cur, other);
        previous_end = end;
        begin = unmerged_begin;
        end = children_begin;
        iters++;
    }
    if (end - begin > 1) {  // the small iterations: one block, no host round trips
        DevBuf tail_state;
        RTB_CUDA(tail_state.alloc(8));
        locb_tail_kernel<<<1, kLocbTailThreads>>>(cur, other, nb.as<uint32_t>(), P.as<uint32_t>(), begin, end, previous_end,
                                                  tail_state.as<uint32_t>());
        uint32_t ts[2] = {0, 0};
        RTB_CUDA(cudaMemcpy(ts, tail_state.p, 8, cudaMemcpyDeviceToHost));
        if (ts[1]) return fail("LOCB: no mutual pair found (degenerate input)");
        if (ts[0] & 1u) std::swap(cur, other);
        iters += ts[0];
    }
    if (cur != out->nodes.as<float4>()) {  // keep the result in out->nodes
        std::swap(out->nodes.p, nodesB.p);
        std::swap(out->nodes.bytes, nodesB.bytes);
    }
    if (iterations) *iterations = iters;
    out->node_count = node_count;
    out->index_count = n;
    RTB_CUDA(cudaDeviceSynchronize());
    return Ok;
}

static ResultCode collapse_device(const float4* d_nodes, uint32_t n_nodes, DevBuf* mnodes, uint32_t* m_count) {
    if (n_nodes == 0) {
        *m_count = 0;
        return Ok;
    }
    // a root that is itself a leaf yields one m-node whose slot 0 is that leaf (mbvh_node.rs:308-323)
    float4 root[2];
    RTB_CUDA(cudaMemcpy(root, d_nodes, 32, cudaMemcpyDeviceToHost));
    int rc, rl;
    memcpy(&rc, &root[0].w, 4);
    memcpy(&rl, &root[1].w, 4);
    if (rc >= 0 || rl < 0) {
        RTMbvhNode m;
        for (int s = 0; s < 4; s++) {
            m.min_x[s] = root[0].x; m.min_y[s] = root[0].y; m.min_z[s] = root[0].z;
            m.max_x[s] = root[1].x; m.max_y[s] = root[1].y; m.max_z[s] = root[1].z;
            m.children[s] = -1;
            m.counts[s] = -1;
        }
        if (rc >= 0 && rl == 0) {  // verbatim: the leaf's primitive offset is reinterpreted as a node index -> node 0
            m.children[0] = rl;
            m.counts[0] = rc;
        }
        RTB_CUDA(mnodes->alloc(128));
        RTB_CUDA(cudaMemcpy(mnodes->p, &m, 128, cudaMemcpyHostToDevice));
        *m_count = 1;
        return Ok;
    }
    DevBuf parent, is_m, sub, mindex, arrived;
    RTB_CUDA(parent.alloc((size_t)n_nodes * 4));
    RTB_CUDA(is_m.alloc(n_nodes));
    RTB_CUDA(sub.alloc((size_t)n_nodes * 4));
    RTB_CUDA(mindex.alloc((size_t)n_nodes * 4));
    RTB_CUDA(arrived.alloc((size_t)n_nodes * 4));
    RTB_CUDA(cudaMemsetAsync(sub.p, 0, (size_t)n_nodes * 4, 0));
    RTB_CUDA(cudaMemsetAsync(arrived.p, 0, (size_t)n_nodes * 4, 0));
    RTB_CUDA(cudaMemsetAsync(parent.p, 0xFF, (size_t)n_nodes * 4, 0));
    parents_kernel<<<blocks(n_nodes, 256), 256>>>(d_nodes, n_nodes, parent.as<int32_t>());
    mroot_flag_kernel<<<blocks(n_nodes, 256), 256>>>(d_nodes, n_nodes, parent.as<int32_t>(), is_m.as<uint8_t>());
    mroot_sub_kernel<<<blocks(n_nodes, 256), 256>>>(d_nodes, n_nodes, parent.as<int32_t>(), is_m.as<uint8_t>(),
                                                    sub.as<uint32_t>(), arrived.as<uint32_t>());
    uint32_t total = 0;
    RTB_CUDA(cudaMemcpy(&total, sub.p, 4, cudaMemcpyDeviceToHost));
    RTB_CUDA(mnodes->alloc((size_t)total * 128));
    mroot_index_kernel<<<blocks(n_nodes, 256), 256>>>(d_nodes, n_nodes, parent.as<int32_t>(), is_m.as<uint8_t>(), sub.as<uint32_t>(),
                                                      mindex.as<uint32_t>());
    collapse_emit_kernel<<<blocks(n_nodes, 256), 256>>>(d_nodes, n_nodes, is_m.as<uint8_t>(), mindex.as<uint32_t>(),
                                                        mnodes->as<float4>());
    RTB_CUDA(cudaGetLastError());
    RTB_CUDA(cudaDeviceSynchronize());
    *m_count = total;
    return Ok;
}

// =================================================================================================
// Host-facing entry points (host arrays in, host mirrors out)
// =================================================================================================
// The tree the last build of this thread left in the arena: create_mbvh right after create_bvh collapses it in place instead
// of uploading the host mirror again.  Valid while the arena has not been reset (epoch) and the mirror is the same array.
struct LastTree {
    uint64_t serial = 0;
    const void* d_nodes = nullptr;
    uint32_t node_count = 0;
    uint64_t epoch = 0;
    Arena::Mark mark;  // end of the build's allocations: every in-place collapse starts from here again
};
static thread_local LastTree g_last_tree;
static thread_local uint64_t g_arena_epoch = 1;
static std::atomic<uint64_t> g_build_serial{0};

static ResultCode need_device() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail("no CUDA device: the builders run on the GPU only (no CPU fallback)");
    }
    // every builder entry point starts with an empty workspace (what earlier builds of this thread left in it is dead:
    // they synchronised before returning and copied their results out)
    const cudaError_t e = arena().reset();
    g_arena_epoch++;
    if (e != cudaSuccess) return fail("builder workspace", e);
    return Ok;
}

static ResultCode download(const DeviceBvh& d, HostBvh* out) {
    if (!out->nodes.resize(d.node_count) || !out->indices.resize(d.index_count)) return fail("host mirror: out of memory");
    if (d.node_count) RTB_CUDA(cudaMemcpyAsync(out->nodes.data(), d.nodes.p, (size_t)d.node_count * 32, cudaMemcpyDeviceToHost, 0));
    if (d.index_count) RTB_CUDA(cudaMemcpyAsync(out->indices.data(), d.indices.p, (size_t)d.index_count * 4, cudaMemcpyDeviceToHost, 0));
    RTB_CUDA(cudaStreamSynchronize(0));
    out->serial = ++g_build_serial;
    g_last_tree = LastTree{out->serial, d.nodes.p, d.node_count, g_arena_epoch, arena().mark()};
    return Ok;
}

ResultCode gpu_build_bvh(const RTAabb* aabbs, size_t prim_count, const float* centers, size_t center_stride,
                         size_t prims_per_leaf, uint32_t bvh_type, HostBvh* out) {
    if (need_device() != Ok) return Error;
    if (prim_count >= (size_t(1) << 31)) return fail("more than 2^31 primitives");
    const uint32_t n = (uint32_t)prim_count;
    const uint32_t cstride = (uint32_t)(center_stride / 4);
    Timer total, dev;
    total.start();
    DevBuf bb, cen;
    RTB_CUDA(bb.alloc((size_t)n * 32));
    RTB_CUDA(cen.alloc((size_t)n * center_stride));
    RTB_CUDA(upload_from_user(cen.p, centers, (size_t)n * center_stride, 0));
    if (aabbs)
        RTB_CUDA(upload_from_user(bb.p, aabbs, (size_t)n * 32, 0));
    dev.start();
    if (!aabbs) point_boxes_kernel<<<blocks(n, 256), 256>>>(cen.as<float>(), cstride, n, bb.as<float4>());
    DeviceBvh d;
    uint32_t iters = 0;
    ResultCode rc;
    if (bvh_type == LocallyOrderedClustered) {
        rc = build_locb_device(bb.as<float4>(), cen.as<float>(), cstride, n, &d, &iters);
        out->build_type = 1;
    } else {  // BvhType::from(u32) maps everything else to BinnedSAH (builders/mod.rs:25-33)
        rc = build_binned_sah_device(bb.as<float4>(), cen.as<float>(), cstride, n, (uint32_t)(prims_per_leaf ? prims_per_leaf : 1), &d);
        out->build_type = 2;
    }
    if (rc != Ok) return rc;
    g_build_stats.device_ms = dev.stop();
    rc = download(d, out);
    g_build_stats.total_ms = total.stop();
    g_build_stats.iterations = iters;
    g_build_stats.node_count = d.node_count;
    return rc;
}

ResultCode gpu_build_bvh_triangles(const float* vertices, size_t vertex_stride, size_t tri_count, size_t prims_per_leaf,
                                   uint32_t bvh_type, HostBvh* out) {
    if (need_device() != Ok) return Error;
    if (tri_count >= (size_t(1) << 31)) return fail("more than 2^31 primitives");
    const uint32_t n = (uint32_t)tri_count;
    Timer total, dev;
    total.start();
    DevBuf verts, bb, cen;
    RTB_CUDA(verts.alloc((size_t)n * 3 * vertex_stride));
    RTB_CUDA(bb.alloc((size_t)n * 32));
    RTB_CUDA(cen.alloc((size_t)n * 12));
    RTB_CUDA(upload_from_user(verts.p, vertices, (size_t)n * 3 * vertex_stride, 0));
    dev.start();
    tri_prims_kernel<<<blocks(n, 256), 256>>>(verts.as<float>(), (uint32_t)(vertex_stride / 4), n, bb.as<float4>(), cen.as<float>());
    DeviceBvh d;
    uint32_t iters = 0;
    ResultCode rc;
    if (bvh_type == LocallyOrderedClustered) {
        rc = build_locb_device(bb.as<float4>(), cen.as<float>(), 3, n, &d, &iters);
        out->build_type = 1;
    } else {
        rc = build_binned_sah_device(bb.as<float4>(), cen.as<float>(), 3, n, (uint32_t)(prims_per_leaf ? prims_per_leaf : 1), &d);
        out->build_type = 2;
    }
    if (rc != Ok) return rc;
    g_build_stats.device_ms = dev.stop();
    rc = download(d, out);
    g_build_stats.total_ms = total.stop();
    g_build_stats.iterations = iters;
    g_build_stats.node_count = d.node_count;
    if (build_trace())
        std::fprintf(stderr, "[rtbvh build] n=%u type=%u device_ms=%.3f (host time inside cudaMalloc %.3f ms) total_ms=%.3f\n", n,
                     bvh_type, g_build_stats.device_ms, g_alloc_host_ms, g_build_stats.total_ms);
    g_alloc_host_ms = 0;
    return rc;
}

ResultCode gpu_collapse(const HostBvh& bvh, HostMbvh* out) {
    out->m_nodes.clear();
    const uint32_t n_nodes = (uint32_t)bvh.nodes.size();
    // the tree this thread built last may still sit in the builder workspace: collapse it where it is (no arena reset, the
    // collapse's own buffers are taken behind it) instead of uploading the 32 B/node mirror again
    const bool resident = n_nodes != 0 && bvh.serial != 0 && g_last_tree.serial == bvh.serial && g_last_tree.node_count == n_nodes &&
                          g_last_tree.epoch == g_arena_epoch;
    if (resident) {
        arena().rewind(g_last_tree.mark);  // repeated collapses of the same tree reuse the same workspace
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("no CUDA device: the builders run on the GPU only (no CPU fallback)");
    } else if (need_device() != Ok) {
        return Error;
    }
    if (n_nodes == 0) return Ok;
    Timer total, dev;
    total.start();
    DevBuf nodes, mnodes;
    const float4* d_nodes = (const float4*)g_last_tree.d_nodes;
    if (!resident) {
        RTB_CUDA(nodes.alloc((size_t)n_nodes * 32));
        RTB_CUDA(upload_from_user(nodes.p, bvh.nodes.data(), (size_t)n_nodes * 32, 0));
        d_nodes = nodes.as<float4>();
    }
    dev.start();
    uint32_t m_count = 0;
    if (collapse_device(d_nodes, n_nodes, &mnodes, &m_count) != Ok) return Error;
    g_build_stats.device_ms = dev.stop();
    if (!out->m_nodes.resize(m_count)) return fail("host mirror: out of memory");
    if (m_count) {
        RTB_CUDA(cudaMemcpyAsync(out->m_nodes.data(), mnodes.p, (size_t)m_count * 128, cudaMemcpyDeviceToHost, 0));
        RTB_CUDA(cudaStreamSynchronize(0));
    }
    g_build_stats.total_ms = total.stop();
    g_build_stats.node_count = m_count;
    g_build_stats.iterations = resident ? 1 : 0;  // 1: collapsed in place from the previous build's device copy
    return Ok;
}

ResultCode gpu_refit(HostBvh* bvh, const RTAabb* aabbs) {
    if (need_device() != Ok) return Error;
    const uint32_t n_nodes = (uint32_t)bvh->nodes.size(), n_idx = (uint32_t)bvh->indices.size();
    if (n_nodes == 0) return Ok;
    Timer total, dev;
    total.start();
    DevBuf nodes, idx, bb, parent, arrived;
    RTB_CUDA(nodes.alloc((size_t)n_nodes * 32));
    RTB_CUDA(idx.alloc((size_t)n_idx * 4));
    RTB_CUDA(bb.alloc((size_t)n_idx * 32));
    RTB_CUDA(parent.alloc((size_t)n_nodes * 4));
    RTB_CUDA(arrived.alloc((size_t)n_nodes * 4));
    RTB_CUDA(upload_from_user(nodes.p, bvh->nodes.data(), (size_t)n_nodes * 32, 0));
    RTB_CUDA(upload_from_user(idx.p, bvh->indices.data(), (size_t)n_idx * 4, 0));
    RTB_CUDA(upload_from_user(bb.p, aabbs, (size_t)n_idx * 32, 0));  // reads prim_count() aabbs (lib.rs:527-530)
    dev.start();
    RTB_CUDA(cudaMemset(arrived.p, 0, (size_t)n_nodes * 4));
    RTB_CUDA(cudaMemset(parent.p, 0xFF, (size_t)n_nodes * 4));
    parents_kernel<<<blocks(n_nodes, 256), 256>>>(nodes.as<float4>(), n_nodes, parent.as<int32_t>());
    refit_kernel<<<blocks(n_nodes, 256), 256>>>(nodes.as<float4>(), n_nodes, parent.as<int32_t>(), idx.as<uint32_t>(),
                                                bb.as<float4>(), arrived.as<uint32_t>());
    RTB_CUDA(cudaGetLastError());
    g_build_stats.device_ms = dev.stop();
    RTB_CUDA(cudaMemcpyAsync(bvh->nodes.data(), nodes.p, (size_t)n_nodes * 32, cudaMemcpyDeviceToHost, 0));
    RTB_CUDA(cudaStreamSynchronize(0));
    g_build_stats.total_ms = total.stop();
    return Ok;
}

// ---- build straight into device-resident trees (no host mirror): the scene path of rtbvh_gpu_scene_build ----------
ResultCode gpu_build_resident(const float* vertices, bool vertices_on_device, size_t vertex_stride, size_t tri_count,
                              size_t prims_per_leaf, uint32_t bvh_type, bool want_mbvh, ResidentTrees* out) {
    if (need_device() != Ok) return Error;
    if (tri_count >= (size_t(1) << 31)) return fail("more than 2^31 primitives");
    const uint32_t n = (uint32_t)tri_count;
    Timer total, dev;
    total.start();
    DevBuf verts, bb, cen;
    const float* d_verts = vertices;
    if (!vertices_on_device) {
        RTB_CUDA(verts.alloc((size_t)n * 3 * vertex_stride));
        RTB_CUDA(cudaMemcpyAsync(verts.p, vertices, (size_t)n * 3 * vertex_stride, cudaMemcpyHostToDevice, 0));
        d_verts = verts.as<float>();
    }
    RTB_CUDA(bb.alloc((size_t)n * 32));
    RTB_CUDA(cen.alloc((size_t)n * 12));
    dev.start();
    tri_prims_kernel<<<blocks(n, 256), 256>>>(d_verts, (uint32_t)(vertex_stride / 4), n, bb.as<float4>(), cen.as<float>());
    DeviceBvh d;
    uint32_t iters = 0;
    ResultCode rc;
    if (bvh_type == LocallyOrderedClustered)
        rc = build_locb_device(bb.as<float4>(), cen.as<float>(), 3, n, &d, &iters);
    else
        rc = build_binned_sah_device(bb.as<float4>(), cen.as<float>(), 3, n, (uint32_t)(prims_per_leaf ? prims_per_leaf : 1), &d);
    if (rc != Ok) return rc;
    DevBuf mnodes;
    uint32_t m_count = 0;
    if (want_mbvh && collapse_device(d.nodes.as<float4>(), d.node_count, &mnodes, &m_count) != Ok) return Error;
    g_build_stats.iterations = iters;
    g_build_stats.node_count = d.node_count;
    out->node_count = d.node_count;
    out->index_count = d.index_count;
    out->m_count = m_count;
    // the trees leave the builder's workspace: exact-size allocations owned by the scene
    RTB_CUDA(detach(d.nodes, (size_t)d.node_count * 32, &out->d_nodes));
    RTB_CUDA(detach(d.indices, (size_t)d.index_count * 4, (void**)&out->d_indices));
    if (want_mbvh) RTB_CUDA(detach(mnodes, (size_t)m_count * 128, &out->d_mnodes));
    if (!vertices_on_device) RTB_CUDA(detach(verts, (size_t)n * 3 * vertex_stride, (void**)&out->d_vertices));
    g_build_stats.device_ms = dev.stop();
    g_build_stats.total_ms = total.stop();
    if (build_trace())
        std::fprintf(stderr, "[rtbvh build] resident n=%u type=%u device_ms=%.3f (host time inside cudaMalloc %.3f ms) total_ms=%.3f\n",
                     n, bvh_type, g_build_stats.device_ms, g_alloc_host_ms, g_build_stats.total_ms);
    g_alloc_host_ms = 0;
    return Ok;
}

ResultCode gpu_trim_workspace() {
    arena().release();
    g_arena_epoch++;
    HostPool::get().trim();
    dev_block_trim();
    return Ok;
}

// ---- device block cache of the resident scenes (build.cuh) ---------------------------------------------------------
namespace {
constexpr size_t kDevCacheBlocks = 24;
struct DevBlock {
    void* p;
    size_t bytes;
    int device;
};
struct DevCache {
    std::mutex mu;
    std::vector<DevBlock> live, idle;
    size_t limit() {
        static const size_t v = [] {
            const char* e = std::getenv("RTBVH_SCENE_CACHE_MB");
            return (size_t)(e ? std::strtoull(e, nullptr, 10) : 24576ull) << 20;
        }();
        return v;
    }
};
DevCache& dev_cache() {
    static DevCache* c = new DevCache;  // leaked on purpose: scenes may be released during static destruction
    return *c;
}
#define g_dev_cache dev_cache()
}  // namespace

cudaError_t dev_block_alloc(void** p, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (bytes == 0) bytes = 16;
    {
        std::lock_guard<std::mutex> lk(g_dev_cache.mu);
        // best fit among the idle blocks of this device that waste at most a quarter (small blocks: at most 1 MiB)
        int best = -1;
        for (size_t i = 0; i < g_dev_cache.idle.size(); i++) {
            const DevBlock& b = g_dev_cache.idle[i];
            if (b.device != dev || b.bytes < bytes || b.bytes > bytes + bytes / 4 + (size_t(1) << 20)) continue;
            if (best < 0 || b.bytes < g_dev_cache.idle[best].bytes) best = (int)i;
        }
        if (best >= 0) {
            const DevBlock b = g_dev_cache.idle[best];
            g_dev_cache.idle.erase(g_dev_cache.idle.begin() + best);
            g_dev_cache.live.push_back(b);
            *p = b.p;
            return cudaSuccess;
        }
    }
    e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {  // out of memory with blocks parked in the cache: give them back and try once more
        cudaGetLastError();
        dev_block_trim();
        e = cudaMalloc(p, bytes);
        if (e != cudaSuccess) return e;
    }
    std::lock_guard<std::mutex> lk(g_dev_cache.mu);
    g_dev_cache.live.push_back(DevBlock{*p, bytes, dev});
    return cudaSuccess;
}

void dev_block_free(void* p) {
    if (!p) return;
    DevBlock blk{nullptr, 0, 0};
    std::vector<void*> drop;
    {
        std::lock_guard<std::mutex> lk(g_dev_cache.mu);
        for (size_t i = 0; i < g_dev_cache.live.size(); i++)
            if (g_dev_cache.live[i].p == p) {
                blk = g_dev_cache.live[i];
                g_dev_cache.live.erase(g_dev_cache.live.begin() + i);
                break;
            }
        if (blk.p && blk.bytes >= (size_t(1) << 20)) {  // only blocks worth keeping
            g_dev_cache.idle.push_back(blk);
            size_t total = 0;
            for (const DevBlock& b : g_dev_cache.idle) total += b.bytes;
            while (!g_dev_cache.idle.empty() && (g_dev_cache.idle.size() > kDevCacheBlocks || total > g_dev_cache.limit())) {
                total -= g_dev_cache.idle.front().bytes;  // oldest first
                drop.push_back(g_dev_cache.idle.front().p);
                g_dev_cache.idle.erase(g_dev_cache.idle.begin());
            }
            p = nullptr;
        }
    }
    // A cached block may still be read by kernels the caller enqueued before freeing the scene; the next owner's work is
    // stream-ordered behind them only on the legacy default stream, so drain the device once here (cudaFree would have).
    if (!p) cudaDeviceSynchronize();
    for (void* d : drop) cudaFree(d);
    if (p) cudaFree(p);
}

void dev_block_trim() {
    std::vector<void*> drop;
    {
        std::lock_guard<std::mutex> lk(g_dev_cache.mu);
        for (const DevBlock& b : g_dev_cache.idle) drop.push_back(b.p);
        g_dev_cache.idle.clear();
    }
    for (void* d : drop) cudaFree(d);
}

// ---- dynamic scenes: refit of device-resident trees ---------------------------------------------
ResidentRefit::~ResidentRefit() {
    cudaFree(parent);
    cudaFree(arrived);
    cudaFree(is_mroot);
    cudaFree(mindex);
    cudaFree(bb);
}

// One-off analysis of the topology (it does not change under refit): parent links and, if the scene holds the
// 4-wide collapse of this tree, which binary nodes are 4-wide roots and where their MbvhNode lives.
static ResultCode resident_refit_prepare(ResidentRefit* c, const float4* d_nodes, uint32_t n_nodes, uint32_t n_prims,
                                         uint32_t m_count, bool want_mbvh, cudaStream_t st) {
    if (c->n_nodes == n_nodes && c->parent && (!want_mbvh || c->is_mroot)) return Ok;
    RTB_CUDA(cudaMalloc(&c->parent, (size_t)n_nodes * 4));
    RTB_CUDA(cudaMalloc(&c->arrived, (size_t)n_nodes * 4));
    RTB_CUDA(cudaMalloc(&c->bb, (size_t)(n_prims ? n_prims : 1) * 32));
    c->n_nodes = n_nodes;
    RTB_CUDA(cudaMemsetAsync(c->parent, 0xFF, (size_t)n_nodes * 4, st));
    parents_kernel<<<blocks(n_nodes, 256), 256, 0, st>>>(d_nodes, n_nodes, c->parent);
    if (want_mbvh) {
        uint32_t* sub = nullptr;
        RTB_CUDA(cudaMalloc(&c->is_mroot, n_nodes));
        RTB_CUDA(cudaMalloc(&c->mindex, (size_t)n_nodes * 4));
        RTB_CUDA(cudaMalloc(&sub, (size_t)n_nodes * 4));
        RTB_CUDA(cudaMemsetAsync(sub, 0, (size_t)n_nodes * 4, st));
        RTB_CUDA(cudaMemsetAsync(c->arrived, 0, (size_t)n_nodes * 4, st));
        mroot_flag_kernel<<<blocks(n_nodes, 256), 256, 0, st>>>(d_nodes, n_nodes, c->parent, c->is_mroot);
        mroot_sub_kernel<<<blocks(n_nodes, 256), 256, 0, st>>>(d_nodes, n_nodes, c->parent, c->is_mroot, sub, c->arrived);
        mroot_index_kernel<<<blocks(n_nodes, 256), 256, 0, st>>>(d_nodes, n_nodes, c->parent, c->is_mroot, sub, c->mindex);
        uint32_t total = 0;
        RTB_CUDA(cudaMemcpyAsync(&total, sub, 4, cudaMemcpyDeviceToHost, st));
        RTB_CUDA(cudaStreamSynchronize(st));
        cudaFree(sub);
        if (total != m_count) return fail("scene refit: the scene's Mbvh is not the 4-wide collapse of its Bvh");
    }
    RTB_CUDA(cudaGetLastError());
    return Ok;
}

ResultCode gpu_refit_resident(ResidentRefit* c, float4* d_nodes, uint32_t n_nodes, const uint32_t* d_indices, uint32_t n_prims,
                              const float* d_vertices, uint32_t vstride, uint32_t tri_count, float4* d_mnodes, uint32_t m_count,
                              cudaStream_t st) {
    if (n_nodes == 0) return Ok;
    float4 root[2];
    RTB_CUDA(cudaMemcpyAsync(root, d_nodes, 32, cudaMemcpyDeviceToHost, st));
    RTB_CUDA(cudaStreamSynchronize(st));
    int rc, rl;
    memcpy(&rc, &root[0].w, 4);
    memcpy(&rl, &root[1].w, 4);
    const bool root_is_leaf = rc >= 0 || rl < 0;
    const bool want_mbvh = d_mnodes != nullptr && !root_is_leaf;
    if (d_mnodes != nullptr && root_is_leaf) return fail("scene refit: single-leaf trees have no 4-wide structure to refresh; recreate the scene");
    if (resident_refit_prepare(c, d_nodes, n_nodes, tri_count, m_count, want_mbvh, st) != Ok) return Error;
    // Primitive::aabb of the bench Triangle (un-padded), then Bvh::refit bottom-up, then merge_nodes' box rules again
    tri_boxes_kernel<<<blocks(tri_count, 256), 256, 0, st>>>(d_vertices, vstride, tri_count, c->bb);
    RTB_CUDA(cudaMemsetAsync(c->arrived, 0, (size_t)n_nodes * 4, st));
    refit_kernel<<<blocks(n_nodes, 256), 256, 0, st>>>(d_nodes, n_nodes, c->parent, d_indices, c->bb, c->arrived);
    if (want_mbvh)
        collapse_emit_kernel<<<blocks(n_nodes, 256), 256, 0, st>>>(d_nodes, n_nodes, c->is_mroot, c->mindex, d_mnodes);
    RTB_CUDA(cudaGetLastError());
    return Ok;
}

}  // namespace rtb
