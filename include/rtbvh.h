/* rtbvh.h — the rtbvh_ffi C ABI (library `rtbvh_rs`), byte-compatible with what cbindgen emits
 * from the reference's rtbvh_ffi/src/lib.rs (rtbvh_ffi/build.rs:21-31, include guard RTBVH_H).
 *
 * Every declaration cites the reference item it replaces.  The types are plain-old-data with the
 * reference's exact layouts (rtbvh_ffi `same_size` test: 32 / 128 / 32 bytes).
 *
 * What runs where in this implementation (rtbvh_b200/librtbvh_rs.so):
 *   create_bvh, create_mbvh, refit      -> hand-written sm_100a CUDA kernels; results are mirrored
 *                                          to host memory so RTBvh.nodes / .indices stay valid host
 *                                          pointers exactly as in the reference.
 *   intersect*, intersect_mbvh*         -> the per-candidate HOST callback cannot cross PCIe, so
 *                                          these walk the host mirror (compatibility shim, not the
 *                                          measured path).  The GPU path for rays is the batch API in
 *                                          rtbvh_gpu.h.
 * No entry point falls back to a CPU build: without a CUDA device create_* / refit return Error.
 */
#ifndef RTBVH_H
#define RTBVH_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* rtbvh_ffi/src/lib.rs:17-25 */
typedef enum ResultCode {
  Ok = 0,
  Error = 1,
  NoPrimitives = 2,
  InequalAabbsAndPrimitives = 3,
  Nan = 4,
} ResultCode;

/* rtbvh_ffi/src/lib.rs:129-133 (#[repr(u32)]) */
enum BvhType
#ifdef __cplusplus
  : uint32_t
#endif
{
  LocallyOrderedClustered = 0,
  BinnedSAH = 1,
};
#ifndef __cplusplus
typedef uint32_t BvhType;
#endif

/* rtbvh_ffi/src/lib.rs:144-151 == rtbvh::Aabb<i32> (src/aabb.rs:13-20).  Input arrays must be
 * 16-byte aligned (the reference reinterprets them as align(16) Aabb, lib.rs:359,452). */
typedef struct RTAabb {
  float min[3];
  int32_t count;      /* >= 0: leaf with `count` primitives; -1: inner node */
  float max[3];
  int32_t left_first; /* leaf: offset into indices; inner: index of the left child (right = +1); -1 invalid */
} RTAabb;

/* rtbvh_ffi/src/lib.rs:175-179 == rtbvh::BvhNode (src/bvh_node.rs:11-14) */
typedef struct RTBvhNode {
  RTAabb aabb;
} RTBvhNode;

/* rtbvh_ffi/src/lib.rs:197-208 == rtbvh::MbvhNode (src/mbvh_node.rs:28-38) */
typedef struct RTMbvhNode {
  float min_x[4];
  float max_x[4];
  float min_y[4];
  float max_y[4];
  float min_z[4];
  float max_z[4];
  int32_t children[4]; /* counts >= 0: offset into indices; counts == -1: m-node index; -1: empty slot */
  int32_t counts[4];
} RTMbvhNode;

/* rtbvh_ffi/src/lib.rs:210-220.  Passed by value.  Default: id = UINT32_MAX, null pointers. */
typedef struct RTBvh {
  uint32_t id;
  uint32_t node_count;
  const RTBvhNode *nodes;
  uint32_t index_count;
  const uint32_t *indices;
} RTBvh;

/* rtbvh_ffi/src/lib.rs:234-244 */
typedef struct RTMbvh {
  uint32_t id;
  uint32_t node_count;
  const RTMbvhNode *nodes;
  uint32_t index_count;
  const uint32_t *indices;
} RTMbvh;

/* Per-candidate callback: (primitive id, inout t, user data) -> true stops the traversal (any-hit).
 * For the packet entry points `t` points at 4 floats. (rtbvh_ffi/src/lib.rs:541-558) */
typedef bool (*RTIntersectCallback)(uint32_t prim_id, float *t, void *user_data);

/* rtbvh_ffi/src/lib.rs:345-389.  Spatial-split SAH is outside the GPU scope: returns Error. */
ResultCode create_spatial_Bvh(const RTAabb *aabbs, size_t prim_count, const float *centers, size_t stride,
                              const float *vertices, size_t vertex_stride, size_t triangle_stride,
                              uint32_t prims_per_leaf, RTBvh *result);

/* rtbvh_ffi/src/lib.rs:428-493.  aabbs may be null (centers then act as point primitives);
 * center_stride is 12 or 16 bytes (anything else: Error; the reference panics);
 * prims_per_leaf 0 means the default 1.  null centers/result -> Error; prim_count 0 -> NoPrimitives. */
ResultCode create_bvh(const RTAabb *aabbs, size_t prim_count, const float *centers, size_t center_stride,
                      size_t prims_per_leaf, BvhType bvh_type, RTBvh *result);

/* rtbvh_ffi/src/lib.rs:499-513.  Looks the tree up by bvh.id and collapses it to 4-wide nodes. */
ResultCode create_mbvh(RTBvh bvh, RTMbvh *mbvh);

/* rtbvh_ffi/src/lib.rs:519-538.  Reads bvh.index_count aabbs. */
ResultCode refit(const RTAabb *aabbs, RTBvh bvh);

/* rtbvh_ffi/src/lib.rs:551-581 */
ResultCode intersect(RTBvh bvh, const float *origin, const float *direction, float *t, void *user_data,
                     RTIntersectCallback intersect);

/* rtbvh_ffi/src/lib.rs:599-686 */
ResultCode intersect_packet(RTBvh bvh, const float *origin_x, const float *origin_y, const float *origin_z,
                            const float *direction_x, const float *direction_y, const float *direction_z, float *t,
                            void *user_data, RTIntersectCallback intersect);

/* rtbvh_ffi/src/lib.rs:700-731 */
ResultCode intersect_mbvh(RTMbvh bvh, const float *origin, const float *direction, float *t, void *user_data,
                          RTIntersectCallback intersect);

/* rtbvh_ffi/src/lib.rs:749-835 */
ResultCode intersect_mbvh_packet(RTMbvh bvh, const float *origin_x, const float *origin_y, const float *origin_z,
                                 const float *direction_x, const float *direction_y, const float *direction_z,
                                 float *t, void *user_data, RTIntersectCallback intersect);

/* rtbvh_ffi/src/lib.rs:838-849.  Ids are never reused; previously returned pointers dangle. */
void free_bvh(RTBvh bvh);
void free_mbvh(RTMbvh bvh);

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* RTBVH_H */
