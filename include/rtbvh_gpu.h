/* rtbvh_gpu.h — batch / device extension of the rtbvh C ABI (SURVEY.md §8b "required extension").
 *
 * The reference's traversal API is a host iterator that yields one candidate primitive at a time to
 * user code (src/iter.rs, src/iter_indices.rs; FFI callback rtbvh_ffi/src/lib.rs:551-835).  A GPU
 * cannot call back into host code per candidate, so the GPU path is a BATCH of the loop every caller
 * of the reference writes (examples/benchmark.rs:25-31 / :55-61, rtbvh_ffi/src/lib.rs:572-576):
 *
 *     for (prim, ray) in tree.iter(ray) { SpatialTriangle::intersect(prim, ray) }       closest hit
 *     ... with `break` on the first success                                              any hit
 *
 * with the canonical triangle tests of src/builders/spatial_sah.rs:131-244.  Visitation order,
 * predicates and floating-point arithmetic are the reference's (no FMA contraction), so t is
 * bit-identical to the host iterators and the primitive id is the lowest id among exactly equal t.
 *
 * Plain C: pointers and sizes only.  `*_device` variants take device pointers and a CUstream /
 * cudaStream_t (as void*) and return without synchronising.
 */
#ifndef RTBVH_GPU_H
#define RTBVH_GPU_H

#include "rtbvh.h"

#ifdef __cplusplus
extern "C" {
#endif

#define RT_NO_HIT 0xFFFFFFFFu

/* First 32 bytes of rtbvh::Ray (src/ray.rs:9-16): origin, t_min, direction, t.  Ray::new's derived
 * fields (inv_direction = 1/direction, signs) are recomputed on the device (src/ray.rs:166-182).
 * Use t_min = 1e-4f and t = 1e34f for Ray::new's defaults. */
typedef struct RTRay {
  float origin[3];
  float t_min;
  float direction[3];
  float t;
} RTRay;

/* Result of one ray: t as left in ray.t by the reference loop (unchanged input t on a miss) and the
 * primitive that produced it (RT_NO_HIT on a miss). */
typedef struct RTHit {
  float t;
  uint32_t prim;
} RTHit;

/* The seven SoA inputs of intersect_packet / intersect_mbvh_packet (rtbvh_ffi/src/lib.rs:599-608);
 * == rtbvh::RayPacket4 (src/ray.rs:47-61) without the derived inv_direction_*. */
typedef struct RTRayPacket4 {
  float origin_x[4], origin_y[4], origin_z[4];
  float direction_x[4], direction_y[4], direction_z[4];
  float t[4];
} RTRayPacket4;

typedef struct RTHitPacket4 {
  float t[4];
  uint32_t prim[4];
} RTHitPacket4;

typedef enum RTTreeKind {
  RT_TREE_BVH = 0,  /* rtbvh::Bvh: BvhIndexIterator / BvhPacketIndexIterator semantics  */
  RT_TREE_MBVH = 1, /* rtbvh::Mbvh: MbvhIndexIterator / MbvhPacketIndexIterator semantics */
} RTTreeKind;

/* A scene resident on one GPU: the tree(s) uploaded UNCHANGED plus the triangles, pre-gathered into
 * leaf order.  Opaque handle (0 is never valid). */
typedef uint64_t RTGpuScene;

/* ---- devices ---------------------------------------------------------------------------------- */
int rtbvh_gpu_device_count(void);                 /* 0 when no CUDA device / driver */
ResultCode rtbvh_gpu_set_device(int device);      /* device used by subsequent calls of this thread */
const char *rtbvh_gpu_last_error(void);           /* message of the last Error on this thread */

/* ---- scenes ----------------------------------------------------------------------------------- */
/* Uploads host trees (either may be null, not both) and triangles.  The RTBvh / RTMbvh structs are
 * trusted like the reference's intersect* trust them (no table lookup): any host arrays in the
 * reference's node formats work, e.g. a reference-built tree.  `vertices`: 3 * triangle_count
 * vertices, `vertex_stride` bytes apart (12 or 16), triangle i = vertices 3i, 3i+1, 3i+2 — the
 * layout of rtbvh_ffi's RTTriangleWrapper (rtbvh_ffi/src/lib.rs:258-330) with triangle_stride =
 * 3 * vertex_stride. */
ResultCode rtbvh_gpu_scene_create(const RTBvh *bvh, const RTMbvh *mbvh, const float *vertices, size_t vertex_stride,
                                  size_t triangle_count, RTGpuScene *scene);
ResultCode rtbvh_gpu_scene_free(RTGpuScene scene);
/* Build straight into a scene: Builder{aabbs: None, primitives: &[Triangle]}.construct_binned_sah /
 * construct_locally_ordered_clustered (src/bvh.rs:87-137) and, if want_mbvh, Mbvh::construct (src/bvh.rs:381-404), with
 * nodes, prim_indices and triangle records LEFT ON THE DEVICE: no host mirror is produced (no D2H of the tree, no second
 * upload), which is what a renderer that only traces on the GPU needs and what makes per-frame rebuilds affordable.
 * The trees are the same, byte for byte, as create_bvh / rtbvh_gpu_create_bvh_triangles + create_mbvh produce;
 * rtbvh_gpu_scene_tree_size / _read_nodes / _read_indices copy them out on demand.  _device takes device vertices. */
ResultCode rtbvh_gpu_scene_build(const float *vertices, size_t vertex_stride, size_t triangle_count, size_t prims_per_leaf,
                                 BvhType type, int want_mbvh, RTGpuScene *scene);
ResultCode rtbvh_gpu_scene_build_device(const float *d_vertices, size_t vertex_stride, size_t triangle_count,
                                        size_t prims_per_leaf, BvhType type, int want_mbvh, RTGpuScene *scene);
ResultCode rtbvh_gpu_scene_tree_size(RTGpuScene scene, RTTreeKind tree, uint32_t *node_count, uint32_t *index_count);
ResultCode rtbvh_gpu_scene_read_indices(RTGpuScene scene, RTTreeKind tree, uint32_t *out, size_t count);
/* The builders keep one device workspace per host thread between calls (it only grows: a 30 M-triangle build leaves
 * ~7 GB reserved) so that repeated builds never wait for the allocator; this gives it back. */
ResultCode rtbvh_gpu_trim_workspace(void);
/* Dynamic scenes (the step after build for animated geometry; Bvh::refit src/bvh.rs:176-205, FFI refit
 * rtbvh_ffi/src/lib.rs:519-538).  New vertex positions for the SAME triangles (same count and order): recomputes the
 * per-triangle boxes (Triangle::aabb, un-padded), refits the scene's Bvh in place (leaf = union of its primitives'
 * boxes, inner = union of its two children, each padded by 1e-4; topology fields kept), refreshes the slot boxes of the
 * scene's Mbvh from the refitted binary nodes — the result equals Mbvh::construct of the refitted Bvh byte for byte
 * (the reference never refreshes m_nodes) — and re-gathers the triangle records.  Nothing leaves the device; the host
 * mirrors behind RTBvh / RTMbvh are not touched (rtbvh_gpu_scene_read_nodes reads the device copy).  The scene must hold
 * the Bvh; an Mbvh in it must be the collapse of that Bvh.  The _device flavour takes device vertices and only
 * enqueues work on `stream`: traversal calls enqueued behind it on that stream see the refitted trees.
 * Ordering: both flavours are ordered against the scene's own host-buffer pipeline — batches submitted before the refit
 * (rtbvh_gpu_*_async included) finish on the old trees, batches submitted after it see the new ones — and consecutive
 * refits of one scene are chained (they share one scratch area), whatever streams they were given.  Device-pointer
 * traversal calls (*_device, *_device_scatter) run on the CALLER's streams and are ordered by the caller like any
 * stream-ordered work: use the refit's stream, or an event.  The blocking flavour returns when the refit has finished. */
ResultCode rtbvh_gpu_scene_refit(RTGpuScene scene, const float *vertices, size_t vertex_stride, size_t triangle_count);
ResultCode rtbvh_gpu_scene_refit_device(RTGpuScene scene, const float *d_vertices, size_t vertex_stride,
                                        size_t triangle_count, void *stream);
ResultCode rtbvh_gpu_scene_read_nodes(RTGpuScene scene, RTTreeKind tree, void *out, size_t bytes);
/* Incoherent batches (shadow / bounce rays): enable = 1 makes the single-ray calls trace every batch in Morton
 * order of (origin, direction) and scatter the results back.  Results are unchanged, bit for bit; only the
 * order in which the device works through the batch changes.  Default 0 (camera rays are coherent already). */
ResultCode rtbvh_gpu_scene_set_ray_sorting(RTGpuScene scene, int enable);
/* Image-ordered batches (primary rays: ray i = pixel (i % row_length, i / row_length), any number of frames back to back):
 * the device-pointer single-ray calls then hand the rays to the warps as 8x8 pixel tiles instead of 64-pixel row segments
 * — a work-order hint like ray sorting, at no cost (index arithmetic in the kernel, no sort): neighbouring lanes walk
 * neighbouring parts of the tree.  Results are unchanged, bit for bit, and stay at their rays' indices.  row_length must be
 * a multiple of 8; rays beyond the last whole band of 8 rows keep the linear order; 0 switches the hint off.  Ignored while
 * ray sorting is on and by the host-buffer calls (their rays arrive in index order). */
ResultCode rtbvh_gpu_scene_set_ray_tiling(RTGpuScene scene, uint32_t row_length);

/* ---- closest hit / any hit, host buffers (H2D + kernels + D2H inside the call) ------------------ */
ResultCode rtbvh_gpu_intersect(RTGpuScene scene, RTTreeKind tree, const RTRay *rays, size_t ray_count, RTHit *hits);
ResultCode rtbvh_gpu_occluded(RTGpuScene scene, RTTreeKind tree, const RTRay *rays, size_t ray_count,
                              uint8_t *occluded);
/* Asynchronous flavour of the two calls above for callers that stream batches (a renderer double-buffers its ray
 * and hit buffers): submit returns at once with a ticket; rtbvh_gpu_wait(scene, ticket) blocks until that batch's
 * records are in its output buffer (ticket 0: everything submitted so far) and reports Error if a traversal stack
 * overflowed.  Batches complete in submission order.  Consecutive submissions share the staging pipeline, so batch
 * k+1 uploads while batch k still traces and downloads.  The host buffers must stay untouched until the wait returns
 * and should be page-locked (rtbvh_gpu_host_alloc, cudaHostAlloc or cudaHostRegister); pageable buffers work but make
 * submit block.  At most 64 tickets may be outstanding per scene (an older one is waited for implicitly). */
ResultCode rtbvh_gpu_intersect_async(RTGpuScene scene, RTTreeKind tree, const RTRay *rays, size_t ray_count, RTHit *hits,
                                     uint64_t *ticket);
ResultCode rtbvh_gpu_occluded_async(RTGpuScene scene, RTTreeKind tree, const RTRay *rays, size_t ray_count,
                                    uint8_t *occluded, uint64_t *ticket);
ResultCode rtbvh_gpu_wait(RTGpuScene scene, uint64_t ticket);
/* Split ray input — the argument shape of the reference's FFI intersect (origin[3], direction[3], t:
 * rtbvh_ffi/src/lib.rs:551-581) for a batch: origins and directions are tightly packed float3 arrays (12 bytes per ray
 * each), t_min and t_max (the initial ray.t) apply to every ray.  Same results as the RTRay calls with those t_min / t;
 * 24 instead of 32 bytes per ray cross PCIe, which is what bounds the host-buffer path.  Ray sorting is not applied. */
ResultCode rtbvh_gpu_intersect_od(RTGpuScene scene, RTTreeKind tree, const float *origins, const float *directions,
                                  size_t ray_count, float t_min, float t_max, RTHit *hits);
ResultCode rtbvh_gpu_occluded_od(RTGpuScene scene, RTTreeKind tree, const float *origins, const float *directions,
                                 size_t ray_count, float t_min, float t_max, uint8_t *occluded);
ResultCode rtbvh_gpu_intersect_od_async(RTGpuScene scene, RTTreeKind tree, const float *origins, const float *directions,
                                        size_t ray_count, float t_min, float t_max, RTHit *hits, uint64_t *ticket);
ResultCode rtbvh_gpu_occluded_od_async(RTGpuScene scene, RTTreeKind tree, const float *origins, const float *directions,
                                       size_t ray_count, float t_min, float t_max, uint8_t *occluded, uint64_t *ticket);
ResultCode rtbvh_gpu_intersect_od_device(RTGpuScene scene, RTTreeKind tree, const float *d_origins,
                                         const float *d_directions, size_t ray_count, float t_min, float t_max,
                                         RTHit *d_hits, void *stream);
ResultCode rtbvh_gpu_host_alloc(size_t bytes, void **ptr); /* page-locked host memory */
ResultCode rtbvh_gpu_host_free(void *ptr);
/* Packets follow SpatialTriangle::intersect4 (eps 1e-6, t >= t_min): pass t_min = 1e-4f to match
 * examples/benchmark.rs:58.  occluded: 4 bytes per packet. */
ResultCode rtbvh_gpu_intersect_packets(RTGpuScene scene, RTTreeKind tree, const RTRayPacket4 *packets,
                                       size_t packet_count, float t_min, RTHitPacket4 *hits);
ResultCode rtbvh_gpu_occluded_packets(RTGpuScene scene, RTTreeKind tree, const RTRayPacket4 *packets,
                                      size_t packet_count, float t_min, uint8_t *occluded);

/* ---- same, device-resident buffers, asynchronous on `stream` ------------------------------------ */
ResultCode rtbvh_gpu_intersect_device(RTGpuScene scene, RTTreeKind tree, const RTRay *d_rays, size_t ray_count,
                                      RTHit *d_hits, void *stream);
ResultCode rtbvh_gpu_occluded_device(RTGpuScene scene, RTTreeKind tree, const RTRay *d_rays, size_t ray_count,
                                     uint8_t *d_occluded, void *stream);
ResultCode rtbvh_gpu_intersect_packets_device(RTGpuScene scene, RTTreeKind tree, const RTRayPacket4 *d_packets,
                                              size_t packet_count, float t_min, RTHitPacket4 *d_hits, void *stream);
ResultCode rtbvh_gpu_occluded_packets_device(RTGpuScene scene, RTTreeKind tree, const RTRayPacket4 *d_packets,
                                             size_t packet_count, float t_min, uint8_t *d_occluded, void *stream);
/* Nonzero if any ray of a previous *_device call on this scene needed more than the 128-entry
 * traversal stack (the reference's stack has 32 entries and panics / is UB beyond, src/iter.rs:25).
 * Synchronises the device.  Host-buffer calls return Error in that case. */
ResultCode rtbvh_gpu_scene_stack_overflowed(RTGpuScene scene, uint32_t *overflowed);

/* ---- multi-GPU gather fused into the traversal kernel ---------------------------------------------------------- */
/* One process per GPU.  Every rank creates a gather buffer (cudaMalloc + cudaIpc handle, 64 bytes), exchanges the
 * handles out of band (e.g. torch.distributed / MPI), opens its peers' buffers, and passes all destinations (own
 * pointer + opened peer pointers, at most 8) to the *_scatter calls: each result record is written the moment its ray
 * finishes to dests[k][dest_offset + i] for every k — P2P stores over NVLink / NVSwitch, overlapped with the traversal,
 * no collective afterwards.  d_hits / d_occluded (local copy in ray order) may be null.  Synchronise the ranks (a
 * barrier after the stream has drained) before reading a gather buffer. */
/* Scene replication (the tree is replicated, rays are sharded).  The scene is built or uploaded ONCE,
 * on one GPU; every other GPU gets a byte-identical replica copied device to device over NVLink / NVSwitch — the trees
 * never travel through host memory or a collective library.
 *   one process per GPU:  rank 0 calls _export (cudaIpc handles of the scene's device arrays + sizes, a 512-byte POD the
 *     caller ships out of band: MPI, torch.distributed, a pipe); every other rank calls _import with its own device
 *     current: the handles are opened, the arrays copied into allocations of the importing rank, the handles closed.  The
 *     exporting scene must stay alive (and must not be refitted) until every importer has returned — one barrier.
 *   one process, several GPUs:  _clone copies the scene onto `device` with cudaMemcpyPeer.
 * The replica is a full scene (traversal, refit, read-back), independent of the original afterwards. */
typedef struct RTGpuSceneExport {
  unsigned char bytes[512];
} RTGpuSceneExport;
ResultCode rtbvh_gpu_scene_export(RTGpuScene scene, RTGpuSceneExport *out);
ResultCode rtbvh_gpu_scene_import(const RTGpuSceneExport *exported, RTGpuScene *scene);
ResultCode rtbvh_gpu_scene_clone(RTGpuScene scene, int device, RTGpuScene *clone);

ResultCode rtbvh_gpu_peer_buffer_create(size_t bytes, void **d_ptr, unsigned char *handle64);
ResultCode rtbvh_gpu_peer_buffer_open(const unsigned char *handle64, void **d_ptr);
ResultCode rtbvh_gpu_peer_buffer_close(void *d_ptr);
ResultCode rtbvh_gpu_peer_buffer_free(void *d_ptr);
/* Step barrier across the ranks, enqueued on `stream` (one tiny kernel, no host involvement): flag_arrays[k] is rank
 * k's flag buffer (a zero-initialised peer buffer of >= 8 * count bytes; own pointer at index `rank`).  Publishes `value`
 * (strictly increasing per call, e.g. step + 1) to every rank and waits until every rank has published it: afterwards
 * all records scattered by kernels enqueued before the barrier on ANY rank are visible in this rank's gather buffer,
 * and every rank has finished the kernels it enqueued before ITS barrier (so with two gather buffers used alternately
 * the buffer of step k-1 may be overwritten by step k+1). */
ResultCode rtbvh_gpu_peer_barrier(void *const *flag_arrays, int count, int rank, uint64_t value, void *stream);
ResultCode rtbvh_gpu_intersect_device_scatter(RTGpuScene scene, RTTreeKind tree, const RTRay *d_rays, size_t ray_count,
                                              RTHit *d_hits, void *const *dests, int dest_count, size_t dest_offset,
                                              void *stream);
ResultCode rtbvh_gpu_occluded_device_scatter(RTGpuScene scene, RTTreeKind tree, const RTRay *d_rays, size_t ray_count,
                                             uint8_t *d_occluded, void *const *dests, int dest_count, size_t dest_offset,
                                             void *stream);

/* ---- builders ---------------------------------------------------------------------------------- */
/* create_bvh (rtbvh.h) is the drop-in builder entry: it takes what rtbvh_ffi takes (aabbs + centers).
 * This variant is Builder{aabbs: None, primitives: &[Triangle]}.construct_* (src/bvh.rs:87-137) for
 * triangle primitives: Primitive::aabb (un-padded grow of the three vertices) and Primitive::center
 * ((v0+v1+v2) * (1/3)) of the bench Triangle (shared/src/lib.rs:27-39) are computed on the device.
 * The result is stored like create_bvh's (free with free_bvh, collapse with create_mbvh). */
ResultCode rtbvh_gpu_create_bvh_triangles(const float *vertices, size_t vertex_stride, size_t triangle_count,
                                          size_t prims_per_leaf, BvhType bvh_type, RTBvh *result);
/* Mbvh::construct (src/bvh.rs:381-404) on the GPU for a binary tree that is NOT in this library's table, e.g. a
 * reference-built spatial-split tree: bvh->nodes / bvh->indices are trusted host pointers.  Free with free_mbvh. */
ResultCode rtbvh_gpu_create_mbvh_from(const RTBvh *bvh, RTMbvh *mbvh);
/* Timing of the last create_bvh / create_mbvh / refit / rtbvh_gpu_create_bvh_triangles on this thread:
 * device_ms = kernels only (CUDA events, inputs resident -> tree resident), total_ms = incl. the H2D of the
 * inputs and the D2H of the host mirror; iterations = LOCB clustering iterations. */
ResultCode rtbvh_gpu_last_build_stats(double *device_ms, double *total_ms, uint32_t *iterations);

/* ---- primary rays generated on the device: nothing but the records crosses PCIe ------------------- */
/* `frames` frames of width x height camera rays — the rays rtbvh_gpu_generate_camera_rays_device writes for frames
 * first_frame .. first_frame + frames - 1, frame after frame, row-major — are generated in the staging slots, traced (in 8x8
 * pixel tiles when width is a multiple of 8) and their records copied to the HOST buffer (hits: width*height*frames RTHit;
 * occluded: as many bytes), chunk by chunk on the scene's pipeline streams.  Returns a ticket at once; rtbvh_gpu_wait as for
 * the *_async calls.  This is the loop of examples/benchmark.rs (generate_ray per pixel, then traverse) without the 24-32
 * bytes per ray of upload the host-ray calls pay: what remains on the bus is 8 bytes (1 byte) per ray of results. */
ResultCode rtbvh_gpu_intersect_camera_async(RTGpuScene scene, RTTreeKind tree, const float pos[3], const float p1[3],
                                            const float right[3], const float up[3], uint32_t width, uint32_t height,
                                            uint64_t jitter_seed, uint64_t first_frame, uint32_t frames, RTHit *hits,
                                            uint64_t *ticket);
ResultCode rtbvh_gpu_occluded_camera_async(RTGpuScene scene, RTTreeKind tree, const float pos[3], const float p1[3],
                                           const float right[3], const float up[3], uint32_t width, uint32_t height,
                                           uint64_t jitter_seed, uint64_t first_frame, uint32_t frames, uint8_t *occluded,
                                           uint64_t *ticket);

/* ---- workload helper: CameraView3D::generate_ray on the device (shared/src/lib.rs:157-165) ----- */
/* Writes width*rows rays for pixel rows [row0, row0+rows): u = (x + jx) / width, v = (y + jy) / height,
 * direction = normalize(p1 + u*right + v*up - pos); (jx, jy) = 0 when jitter_seed == 0, else
 * splitmix64-hashed sub-pixel offsets of (jitter_seed, frame, pixel).  t_min = 1e-4, t = 1e34. */
ResultCode rtbvh_gpu_generate_camera_rays_device(const float pos[3], const float p1[3], const float right[3],
                                                 const float up[3], uint32_t width, uint32_t height, uint32_t row0,
                                                 uint32_t rows, uint64_t jitter_seed, uint64_t frame, RTRay *d_rays,
                                                 void *stream);

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* RTBVH_GPU_H */
