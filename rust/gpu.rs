//! `src/gpu.rs` — B200 back end for the `rtbvh` crate (drop this file into the reference's `src/`,
//! apply `bvh_rs.patch`, add `build.rs`).
//!
//! STATUS: written against `include/rtbvh.h` + `include/rtbvh_gpu.h` of this repository and the crate's public surface
//! (`src/bvh.rs:10-56, 143-147, 320-324`), **not compiled here** — the image has no `rustc`/`cargo`.
//! `tests/test_rust_shim_decls.py` checks mechanically that every `extern "C"` declaration below matches the C
//! headers (name, arity, argument and return types); the C++ mirror `include/rtbvh.hpp` is the compiled and tested
//! equivalent of the safe layer.
//!
//! What it does:
//! * `sys` — the raw C ABI of `librtbvh_rs.so` (all entry points of both headers).
//! * `build_on_gpu` / `collapse_on_gpu` — what `Builder::construct_binned_sah`,
//!   `Builder::construct_locally_ordered_clustered` (`src/bvh.rs:87-137`) and `Mbvh::construct` (`src/bvh.rs:381-404`)
//!   call instead of the CPU builders.  Validation and `BuildError`s stay in `bvh.rs`, unchanged.  The trees come back
//!   in the crate's own node formats, so `Bvh::nodes()`, `validate()`, `into_raw()`, serde and the host iterators
//!   (`src/iter.rs`, `src/iter_indices.rs`) keep working on them untouched.
//! * `GpuScene` — the batched form of the loop every caller writes around the iterators
//!   (`examples/benchmark.rs:25-31`, `:55-61`): closest hit, any hit, packets, dynamic refit, asynchronous batches.
//!
//! There is no CPU fallback: without a CUDA device the constructors panic with the library's error string
//! (`BuildError` has no variant for it, `src/bvh.rs:26-29`) and `GpuScene` calls return `GpuError`.

use std::ffi::CStr;
use std::marker::PhantomData;
use std::num::NonZeroUsize;
use std::ops::{Deref, DerefMut};
use std::os::raw::{c_char, c_int, c_uchar, c_void};

use rayon::prelude::*;

use crate::{Aabb, BuildError, BuildType, Bvh, BvhNode, Mbvh, MbvhNode, Primitive, Ray, RayPacket4, SpatialTriangle};

/// Raw C ABI.  Layouts: `include/rtbvh.h` (== cbindgen output of `rtbvh_ffi/src/lib.rs`) and `include/rtbvh_gpu.h`.
#[allow(non_camel_case_types, non_snake_case, dead_code)]
pub mod sys {
    use super::*;

    /// `rtbvh_ffi/src/lib.rs:17-25`.  A transparent integer, not a Rust enum: a value outside the list must not be UB.
    #[repr(transparent)]
    #[derive(Debug, Copy, Clone, PartialEq, Eq, Hash)]
    pub struct ResultCode(pub c_int);
    impl ResultCode {
        pub const OK: ResultCode = ResultCode(0);
        pub const ERROR: ResultCode = ResultCode(1);
        pub const NO_PRIMITIVES: ResultCode = ResultCode(2);
        pub const INEQUAL_AABBS_AND_PRIMITIVES: ResultCode = ResultCode(3);
        pub const NAN: ResultCode = ResultCode(4);
    }

    /// `rtbvh_ffi/src/lib.rs:129-133` (`#[repr(u32)]`).
    pub type BvhType = u32;
    pub const LOCALLY_ORDERED_CLUSTERED: BvhType = 0;
    pub const BINNED_SAH: BvhType = 1;

    /// `RTTreeKind` (C enum, int sized).
    pub type RTTreeKind = c_int;
    pub const RT_TREE_BVH: RTTreeKind = 0;
    pub const RT_TREE_MBVH: RTTreeKind = 1;

    pub const RT_NO_HIT: u32 = 0xFFFF_FFFF;

    /// `RTAabb` is `Aabb<i32>` (`src/aabb.rs:13-20`), `RTBvhNode` is `BvhNode`, `RTMbvhNode` is `MbvhNode`
    /// (`rtbvh_ffi` `same_size` test: 32 / 32 / 128 bytes), so the crate's own types cross the boundary.
    pub type RTAabb = Aabb<i32>;
    pub type RTBvhNode = BvhNode;
    pub type RTMbvhNode = MbvhNode;

    /// `rtbvh_ffi/src/lib.rs:210-220`; passed by value.
    #[repr(C)]
    #[derive(Debug, Copy, Clone)]
    pub struct RTBvh {
        pub id: u32,
        pub node_count: u32,
        pub nodes: *const RTBvhNode,
        pub index_count: u32,
        pub indices: *const u32,
    }

    /// `rtbvh_ffi/src/lib.rs:234-244`.
    #[repr(C)]
    #[derive(Debug, Copy, Clone)]
    pub struct RTMbvh {
        pub id: u32,
        pub node_count: u32,
        pub nodes: *const RTMbvhNode,
        pub index_count: u32,
        pub indices: *const u32,
    }

    impl Default for RTBvh {
        fn default() -> Self {
            RTBvh { id: u32::MAX, node_count: 0, nodes: std::ptr::null(), index_count: 0, indices: std::ptr::null() }
        }
    }
    impl Default for RTMbvh {
        fn default() -> Self {
            RTMbvh { id: u32::MAX, node_count: 0, nodes: std::ptr::null(), index_count: 0, indices: std::ptr::null() }
        }
    }

    /// First 32 bytes of `Ray` (`src/ray.rs:9-16`); the derived fields are recomputed on the device.
    #[repr(C)]
    #[derive(Debug, Copy, Clone, PartialEq)]
    pub struct RTRay {
        pub origin: [f32; 3],
        pub t_min: f32,
        pub direction: [f32; 3],
        pub t: f32,
    }

    #[repr(C)]
    #[derive(Debug, Copy, Clone, PartialEq)]
    pub struct RTHit {
        pub t: f32,
        pub prim: u32,
    }

    /// `RayPacket4` (`src/ray.rs:47-61`) without `inv_direction_*`.
    #[repr(C)]
    #[derive(Debug, Copy, Clone, PartialEq)]
    pub struct RTRayPacket4 {
        pub origin_x: [f32; 4],
        pub origin_y: [f32; 4],
        pub origin_z: [f32; 4],
        pub direction_x: [f32; 4],
        pub direction_y: [f32; 4],
        pub direction_z: [f32; 4],
        pub t: [f32; 4],
    }

    #[repr(C)]
    #[derive(Debug, Copy, Clone, PartialEq)]
    pub struct RTHitPacket4 {
        pub t: [f32; 4],
        pub prim: [u32; 4],
    }

    pub type RTGpuScene = u64;

    /// What `rtbvh_gpu_scene_export` fills and `rtbvh_gpu_scene_import` reads: cudaIpc handles + sizes, shipped between the
    /// per-GPU processes by whatever the application already uses (MPI, a pipe).
    #[repr(C)]
    #[derive(Copy, Clone)]
    pub struct RTGpuSceneExport {
        pub bytes: [u8; 512],
    }
    /// `(prim_id, inout t, user_data) -> stop` (`rtbvh_ffi/src/lib.rs:541-558`).
    pub type RTIntersectCallback = Option<unsafe extern "C" fn(u32, *mut f32, *mut c_void) -> bool>;

    #[link(name = "rtbvh_rs")]
    extern "C" {
        // ---- include/rtbvh.h: the reference's ten entry points -----------------------------------------------
        pub fn create_spatial_Bvh(aabbs: *const RTAabb, prim_count: usize, centers: *const f32, stride: usize, vertices: *const f32, vertex_stride: usize, triangle_stride: usize, prims_per_leaf: u32, result: *mut RTBvh) -> ResultCode;
        pub fn create_bvh(aabbs: *const RTAabb, prim_count: usize, centers: *const f32, center_stride: usize, prims_per_leaf: usize, bvh_type: BvhType, result: *mut RTBvh) -> ResultCode;
        pub fn create_mbvh(bvh: RTBvh, mbvh: *mut RTMbvh) -> ResultCode;
        pub fn refit(aabbs: *const RTAabb, bvh: RTBvh) -> ResultCode;
        pub fn intersect(bvh: RTBvh, origin: *const f32, direction: *const f32, t: *mut f32, user_data: *mut c_void, intersect: RTIntersectCallback) -> ResultCode;
        pub fn intersect_packet(bvh: RTBvh, origin_x: *const f32, origin_y: *const f32, origin_z: *const f32, direction_x: *const f32, direction_y: *const f32, direction_z: *const f32, t: *mut f32, user_data: *mut c_void, intersect: RTIntersectCallback) -> ResultCode;
        pub fn intersect_mbvh(bvh: RTMbvh, origin: *const f32, direction: *const f32, t: *mut f32, user_data: *mut c_void, intersect: RTIntersectCallback) -> ResultCode;
        pub fn intersect_mbvh_packet(bvh: RTMbvh, origin_x: *const f32, origin_y: *const f32, origin_z: *const f32, direction_x: *const f32, direction_y: *const f32, direction_z: *const f32, t: *mut f32, user_data: *mut c_void, intersect: RTIntersectCallback) -> ResultCode;
        pub fn free_bvh(bvh: RTBvh);
        pub fn free_mbvh(bvh: RTMbvh);

        // ---- include/rtbvh_gpu.h: devices ----------------------------------------------------------------------
        pub fn rtbvh_gpu_device_count() -> c_int;
        pub fn rtbvh_gpu_set_device(device: c_int) -> ResultCode;
        pub fn rtbvh_gpu_last_error() -> *const c_char;

        // ---- scenes --------------------------------------------------------------------------------------------
        pub fn rtbvh_gpu_scene_create(bvh: *const RTBvh, mbvh: *const RTMbvh, vertices: *const f32, vertex_stride: usize, triangle_count: usize, scene: *mut RTGpuScene) -> ResultCode;
        pub fn rtbvh_gpu_scene_free(scene: RTGpuScene) -> ResultCode;
        pub fn rtbvh_gpu_scene_build(vertices: *const f32, vertex_stride: usize, triangle_count: usize, prims_per_leaf: usize, type_: BvhType, want_mbvh: c_int, scene: *mut RTGpuScene) -> ResultCode;
        pub fn rtbvh_gpu_scene_build_device(d_vertices: *const f32, vertex_stride: usize, triangle_count: usize, prims_per_leaf: usize, type_: BvhType, want_mbvh: c_int, scene: *mut RTGpuScene) -> ResultCode;
        pub fn rtbvh_gpu_scene_tree_size(scene: RTGpuScene, tree: RTTreeKind, node_count: *mut u32, index_count: *mut u32) -> ResultCode;
        pub fn rtbvh_gpu_scene_read_indices(scene: RTGpuScene, tree: RTTreeKind, out: *mut u32, count: usize) -> ResultCode;
        pub fn rtbvh_gpu_trim_workspace() -> ResultCode;
        pub fn rtbvh_gpu_scene_refit(scene: RTGpuScene, vertices: *const f32, vertex_stride: usize, triangle_count: usize) -> ResultCode;
        pub fn rtbvh_gpu_scene_refit_device(scene: RTGpuScene, d_vertices: *const f32, vertex_stride: usize, triangle_count: usize, stream: *mut c_void) -> ResultCode;
        pub fn rtbvh_gpu_scene_read_nodes(scene: RTGpuScene, tree: RTTreeKind, out: *mut c_void, bytes: usize) -> ResultCode;
        pub fn rtbvh_gpu_scene_set_ray_sorting(scene: RTGpuScene, enable: c_int) -> ResultCode;
        pub fn rtbvh_gpu_scene_set_ray_tiling(scene: RTGpuScene, row_length: u32) -> ResultCode;

        // ---- closest hit / any hit, host buffers ---------------------------------------------------------------
        pub fn rtbvh_gpu_intersect(scene: RTGpuScene, tree: RTTreeKind, rays: *const RTRay, ray_count: usize, hits: *mut RTHit) -> ResultCode;
        pub fn rtbvh_gpu_occluded(scene: RTGpuScene, tree: RTTreeKind, rays: *const RTRay, ray_count: usize, occluded: *mut u8) -> ResultCode;
        pub fn rtbvh_gpu_intersect_async(scene: RTGpuScene, tree: RTTreeKind, rays: *const RTRay, ray_count: usize, hits: *mut RTHit, ticket: *mut u64) -> ResultCode;
        pub fn rtbvh_gpu_occluded_async(scene: RTGpuScene, tree: RTTreeKind, rays: *const RTRay, ray_count: usize, occluded: *mut u8, ticket: *mut u64) -> ResultCode;
        pub fn rtbvh_gpu_wait(scene: RTGpuScene, ticket: u64) -> ResultCode;
        pub fn rtbvh_gpu_intersect_od(scene: RTGpuScene, tree: RTTreeKind, origins: *const f32, directions: *const f32, ray_count: usize, t_min: f32, t_max: f32, hits: *mut RTHit) -> ResultCode;
        pub fn rtbvh_gpu_occluded_od(scene: RTGpuScene, tree: RTTreeKind, origins: *const f32, directions: *const f32, ray_count: usize, t_min: f32, t_max: f32, occluded: *mut u8) -> ResultCode;
        pub fn rtbvh_gpu_intersect_od_async(scene: RTGpuScene, tree: RTTreeKind, origins: *const f32, directions: *const f32, ray_count: usize, t_min: f32, t_max: f32, hits: *mut RTHit, ticket: *mut u64) -> ResultCode;
        pub fn rtbvh_gpu_occluded_od_async(scene: RTGpuScene, tree: RTTreeKind, origins: *const f32, directions: *const f32, ray_count: usize, t_min: f32, t_max: f32, occluded: *mut u8, ticket: *mut u64) -> ResultCode;
        pub fn rtbvh_gpu_intersect_od_device(scene: RTGpuScene, tree: RTTreeKind, d_origins: *const f32, d_directions: *const f32, ray_count: usize, t_min: f32, t_max: f32, d_hits: *mut RTHit, stream: *mut c_void) -> ResultCode;
        pub fn rtbvh_gpu_host_alloc(bytes: usize, ptr: *mut *mut c_void) -> ResultCode;
        pub fn rtbvh_gpu_host_free(ptr: *mut c_void) -> ResultCode;
        pub fn rtbvh_gpu_intersect_packets(scene: RTGpuScene, tree: RTTreeKind, packets: *const RTRayPacket4, packet_count: usize, t_min: f32, hits: *mut RTHitPacket4) -> ResultCode;
        pub fn rtbvh_gpu_occluded_packets(scene: RTGpuScene, tree: RTTreeKind, packets: *const RTRayPacket4, packet_count: usize, t_min: f32, occluded: *mut u8) -> ResultCode;

        // ---- device-resident buffers, asynchronous on `stream` (a cudaStream_t) ---------------------------------
        pub fn rtbvh_gpu_intersect_device(scene: RTGpuScene, tree: RTTreeKind, d_rays: *const RTRay, ray_count: usize, d_hits: *mut RTHit, stream: *mut c_void) -> ResultCode;
        pub fn rtbvh_gpu_occluded_device(scene: RTGpuScene, tree: RTTreeKind, d_rays: *const RTRay, ray_count: usize, d_occluded: *mut u8, stream: *mut c_void) -> ResultCode;
        pub fn rtbvh_gpu_intersect_packets_device(scene: RTGpuScene, tree: RTTreeKind, d_packets: *const RTRayPacket4, packet_count: usize, t_min: f32, d_hits: *mut RTHitPacket4, stream: *mut c_void) -> ResultCode;
        pub fn rtbvh_gpu_occluded_packets_device(scene: RTGpuScene, tree: RTTreeKind, d_packets: *const RTRayPacket4, packet_count: usize, t_min: f32, d_occluded: *mut u8, stream: *mut c_void) -> ResultCode;
        pub fn rtbvh_gpu_scene_stack_overflowed(scene: RTGpuScene, overflowed: *mut u32) -> ResultCode;

        // ---- multi-GPU gather fused into the traversal kernel ---------------------------------------------------
        pub fn rtbvh_gpu_scene_export(scene: RTGpuScene, out: *mut RTGpuSceneExport) -> ResultCode;
        pub fn rtbvh_gpu_scene_import(exported: *const RTGpuSceneExport, scene: *mut RTGpuScene) -> ResultCode;
        pub fn rtbvh_gpu_scene_clone(scene: RTGpuScene, device: c_int, clone: *mut RTGpuScene) -> ResultCode;
        pub fn rtbvh_gpu_peer_buffer_create(bytes: usize, d_ptr: *mut *mut c_void, handle64: *mut c_uchar) -> ResultCode;
        pub fn rtbvh_gpu_peer_buffer_open(handle64: *const c_uchar, d_ptr: *mut *mut c_void) -> ResultCode;
        pub fn rtbvh_gpu_peer_buffer_close(d_ptr: *mut c_void) -> ResultCode;
        pub fn rtbvh_gpu_peer_buffer_free(d_ptr: *mut c_void) -> ResultCode;
        pub fn rtbvh_gpu_peer_barrier(flag_arrays: *const *mut c_void, count: c_int, rank: c_int, value: u64, stream: *mut c_void) -> ResultCode;
        pub fn rtbvh_gpu_intersect_device_scatter(scene: RTGpuScene, tree: RTTreeKind, d_rays: *const RTRay, ray_count: usize, d_hits: *mut RTHit, dests: *const *mut c_void, dest_count: c_int, dest_offset: usize, stream: *mut c_void) -> ResultCode;
        pub fn rtbvh_gpu_occluded_device_scatter(scene: RTGpuScene, tree: RTTreeKind, d_rays: *const RTRay, ray_count: usize, d_occluded: *mut u8, dests: *const *mut c_void, dest_count: c_int, dest_offset: usize, stream: *mut c_void) -> ResultCode;

        // ---- builders --------------------------------------------------------------------------------------------
        pub fn rtbvh_gpu_create_bvh_triangles(vertices: *const f32, vertex_stride: usize, triangle_count: usize, prims_per_leaf: usize, bvh_type: BvhType, result: *mut RTBvh) -> ResultCode;
        pub fn rtbvh_gpu_create_mbvh_from(bvh: *const RTBvh, mbvh: *mut RTMbvh) -> ResultCode;
        pub fn rtbvh_gpu_last_build_stats(device_ms: *mut f64, total_ms: *mut f64, iterations: *mut u32) -> ResultCode;

        // ---- workload helper --------------------------------------------------------------------------------------
        pub fn rtbvh_gpu_intersect_camera_async(scene: RTGpuScene, tree: RTTreeKind, pos: *const f32, p1: *const f32, right: *const f32, up: *const f32, width: u32, height: u32, jitter_seed: u64, first_frame: u64, frames: u32, hits: *mut RTHit, ticket: *mut u64) -> ResultCode;
        pub fn rtbvh_gpu_occluded_camera_async(scene: RTGpuScene, tree: RTTreeKind, pos: *const f32, p1: *const f32, right: *const f32, up: *const f32, width: u32, height: u32, jitter_seed: u64, first_frame: u64, frames: u32, occluded: *mut u8, ticket: *mut u64) -> ResultCode;
        pub fn rtbvh_gpu_generate_camera_rays_device(pos: *const f32, p1: *const f32, right: *const f32, up: *const f32, width: u32, height: u32, row0: u32, rows: u32, jitter_seed: u64, frame: u64, d_rays: *mut RTRay, stream: *mut c_void) -> ResultCode;
    }
}

use sys::ResultCode;

// ------------------------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------------------------

/// A failed `rtbvh_gpu_*` call: the `ResultCode` and the library's per-thread message (`rtbvh_gpu_last_error`).
#[derive(Debug, Clone, PartialEq, Eq)]
pub struct GpuError {
    pub code: i32,
    pub message: String,
}

impl std::fmt::Display for GpuError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "rtbvh gpu error (ResultCode {}): {}", self.code, self.message)
    }
}

impl std::error::Error for GpuError {}

fn last_error() -> String {
    // The pointer is owned by the library (thread local) and valid until the next call on this thread.
    let p = unsafe { sys::rtbvh_gpu_last_error() };
    if p.is_null() {
        String::new()
    } else {
        unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned()
    }
}

fn check(rc: ResultCode) -> Result<(), GpuError> {
    if rc == ResultCode::OK {
        Ok(())
    } else {
        Err(GpuError { code: rc.0 as i32, message: last_error() })
    }
}

/// Number of CUDA devices the library sees (0: no driver / no device; every other call then fails).
pub fn device_count() -> usize {
    unsafe { sys::rtbvh_gpu_device_count() }.max(0) as usize
}

/// Device used by subsequent calls of this thread (one process per GPU: call once with the local rank).
pub fn set_device(device: usize) -> Result<(), GpuError> {
    check(unsafe { sys::rtbvh_gpu_set_device(device as c_int) })
}

/// Gives the per-thread builder workspace back to the driver.
pub fn trim_workspace() -> Result<(), GpuError> {
    check(unsafe { sys::rtbvh_gpu_trim_workspace() })
}

/// `(device_ms, total_ms, locb_iterations)` of the last build / collapse / refit on this thread.
pub fn last_build_stats() -> Result<(f64, f64, u32), GpuError> {
    let (mut d, mut t, mut i) = (0f64, 0f64, 0u32);
    check(unsafe { sys::rtbvh_gpu_last_build_stats(&mut d, &mut t, &mut i) })?;
    Ok((d, t, i))
}

// ------------------------------------------------------------------------------------------------------------------
// builders: what src/bvh.rs calls
// ------------------------------------------------------------------------------------------------------------------

fn bvh_type_of(build_type: BuildType) -> sys::BvhType {
    match build_type {
        BuildType::LocallyOrderedClustered => sys::LOCALLY_ORDERED_CLUSTERED,
        BuildType::BinnedSAH => sys::BINNED_SAH,
        other => panic!("rtbvh: {:?} trees are not built on the GPU", other),
    }
}

/// Copies a library-owned tree into crate-owned vectors and releases the library's entry.
unsafe fn take_bvh(out: sys::RTBvh, build_type: BuildType) -> Bvh {
    let nodes = std::slice::from_raw_parts(out.nodes, out.node_count as usize).to_vec();
    let prim_indices = std::slice::from_raw_parts(out.indices, out.index_count as usize).to_vec();
    sys::free_bvh(out);
    Bvh { nodes, prim_indices, build_type }
}

/// The build behind `Builder::construct_binned_sah` / `construct_locally_ordered_clustered`
/// (`src/bvh.rs:87-137`), called by them *after* their own validation.  Like the reference builders it sees only
/// `Primitive::center()` and the aabbs (`src/builders/binned_sah.rs:80-111`, `locb.rs:18-46`); both are gathered into
/// flat arrays (12 + 32 bytes per primitive) with rayon, exactly what `rtbvh_ffi::create_bvh` receives
/// (`rtbvh_ffi/src/lib.rs:428-493`).  `primitives_per_leaf` is ignored by LOCB, as in the reference.
pub(crate) fn build_on_gpu<T: Primitive<i32>>(
    aabbs: &[Aabb<i32>],
    primitives: &[T],
    primitives_per_leaf: Option<NonZeroUsize>,
    build_type: BuildType,
) -> Result<Bvh, BuildError> {
    debug_assert_eq!(aabbs.len(), primitives.len());
    let centers: Vec<[f32; 3]> = primitives
        .par_iter()
        .map(|p| {
            let c = p.center();
            [c.x, c.y, c.z]
        })
        .collect();
    let mut out = sys::RTBvh::default();
    // `Aabb` is `repr(C, align(16))`, so the slice satisfies the ABI's alignment requirement as it is.
    let rc = unsafe {
        sys::create_bvh(
            aabbs.as_ptr(),
            primitives.len(),
            centers.as_ptr() as *const f32,
            std::mem::size_of::<[f32; 3]>(),
            primitives_per_leaf.map_or(0, NonZeroUsize::get),
            bvh_type_of(build_type),
            &mut out,
        )
    };
    match rc {
        ResultCode::OK => Ok(unsafe { take_bvh(out, build_type) }),
        ResultCode::NO_PRIMITIVES => Err(BuildError::NoPrimitives),
        ResultCode::INEQUAL_AABBS_AND_PRIMITIVES => Err(BuildError::InequalAabbsAndPrimitives(aabbs.len(), primitives.len())),
        other => panic!("rtbvh: GPU build failed (ResultCode {}): {}", other.0, last_error()),
    }
}

/// `Builder{aabbs: None, ..}` for triangle primitives without the host-side gather of aabbs and centers: both are
/// computed on the device from the vertices (`Triangle::aabb` / `Triangle::center` of `shared/src/lib.rs:27-39`:
/// un-padded box of the three vertices, `(v0 + v1 + v2) * (1/3)`).  Only valid for primitives whose `Primitive`
/// impl is that canonical one.
pub fn build_triangles_on_gpu<T: SpatialTriangle + Sync>(
    triangles: &[T],
    primitives_per_leaf: Option<NonZeroUsize>,
    build_type: BuildType,
) -> Result<Bvh, BuildError> {
    if triangles.is_empty() {
        return Err(BuildError::NoPrimitives);
    }
    let vertices = gather_vertices(triangles);
    let mut out = sys::RTBvh::default();
    let rc = unsafe {
        sys::rtbvh_gpu_create_bvh_triangles(
            vertices.as_ptr() as *const f32,
            std::mem::size_of::<[f32; 3]>(),
            triangles.len(),
            primitives_per_leaf.map_or(0, NonZeroUsize::get),
            bvh_type_of(build_type),
            &mut out,
        )
    };
    match rc {
        ResultCode::OK => Ok(unsafe { take_bvh(out, build_type) }),
        ResultCode::NO_PRIMITIVES => Err(BuildError::NoPrimitives),
        other => panic!("rtbvh: GPU build failed (ResultCode {}): {}", other.0, last_error()),
    }
}

/// The collapse behind `Mbvh::construct` (`src/bvh.rs:381-404`, `MbvhNode::merge_nodes` `src/mbvh_node.rs:297-411`):
/// byte-identical 4-wide nodes, produced on the device from the crate-owned binary tree.
pub(crate) fn collapse_on_gpu(bvh: &Bvh) -> Mbvh {
    if bvh.nodes.is_empty() {
        return Mbvh::default();
    }
    let src = raw_bvh(bvh);
    let mut out = sys::RTMbvh::default();
    let rc = unsafe { sys::rtbvh_gpu_create_mbvh_from(&src, &mut out) };
    if rc != ResultCode::OK {
        panic!("rtbvh: GPU collapse failed (ResultCode {}): {}", rc.0, last_error());
    }
    let m_nodes = unsafe { std::slice::from_raw_parts(out.nodes, out.node_count as usize) }.to_vec();
    unsafe { sys::free_mbvh(out) };
    Mbvh { nodes: bvh.nodes.clone(), m_nodes, prim_indices: bvh.prim_indices.clone() }
}

fn raw_bvh(bvh: &Bvh) -> sys::RTBvh {
    sys::RTBvh {
        id: u32::MAX, // not in the library's table: the pointers are trusted, like the reference's intersect* do
        node_count: bvh.nodes.len() as u32,
        nodes: bvh.nodes.as_ptr(),
        index_count: bvh.prim_indices.len() as u32,
        indices: bvh.prim_indices.as_ptr(),
    }
}

fn raw_mbvh(mbvh: &Mbvh) -> sys::RTMbvh {
    sys::RTMbvh {
        id: u32::MAX,
        node_count: mbvh.m_nodes.len() as u32,
        nodes: mbvh.m_nodes.as_ptr(),
        index_count: mbvh.prim_indices.len() as u32,
        indices: mbvh.prim_indices.as_ptr(),
    }
}

fn gather_vertices<T: SpatialTriangle + Sync>(triangles: &[T]) -> Vec<[f32; 3]> {
    let mut vertices = vec![[0f32; 3]; 3 * triangles.len()];
    vertices.par_chunks_mut(3).zip(triangles.par_iter()).for_each(|(dst, tri)| {
        let (a, b, c) = (tri.vertex0(), tri.vertex1(), tri.vertex2());
        dst[0] = [a.x, a.y, a.z];
        dst[1] = [b.x, b.y, b.z];
        dst[2] = [c.x, c.y, c.z];
    });
    vertices
}

// ------------------------------------------------------------------------------------------------------------------
// batched traversal
// ------------------------------------------------------------------------------------------------------------------

/// Which of a scene's trees a batch walks: `Bvh` follows `BvhIndexIterator` / `BvhPacketIndexIterator`,
/// `Mbvh` follows `MbvhIndexIterator` / `MbvhPacketIndexIterator` (`src/iter_indices.rs`), visit for visit.
#[derive(Debug, Copy, Clone, PartialEq, Eq, Hash)]
pub enum Tree {
    Bvh,
    Mbvh,
}

impl Tree {
    fn raw(self) -> sys::RTTreeKind {
        match self {
            Tree::Bvh => sys::RT_TREE_BVH,
            Tree::Mbvh => sys::RT_TREE_MBVH,
        }
    }
}

/// Result of one ray: `t` as the reference loop leaves it in `ray.t` (the input `t` on a miss) and the primitive
/// that produced it (lowest id among exactly equal `t`), `None` on a miss.
#[derive(Debug, Copy, Clone, PartialEq)]
pub struct Hit {
    pub t: f32,
    pub prim: Option<u32>,
}

impl From<sys::RTHit> for Hit {
    fn from(h: sys::RTHit) -> Self {
        Hit { t: h.t, prim: if h.prim == sys::RT_NO_HIT { None } else { Some(h.prim) } }
    }
}

impl From<&Ray> for sys::RTRay {
    fn from(r: &Ray) -> Self {
        sys::RTRay {
            origin: [r.origin.x, r.origin.y, r.origin.z],
            t_min: r.t_min,
            direction: [r.direction.x, r.direction.y, r.direction.z],
            t: r.t,
        }
    }
}

impl From<&RayPacket4> for sys::RTRayPacket4 {
    fn from(p: &RayPacket4) -> Self {
        sys::RTRayPacket4 {
            origin_x: p.origin_x.into(),
            origin_y: p.origin_y.into(),
            origin_z: p.origin_z.into(),
            direction_x: p.direction_x.into(),
            direction_y: p.direction_y.into(),
            direction_z: p.direction_z.into(),
            t: p.t.into(),
        }
    }
}

/// A scene resident on one GPU: the tree(s), uploaded unchanged, and the triangles.
///
/// ```ignore
/// // examples/benchmark.rs:25-31 — `for (triangle, r) in bvh.traverse_iter(&mut ray, &triangles) { triangle.intersect(r); }`
/// let scene = GpuScene::new(Some(&bvh), Some(&mbvh), &triangles)?;
/// let hits = scene.intersect(Tree::Mbvh, &mut rays)?;      // ray.t updated, hits[i].prim = the triangle
/// ```
pub struct GpuScene {
    handle: sys::RTGpuScene,
    triangle_count: usize,
    // Scenes keep per-thread device selection and a copy pipeline: usable from several threads (the library
    // serialises them), but the handle is released exactly once.
    _not_copy: PhantomData<*const ()>,
}

unsafe impl Send for GpuScene {}
unsafe impl Sync for GpuScene {}

impl GpuScene {
    /// Uploads crate-built (or deserialised, or reference-built) trees unchanged, plus the triangles.
    /// Either tree may be `None`, not both.  `triangles[i]` must be primitive `i` of the build.
    pub fn new<T: SpatialTriangle + Sync>(bvh: Option<&Bvh>, mbvh: Option<&Mbvh>, triangles: &[T]) -> Result<Self, GpuError> {
        let vertices = gather_vertices(triangles);
        let rb = bvh.map(raw_bvh);
        let rm = mbvh.map(raw_mbvh);
        let mut handle: sys::RTGpuScene = 0;
        check(unsafe {
            sys::rtbvh_gpu_scene_create(
                rb.as_ref().map_or(std::ptr::null(), |b| b as *const sys::RTBvh),
                rm.as_ref().map_or(std::ptr::null(), |m| m as *const sys::RTMbvh),
                vertices.as_ptr() as *const f32,
                std::mem::size_of::<[f32; 3]>(),
                triangles.len(),
                &mut handle,
            )
        })?;
        Ok(GpuScene { handle, triangle_count: triangles.len(), _not_copy: PhantomData })
    }

    /// Build + collapse + triangle records, all left on the device (no host mirror): the trees are byte for byte
    /// what `Builder::construct_*` + `Mbvh::construct` return; `read_bvh` / `read_mbvh` copy them out on demand.
    pub fn build<T: SpatialTriangle + Sync>(
        triangles: &[T],
        primitives_per_leaf: Option<NonZeroUsize>,
        build_type: BuildType,
        want_mbvh: bool,
    ) -> Result<Self, GpuError> {
        let vertices = gather_vertices(triangles);
        let mut handle: sys::RTGpuScene = 0;
        check(unsafe {
            sys::rtbvh_gpu_scene_build(
                vertices.as_ptr() as *const f32,
                std::mem::size_of::<[f32; 3]>(),
                triangles.len(),
                primitives_per_leaf.map_or(0, NonZeroUsize::get),
                bvh_type_of(build_type),
                want_mbvh as c_int,
                &mut handle,
            )
        })?;
        Ok(GpuScene { handle, triangle_count: triangles.len(), _not_copy: PhantomData })
    }

    pub fn triangle_count(&self) -> usize {
        self.triangle_count
    }

    /// Multi-GPU, one process per GPU: the scene is built once; `export` yields 512 bytes (cudaIpc handles + sizes) that any
    /// channel carries to the other ranks, `import` turns them into a byte-identical replica on the caller's current GPU,
    /// copied device to device over NVLink.  Keep the exporting scene alive until every importer has returned.
    pub fn export(&self) -> Result<[u8; 512], GpuError> {
        let mut x = sys::RTGpuSceneExport { bytes: [0u8; 512] };
        check(unsafe { sys::rtbvh_gpu_scene_export(self.handle, &mut x) })?;
        Ok(x.bytes)
    }

    pub fn import(exported: &[u8; 512], triangle_count: usize) -> Result<Self, GpuError> {
        let x = sys::RTGpuSceneExport { bytes: *exported };
        let mut handle: sys::RTGpuScene = 0;
        check(unsafe { sys::rtbvh_gpu_scene_import(&x, &mut handle) })?;
        Ok(GpuScene { handle, triangle_count, _not_copy: PhantomData })
    }

    /// The same inside one process that drives several GPUs (`cudaMemcpyPeer`).
    pub fn clone_to_device(&self, device: i32) -> Result<Self, GpuError> {
        let mut handle: sys::RTGpuScene = 0;
        check(unsafe { sys::rtbvh_gpu_scene_clone(self.handle, device as c_int, &mut handle) })?;
        Ok(GpuScene { handle, triangle_count: self.triangle_count, _not_copy: PhantomData })
    }

    /// Dynamic scenes: new positions for the same triangles.  `Bvh::refit` (`src/bvh.rs:176-205`) on the device,
    /// followed by what the reference never does: the Mbvh's slot boxes are refreshed from the refitted binary tree
    /// (equal to `Mbvh::construct` of it).
    pub fn refit<T: SpatialTriangle + Sync>(&mut self, triangles: &[T]) -> Result<(), GpuError> {
        let vertices = gather_vertices(triangles);
        check(unsafe {
            sys::rtbvh_gpu_scene_refit(self.handle, vertices.as_ptr() as *const f32, std::mem::size_of::<[f32; 3]>(), triangles.len())
        })
    }

    /// Trace incoherent batches (shadow / bounce rays) in Morton order of (origin, direction); results unchanged.
    pub fn set_ray_sorting(&mut self, enable: bool) -> Result<(), GpuError> {
        check(unsafe { sys::rtbvh_gpu_scene_set_ray_sorting(self.handle, enable as c_int) })
    }

    /// Image-ordered batches (primary rays, `row_length` pixels per row): the device-pointer calls trace 8x8 pixel tiles
    /// (work order only; results unchanged; 0 = off).
    pub fn set_ray_tiling(&mut self, row_length: u32) -> Result<(), GpuError> {
        check(unsafe { sys::rtbvh_gpu_scene_set_ray_tiling(self.handle, row_length) })
    }

    /// The device copy of the binary tree (after `build` or `refit`), in the crate's format.
    pub fn read_bvh(&self, build_type: BuildType) -> Result<Bvh, GpuError> {
        let (mut n, mut k) = (0u32, 0u32);
        check(unsafe { sys::rtbvh_gpu_scene_tree_size(self.handle, sys::RT_TREE_BVH, &mut n, &mut k) })?;
        let mut nodes = vec![BvhNode::default(); n as usize];
        let mut prim_indices = vec![0u32; k as usize];
        check(unsafe {
            sys::rtbvh_gpu_scene_read_nodes(self.handle, sys::RT_TREE_BVH, nodes.as_mut_ptr() as *mut c_void, nodes.len() * std::mem::size_of::<BvhNode>())
        })?;
        check(unsafe { sys::rtbvh_gpu_scene_read_indices(self.handle, sys::RT_TREE_BVH, prim_indices.as_mut_ptr(), prim_indices.len()) })?;
        Ok(Bvh { nodes, prim_indices, build_type })
    }

    /// The device copy of the 4-wide tree.  `Mbvh::nodes()` (the binary nodes the reference clones into every Mbvh,
    /// `src/bvh.rs:399-403`) is filled from the scene's Bvh when it holds one.
    pub fn read_mbvh(&self) -> Result<Mbvh, GpuError> {
        let (mut n, mut k) = (0u32, 0u32);
        check(unsafe { sys::rtbvh_gpu_scene_tree_size(self.handle, sys::RT_TREE_MBVH, &mut n, &mut k) })?;
        let mut m_nodes = vec![MbvhNode::default(); n as usize];
        let mut prim_indices = vec![0u32; k as usize];
        check(unsafe {
            sys::rtbvh_gpu_scene_read_nodes(self.handle, sys::RT_TREE_MBVH, m_nodes.as_mut_ptr() as *mut c_void, m_nodes.len() * std::mem::size_of::<MbvhNode>())
        })?;
        check(unsafe { sys::rtbvh_gpu_scene_read_indices(self.handle, sys::RT_TREE_MBVH, prim_indices.as_mut_ptr(), prim_indices.len()) })?;
        let nodes = self.read_bvh(BuildType::None).map(|b| b.nodes).unwrap_or_default();
        Ok(Mbvh { nodes, m_nodes, prim_indices })
    }

    /// Closest hit for a batch — the loop `for (tri, r) in tree.traverse_iter(&mut ray, &triangles) { tri.intersect(r); }`
    /// for every ray.  `ray.t` is updated like the loop would (bit-identical `t`); the returned records add the
    /// primitive id the reference leaves to the caller.
    pub fn intersect(&self, tree: Tree, rays: &mut [Ray]) -> Result<Vec<Hit>, GpuError> {
        let packed: Vec<sys::RTRay> = rays.par_iter().map(sys::RTRay::from).collect();
        let mut raw = vec![sys::RTHit { t: 0.0, prim: sys::RT_NO_HIT }; rays.len()];
        check(unsafe { sys::rtbvh_gpu_intersect(self.handle, tree.raw(), packed.as_ptr(), packed.len(), raw.as_mut_ptr()) })?;
        rays.par_iter_mut().zip(raw.par_iter()).for_each(|(r, h)| r.t = h.t);
        Ok(raw.into_iter().map(Hit::from).collect())
    }

    /// Any hit — the same loop with `break` on the first successful test (the FFI callback returning `true`,
    /// `rtbvh_ffi/src/lib.rs:572-576`).  `rays` are not modified.
    pub fn occluded(&self, tree: Tree, rays: &[Ray]) -> Result<Vec<bool>, GpuError> {
        let packed: Vec<sys::RTRay> = rays.par_iter().map(sys::RTRay::from).collect();
        let mut raw = vec![0u8; rays.len()];
        check(unsafe { sys::rtbvh_gpu_occluded(self.handle, tree.raw(), packed.as_ptr(), packed.len(), raw.as_mut_ptr()) })?;
        Ok(raw.into_iter().map(|b| b != 0).collect())
    }

    /// Split input: tightly packed origins and directions (the argument shape of the FFI's `intersect`), one
    /// `t_min` / initial `t` for the whole batch (`Ray::DEFAULT_T_MIN`, `Ray::DEFAULT_T_MAX` for `Ray::new` rays).
    /// 24 instead of 32 bytes per ray cross PCIe, which is what bounds the host-buffer path.
    pub fn intersect_od(&self, tree: Tree, origins: &[[f32; 3]], directions: &[[f32; 3]], t_min: f32, t_max: f32) -> Result<Vec<Hit>, GpuError> {
        assert_eq!(origins.len(), directions.len());
        let mut raw = vec![sys::RTHit { t: 0.0, prim: sys::RT_NO_HIT }; origins.len()];
        check(unsafe {
            sys::rtbvh_gpu_intersect_od(self.handle, tree.raw(), origins.as_ptr() as *const f32, directions.as_ptr() as *const f32, origins.len(), t_min, t_max, raw.as_mut_ptr())
        })?;
        Ok(raw.into_iter().map(Hit::from).collect())
    }

    /// Packets of four rays, `SpatialTriangle::intersect4` semantics (`src/builders/spatial_sah.rs:165-244`:
    /// determinant eps 1e-6, `t >= t_min`); pass `t_min = 1e-4` for `examples/benchmark.rs:58`.  `packet.t` is updated
    /// per lane; returns the primitive per lane.
    pub fn intersect_packets(&self, tree: Tree, packets: &mut [RayPacket4], t_min: f32) -> Result<Vec<[Option<u32>; 4]>, GpuError> {
        let packed: Vec<sys::RTRayPacket4> = packets.par_iter().map(sys::RTRayPacket4::from).collect();
        let mut raw = vec![sys::RTHitPacket4 { t: [0.0; 4], prim: [sys::RT_NO_HIT; 4] }; packets.len()];
        check(unsafe { sys::rtbvh_gpu_intersect_packets(self.handle, tree.raw(), packed.as_ptr(), packed.len(), t_min, raw.as_mut_ptr()) })?;
        packets.par_iter_mut().zip(raw.par_iter()).for_each(|(p, h)| p.t = glam::Vec4::from(h.t));
        Ok(raw
            .into_iter()
            .map(|h| {
                let mut lanes = [None; 4];
                for (lane, prim) in lanes.iter_mut().zip(h.prim.iter()) {
                    if *prim != sys::RT_NO_HIT {
                        *lane = Some(*prim);
                    }
                }
                lanes
            })
            .collect())
    }

    /// Any hit for packets: one flag per lane.
    pub fn occluded_packets(&self, tree: Tree, packets: &[RayPacket4], t_min: f32) -> Result<Vec<[bool; 4]>, GpuError> {
        let packed: Vec<sys::RTRayPacket4> = packets.par_iter().map(sys::RTRayPacket4::from).collect();
        let mut raw = vec![0u8; 4 * packets.len()];
        check(unsafe { sys::rtbvh_gpu_occluded_packets(self.handle, tree.raw(), packed.as_ptr(), packed.len(), t_min, raw.as_mut_ptr()) })?;
        Ok(raw.chunks_exact(4).map(|c| [c[0] != 0, c[1] != 0, c[2] != 0, c[3] != 0]).collect())
    }

    /// Streaming flavour for renderers that double-buffer their batches: returns at once; batch k+1 uploads while
    /// batch k still traces and downloads.  Both buffers are page-locked (`PinnedBuf`) and stay borrowed until
    /// `Pending::wait` (or its drop) returns.
    pub fn submit<'a>(&'a self, tree: Tree, rays: &'a PinnedBuf<sys::RTRay>, hits: &'a mut PinnedBuf<sys::RTHit>) -> Result<Pending<'a>, GpuError> {
        assert!(hits.len() >= rays.len());
        let mut ticket = 0u64;
        check(unsafe { sys::rtbvh_gpu_intersect_async(self.handle, tree.raw(), rays.as_ptr(), rays.len(), hits.as_mut_ptr(), &mut ticket) })?;
        Ok(Pending { scene: self, ticket, done: false, _buffers: PhantomData })
    }

    /// True if a ray of an earlier device-side call needed a deeper traversal stack than the kernel has (the
    /// reference's 32-entry stack panics / is UB there, `src/iter.rs:25`).  Host-buffer calls report it as an error.
    pub fn stack_overflowed(&self) -> Result<bool, GpuError> {
        let mut flag = 0u32;
        check(unsafe { sys::rtbvh_gpu_scene_stack_overflowed(self.handle, &mut flag) })?;
        Ok(flag != 0)
    }

    /// The raw handle, for the `*_device` / `*_scatter` entry points in `sys` (device pointers and a `cudaStream_t`).
    pub fn raw_handle(&self) -> sys::RTGpuScene {
        self.handle
    }
}

impl Drop for GpuScene {
    fn drop(&mut self) {
        unsafe { sys::rtbvh_gpu_scene_free(self.handle) };
    }
}

/// A submitted batch; borrows the scene and both buffers until it has been waited for.
pub struct Pending<'a> {
    scene: &'a GpuScene,
    ticket: u64,
    done: bool,
    _buffers: PhantomData<&'a mut ()>,
}

impl<'a> Pending<'a> {
    /// Blocks until this batch's hit records are in its output buffer.
    pub fn wait(mut self) -> Result<(), GpuError> {
        self.done = true;
        check(unsafe { sys::rtbvh_gpu_wait(self.scene.handle, self.ticket) })
    }
}

impl<'a> Drop for Pending<'a> {
    fn drop(&mut self) {
        if !self.done {
            unsafe { sys::rtbvh_gpu_wait(self.scene.handle, self.ticket) };
        }
    }
}

/// Page-locked host memory from the library (`rtbvh_gpu_host_alloc`), zero-initialised, for the streaming calls.
pub struct PinnedBuf<T: Copy> {
    ptr: *mut T,
    len: usize,
}

unsafe impl<T: Copy + Send> Send for PinnedBuf<T> {}
unsafe impl<T: Copy + Sync> Sync for PinnedBuf<T> {}

impl<T: Copy> PinnedBuf<T> {
    /// `T` must be valid for the all-zero bit pattern (`RTRay`, `RTHit`, `f32`, `u8` are).
    pub fn zeroed(len: usize) -> Result<Self, GpuError> {
        let bytes = len.max(1) * std::mem::size_of::<T>();
        let mut p: *mut c_void = std::ptr::null_mut();
        check(unsafe { sys::rtbvh_gpu_host_alloc(bytes, &mut p) })?;
        unsafe { std::ptr::write_bytes(p as *mut u8, 0, bytes) };
        Ok(PinnedBuf { ptr: p as *mut T, len })
    }
}

impl<T: Copy> Deref for PinnedBuf<T> {
    type Target = [T];
    fn deref(&self) -> &[T] {
        unsafe { std::slice::from_raw_parts(self.ptr, self.len) }
    }
}

impl<T: Copy> DerefMut for PinnedBuf<T> {
    fn deref_mut(&mut self) -> &mut [T] {
        unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) }
    }
}

impl<T: Copy> Drop for PinnedBuf<T> {
    fn drop(&mut self) {
        unsafe { sys::rtbvh_gpu_host_free(self.ptr as *mut c_void) };
    }
}

// ------------------------------------------------------------------------------------------------------------------
// the crate's own tests, pointed at the GPU path (run with `cargo test` on a machine with a B200 and cargo)
// ------------------------------------------------------------------------------------------------------------------

#[cfg(test)]
mod tests {
    use super::*;
    use crate::*;
    use glam::*;

    #[derive(Debug, Copy, Clone)]
    struct Tri(Vec3, Vec3, Vec3);

    impl Primitive for Tri {
        fn center(&self) -> Vec3 {
            (self.0 + self.1 + self.2) * (1.0 / 3.0)
        }
        fn aabb(&self) -> Aabb {
            let mut bb = Aabb::new();
            bb.grow(self.0);
            bb.grow(self.1);
            bb.grow(self.2);
            bb
        }
    }

    impl SpatialTriangle for Tri {
        fn vertex0(&self) -> Vec3 {
            self.0
        }
        fn vertex1(&self) -> Vec3 {
            self.1
        }
        fn vertex2(&self) -> Vec3 {
            self.2
        }
    }

    fn quad() -> Vec<Tri> {
        let v = [vec3(-1.0, -1.0, 1.0), vec3(1.0, -1.0, 1.0), vec3(1.0, 1.0, 1.0), vec3(-1.0, 1.0, 1.0)];
        vec![Tri(v[0], v[1], v[2]), Tri(v[0], v[2], v[3])]
    }

    /// The host iterator loop and the batched call must agree bit for bit (t) and on the primitive.
    #[test]
    fn batch_equals_iterator_loop() {
        let tris = quad();
        let bvh = Builder { aabbs: None, primitives: &tris, primitives_per_leaf: None }.construct_binned_sah().unwrap();
        assert!(bvh.validate(tris.len()));
        let mbvh = Mbvh::construct(&bvh);
        let scene = GpuScene::new(Some(&bvh), Some(&mbvh), &tris).unwrap();

        let mut rays: Vec<Ray> = (0..64)
            .map(|i| Ray::new(vec3(-0.9 + 0.028 * i as f32, 0.3, 0.0), vec3(0.0, 0.0, 1.0)))
            .collect();
        let mut expect = rays.clone();
        let mut prims = vec![None; rays.len()];
        for (k, ray) in expect.iter_mut().enumerate() {
            let mut best = None;
            for (id, r) in mbvh.traverse_iter_indices(ray) {
                if tris[id as usize].intersect(r) {
                    best = Some(id);
                }
            }
            prims[k] = best;
        }
        let hits = scene.intersect(Tree::Mbvh, &mut rays).unwrap();
        for k in 0..rays.len() {
            assert_eq!(rays[k].t.to_bits(), expect[k].t.to_bits());
            assert_eq!(hits[k].prim, prims[k]);
        }
        let occluded = scene.occluded(Tree::Bvh, &expect.iter().map(|r| Ray::new(r.origin, r.direction)).collect::<Vec<_>>()).unwrap();
        assert!(occluded.iter().zip(prims.iter()).all(|(o, p)| *o == p.is_some()));
    }

    #[test]
    fn resident_build_matches_host_mirrored_build() {
        let tris = quad();
        let bvh = Builder { aabbs: None, primitives: &tris, primitives_per_leaf: None }.construct_binned_sah().unwrap();
        let scene = GpuScene::build(&tris, None, BuildType::BinnedSAH, true).unwrap();
        let resident = scene.read_bvh(BuildType::BinnedSAH).unwrap();
        assert_eq!(resident.indices(), bvh.indices());
        assert_eq!(resident.nodes().len(), bvh.nodes().len());
    }
}
