"""ctypes binding of rtbvh_b200/librtbvh_rs.so — the Python-side stand-in for the reference's Rust surface.

The product is the CUDA library behind the C ABI of include/rtbvh.h + include/rtbvh_gpu.h; this module only
marshals numpy / torch buffers into those calls, with the reference's names:

    Builder(aabbs, primitives(centers), primitives_per_leaf).construct_binned_sah()      src/bvh.rs:87-111
                                                    .construct_locally_ordered_clustered()  src/bvh.rs:113-137
    Bvh.nodes / .indices / .refit / .validate                                         src/bvh.rs:159-284
    Mbvh.construct(bvh) / Mbvh.from_bvh                                               src/bvh.rs:381-404,446-450
    Scene(...).intersect / .occluded / packets: the batched traversal loop (include/rtbvh_gpu.h)

There is NO CPU fallback: if the shared library is missing or no CUDA device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTBVH_LIB") or os.path.join(_HERE, "librtbvh_rs.so")  # RTBVH_LIB: A/B builds only

NODE_DTYPE = np.dtype([("min", "<f4", 3), ("count", "<i4"), ("max", "<f4", 3), ("left_first", "<i4")])
MNODE_DTYPE = np.dtype(
    [("min_x", "<f4", 4), ("max_x", "<f4", 4), ("min_y", "<f4", 4), ("max_y", "<f4", 4), ("min_z", "<f4", 4),
     ("max_z", "<f4", 4), ("children", "<i4", 4), ("counts", "<i4", 4)]
)
RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("t_min", "<f4"), ("direction", "<f4", 3), ("t", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("prim", "<u4")])
PACKET_DTYPE = np.dtype(
    [("origin_x", "<f4", 4), ("origin_y", "<f4", 4), ("origin_z", "<f4", 4), ("direction_x", "<f4", 4),
     ("direction_y", "<f4", 4), ("direction_z", "<f4", 4), ("t", "<f4", 4)]
)
HIT4_DTYPE = np.dtype([("t", "<f4", 4), ("prim", "<u4", 4)])

NO_HIT = 0xFFFFFFFF
OK, ERROR, NO_PRIMITIVES, INEQUAL_AABBS_AND_PRIMITIVES, NAN = range(5)  # ResultCode, rtbvh_ffi/src/lib.rs:17-25
LOCALLY_ORDERED_CLUSTERED, BINNED_SAH = 0, 1                            # BvhType, rtbvh_ffi/src/lib.rs:129-133
TREE_BVH, TREE_MBVH = 0, 1

# every symbol the two headers declare (tests check the library exports all of them)
LEGACY_SYMBOLS = ("create_spatial_Bvh", "create_bvh", "create_mbvh", "refit", "intersect", "intersect_packet",
                  "intersect_mbvh", "intersect_mbvh_packet", "free_bvh", "free_mbvh")
GPU_SYMBOLS = ("rtbvh_gpu_device_count", "rtbvh_gpu_set_device", "rtbvh_gpu_last_error", "rtbvh_gpu_scene_create",
               "rtbvh_gpu_scene_free", "rtbvh_gpu_intersect", "rtbvh_gpu_occluded", "rtbvh_gpu_intersect_packets",
               "rtbvh_gpu_occluded_packets", "rtbvh_gpu_intersect_device", "rtbvh_gpu_occluded_device",
               "rtbvh_gpu_intersect_packets_device", "rtbvh_gpu_occluded_packets_device",
               "rtbvh_gpu_scene_stack_overflowed", "rtbvh_gpu_generate_camera_rays_device", "rtbvh_gpu_intersect_camera_async", "rtbvh_gpu_occluded_camera_async",
               "rtbvh_gpu_create_bvh_triangles", "rtbvh_gpu_last_build_stats", "rtbvh_gpu_scene_set_ray_sorting", "rtbvh_gpu_scene_set_ray_tiling", "rtbvh_gpu_create_mbvh_from", "rtbvh_gpu_peer_buffer_create",
               "rtbvh_gpu_peer_buffer_open", "rtbvh_gpu_peer_buffer_close", "rtbvh_gpu_peer_buffer_free",
               "rtbvh_gpu_intersect_device_scatter", "rtbvh_gpu_occluded_device_scatter", "rtbvh_gpu_peer_barrier",
               "rtbvh_gpu_intersect_async", "rtbvh_gpu_occluded_async", "rtbvh_gpu_wait", "rtbvh_gpu_host_alloc",
               "rtbvh_gpu_host_free", "rtbvh_gpu_scene_refit", "rtbvh_gpu_scene_refit_device", "rtbvh_gpu_scene_read_nodes",
               "rtbvh_gpu_intersect_od", "rtbvh_gpu_occluded_od", "rtbvh_gpu_intersect_od_async", "rtbvh_gpu_occluded_od_async",
               "rtbvh_gpu_intersect_od_device", "rtbvh_gpu_trim_workspace", "rtbvh_gpu_scene_build", "rtbvh_gpu_scene_build_device", "rtbvh_gpu_scene_tree_size", "rtbvh_gpu_scene_read_indices",
               "rtbvh_gpu_scene_export", "rtbvh_gpu_scene_import", "rtbvh_gpu_scene_clone")


class RTBvh(C.Structure):  # rtbvh_ffi/src/lib.rs:210-220
    _fields_ = [("id", C.c_uint32), ("node_count", C.c_uint32), ("nodes", C.c_void_p), ("index_count", C.c_uint32),
                ("indices", C.c_void_p)]


class RTMbvh(C.Structure):  # rtbvh_ffi/src/lib.rs:234-244
    _fields_ = [("id", C.c_uint32), ("node_count", C.c_uint32), ("nodes", C.c_void_p), ("index_count", C.c_uint32),
                ("indices", C.c_void_p)]


assert C.sizeof(RTBvh) == 32 and C.sizeof(RTMbvh) == 32

CALLBACK = C.CFUNCTYPE(C.c_bool, C.c_uint32, C.POINTER(C.c_float), C.c_void_p)


class RtbvhError(RuntimeError):
    def __init__(self, code: int, msg: str = ""):
        super().__init__(f"ResultCode {code}: {msg}")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    """Loads the CUDA library.  Fails loudly: there is no other implementation to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). rtbvh_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, sz, u32, u64, f32 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_float
    rc = C.c_int
    L.create_spatial_Bvh.restype = rc
    L.create_spatial_Bvh.argtypes = [vp, sz, vp, sz, vp, sz, sz, u32, C.POINTER(RTBvh)]
    L.create_bvh.restype = rc
    L.create_bvh.argtypes = [vp, sz, vp, sz, sz, u32, C.POINTER(RTBvh)]
    L.create_mbvh.restype = rc
    L.create_mbvh.argtypes = [RTBvh, C.POINTER(RTMbvh)]
    L.refit.restype = rc
    L.refit.argtypes = [vp, RTBvh]
    L.intersect.restype = rc
    L.intersect.argtypes = [RTBvh, vp, vp, C.POINTER(f32), vp, CALLBACK]
    L.intersect_mbvh.restype = rc
    L.intersect_mbvh.argtypes = [RTMbvh, vp, vp, C.POINTER(f32), vp, CALLBACK]
    L.intersect_packet.restype = rc
    L.intersect_packet.argtypes = [RTBvh, vp, vp, vp, vp, vp, vp, vp, vp, CALLBACK]
    L.intersect_mbvh_packet.restype = rc
    L.intersect_mbvh_packet.argtypes = [RTMbvh, vp, vp, vp, vp, vp, vp, vp, vp, CALLBACK]
    L.free_bvh.restype = None
    L.free_bvh.argtypes = [RTBvh]
    L.free_mbvh.restype = None
    L.free_mbvh.argtypes = [RTMbvh]
    L.rtbvh_gpu_device_count.restype = C.c_int
    L.rtbvh_gpu_set_device.restype = rc
    L.rtbvh_gpu_set_device.argtypes = [C.c_int]
    L.rtbvh_gpu_last_error.restype = C.c_char_p
    L.rtbvh_gpu_scene_create.restype = rc
    L.rtbvh_gpu_scene_create.argtypes = [C.POINTER(RTBvh), C.POINTER(RTMbvh), vp, sz, sz, C.POINTER(u64)]
    L.rtbvh_gpu_scene_set_ray_sorting.restype = rc
    L.rtbvh_gpu_scene_set_ray_sorting.argtypes = [u64, C.c_int]
    L.rtbvh_gpu_scene_set_ray_tiling.restype = rc
    L.rtbvh_gpu_scene_set_ray_tiling.argtypes = [u64, C.c_uint32]
    L.rtbvh_gpu_scene_free.restype = rc
    L.rtbvh_gpu_scene_free.argtypes = [u64]
    L.rtbvh_gpu_intersect.restype = rc
    L.rtbvh_gpu_intersect.argtypes = [u64, C.c_int, vp, sz, vp]
    L.rtbvh_gpu_occluded.restype = rc
    L.rtbvh_gpu_occluded.argtypes = [u64, C.c_int, vp, sz, vp]
    L.rtbvh_gpu_intersect_packets.restype = rc
    L.rtbvh_gpu_intersect_packets.argtypes = [u64, C.c_int, vp, sz, f32, vp]
    L.rtbvh_gpu_occluded_packets.restype = rc
    L.rtbvh_gpu_occluded_packets.argtypes = [u64, C.c_int, vp, sz, f32, vp]
    L.rtbvh_gpu_intersect_device.restype = rc
    L.rtbvh_gpu_intersect_device.argtypes = [u64, C.c_int, vp, sz, vp, vp]
    L.rtbvh_gpu_occluded_device.restype = rc
    L.rtbvh_gpu_occluded_device.argtypes = [u64, C.c_int, vp, sz, vp, vp]
    L.rtbvh_gpu_intersect_packets_device.restype = rc
    L.rtbvh_gpu_intersect_packets_device.argtypes = [u64, C.c_int, vp, sz, f32, vp, vp]
    L.rtbvh_gpu_occluded_packets_device.restype = rc
    L.rtbvh_gpu_occluded_packets_device.argtypes = [u64, C.c_int, vp, sz, f32, vp, vp]
    L.rtbvh_gpu_scene_stack_overflowed.restype = rc
    L.rtbvh_gpu_scene_stack_overflowed.argtypes = [u64, C.POINTER(u32)]
    L.rtbvh_gpu_create_bvh_triangles.restype = rc
    L.rtbvh_gpu_create_bvh_triangles.argtypes = [vp, sz, sz, sz, u32, C.POINTER(RTBvh)]
    L.rtbvh_gpu_peer_buffer_create.restype = rc
    L.rtbvh_gpu_peer_buffer_create.argtypes = [sz, C.POINTER(vp), vp]
    L.rtbvh_gpu_scene_export.restype = rc
    L.rtbvh_gpu_scene_export.argtypes = [u64, vp]
    L.rtbvh_gpu_scene_import.restype = rc
    L.rtbvh_gpu_scene_import.argtypes = [vp, C.POINTER(u64)]
    L.rtbvh_gpu_scene_clone.restype = rc
    L.rtbvh_gpu_scene_clone.argtypes = [u64, C.c_int, C.POINTER(u64)]
    L.rtbvh_gpu_peer_buffer_open.restype = rc
    L.rtbvh_gpu_peer_buffer_open.argtypes = [vp, C.POINTER(vp)]
    L.rtbvh_gpu_peer_buffer_close.restype = rc
    L.rtbvh_gpu_peer_buffer_close.argtypes = [vp]
    L.rtbvh_gpu_peer_buffer_free.restype = rc
    L.rtbvh_gpu_peer_buffer_free.argtypes = [vp]
    L.rtbvh_gpu_intersect_async.restype = rc
    L.rtbvh_gpu_intersect_async.argtypes = [u64, C.c_int, vp, sz, vp, C.POINTER(u64)]
    L.rtbvh_gpu_occluded_async.restype = rc
    L.rtbvh_gpu_occluded_async.argtypes = [u64, C.c_int, vp, sz, vp, C.POINTER(u64)]
    f32 = C.c_float
    for name in ("rtbvh_gpu_intersect_od", "rtbvh_gpu_occluded_od"):
        getattr(L, name).restype = rc
        getattr(L, name).argtypes = [u64, C.c_int, vp, vp, sz, f32, f32, vp]
    for name in ("rtbvh_gpu_intersect_od_async", "rtbvh_gpu_occluded_od_async"):
        getattr(L, name).restype = rc
        getattr(L, name).argtypes = [u64, C.c_int, vp, vp, sz, f32, f32, vp, C.POINTER(u64)]
    L.rtbvh_gpu_intersect_od_device.restype = rc
    L.rtbvh_gpu_intersect_od_device.argtypes = [u64, C.c_int, vp, vp, sz, f32, f32, vp, vp]
    L.rtbvh_gpu_wait.restype = rc
    L.rtbvh_gpu_wait.argtypes = [u64, u64]
    L.rtbvh_gpu_host_alloc.restype = rc
    L.rtbvh_gpu_host_alloc.argtypes = [sz, C.POINTER(vp)]
    L.rtbvh_gpu_host_free.restype = rc
    L.rtbvh_gpu_host_free.argtypes = [vp]
    L.rtbvh_gpu_scene_build.restype = rc
    L.rtbvh_gpu_scene_build.argtypes = [vp, sz, sz, sz, u32, C.c_int, C.POINTER(u64)]
    L.rtbvh_gpu_scene_build_device.restype = rc
    L.rtbvh_gpu_scene_build_device.argtypes = [vp, sz, sz, sz, u32, C.c_int, C.POINTER(u64)]
    L.rtbvh_gpu_scene_tree_size.restype = rc
    L.rtbvh_gpu_scene_tree_size.argtypes = [u64, C.c_int, C.POINTER(u32), C.POINTER(u32)]
    L.rtbvh_gpu_scene_read_indices.restype = rc
    L.rtbvh_gpu_scene_read_indices.argtypes = [u64, C.c_int, vp, sz]
    L.rtbvh_gpu_scene_refit.restype = rc
    L.rtbvh_gpu_scene_refit.argtypes = [u64, vp, sz, sz]
    L.rtbvh_gpu_scene_refit_device.restype = rc
    L.rtbvh_gpu_scene_refit_device.argtypes = [u64, vp, sz, sz, vp]
    L.rtbvh_gpu_scene_read_nodes.restype = rc
    L.rtbvh_gpu_scene_read_nodes.argtypes = [u64, C.c_int, vp, sz]
    L.rtbvh_gpu_peer_barrier.restype = rc
    L.rtbvh_gpu_peer_barrier.argtypes = [C.POINTER(vp), C.c_int, C.c_int, u64, vp]
    L.rtbvh_gpu_intersect_device_scatter.restype = rc
    L.rtbvh_gpu_intersect_device_scatter.argtypes = [u64, C.c_int, vp, sz, vp, C.POINTER(vp), C.c_int, sz, vp]
    L.rtbvh_gpu_occluded_device_scatter.restype = rc
    L.rtbvh_gpu_occluded_device_scatter.argtypes = [u64, C.c_int, vp, sz, vp, C.POINTER(vp), C.c_int, sz, vp]
    L.rtbvh_gpu_create_mbvh_from.restype = rc
    L.rtbvh_gpu_create_mbvh_from.argtypes = [C.POINTER(RTBvh), C.POINTER(RTMbvh)]
    L.rtbvh_gpu_last_build_stats.restype = rc
    L.rtbvh_gpu_last_build_stats.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(u32)]
    L.rtbvh_gpu_generate_camera_rays_device.restype = rc
    L.rtbvh_gpu_generate_camera_rays_device.argtypes = [vp, vp, vp, vp, u32, u32, u32, u32, u64, u64, vp, vp]
    for fn in (L.rtbvh_gpu_intersect_camera_async, L.rtbvh_gpu_occluded_camera_async):
        fn.restype = rc
        fn.argtypes = [u64, C.c_int, vp, vp, vp, vp, u32, u32, u64, u64, u32, vp, C.POINTER(u64)]
    _lib = L
    return L


def last_error() -> str:
    return (lib().rtbvh_gpu_last_error() or b"").decode()


def _check(code: int):
    if code != OK:
        raise RtbvhError(code, last_error())


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _view(ptr, count, dtype):
    """numpy view (no copy) of library-owned host memory, like the slices the reference hands out."""
    if not ptr or count == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count)


def device_count() -> int:
    return int(lib().rtbvh_gpu_device_count())


def set_device(device: int):
    _check(lib().rtbvh_gpu_set_device(device))


class Bvh:
    """rtbvh::Bvh (src/bvh.rs:143-284) over an RTBvh handle or over caller-owned arrays."""

    def __init__(self, rt: RTBvh, keep=None, owned: bool = False):
        self.rt, self._keep, self._owned = rt, keep, owned

    @classmethod
    def from_arrays(cls, nodes: np.ndarray, indices: np.ndarray) -> "Bvh":
        """Wraps a tree built elsewhere (e.g. by the reference): the struct is trusted, as in the reference."""
        nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
        indices = np.ascontiguousarray(indices, dtype=np.uint32)
        rt = RTBvh(0xFFFFFFFF, len(nodes), nodes.ctypes.data, len(indices), indices.ctypes.data)
        return cls(rt, keep=(nodes, indices))

    @property
    def nodes(self) -> np.ndarray:
        return _view(self.rt.nodes, self.rt.node_count, NODE_DTYPE)

    @property
    def indices(self) -> np.ndarray:
        return _view(self.rt.indices, self.rt.index_count, np.uint32)

    def prim_count(self) -> int:
        return int(self.rt.index_count)

    def refit(self, aabbs: np.ndarray):
        aabbs = np.ascontiguousarray(aabbs, dtype=NODE_DTYPE)
        _check(lib().refit(_p(aabbs), self.rt))

    def validate(self, prim_count: int) -> bool:  # src/bvh.rs:232-244
        nodes, idx = self.nodes, self.indices
        found = np.zeros(prim_count, dtype=bool)
        stack = [0] if len(nodes) else []
        while stack:
            nd = nodes[stack.pop()]
            if nd["left_first"] < 0:
                continue
            if nd["count"] >= 0:
                found[idx[nd["left_first"]:nd["left_first"] + nd["count"]]] = True
            else:
                stack += [int(nd["left_first"]), int(nd["left_first"]) + 1]
        return bool(found.all())

    def free(self):
        if self._owned:
            lib().free_bvh(self.rt)
            self._owned = False


class Mbvh:
    """rtbvh::Mbvh (src/bvh.rs:320-450)."""

    def __init__(self, rt: RTMbvh, keep=None, owned: bool = False):
        self.rt, self._keep, self._owned = rt, keep, owned

    @classmethod
    def construct(cls, bvh: Bvh) -> "Mbvh":  # Mbvh::construct / From<Bvh>
        out = RTMbvh(0xFFFFFFFF, 0, None, 0, None)
        if bvh.rt.id == 0xFFFFFFFF:  # caller-owned arrays (e.g. a reference-built tree): collapse by pointer
            _check(lib().rtbvh_gpu_create_mbvh_from(C.byref(bvh.rt), C.byref(out)))
        else:
            _check(lib().create_mbvh(bvh.rt, C.byref(out)))
        return cls(out, owned=True)

    from_bvh = construct

    @classmethod
    def from_arrays(cls, nodes: np.ndarray, indices: np.ndarray) -> "Mbvh":
        nodes = np.ascontiguousarray(nodes, dtype=MNODE_DTYPE)
        indices = np.ascontiguousarray(indices, dtype=np.uint32)
        rt = RTMbvh(0xFFFFFFFF, len(nodes), nodes.ctypes.data, len(indices), indices.ctypes.data)
        return cls(rt, keep=(nodes, indices))

    @property
    def nodes(self) -> np.ndarray:
        return _view(self.rt.nodes, self.rt.node_count, MNODE_DTYPE)

    quad_nodes = nodes

    @property
    def indices(self) -> np.ndarray:
        return _view(self.rt.indices, self.rt.index_count, np.uint32)

    def free(self):
        if self._owned:
            lib().free_mbvh(self.rt)
            self._owned = False


class Builder:
    """rtbvh::Builder { aabbs, primitives, primitives_per_leaf } (src/bvh.rs:49-138), through create_bvh
    (rtbvh_ffi/src/lib.rs:428-493): `primitives` are given by their centers (Primitive::center)."""

    def __init__(self, aabbs, centers, primitives_per_leaf: int | None = None):
        self.aabbs = None if aabbs is None else np.ascontiguousarray(aabbs, dtype=NODE_DTYPE)
        self.centers = None if centers is None else np.ascontiguousarray(centers, dtype=np.float32)
        self.primitives_per_leaf = primitives_per_leaf or 0

    def _construct(self, kind: int) -> Bvh:
        c = self.centers
        n = 0 if c is None else (c.shape[0] if c.ndim == 2 else c.size // 3)
        stride = 12 if c is None or c.ndim != 2 else c.shape[1] * 4
        if self.aabbs is not None and len(self.aabbs) != n:
            raise RtbvhError(INEQUAL_AABBS_AND_PRIMITIVES, f"#Aabbs({len(self.aabbs)}) != #Primitives({n})")
        out = RTBvh(0xFFFFFFFF, 0, None, 0, None)
        dummy = np.zeros(4, np.float32)
        _check(lib().create_bvh(_p(self.aabbs), n, _p(c if c is not None and c.size else (dummy if c is not None else None)),
                                stride, self.primitives_per_leaf, kind, C.byref(out)))
        return Bvh(out, owned=True)

    def construct_binned_sah(self) -> Bvh:
        return self._construct(BINNED_SAH)

    def construct_locally_ordered_clustered(self) -> Bvh:
        return self._construct(LOCALLY_ORDERED_CLUSTERED)


def build_triangles(vertices: np.ndarray, bvh_type: int = BINNED_SAH, primitives_per_leaf: int | None = None) -> Bvh:
    """Builder{aabbs: None, primitives: triangles}.construct_*: aabbs and centers computed on the device."""
    v = np.ascontiguousarray(vertices, dtype=np.float32)
    stride = 16 if v.shape[-1] == 4 else 12
    v = v.reshape(-1, stride // 4)
    out = RTBvh(0xFFFFFFFF, 0, None, 0, None)
    _check(lib().rtbvh_gpu_create_bvh_triangles(_p(v), stride, v.shape[0] // 3, primitives_per_leaf or 0, bvh_type,
                                                C.byref(out)))
    return Bvh(out, owned=True)


def last_build_stats() -> dict:
    d, t, it = C.c_double(), C.c_double(), C.c_uint32()
    lib().rtbvh_gpu_last_build_stats(C.byref(d), C.byref(t), C.byref(it))
    return {"device_ms": d.value, "total_ms": t.value, "iterations": it.value}


def _dev_ptr(x):
    """Device pointer of a torch CUDA tensor (or a raw int)."""
    return C.c_void_p(x if isinstance(x, int) else x.data_ptr())


class Scene:
    """A tree (or a Bvh + its Mbvh) plus triangles resident on the current GPU (rtbvh_gpu_scene_create)."""

    def __init__(self, vertices: np.ndarray, bvh: Bvh | None = None, mbvh: Mbvh | None = None):
        v = np.ascontiguousarray(vertices, dtype=np.float32)
        stride = 16 if v.shape[-1] == 4 else 12
        v = v.reshape(-1, stride // 4)
        assert v.shape[0] % 3 == 0
        self.handle = C.c_uint64(0)
        self.n_nodes = int(bvh.rt.node_count) if bvh else 0
        self.n_mnodes = int(mbvh.rt.node_count) if mbvh else 0
        _check(lib().rtbvh_gpu_scene_create(C.byref(bvh.rt) if bvh else None, C.byref(mbvh.rt) if mbvh else None,
                                            _p(v), stride, v.shape[0] // 3, C.byref(self.handle)))

    @classmethod
    def build(cls, vertices, bvh_type: int = 1, prims_per_leaf: int = 1, mbvh: bool = True, n_tris: int | None = None,
              vertex_stride: int = 12) -> "Scene":
        """rtbvh_gpu_scene_build(_device): trees built and kept on the device.  `vertices`: numpy [n, 3, 3|4] (host) or a
        torch CUDA tensor / raw device pointer (then pass n_tris and vertex_stride)."""
        self = cls.__new__(cls)
        self.handle = C.c_uint64(0)
        if isinstance(vertices, np.ndarray):
            v = np.ascontiguousarray(vertices, dtype=np.float32)
            stride = 16 if v.shape[-1] == 4 else 12
            v = v.reshape(-1, stride // 4)
            _check(lib().rtbvh_gpu_scene_build(_p(v), stride, v.shape[0] // 3, prims_per_leaf, bvh_type, int(mbvh),
                                               C.byref(self.handle)))
        else:
            _check(lib().rtbvh_gpu_scene_build_device(_dev_ptr(vertices), vertex_stride, n_tris, prims_per_leaf, bvh_type,
                                                      int(mbvh), C.byref(self.handle)))
        nn, ni = C.c_uint32(0), C.c_uint32(0)
        _check(lib().rtbvh_gpu_scene_tree_size(self.handle, TREE_BVH, C.byref(nn), C.byref(ni)))
        self.n_nodes, self.n_indices, self.n_mnodes = nn.value, ni.value, 0
        if mbvh:
            _check(lib().rtbvh_gpu_scene_tree_size(self.handle, TREE_MBVH, C.byref(nn), None))
            self.n_mnodes = nn.value
        return self

    def _adopt_sizes(self):
        nn, ni = C.c_uint32(0), C.c_uint32(0)
        self.n_nodes = self.n_indices = self.n_mnodes = 0
        if lib().rtbvh_gpu_scene_tree_size(self.handle, TREE_BVH, C.byref(nn), C.byref(ni)) == 0:
            self.n_nodes, self.n_indices = nn.value, ni.value
        if lib().rtbvh_gpu_scene_tree_size(self.handle, TREE_MBVH, C.byref(nn), C.byref(ni)) == 0:
            self.n_mnodes, self.n_indices = nn.value, ni.value
        return self

    def export_bytes(self) -> bytes:
        """rtbvh_gpu_scene_export: 512 bytes (cudaIpc handles + sizes) another process turns into a replica with
        Scene.import_bytes.  Keep this scene alive until every importer has returned."""
        blob = C.create_string_buffer(512)
        _check(lib().rtbvh_gpu_scene_export(self.handle, blob))
        return blob.raw

    @classmethod
    def import_bytes(cls, blob: bytes) -> "Scene":
        """rtbvh_gpu_scene_import: replica on the current device, copied device to device from the exporting process's GPU."""
        self = cls.__new__(cls)
        self.handle = C.c_uint64(0)
        buf = C.create_string_buffer(bytes(blob), 512)
        _check(lib().rtbvh_gpu_scene_import(buf, C.byref(self.handle)))
        return self._adopt_sizes()

    def clone(self, device: int) -> "Scene":
        """rtbvh_gpu_scene_clone: replica on another GPU of this process (cudaMemcpyPeer)."""
        other = Scene.__new__(Scene)
        other.handle = C.c_uint64(0)
        _check(lib().rtbvh_gpu_scene_clone(self.handle, int(device), C.byref(other.handle)))
        return other._adopt_sizes()

    def read_indices(self, tree: int = TREE_BVH) -> np.ndarray:
        ni = C.c_uint32(0)
        _check(lib().rtbvh_gpu_scene_tree_size(self.handle, tree, None, C.byref(ni)))
        out = np.zeros(ni.value, dtype=np.uint32)
        _check(lib().rtbvh_gpu_scene_read_indices(self.handle, tree, out.ctypes.data_as(C.c_void_p), ni.value))
        return out

    def set_ray_sorting(self, enable: bool = True):
        """Trace every single-ray batch in Morton order of (origin, direction); results are unchanged."""
        _check(lib().rtbvh_gpu_scene_set_ray_sorting(self.handle, int(enable)))

    def set_ray_tiling(self, row_length: int = 0):
        """Image-ordered batches: the device-pointer calls trace 8x8 pixel tiles (work order only; 0 = off)."""
        _check(lib().rtbvh_gpu_scene_set_ray_tiling(self.handle, int(row_length)))

    def free(self):
        if self.handle.value:
            lib().rtbvh_gpu_scene_free(self.handle)
            self.handle = C.c_uint64(0)

    # ---- host buffers ---------------------------------------------------------------------
    def intersect(self, rays: np.ndarray, tree: int = TREE_MBVH, out: np.ndarray | None = None) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = out if out is not None else np.empty(len(rays), dtype=HIT_DTYPE)
        _check(lib().rtbvh_gpu_intersect(self.handle, tree, _p(rays), len(rays), _p(hits)))
        return hits

    def occluded(self, rays: np.ndarray, tree: int = TREE_MBVH) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        occ = np.empty(len(rays), dtype=np.uint8)
        _check(lib().rtbvh_gpu_occluded(self.handle, tree, _p(rays), len(rays), _p(occ)))
        return occ

    def intersect_packets(self, packets: np.ndarray, tree: int = TREE_MBVH, t_min: float = 1e-4) -> np.ndarray:
        packets = np.ascontiguousarray(packets, dtype=PACKET_DTYPE)
        hits = np.empty(len(packets), dtype=HIT4_DTYPE)
        _check(lib().rtbvh_gpu_intersect_packets(self.handle, tree, _p(packets), len(packets), t_min, _p(hits)))
        return hits

    def occluded_packets(self, packets: np.ndarray, tree: int = TREE_MBVH, t_min: float = 1e-4) -> np.ndarray:
        packets = np.ascontiguousarray(packets, dtype=PACKET_DTYPE)
        occ = np.empty((len(packets), 4), dtype=np.uint8)
        _check(lib().rtbvh_gpu_occluded_packets(self.handle, tree, _p(packets), len(packets), t_min, _p(occ)))
        return occ

    # ---- raw host pointers (pinned torch tensors etc.) ----------------------------------------
    def intersect_ptr(self, rays_ptr: int, n: int, hits_ptr: int, tree: int = TREE_MBVH):
        _check(lib().rtbvh_gpu_intersect(self.handle, tree, C.c_void_p(rays_ptr), n, C.c_void_p(hits_ptr)))

    # ---- dynamic scenes: refit in place from new vertex positions --------------------------------
    def refit(self, verts: np.ndarray):
        """rtbvh_gpu_scene_refit: verts float32 [n, 3, 3] (or [n, 3, 4]) of the same triangles, moved."""
        v = np.ascontiguousarray(verts, dtype=np.float32)
        stride = 16 if v.shape[-1] == 4 else 12
        v = v.reshape(-1, stride // 4)
        _check(lib().rtbvh_gpu_scene_refit(self.handle, v.ctypes.data_as(C.c_void_p), stride, v.shape[0] // 3))

    def refit_device(self, d_verts, n_tris: int, vertex_stride: int = 12, stream: int = 0):
        _check(lib().rtbvh_gpu_scene_refit_device(self.handle, _dev_ptr(d_verts), vertex_stride, n_tris, C.c_void_p(stream)))

    def read_nodes(self, tree: int) -> np.ndarray:
        """The device copy of the scene's nodes (after refits it differs from the host mirror)."""
        dt, n = (MNODE_DTYPE, self.n_mnodes) if tree == TREE_MBVH else (NODE_DTYPE, self.n_nodes)
        out = np.zeros(n, dtype=dt)
        _check(lib().rtbvh_gpu_scene_read_nodes(self.handle, tree, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    # ---- asynchronous host-buffer calls: submit -> ticket, wait(ticket) ------------------------
    def intersect_async(self, rays_ptr: int, n: int, hits_ptr: int, tree: int = TREE_MBVH) -> int:
        t = C.c_uint64(0)
        _check(lib().rtbvh_gpu_intersect_async(self.handle, tree, C.c_void_p(rays_ptr), n, C.c_void_p(hits_ptr), C.byref(t)))
        return t.value

    def occluded_async(self, rays_ptr: int, n: int, occ_ptr: int, tree: int = TREE_MBVH) -> int:
        t = C.c_uint64(0)
        _check(lib().rtbvh_gpu_occluded_async(self.handle, tree, C.c_void_p(rays_ptr), n, C.c_void_p(occ_ptr), C.byref(t)))
        return t.value

    # ---- split ray input: origins / directions as packed float3 arrays, common t_min / t_max ----
    def intersect_od(self, origins: np.ndarray, directions: np.ndarray, tree: int = TREE_MBVH, t_min: float = 1e-4,
                     t_max: float = 1e34, any_hit: bool = False) -> np.ndarray:
        o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, dtype=np.float32).reshape(-1, 3)
        assert o.shape == d.shape
        out = np.zeros(len(o), dtype=np.uint8 if any_hit else HIT_DTYPE)
        fn = lib().rtbvh_gpu_occluded_od if any_hit else lib().rtbvh_gpu_intersect_od
        _check(fn(self.handle, tree, _p(o), _p(d), len(o), t_min, t_max, out.ctypes.data_as(C.c_void_p)))
        return out

    def intersect_od_async(self, origins_ptr: int, directions_ptr: int, n: int, hits_ptr: int, tree: int = TREE_MBVH,
                           t_min: float = 1e-4, t_max: float = 1e34) -> int:
        t = C.c_uint64(0)
        _check(lib().rtbvh_gpu_intersect_od_async(self.handle, tree, C.c_void_p(origins_ptr), C.c_void_p(directions_ptr), n,
                                                  t_min, t_max, C.c_void_p(hits_ptr), C.byref(t)))
        return t.value

    def intersect_od_ptr(self, origins_ptr: int, directions_ptr: int, n: int, hits_ptr: int, tree: int = TREE_MBVH,
                         t_min: float = 1e-4, t_max: float = 1e34):
        _check(lib().rtbvh_gpu_intersect_od(self.handle, tree, C.c_void_p(origins_ptr), C.c_void_p(directions_ptr), n, t_min,
                                            t_max, C.c_void_p(hits_ptr)))

    def intersect_od_device(self, d_origins, d_directions, n: int, d_hits, tree: int = TREE_MBVH, t_min: float = 1e-4,
                            t_max: float = 1e34, stream: int = 0):
        _check(lib().rtbvh_gpu_intersect_od_device(self.handle, tree, _dev_ptr(d_origins), _dev_ptr(d_directions), n, t_min,
                                                   t_max, _dev_ptr(d_hits), C.c_void_p(stream)))

    def intersect_camera_async(self, cam: dict, frames: int, hits_ptr: int, tree: int = TREE_MBVH, jitter_seed: int = 0,
                               first_frame: int = 0, any_hit: bool = False) -> int:
        """rtbvh_gpu_intersect_camera_async / _occluded_camera_async: primary rays generated on the device, records to the
        host buffer at `hits_ptr`; returns a ticket."""
        keep = [np.ascontiguousarray(cam[k], dtype=np.float32) for k in ("pos", "p1", "right", "up")]
        t = C.c_uint64(0)
        fn = lib().rtbvh_gpu_occluded_camera_async if any_hit else lib().rtbvh_gpu_intersect_camera_async
        _check(fn(self.handle, tree, _p(keep[0]), _p(keep[1]), _p(keep[2]), _p(keep[3]), cam["width"], cam["height"], jitter_seed,
                  first_frame, frames, C.c_void_p(hits_ptr), C.byref(t)))
        return t.value

    def wait(self, ticket: int = 0):
        _check(lib().rtbvh_gpu_wait(self.handle, ticket))

    # ---- device-resident, asynchronous on `stream` (a cudaStream_t as int) --------------------
    def intersect_device(self, d_rays, n: int, d_hits, tree: int = TREE_MBVH, stream: int = 0):
        _check(lib().rtbvh_gpu_intersect_device(self.handle, tree, _dev_ptr(d_rays), n, _dev_ptr(d_hits), C.c_void_p(stream)))

    def occluded_device(self, d_rays, n: int, d_occ, tree: int = TREE_MBVH, stream: int = 0):
        _check(lib().rtbvh_gpu_occluded_device(self.handle, tree, _dev_ptr(d_rays), n, _dev_ptr(d_occ), C.c_void_p(stream)))

    def intersect_packets_device(self, d_packets, n: int, d_hits, tree: int = TREE_MBVH, t_min: float = 1e-4, stream: int = 0):
        _check(lib().rtbvh_gpu_intersect_packets_device(self.handle, tree, _dev_ptr(d_packets), n, t_min, _dev_ptr(d_hits),
                                                        C.c_void_p(stream)))

    def occluded_packets_device(self, d_packets, n: int, d_occ, tree: int = TREE_MBVH, t_min: float = 1e-4, stream: int = 0):
        _check(lib().rtbvh_gpu_occluded_packets_device(self.handle, tree, _dev_ptr(d_packets), n, t_min, _dev_ptr(d_occ),
                                                       C.c_void_p(stream)))

    # ---- multi-GPU: gather fused into the kernel (P2P stores into peer buffers) ----------------
    def intersect_device_scatter(self, d_rays, n: int, dests: list[int], dest_offset: int, d_hits=None,
                                 tree: int = TREE_MBVH, stream: int = 0):
        arr = (C.c_void_p * len(dests))(*dests)
        _check(lib().rtbvh_gpu_intersect_device_scatter(self.handle, tree, _dev_ptr(d_rays), n,
                                                        _dev_ptr(d_hits) if d_hits is not None else None, arr, len(dests),
                                                        dest_offset, C.c_void_p(stream)))

    def occluded_device_scatter(self, d_rays, n: int, dests: list[int], dest_offset: int, d_occ=None,
                                tree: int = TREE_MBVH, stream: int = 0):
        arr = (C.c_void_p * len(dests))(*dests)
        _check(lib().rtbvh_gpu_occluded_device_scatter(self.handle, tree, _dev_ptr(d_rays), n,
                                                       _dev_ptr(d_occ) if d_occ is not None else None, arr, len(dests),
                                                       dest_offset, C.c_void_p(stream)))

    def stack_overflowed(self) -> bool:
        v = C.c_uint32(0)
        _check(lib().rtbvh_gpu_scene_stack_overflowed(self.handle, C.byref(v)))
        return bool(v.value)


class PeerBuffer:
    """A gather buffer other ranks can write into: cudaMalloc + cudaIpc handle (64 bytes, exchange out of band)."""

    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        self.handle = (C.c_ubyte * 64)()
        self.nbytes = nbytes
        _check(lib().rtbvh_gpu_peer_buffer_create(nbytes, C.byref(self.ptr), self.handle))

    def handle_bytes(self) -> bytes:
        return bytes(self.handle)

    @staticmethod
    def open(handle: bytes) -> int:
        h = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        _check(lib().rtbvh_gpu_peer_buffer_open(h, C.byref(p)))
        return p.value

    @staticmethod
    def close(ptr: int):
        _check(lib().rtbvh_gpu_peer_buffer_close(C.c_void_p(ptr)))

    def free(self):
        if self.ptr:
            lib().rtbvh_gpu_peer_buffer_free(self.ptr)
            self.ptr = C.c_void_p()


class _DevView:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


def device_view(ptr: int, nbytes: int):
    """A torch uint8 tensor aliasing raw device memory (e.g. a gather buffer), for stream-ordered consumers."""
    import torch
    return torch.as_tensor(_DevView(ptr, nbytes), device="cuda")


def trim_workspace():
    """Frees this thread's builder workspace (rtbvh_gpu_trim_workspace)."""
    lib().rtbvh_gpu_trim_workspace.restype = C.c_int
    _check(lib().rtbvh_gpu_trim_workspace())


def peer_barrier(flag_arrays: list[int], rank: int, value: int, stream: int = 0):
    """Cross-rank step barrier enqueued on `stream` (rtbvh_gpu_peer_barrier)."""
    arr = (C.c_void_p * len(flag_arrays))(*flag_arrays)
    _check(lib().rtbvh_gpu_peer_barrier(arr, len(flag_arrays), rank, value, C.c_void_p(stream)))


def generate_camera_rays_device(cam: dict, row0: int, rows: int, d_rays, jitter_seed: int = 0, frame: int = 0,
                                stream: int = 0):
    """CameraView3D::generate_ray on the device (shared/src/lib.rs:157-165) for pixel rows [row0, row0+rows)."""
    f = lambda k: _p(np.ascontiguousarray(cam[k], dtype=np.float32))
    keep = [np.ascontiguousarray(cam[k], dtype=np.float32) for k in ("pos", "p1", "right", "up")]
    _check(lib().rtbvh_gpu_generate_camera_rays_device(_p(keep[0]), _p(keep[1]), _p(keep[2]), _p(keep[3]), cam["width"],
                                                       cam["height"], row0, rows, jitter_seed, frame, _dev_ptr(d_rays),
                                                       C.c_void_p(stream)))


def intersect_callback(tree, origin, direction, t: float, cb) -> tuple[int, float]:
    """Legacy per-candidate entry points intersect / intersect_mbvh (rtbvh_ffi/src/lib.rs:551-581, :700-731)."""
    o = np.asarray(origin, dtype=np.float32)
    d = np.asarray(direction, dtype=np.float32)
    tv = C.c_float(t)
    fn = lib().intersect_mbvh if isinstance(tree, Mbvh) else lib().intersect
    code = fn(tree.rt, _p(o), _p(d), C.byref(tv), None, cb)
    return code, tv.value


def intersect_packet_callback(tree, packet: np.ndarray, cb) -> tuple[int, np.ndarray]:
    """intersect_packet / intersect_mbvh_packet (rtbvh_ffi/src/lib.rs:599-686, :749-835) for one PACKET_DTYPE record."""
    p = np.ascontiguousarray(packet, dtype=PACKET_DTYPE).reshape(1)
    cols = [np.ascontiguousarray(p[k][0]) for k in ("origin_x", "origin_y", "origin_z", "direction_x", "direction_y", "direction_z")]
    t = np.ascontiguousarray(p["t"][0]).copy()
    fn = lib().intersect_mbvh_packet if isinstance(tree, Mbvh) else lib().intersect_packet
    code = fn(tree.rt, *[_p(c) for c in cols], _p(t), None, cb)
    return code, t
