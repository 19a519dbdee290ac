"""ctypes front-end for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  Nothing under rtbvh_b200/ imports this module.

The numpy dtypes below are the wire formats shared with the product's C ABI (include/rtbvh.h,
include/rtbvh_gpu.h): BvhNode 32 B (reference src/bvh_node.rs:11-14 / src/aabb.rs:13-20), MbvhNode 128 B
(src/mbvh_node.rs:28-38), RTRay 32 B (first 32 bytes of src/ray.rs:9-16), RTHit 8 B.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

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
assert NODE_DTYPE.itemsize == 32 and MNODE_DTYPE.itemsize == 128 and RAY_DTYPE.itemsize == 32
assert HIT_DTYPE.itemsize == 8 and PACKET_DTYPE.itemsize == 112 and HIT4_DTYPE.itemsize == 32

NO_HIT = 0xFFFFFFFF
LOCB, BINNED_SAH = 0, 1  # rtbvh_ffi BvhType (rtbvh_ffi/src/lib.rs:129-133)
COUNTER_NAMES = ("node_visits", "inner_visits", "prim_tests", "max_stack", "overflow32")


def build_lib(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile (no-op when it is up to date)."""
    src = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "rtbvh_oracle.hpp", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build_lib()
    L = C.CDLL(_LIB_PATH)
    vp, sz, u32p, f32p, u8p, u64p = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.POINTER(
        C.c_uint8), C.POINTER(C.c_uint64)
    L.rto_num_threads.restype = C.c_int
    L.rto_morton_split.restype = C.c_uint32
    L.rto_morton_split.argtypes = [C.c_uint32]
    L.rto_morton_encode.restype = C.c_uint32
    L.rto_morton_encode.argtypes = [vp, vp]
    for n in ("rto_sizeof_aabb", "rto_sizeof_bvh_node", "rto_sizeof_mbvh_node"):
        getattr(L, n).restype = sz
    L.rto_prefix_sum_u32.restype = C.c_uint32
    L.rto_prefix_sum_u32.argtypes = [vp, sz, vp]
    L.rto_prefix_sum_i32.restype = C.c_int32
    L.rto_prefix_sum_i32.argtypes = [vp, sz, vp]
    L.rto_prefix_sum_u64.restype = C.c_uint64
    L.rto_prefix_sum_u64.argtypes = [vp, sz, vp]
    L.rto_move_backward_u32.argtypes = [vp, vp, vp]
    L.rto_partition_lt.restype = sz
    L.rto_partition_lt.argtypes = [vp, sz, C.c_uint32]
    L.rto_prims_from_triangles.argtypes = [vp, sz, C.c_float, vp, vp]
    L.rto_aabb_centers.argtypes = [vp, sz, vp]
    L.rto_bvh_build.restype = C.c_int
    L.rto_bvh_build.argtypes = [C.c_int, vp, sz, vp, sz, sz, sz, C.c_int, C.POINTER(vp), C.POINTER(C.c_double),
                                C.POINTER(C.c_double)]
    L.rto_bvh_build_spatial.restype = C.c_int
    L.rto_bvh_build_spatial.argtypes = [vp, sz, sz, C.c_int, C.POINTER(vp), C.POINTER(C.c_double), vp]
    L.rto_bvh_from_raw.restype = vp
    L.rto_bvh_from_raw.argtypes = [vp, sz, vp, sz]
    L.rto_bvh_free.argtypes = [vp]
    L.rto_bvh_node_count.restype = sz
    L.rto_bvh_node_count.argtypes = [vp]
    L.rto_bvh_nodes.restype = vp
    L.rto_bvh_nodes.argtypes = [vp]
    L.rto_bvh_index_count.restype = sz
    L.rto_bvh_index_count.argtypes = [vp]
    L.rto_bvh_indices.restype = vp
    L.rto_bvh_indices.argtypes = [vp]
    L.rto_bvh_validate.restype = C.c_int
    L.rto_bvh_validate.argtypes = [vp, sz]
    L.rto_bvh_sah_cost.restype = C.c_double
    L.rto_bvh_sah_cost.argtypes = [vp]
    L.rto_sah_cost_raw.restype = C.c_double
    L.rto_sah_cost_raw.argtypes = [vp, sz]
    L.rto_bvh_refit.argtypes = [vp, vp]
    L.rto_bvh_depth_stats.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    L.rto_mbvh_construct.restype = vp
    L.rto_mbvh_construct.argtypes = [vp, C.POINTER(C.c_double)]
    L.rto_mbvh_free.argtypes = [vp]
    L.rto_mbvh_node_count.restype = sz
    L.rto_mbvh_node_count.argtypes = [vp]
    L.rto_mbvh_nodes.restype = vp
    L.rto_mbvh_nodes.argtypes = [vp]
    L.rto_mbvh_index_count.restype = sz
    L.rto_mbvh_index_count.argtypes = [vp]
    L.rto_mbvh_indices.restype = vp
    L.rto_mbvh_indices.argtypes = [vp]
    L.rto_trace.restype = C.c_double
    L.rto_trace.argtypes = [C.c_int, C.c_int, vp, sz, vp, vp, vp, sz, vp, vp, vp, C.c_int]
    L.rto_trace_packet.restype = C.c_double
    L.rto_trace_packet.argtypes = [C.c_int, C.c_int, vp, sz, vp, vp, vp, sz, C.c_float, vp, vp, vp, C.c_int]
    L.rto_brute_force.argtypes = [vp, sz, vp, sz, vp, C.c_int]
    L.rto_intersect_cb.restype = C.c_int
    L.rto_intersect_cb.argtypes = [C.c_int, vp, sz, vp, vp, vp, vp, vp, vp]
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _copy(ptr, count, dtype):
    if count == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count).copy()


def num_threads() -> int:
    return int(lib().rto_num_threads())


def prims_from_triangles(verts: np.ndarray, pad: float = 0.0):
    """aabbs (NODE_DTYPE, count/left_first = 0) and centers [n,3] as the bench Triangle reports them
    (reference shared/src/lib.rs:27-39)."""
    verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 9)
    n = verts.shape[0]
    aabbs = np.zeros(n, dtype=NODE_DTYPE)
    centers = np.zeros((n, 3), dtype=np.float32)
    lib().rto_prims_from_triangles(_p(verts), n, pad, _p(aabbs), _p(centers))
    return aabbs, centers


def aabb_centers(aabbs: np.ndarray) -> np.ndarray:
    out = np.zeros((len(aabbs), 3), dtype=np.float32)
    lib().rto_aabb_centers(_p(aabbs), len(aabbs), _p(out))
    return out


class Bvh:
    """Host copy of a built tree: nodes (NODE_DTYPE) + prim_indices (u32)."""

    def __init__(self, nodes, indices, build_ms=0.0, kappa=0.0):
        self.nodes, self.indices, self.build_ms, self.kappa = nodes, indices, build_ms, kappa

    def _handle(self):
        return lib().rto_bvh_from_raw(_p(self.nodes), len(self.nodes), _p(self.indices), len(self.indices))

    def validate(self, prim_count: int) -> bool:
        h = self._handle()
        try:
            return bool(lib().rto_bvh_validate(h, prim_count))
        finally:
            lib().rto_bvh_free(h)

    def sah_cost(self) -> float:
        return float(lib().rto_sah_cost_raw(_p(self.nodes), len(self.nodes)))

    def depth_stats(self):
        h = self._handle()
        mean, mx, leaves = C.c_double(), C.c_uint32(), C.c_uint64()
        lib().rto_bvh_depth_stats(h, C.byref(mean), C.byref(mx), C.byref(leaves))
        lib().rto_bvh_free(h)
        return mean.value, mx.value, leaves.value

    def refit(self, aabbs: np.ndarray) -> "Bvh":
        h = self._handle()
        lib().rto_bvh_refit(h, _p(np.ascontiguousarray(aabbs)))
        out = Bvh(_copy(lib().rto_bvh_nodes(h), lib().rto_bvh_node_count(h), NODE_DTYPE), self.indices.copy())
        lib().rto_bvh_free(h)
        return out

    def collapse(self) -> "Mbvh":
        h = self._handle()
        ms = C.c_double()
        m = lib().rto_mbvh_construct(h, C.byref(ms))
        out = Mbvh(_copy(lib().rto_mbvh_nodes(m), lib().rto_mbvh_node_count(m), MNODE_DTYPE),
                   _copy(lib().rto_mbvh_indices(m), lib().rto_mbvh_index_count(m), np.uint32), ms.value)
        lib().rto_mbvh_free(m)
        lib().rto_bvh_free(h)
        return out


class Mbvh:
    def __init__(self, nodes, indices, collapse_ms=0.0):
        self.nodes, self.indices, self.collapse_ms = nodes, indices, collapse_ms


def build(bvh_type: int, aabbs, centers: np.ndarray, prims_per_leaf: int = 1, parallel: bool = False,
          prim_count: int | None = None):
    """rtbvh_ffi create_bvh semantics (rtbvh_ffi/src/lib.rs:428-493).  Returns (result_code, Bvh | None)."""
    if centers is None:
        return 1, None
    centers = np.ascontiguousarray(centers, dtype=np.float32)
    stride = centers.shape[1] * 4 if centers.ndim == 2 else 12
    n = prim_count if prim_count is not None else (centers.shape[0] if centers.ndim == 2 else centers.size // 3)
    if aabbs is not None:
        aabbs = np.ascontiguousarray(aabbs, dtype=NODE_DTYPE)
    h, ms, kappa = C.c_void_p(), C.c_double(), C.c_double()
    dummy = np.zeros(4, dtype=np.float32)
    rc = lib().rto_bvh_build(bvh_type, _p(aabbs) if aabbs is not None else None, 0 if aabbs is None else len(aabbs),
                             _p(centers) if centers.size else _p(dummy), stride, n, prims_per_leaf, int(parallel),
                             C.byref(h), C.byref(ms), C.byref(kappa))
    if rc != 0:
        return rc, None
    L = lib()
    out = Bvh(_copy(L.rto_bvh_nodes(h), L.rto_bvh_node_count(h), NODE_DTYPE),
              _copy(L.rto_bvh_indices(h), L.rto_bvh_index_count(h), np.uint32), ms.value, kappa.value)
    L.rto_bvh_free(h)
    return 0, out


def build_spatial(verts: np.ndarray, prims_per_leaf: int = 1, fix_child_ranges: bool = True):
    """Builder::construct_spatial_sah (src/bvh.rs:58-85, src/builders/spatial_sah.rs:925-1033) on the CPU.
    Returns (result_code, Bvh | None); Bvh.stats = (spatial splits, object splits, references used)."""
    verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 9)
    h, ms = C.c_void_p(), C.c_double()
    stats = np.zeros(3, dtype=np.uint64)
    rc = lib().rto_bvh_build_spatial(_p(verts), verts.shape[0], prims_per_leaf, int(fix_child_ranges), C.byref(h), C.byref(ms),
                                     _p(stats))
    if rc != 0:
        return rc, None
    L = lib()
    out = Bvh(_copy(L.rto_bvh_nodes(h), L.rto_bvh_node_count(h), NODE_DTYPE),
              _copy(L.rto_bvh_indices(h), L.rto_bvh_index_count(h), np.uint32), ms.value, 0.0)
    out.stats = tuple(int(x) for x in stats)
    L.rto_bvh_free(h)
    return 0, out


def _tree_args(tree):
    kind = 1 if tree.nodes.dtype == MNODE_DTYPE else 0
    return kind, _p(tree.nodes), len(tree.nodes), _p(tree.indices)


def trace(tree, verts: np.ndarray, rays: np.ndarray, mode: str = "closest", threads: int = 0, counters: bool = False):
    """Per-ray loop of examples/benchmark.rs:15-41 with hit-id tracking.
    Returns (hits | occluded, elapsed_ms, counters dict | None)."""
    verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 9)
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    kind, nodes, n_nodes, idx = _tree_args(tree)
    n = len(rays)
    hits = np.zeros(n, dtype=HIT_DTYPE)
    occ = np.zeros(n, dtype=np.uint8)
    cnt = np.zeros(5, dtype=np.uint64) if counters else None
    threads = threads or num_threads()
    ms = lib().rto_trace(kind, 0 if mode == "closest" else 1, nodes, n_nodes, idx, _p(verts), _p(rays), n, _p(hits),
                         _p(occ), _p(cnt), threads)
    cd = dict(zip(COUNTER_NAMES, (int(x) for x in cnt))) if counters else None
    return (hits if mode == "closest" else occ), ms, cd


def trace_packets(tree, verts, packets, t_min: float = 1e-4, mode: str = "closest", threads: int = 0,
                  counters: bool = False):
    verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 9)
    packets = np.ascontiguousarray(packets, dtype=PACKET_DTYPE)
    kind, nodes, n_nodes, idx = _tree_args(tree)
    n = len(packets)
    hits = np.zeros(n, dtype=HIT4_DTYPE)
    occ = np.zeros((n, 4), dtype=np.uint8)
    cnt = np.zeros(5, dtype=np.uint64) if counters else None
    threads = threads or num_threads()
    ms = lib().rto_trace_packet(kind, 0 if mode == "closest" else 1, nodes, n_nodes, idx, _p(verts), _p(packets), n,
                                t_min, _p(hits), _p(occ), _p(cnt), threads)
    cd = dict(zip(COUNTER_NAMES, (int(x) for x in cnt))) if counters else None
    return (hits if mode == "closest" else occ), ms, cd


def brute_force(verts, rays, threads: int = 0):
    verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 9)
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    hits = np.zeros(len(rays), dtype=HIT_DTYPE)
    lib().rto_brute_force(_p(verts), verts.shape[0], _p(rays), len(rays), _p(hits), threads or num_threads())
    return hits


CALLBACK = C.CFUNCTYPE(C.c_bool, C.c_uint32, C.POINTER(C.c_float), C.c_void_p)


def intersect_cb(tree, origin, direction, t: float, cb):
    """Legacy single-ray callback walk (rtbvh_ffi/src/lib.rs:551-581, :700-731). Returns (code, t)."""
    kind, nodes, n_nodes, idx = _tree_args(tree)
    o = np.asarray(origin, dtype=np.float32)
    d = np.asarray(direction, dtype=np.float32)
    tv = C.c_float(t)
    rc = lib().rto_intersect_cb(kind, nodes, n_nodes, idx, _p(o), _p(d), C.byref(tv), None, C.cast(cb, C.c_void_p))
    return rc, tv.value
