"""The reference's own known-answer / contract tests, ported against the CPU oracle.

Each test names the reference test it restates.  These pin the oracle; they run without a GPU.
"""
import ctypes as C

import numpy as np
import pytest


# src/morton.rs:109-115  morton_split_works
def test_morton_split_works(O):
    L = O.lib()
    assert L.rto_morton_split(2) == 8
    assert L.rto_morton_split(4) == 64
    assert L.rto_morton_split(8) == 512
    assert L.rto_morton_split(32) == 32768


def test_morton_split_equals_part1by2(O):
    # SURVEY a-17: equal to the classic 0x09249249 part-by-2 for all 10-bit inputs
    def part(v):
        v &= 0x3FF
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    L = O.lib()
    for v in range(1024):
        assert L.rto_morton_split(v) == part(v)


# src/utils.rs:295-345  prefix_sum_{u32,usize,i32}_works, prefix_sum_zero
@pytest.mark.parametrize("fn,dt", [("rto_prefix_sum_u32", np.uint32), ("rto_prefix_sum_u64", np.uint64),
                                   ("rto_prefix_sum_i32", np.int32)])
def test_prefix_sum_works(O, fn, dt):
    inp = np.array([1, 2, 3, 4, 5, 6], dtype=dt)
    out = np.zeros(6, dtype=dt)
    total = getattr(O.lib(), fn)(O._p(inp), 6, O._p(out))
    assert out.tolist() == [1, 3, 6, 10, 15, 21]
    assert total == 21


def test_prefix_sum_zero(O):
    inp = np.array([7, 2, 3], dtype=np.uint32)
    out = np.zeros(3, dtype=np.uint32)
    assert O.lib().rto_prefix_sum_u32(O._p(inp), 0, O._p(out)) == 7  # count == 0 returns first[0]
    assert out.tolist() == [0, 0, 0]


# src/utils.rs:347-368  test_move_backwards
def test_move_backwards(O):
    a = np.array([1, 2, 3, 4, 5, 6, 0, 0], dtype=np.uint32)
    base = a.ctypes.data
    O.lib().rto_move_backward_u32(C.c_void_p(base), C.c_void_p(base + 6 * 4), C.c_void_p(base + 8 * 4))
    assert a.tolist() == [1, 2, 1, 2, 3, 4, 5, 6]


def test_partition_is_the_swap_partition(O):
    # src/utils.rs:76-96: `true` elements keep their order on the left, the right side is permuted by swaps
    a = np.array([5, 1, 7, 2, 9, 3], dtype=np.uint32)
    k = O.lib().rto_partition_lt(O._p(a), 6, 4)
    assert k == 3 and a[:3].tolist() == [1, 2, 3] and sorted(a[3:].tolist()) == [5, 7, 9]
    assert a.tolist() == [1, 2, 3, 5, 9, 7]


# rtbvh_ffi/src/lib.rs:856-866  same_size
def test_same_size(O):
    L = O.lib()
    assert L.rto_sizeof_aabb() == 32 and L.rto_sizeof_bvh_node() == 32 and L.rto_sizeof_mbvh_node() == 128


def _quad_flat():
    v = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], dtype=np.float32)
    return np.stack([v[[0, 1, 2]], v[[0, 2, 3]]])


# src/lib.rs:28-64  test_invalid_input
def test_invalid_input(O):
    assert O.build(O.BINNED_SAH, None, np.zeros((0, 3), np.float32))[0] == 2  # NoPrimitives
    tri = np.zeros((1, 3, 3), np.float32)
    aabbs, centers = O.prims_from_triangles(tri)
    rc, bvh = O.build(O.BINNED_SAH, aabbs, centers)
    assert rc == 0 and len(bvh.nodes) == 1  # one degenerate primitive is Ok
    rc, _ = O.build(O.BINNED_SAH, np.zeros(0, O.NODE_DTYPE), centers)
    assert rc == 3  # InequalAabbsAndPrimitives(0, 1)


# src/lib.rs:66-124  test_sah / test_locb
@pytest.mark.parametrize("kind", [0, 1])
def test_two_triangle_quad_builds(O, kind):
    aabbs, centers = O.prims_from_triangles(_quad_flat())
    rc, bvh = O.build(kind, aabbs, centers)
    assert rc == 0 and bvh.validate(2)


# src/lib.rs:246-311  five_triangle_test_case (must not panic for leaf sizes 1..=10, then Mbvh::from)
def test_five_triangle_case(O):
    t = np.array([
        [[128.79, -1422.82, 0.16], [128.5, -1426.88, 0.16], [128.79, -1426.9067, 0.16]],
        [[129.8, -1422.8629, 0.16], [128.79, -1422.82, 0.16], [128.79, -1426.9067, 0.16]],
        [[129.8, -1422.8629, 0.16], [128.79, -1426.9067, 0.16], [129.8, -1427.0, 0.16]],
        [[130.2, -1422.88, 0.16], [129.8, -1422.8629, 0.16], [129.8, -1427.0, 0.16]],
        [[130.2, -1422.88, 0.16], [129.8, -1427.0, 0.16], [130.2, -1423.13, 0.16]],
    ], dtype=np.float32)
    aabbs, centers = O.prims_from_triangles(t)
    for leaf in range(1, 11):
        rc, bvh = O.build(O.BINNED_SAH, aabbs, centers, leaf)
        assert rc == 0 and bvh.validate(5)
        m = bvh.collapse()
        assert len(m.nodes) >= 1


# src/builders/binned_sah.rs:408-458, src/builders/locb.rs:337-388  no_primitives + test_*_build on the teapot
@pytest.mark.parametrize("name", ["sah", "locb"])
def test_teapot_structure(O, teapot, teapot_trees, name):
    bvh, m = teapot_trees[name]
    n = len(teapot["tris"])
    assert n <= len(bvh.nodes) <= 2 * n
    root = bvh.nodes[0]
    assert np.all(root["min"] <= root["max"])  # bounds.is_valid()
    assert bvh.validate(n)
    v = teapot["tris"].reshape(-1, 3)
    assert np.all(v > root["min"]) and np.all(v < root["max"])  # bounds.contains(vertex) is strict
    if name == "locb":
        assert len(bvh.nodes) == 2 * n - 1 and np.all(bvh.nodes["count"][bvh.nodes["count"] >= 0] == 1)


def test_teapot_numbers_match_survey(teapot_trees):
    # SURVEY.md §8 (marked †): independent throw-away restatement made during the survey
    bvh, m = teapot_trees["sah"]
    leaves = bvh.nodes["count"][bvh.nodes["count"] >= 0]
    assert len(bvh.nodes) == 12307 and len(leaves) == 6154
    assert np.bincount(leaves).tolist() == [0, 6015, 112, 27]
    assert len(m.nodes) == 3061
    used = (m.nodes["children"] >= 0).sum(axis=1)
    assert np.bincount(used).tolist() == [0, 0, 1382, 266, 1413]
    assert abs(bvh.sah_cost() - 25.73) < 0.01
    assert abs(teapot_trees["locb"][0].sah_cost() - 27.3) < 0.05
    assert bvh.depth_stats()[1] == 16


def _ffi_quad(O, W):
    tris = W.quad()
    aabbs, _ = O.prims_from_triangles(tris, pad=1e-4)  # aabb!(v0, v1, v2)
    centers = O.aabb_centers(aabbs)                    # bb.center()
    return tris, aabbs, centers


def _intersect_test_cb(O, tris, origin, direction):
    """rtbvh_ffi/src/lib.rs:1028-1060 intersect_test: Moeller-Trumbore, `t_val > 1e-5 && t_val < *t`, returns false."""
    o = np.asarray(origin, np.float32)
    d = np.asarray(direction, np.float32)
    f32 = np.float32

    def dot(a, b):
        return f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))

    def cross(a, b):
        return np.array([f32(a[1] * b[2]) - f32(b[1] * a[2]), f32(a[2] * b[0]) - f32(b[2] * a[0]),
                         f32(a[0] * b[1]) - f32(b[0] * a[1])], dtype=f32)

    def cb(prim, t_ptr, user):
        v0, v1, v2 = tris[prim]
        e1, e2 = v1 - v0, v2 - v0
        h = cross(d, e2)
        a = dot(e1, h)
        if -1e-5 < a < 1e-5:
            return False
        f = f32(1.0) / a
        s = o - v0
        u = f32(f * dot(s, h))
        if not (0.0 <= u <= 1.0):
            return False
        q = cross(s, e1)
        v = f32(f * dot(d, q))
        if v < 0 or u + v > 1.0:
            return False
        tv = f32(f * dot(e2, q))
        if tv > 1e-5 and tv < t_ptr[0]:
            t_ptr[0] = tv
        return False
    return O.CALLBACK(cb)


# rtbvh_ffi/src/lib.rs:946-1019  intersect — the only numeric traversal KAT of the reference
def test_ffi_intersect_kat(O, W):
    tris, aabbs, centers = _ffi_quad(O, W)
    rc, bvh = O.build(O.BINNED_SAH, aabbs, centers, 1)
    assert rc == 0
    m = bvh.collapse()
    cb = _intersect_test_cb(O, tris, (0, 0, 0), (0, 0, 1))
    eps = np.finfo(np.float32).eps
    for tree in (bvh, m):
        rc, t = O.intersect_cb(tree, (0, 0, 0), (0, 0, 1), 1e26, cb)
        assert rc == 0 and abs(t - 1.0) < eps
    # NaN origin -> ResultCode::Nan (lib.rs:562-564)
    assert O.intersect_cb(bvh, (np.nan, 0, 0), (0, 0, 1), 1e26, cb)[0] == 4
    # the batch flavour agrees and reports which triangle (id 0 and 1 share the diagonal: lowest id wins)
    rays = W.make_rays(np.zeros((1, 3), np.float32), np.array([[0, 0, 1]], np.float32), t_max=np.float32(1e26))
    for tree in (bvh, m):
        hits, _, _ = O.trace(tree, tris, rays)
        assert abs(hits["t"][0] - 1.0) < eps and hits["prim"][0] == 0


# rtbvh_ffi/src/lib.rs:869-943  create_delete (27 "triangles" on a 10x10 grid of points)
def test_ffi_create_delete_codes(O):
    verts = np.array([[x, y, 0] for x in range(10) for y in range(10)], dtype=np.float32)
    tris = verts[:81].reshape(27, 3, 3)
    aabbs, _ = O.prims_from_triangles(tris, pad=1e-4)
    centers = O.aabb_centers(aabbs)
    assert O.build(O.BINNED_SAH, None, None, prim_count=27)[0] == 1             # null centers -> Error
    c16 = np.zeros((27, 4), np.float32)
    c16[:, :3] = centers
    rc, bvh = O.build(O.BINNED_SAH, None, c16, 1)                                # null aabbs is fine, stride 16
    assert rc == 0 and bvh.validate(27)
    rc, bvh = O.build(O.BINNED_SAH, aabbs, c16, 1)
    assert rc == 0 and bvh.validate(27)
    assert len(bvh.collapse().nodes) >= 1


def test_threaded_builder_is_the_serial_builder_up_to_numbering(O, W):
    """oracle build(parallel=True) = the reference's threaded scheduling (subtrees > 1024 primitives on other threads,
    src/utils.rs:189-289): same topology, boxes and leaves as the deterministic builder; only the node numbering may differ."""
    tris = W.soup(60_000, seed=11)
    aabbs, centers = O.prims_from_triangles(tris)
    rc0, a = O.build(O.BINNED_SAH, aabbs, centers, 1)
    rc1, b = O.build(O.BINNED_SAH, aabbs, centers, 1, parallel=True)
    assert rc0 == 0 and rc1 == 0
    assert len(a.nodes) == len(b.nodes)
    assert abs(a.sah_cost() - b.sah_cost()) <= 1e-9 * a.sah_cost()

    def signature(t):  # multiset of nodes as (box bits, count, first primitive id of a leaf)
        n = t.nodes
        rows = []
        for k in range(len(n)):
            cnt = int(n["count"][k])
            first = int(t.indices[int(n["left_first"][k])]) if cnt > 0 else -1
            rows.append((n["min"][k].tobytes(), n["max"][k].tobytes(), cnt, first))
        return sorted(rows)

    assert signature(a) == signature(b)
