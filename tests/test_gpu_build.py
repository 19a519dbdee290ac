"""GPU builders (create_bvh / create_mbvh / refit / rtbvh_gpu_create_bvh_triangles) against the CPU oracle.

Contract (BASELINE.json north_star): GPU-built trees come within 3 % of the reference construct_binned_sah SAH
cost and give identical hit ids on a 1 M-ray probe.  The builders reproduce the reference's arithmetic and
decisions, so the tests demand more: binned SAH trees are ISOMORPHIC to the oracle's (same topology, bit-equal
boxes, same leaf primitive sets; only node numbering / in-leaf order differ, which the reference itself leaves
to thread scheduling), LOCB trees and collapsed Mbvh node arrays are BYTE-IDENTICAL."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SAH_TOL = 1.03  # contractual bound; assertions below are exact


@pytest.fixture(scope="module")
def A():
    from rtbvh_b200 import api
    if api.device_count() == 0:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the B200 box")
    return api


def assert_isomorphic(a_nodes, a_idx, b_nodes, b_idx):
    """Same tree up to node numbering and primitive order inside leaves; boxes compared bit for bit."""
    assert len(a_nodes) == len(b_nodes)
    stack = [(0, 0)]
    visited = 0
    while stack:
        x, y = stack.pop()
        na, nb = a_nodes[x], b_nodes[y]
        visited += 1
        assert na["min"].tobytes() == nb["min"].tobytes() and na["max"].tobytes() == nb["max"].tobytes(), (x, y, na, nb)
        assert na["count"] == nb["count"], (x, y, na, nb)
        if na["count"] >= 0:
            pa = np.sort(a_idx[na["left_first"]:na["left_first"] + na["count"]])
            pb = np.sort(b_idx[nb["left_first"]:nb["left_first"] + nb["count"]])
            assert np.array_equal(pa, pb)
        else:
            stack.append((int(na["left_first"]), int(nb["left_first"])))
            stack.append((int(na["left_first"]) + 1, int(nb["left_first"]) + 1))
    assert visited == len(a_nodes)


def _scenes(W):
    return {"teapot": W.teapot(), "soup20k": W.soup(20_000), "field": W.heightfield(60, 60)}


@pytest.mark.parametrize("scene", ["teapot", "soup20k", "field"])
@pytest.mark.parametrize("leaf", [1, 4])
def test_binned_sah_isomorphic_to_oracle(A, O, W, scene, leaf):
    tris = _scenes(W)[scene]
    aabbs, centers = O.prims_from_triangles(tris)
    rc, want = O.build(O.BINNED_SAH, aabbs, centers, leaf)
    assert rc == 0
    got = A.Builder(aabbs, centers, leaf).construct_binned_sah()
    assert got.validate(len(tris))
    assert_isomorphic(got.nodes, got.indices, want.nodes, want.indices)
    sah_g = O.Bvh(got.nodes.copy(), got.indices.copy()).sah_cost()
    assert sah_g <= want.sah_cost() * SAH_TOL and abs(sah_g - want.sah_cost()) < 1e-9 * want.sah_cost()
    # level-order numbering: children after parents (Bvh::refit relies on it, src/bvh.rs:177)
    inner = got.nodes["count"] < 0
    assert np.all(got.nodes["left_first"][inner] > np.nonzero(inner)[0])
    got.free()


@pytest.mark.parametrize("scene", ["teapot", "soup20k", "field"])
def test_locb_identical_to_oracle(A, O, W, scene):
    tris = _scenes(W)[scene]
    aabbs, centers = O.prims_from_triangles(tris)
    rc, want = O.build(O.LOCB, aabbs, centers)
    got = A.Builder(aabbs, centers).construct_locally_ordered_clustered()
    assert np.array_equal(got.indices, want.indices)
    assert got.nodes.tobytes() == want.nodes.tobytes()
    assert A.last_build_stats()["iterations"] > 10
    got.free()


@pytest.mark.parametrize("kind", [0, 1])
def test_collapse_identical_to_oracle(A, O, W, teapot, kind):
    bvh = A.Builder(teapot["aabbs"], teapot["centers"], 1)._construct(kind)
    m = A.Mbvh.construct(bvh)
    want = O.Bvh(bvh.nodes.copy(), bvh.indices.copy()).collapse()  # merge_nodes on the very same binary tree
    assert len(m.nodes) == len(want.nodes)
    assert m.nodes.tobytes() == want.nodes.tobytes()
    assert np.array_equal(m.indices, bvh.indices)
    m.free()
    bvh.free()


def test_triangle_front_end_equals_create_bvh(A, O, W, teapot):
    a = A.build_triangles(teapot["tris"], A.BINNED_SAH, 1)
    b = A.Builder(teapot["aabbs"], teapot["centers"], 1).construct_binned_sah()
    assert a.nodes.tobytes() == b.nodes.tobytes() and np.array_equal(a.indices, b.indices)
    a4 = np.zeros((len(teapot["tris"]), 3, 4), np.float32)  # 16-byte vertex stride
    a4[:, :, :3] = teapot["tris"]
    c = A.build_triangles(a4, A.LOCALLY_ORDERED_CLUSTERED)
    d = A.Builder(teapot["aabbs"], teapot["centers"]).construct_locally_ordered_clustered()
    assert c.nodes.tobytes() == d.nodes.tobytes()
    for t in (a, b, c, d):
        t.free()


def test_probe_ids_match_reference_built_tree(A, O, W):
    """1 M-ray probe: the GPU-built tree and the oracle-built tree give identical hit ids and t (both traced on the GPU;
    the oracle tree's GPU traversal is itself checked against the CPU in test_gpu_traversal)."""
    tris = W.soup(100_000)
    aabbs, centers = O.prims_from_triangles(tris)
    rays = np.concatenate([W.camera_rays(W.soup_camera(800, 800), jitter_seed=9), W.random_rays(360_000, *W.bounds(tris))])
    assert len(rays) == 1_000_000
    for kind, okind in ((A.BINNED_SAH, O.BINNED_SAH), (A.LOCALLY_ORDERED_CLUSTERED, O.LOCB)):
        g = A.Builder(aabbs, centers, 1)._construct(kind)
        gm = A.Mbvh.construct(g)
        rc, o = O.build(okind, aabbs, centers, 1)
        om = o.collapse()
        sg = A.Scene(tris, bvh=g, mbvh=gm)
        so = A.Scene(tris, bvh=A.Bvh.from_arrays(o.nodes, o.indices), mbvh=A.Mbvh.from_arrays(om.nodes, om.indices))
        for tree in (A.TREE_BVH, A.TREE_MBVH):
            hg, ho = sg.intersect(rays, tree), so.intersect(rays, tree)
            assert np.array_equal(hg["prim"], ho["prim"]) and np.array_equal(hg["t"], ho["t"])
        # and against the CPU oracle on a sample
        sample = rays[::50]
        want, _, _ = O.trace(om, tris, sample)
        assert np.array_equal(sg.intersect(sample, A.TREE_MBVH), want)
        sg.free(); so.free(); gm.free(); g.free()


def test_refit_identical_to_oracle(A, O, W, teapot):
    bvh = A.Builder(teapot["aabbs"], teapot["centers"], 1).construct_binned_sah()
    before = O.Bvh(bvh.nodes.copy(), bvh.indices.copy())
    moved = teapot["aabbs"].copy()
    rng = np.random.default_rng(3)
    d = rng.uniform(-0.05, 0.05, (len(moved), 3)).astype(np.float32)
    moved["min"] += d
    moved["max"] += d
    bvh.refit(moved)
    want = before.refit(moved)
    assert bvh.nodes.tobytes() == want.nodes.tobytes()
    bvh.free()


# ---- the reference's own contract tests, through this library ------------------------------------
def test_ffi_create_delete(A, O):
    # rtbvh_ffi/src/lib.rs:869-943
    verts = np.array([[x, y, 0] for x in range(10) for y in range(10)], dtype=np.float32)
    aabbs, _ = O.prims_from_triangles(verts[:81].reshape(27, 3, 3), pad=1e-4)
    c16 = np.zeros((27, 4), np.float32)
    c16[:, :3] = O.aabb_centers(aabbs)
    L = A.lib()
    out = A.RTBvh(0xFFFFFFFF, 0, None, 0, None)
    assert L.create_bvh(None, 27, None, 16, 1, A.BINNED_SAH, C.byref(out)) == A.ERROR
    assert L.create_bvh(None, 27, A._p(c16), 16, 1, A.BINNED_SAH, C.byref(out)) == A.OK
    L.free_bvh(out)
    assert L.create_bvh(A._p(aabbs), 27, A._p(c16), 16, 1, A.BINNED_SAH, C.byref(out)) == A.OK
    m = A.RTMbvh(0xFFFFFFFF, 0, None, 0, None)
    assert L.create_mbvh(out, C.byref(m)) == A.OK and m.node_count >= 1
    first_id = out.id
    L.free_bvh(out)
    L.free_mbvh(m)
    assert L.create_bvh(A._p(aabbs), 27, A._p(c16), 16, 1, A.BINNED_SAH, C.byref(out)) == A.OK
    assert out.id > first_id  # ids are never reused (lib.rs:46-60)
    L.free_bvh(out)


def test_ffi_intersect_kat_end_to_end(A, O, W):
    # rtbvh_ffi/src/lib.rs:946-1019 with the tree built by create_bvh on the GPU
    tris = W.quad()
    aabbs, _ = O.prims_from_triangles(tris, pad=1e-4)
    bvh = A.Builder(aabbs, O.aabb_centers(aabbs), 1).construct_binned_sah()
    m = A.Mbvh.construct(bvh)
    sc = A.Scene(tris, bvh=bvh, mbvh=m)
    rays = W.make_rays(np.zeros((1, 3), np.float32), np.array([[0, 0, 1]], np.float32), t_max=np.float32(1e26))
    for kind in (A.TREE_BVH, A.TREE_MBVH):
        h = sc.intersect(rays, kind)
        assert abs(h["t"][0] - 1.0) < np.finfo(np.float32).eps and h["prim"][0] == 0
    sc.free(); m.free(); bvh.free()


def test_invalid_input_and_small_inputs(A, O):
    # src/lib.rs:28-64 test_invalid_input
    with pytest.raises(A.RtbvhError) as e:
        A.Builder(None, np.zeros((0, 3), np.float32)).construct_binned_sah()
    assert e.value.code == A.NO_PRIMITIVES
    tri = np.zeros((1, 3, 3), np.float32)
    aabbs, centers = O.prims_from_triangles(tri)
    b = A.Builder(aabbs, centers).construct_binned_sah()  # one degenerate primitive is Ok
    rc, want = O.build(O.BINNED_SAH, aabbs, centers)
    assert b.nodes.tobytes() == want.nodes.tobytes()
    with pytest.raises(A.RtbvhError) as e:
        A.Builder(np.zeros(0, A.NODE_DTYPE), centers).construct_binned_sah()
    assert e.value.code == A.INEQUAL_AABBS_AND_PRIMITIVES
    # LOCB with 1 and 2 primitives: a single root leaf (locb.rs:258-269); 3 primitives: a real tree
    for n in (1, 2, 3):
        t = np.random.default_rng(n).random((n, 3, 3)).astype(np.float32)
        ab, ce = O.prims_from_triangles(t)
        g = A.Builder(ab, ce).construct_locally_ordered_clustered()
        rc, w = O.build(O.LOCB, ab, ce)
        assert g.nodes.tobytes() == w.nodes.tobytes() and np.array_equal(g.indices, w.indices)
        assert A.Mbvh.construct(g).nodes.tobytes() == w.collapse().nodes.tobytes()


def test_five_triangle_case(A, O):
    # src/lib.rs:246-311: leaf sizes 1..=10 on 5 coplanar near-degenerate triangles, then Mbvh::from
    t = np.array([
        [[128.79, -1422.82, 0.16], [128.5, -1426.88, 0.16], [128.79, -1426.9067, 0.16]],
        [[129.8, -1422.8629, 0.16], [128.79, -1422.82, 0.16], [128.79, -1426.9067, 0.16]],
        [[129.8, -1422.8629, 0.16], [128.79, -1426.9067, 0.16], [129.8, -1427.0, 0.16]],
        [[130.2, -1422.88, 0.16], [129.8, -1422.8629, 0.16], [129.8, -1427.0, 0.16]],
        [[130.2, -1422.88, 0.16], [129.8, -1427.0, 0.16], [130.2, -1423.13, 0.16]],
    ], dtype=np.float32)
    for leaf in range(1, 11):
        g = A.build_triangles(t, A.BINNED_SAH, leaf)
        aabbs, centers = O.prims_from_triangles(t)
        rc, w = O.build(O.BINNED_SAH, aabbs, centers, leaf)
        assert_isomorphic(g.nodes, g.indices, w.nodes, w.indices)
        m = A.Mbvh.construct(g)
        assert m.nodes.tobytes() == O.Bvh(g.nodes.copy(), g.indices.copy()).collapse().nodes.tobytes()


def test_duplicate_primitives_force_multi_prim_leaves(A, O, W):
    # identical centroids cannot be separated: leaves keep several primitives although primitives_per_leaf = 1, the
    # small-subtree path reserves more node slots than it uses and the node array must be compacted
    base = W.soup(3000, seed=11)
    tris = np.repeat(base, 3, axis=0)
    aabbs, centers = O.prims_from_triangles(tris)
    for leaf in (1, 2, 5):
        g = A.Builder(aabbs, centers, leaf).construct_binned_sah()
        rc, w = O.build(O.BINNED_SAH, aabbs, centers, leaf)
        assert g.validate(len(tris))
        assert_isomorphic(g.nodes, g.indices, w.nodes, w.indices)
        m = A.Mbvh.construct(g)
        assert m.nodes.tobytes() == O.Bvh(g.nodes.copy(), g.indices.copy()).collapse().nodes.tobytes()


def test_point_primitives_without_aabbs(A, O):
    # create_bvh with aabbs == null: centers act as point primitives (rtbvh_ffi/src/lib.rs:396-422)
    pts = np.random.default_rng(5).random((5000, 3)).astype(np.float32)
    for kind, okind in ((A.BINNED_SAH, O.BINNED_SAH), (A.LOCALLY_ORDERED_CLUSTERED, O.LOCB)):
        g = A.Builder(None, pts, 2)._construct(kind)
        rc, w = O.build(okind, None, pts, 2)
        assert_isomorphic(g.nodes, g.indices, w.nodes, w.indices)


def test_full_size_build_soup_1m(A, O, W):
    """BASELINE config 2 geometry at full size: structural properties + SAH equal to the oracle's."""
    tris = W.soup(1 << 20)
    g = A.build_triangles(tris, A.BINNED_SAH, 1)
    stats = A.last_build_stats()
    assert g.validate(len(tris))
    nodes = g.nodes
    inner = np.nonzero(nodes["count"] < 0)[0]
    l = nodes["left_first"][inner]
    # every inner box encloses its children's boxes up to quirk Q3 (non-conservative left boxes are rare)
    bad = (nodes["min"][l] < nodes["min"][inner] - 2e-4).any(axis=1) | (nodes["max"][l + 1] > nodes["max"][inner] + 2e-4).any(axis=1)
    assert bad.mean() < 0.01
    aabbs, centers = O.prims_from_triangles(tris)
    rc, w = O.build(O.BINNED_SAH, aabbs, centers, 1)
    assert len(nodes) == len(w.nodes)
    sah_g = O.Bvh(nodes.copy(), g.indices.copy()).sah_cost()
    assert abs(sah_g - w.sah_cost()) < 1e-9 * w.sah_cost()
    m = A.Mbvh.construct(g)
    assert len(m.nodes) == len(w.collapse().nodes)
    print("build stats", stats)


@pytest.mark.gpu
def test_level_loop_variants_build_identical_trees():
    """The binned-SAH level loop has three switchable pieces (block-chunk partition vs CUB scan-by-key, split merged into
    the warp-task launch, one-block scan + emit); every combination must produce the legacy path's tree byte for byte
    (a stable partition and an exclusive scan have one result).  The knobs are read once per process, so
    scripts/partition_ab.py runs each mode in a child process."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, AB_QUICK="1")
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "partition_ab.py")], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    summary = json.loads(r.stdout.strip().splitlines()[-1])
    assert set(summary["identical_to_legacy"]) == {"legacy", "part", "part+merge", "part+scanemit", "all"}
    assert all(summary["identical_to_legacy"].values()), summary
