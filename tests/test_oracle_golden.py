"""Oracle regression vectors (tests/golden/oracle_golden.npz, made by tests/golden/make_fixtures.py) and
oracle-vs-brute-force cross checks.  CPU only."""
import hashlib
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.npz")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["sah", "locb"])
def test_oracle_matches_golden(O, W, teapot, teapot_trees, name):
    g = np.load(GOLDEN)
    bvh, m = teapot_trees[name]
    assert sha(bvh.nodes) == str(g[name + "_nodes_sha"])
    assert sha(bvh.indices) == str(g[name + "_indices_sha"])
    assert sha(m.nodes) == str(g[name + "_mnodes_sha"])
    assert bvh.sah_cost() == float(g[name + "_sah"])
    cam = W.benchmark_camera(256, 256)
    rays = np.concatenate([W.camera_rays(cam), W.random_rays(16384, *W.bounds(teapot["tris"]))])
    h2, _, _ = O.trace(bvh, teapot["tris"], rays)
    h4, _, _ = O.trace(m, teapot["tris"], rays)
    assert np.array_equal(h2, g[name + "_bvh_hits"]) and np.array_equal(h4, g[name + "_mbvh_hits"])
    # any hit and packets of four (closest + any), pinned on a subset of the same rays
    sub = rays[::5][: 16384 // 4 * 4]
    packets = W.pack4(sub)
    for tag, tree in (("bvh", bvh), ("mbvh", m)):
        assert np.array_equal(O.trace(tree, teapot["tris"], sub, mode="any")[0], g[f"{name}_{tag}_any"])
        assert np.array_equal(O.trace_packets(tree, teapot["tris"], packets)[0], g[f"{name}_{tag}_packet_hits"])
        assert np.array_equal(O.trace_packets(tree, teapot["tris"], packets, mode="any")[0], g[f"{name}_{tag}_packet_any"])


def test_refit_and_spatial_builder_match_golden(O, teapot, teapot_trees):
    g = np.load(GOLDEN)
    bvh, _ = teapot_trees["sah"]
    moved = teapot["aabbs"].copy()
    moved["min"] += np.float32(0.25)
    moved["max"] += np.float32(0.5)
    assert sha(bvh.refit(moved).nodes) == str(g["refit_nodes_sha"])
    rc, sp = O.build_spatial(teapot["tris"][:2000], 1)
    assert rc == 0
    assert sha(sp.nodes) == str(g["spatial_nodes_sha"]) and sha(sp.indices) == str(g["spatial_indices_sha"])


def test_locb_tree_equals_brute_force(O, W, teapot, teapot_trees):
    # LOCB boxes are plain unions (conservative), so tree traversal must equal the brute-force arbiter exactly
    tris = teapot["tris"]
    rays = np.concatenate([W.camera_rays(W.benchmark_camera(96, 96)), W.random_rays(4096, *W.bounds(tris))])
    bf = O.brute_force(tris, rays)
    for tree in teapot_trees["locb"]:
        hits, _, _ = O.trace(tree, tris, rays)
        assert np.array_equal(hits["t"], bf["t"]) and np.array_equal(hits["prim"], bf["prim"])


def test_sah_tree_vs_brute_force_documents_q3(O, W, teapot, teapot_trees):
    # quirk Q3 (binned_sah.rs:232-235) can make left boxes non-conservative: the tree may miss hits the
    # brute force finds, never the other way round (t_tree >= t_bf), and only on a tiny fraction of rays.
    tris = teapot["tris"]
    rays = W.camera_rays(W.benchmark_camera(128, 128))
    bf = O.brute_force(tris, rays)
    for tree in teapot_trees["sah"]:
        hits, _, _ = O.trace(tree, tris, rays)
        assert np.all(hits["t"] >= bf["t"])
        assert (hits["t"] != bf["t"]).mean() < 1e-3


def test_any_hit_equals_closest_hit_predicate(O, W, teapot, teapot_trees):
    tris = teapot["tris"]
    rays = W.random_rays(8192, *W.bounds(tris))
    for tree in teapot_trees["locb"]:
        hits, _, _ = O.trace(tree, tris, rays)
        occ, _, _ = O.trace(tree, tris, rays, mode="any")
        assert np.array_equal(occ.astype(bool), hits["prim"] != O.NO_HIT)


def test_packets_agree_with_single_rays_on_conservative_tree(O, W, teapot, teapot_trees):
    # packet triangle test uses different constants (quirk Q7: eps 1e-6, t >= t_min) so t may differ on
    # grazing hits; on this camera every lane must agree to 1e-5 relative and ids must match where t matches
    tris = teapot["tris"]
    rays = W.camera_rays(W.benchmark_camera(128, 128))
    packets = W.pack4(rays)
    for tree in teapot_trees["locb"]:
        h1, _, _ = O.trace(tree, tris, rays)
        h4, _, _ = O.trace_packets(tree, tris, packets)
        t4 = h4["t"].reshape(-1)
        p4 = h4["prim"].reshape(-1)
        same = t4 == h1["t"]
        assert same.mean() > 0.999
        assert np.array_equal(p4[same], h1["prim"][same])
        occ, _, _ = O.trace_packets(tree, tris, packets, mode="any")
        assert np.array_equal(occ.reshape(-1).astype(bool), p4 != O.NO_HIT)


def test_counters_and_stack_depth(O, W, teapot, teapot_trees):
    tris = teapot["tris"]
    rays = W.camera_rays(W.benchmark_camera(200, 200))
    bvh, m = teapot_trees["sah"]
    _, _, c2 = O.trace(bvh, tris, rays, counters=True)
    _, _, c4 = O.trace(m, tris, rays, counters=True)
    n = len(rays)
    assert 25 < c2["node_visits"] / n < 35 and 7 < c4["node_visits"] / n < 10  # 29.8 / 8.6
    assert c2["max_stack"] <= 32 and c4["max_stack"] <= 32 and c2["overflow32"] == 0


def test_refit_keeps_topology_and_bounds(O, teapot, teapot_trees):
    bvh, _ = teapot_trees["sah"]
    moved = teapot["aabbs"].copy()
    moved["min"] += np.float32(0.25)
    moved["max"] += np.float32(0.25)
    r = bvh.refit(moved)
    assert np.array_equal(r.nodes["count"], bvh.nodes["count"]) and np.array_equal(r.nodes["left_first"], bvh.nodes["left_first"])
    assert r.validate(len(moved))
    # every leaf box contains its primitives, every inner box its children
    for k, nd in enumerate(r.nodes):
        if nd["count"] >= 0:
            ids = r.indices[nd["left_first"]:nd["left_first"] + nd["count"]]
            assert np.all(moved["min"][ids] >= nd["min"]) and np.all(moved["max"][ids] <= nd["max"])
        else:
            for c in (nd["left_first"], nd["left_first"] + 1):
                assert np.all(r.nodes[c]["min"] >= nd["min"]) and np.all(r.nodes[c]["max"] <= nd["max"])


def test_leaf_depth_stats_for_the_build_roofline(O, W, teapot, teapot_trees):
    """bench.py's builder roofline uses D-bar = mean leaf depth per primitive: the level-by-level
    numpy walk must equal a plain recursive walk, and reproduce the survey's teapot figures (6 154 leaves, depth 16)."""
    nodes = teapot_trees["sah"][0].nodes
    st = W.leaf_depth_stats(nodes)
    acc = []
    stack = [(0, 0)]
    while stack:
        i, d = stack.pop()
        lf, cnt = int(nodes[i]["left_first"]), int(nodes[i]["count"])
        if lf < 0:
            continue
        if cnt >= 0:
            acc.append((d, cnt))
        else:
            stack += [(lf, d + 1), (lf + 1, d + 1)]
    assert st["leaves"] == len(acc) == 6154 and st["prims"] == sum(c for _, c in acc) == 6320
    assert st["max_depth"] == max(d for d, _ in acc) == 16
    assert abs(st["mean_leaf_depth_per_prim"] - sum(d * c for d, c in acc) / 6320) < 1e-12
    assert abs(W.binned_sah_bytes_per_tri(st["mean_leaf_depth_per_prim"]) - (148 + 60 * st["mean_leaf_depth_per_prim"])) < 1e-9
    # a single-leaf tree: depth 0
    one = nodes[:1].copy()
    one["count"], one["left_first"] = 3, 0
    assert W.leaf_depth_stats(one) == {"leaves": 1, "prims": 3, "max_depth": 0, "mean_leaf_depth": 0.0, "mean_leaf_depth_per_prim": 0.0}


@pytest.mark.parametrize("n", [1, 2, 3, 17, 1000, 20000])
def test_locb_trees_equal_brute_force_on_random_soups(O, W, n):
    """Size-independent property: a LOCB tree has conservative boxes (plain unions, no quirk Q3), so
    closest hit through the Bvh and through its collapsed Mbvh must equal the brute-force arbiter record for record
    (same triangle test, lowest id on exact ties), for ragged sizes down to a single triangle."""
    tris = W.soup(n, seed=0xC0FFEE + n)
    aabbs, centers = O.prims_from_triangles(tris)
    rc, bvh = O.build(O.LOCB, aabbs, centers, 1)
    assert rc == 0 and bvh.validate(n)
    m = bvh.collapse()
    lo, hi = W.bounds(tris)
    rays = W.random_rays(4000, lo, hi, seed=7 + n)
    bf = O.brute_force(tris, rays)
    for tree in (bvh, m):
        got, _, _ = O.trace(tree, tris, rays)
        assert np.array_equal(got, bf)
        occ, _, _ = O.trace(tree, tris, rays, mode="any")
        assert np.array_equal(occ.astype(bool), bf["prim"] != 0xFFFFFFFF)
