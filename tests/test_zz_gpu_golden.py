"""The CUDA path against the COMMITTED golden vectors (tests/golden/oracle_golden.npz, made by make_fixtures.py from the
pinned CPU oracle): nothing is traced or built by the oracle at run time here except the reference-format teapot trees that
are uploaded unchanged.  (The file name sorts last on purpose: these checks repeat, against frozen data, what
test_gpu_traversal.py / test_gpu_build.py check against the live oracle.)"""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.npz")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def A():
    from rtbvh_b200 import api
    if api.device_count() == 0:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the B200 box")
    return api


@pytest.mark.parametrize("name", ["sah", "locb"])
def test_traversal_equals_golden_vectors(A, W, teapot, teapot_trees, name):
    g = np.load(GOLDEN)
    tris = teapot["tris"]
    bvh, m = teapot_trees[name]
    assert sha(bvh.nodes) == str(g[name + "_nodes_sha"]) and sha(m.nodes) == str(g[name + "_mnodes_sha"])  # the pinned trees
    rays = np.concatenate([W.camera_rays(W.benchmark_camera(256, 256)), W.random_rays(16384, *W.bounds(tris))])
    sub = rays[::5][: 16384 // 4 * 4]
    packets = W.pack4(sub)
    sc = A.Scene(tris, bvh=A.Bvh.from_arrays(bvh.nodes, bvh.indices), mbvh=A.Mbvh.from_arrays(m.nodes, m.indices))
    try:
        for kind, tag in ((A.TREE_BVH, "bvh"), (A.TREE_MBVH, "mbvh")):
            got = sc.intersect(rays, kind)
            want = g[f"{name}_{tag}_hits"]
            assert np.array_equal(got["prim"], want["prim"]) and np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
            assert np.array_equal(sc.occluded(sub, kind), g[f"{name}_{tag}_any"])
            got4 = sc.intersect_packets(packets, kind)
            want4 = g[f"{name}_{tag}_packet_hits"]
            assert np.array_equal(got4["prim"], want4["prim"]) and np.array_equal(got4["t"].view(np.uint32), want4["t"].view(np.uint32))
            assert np.array_equal(sc.occluded_packets(packets, kind), g[f"{name}_{tag}_packet_any"])
    finally:
        sc.free()


def test_gpu_builders_equal_golden_vectors(A, O, teapot):
    g = np.load(GOLDEN)
    # LOCB and its collapse are byte-identical to the pinned trees
    locb = A.Builder(teapot["aabbs"], teapot["centers"]).construct_locally_ordered_clustered()
    assert sha(locb.nodes) == str(g["locb_nodes_sha"]) and sha(locb.indices) == str(g["locb_indices_sha"])
    ml = A.Mbvh.construct(locb)
    assert sha(ml.nodes) == str(g["locb_mnodes_sha"])
    # binned SAH: isomorphic to the pinned tree (level-order numbering): same node count, SAH cost equal to the last digit
    b = A.Builder(teapot["aabbs"], teapot["centers"], 1).construct_binned_sah()
    assert len(b.nodes) == int(g["sah_node_count"])
    assert abs(O.Bvh(b.nodes.copy(), b.indices.copy()).sah_cost() - float(g["sah_sah"])) < 1e-9 * float(g["sah_sah"])
    mb = A.Mbvh.construct(b)
    assert len(mb.nodes) == int(g["sah_mnode_count"])
    for t in (ml, mb, locb, b):
        t.free()
