"""Spatial-split SAH builder restated in the oracle (reference src/builders/spatial_sah.rs).  CPU only.
The north star never builds such a tree on the GPU: a reference-built one is uploaded unchanged (config 5)."""
import numpy as np
import pytest


# src/builders/spatial_sah.rs:1043-1093  no_primitives + test_spatial_sah_build on the teapot
@pytest.mark.parametrize("fix", [True, False])
def test_teapot_spatial_build(O, W, teapot, fix):
    tris = teapot["tris"]
    n = len(tris)
    rc, bvh = O.build_spatial(tris, 1, fix_child_ranges=fix)
    assert rc == 0
    assert n <= len(bvh.nodes) <= 2 * n
    root = bvh.nodes[0]
    assert np.all(root["min"] <= root["max"])
    assert bvh.validate(n)
    v = tris.reshape(-1, 3)
    assert np.all(v > root["min"]) and np.all(v < root["max"])
    assert len(bvh.indices) == n + int(n * 0.75)  # quirk Q9: N + floor(0.75 N) entries, unused tail is zero
    assert np.all(bvh.indices[bvh.stats[2]:] == 0)
    assert O.build_spatial(np.zeros((0, 3, 3), np.float32))[0] == 2  # NoPrimitives


# src/lib.rs:126-154 test_spatial (two-triangle quad)
def test_spatial_quad(O):
    v = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], dtype=np.float32)
    rc, bvh = O.build_spatial(np.stack([v[[0, 1, 2]], v[[0, 2, 3]]]))
    assert rc == 0 and bvh.validate(2)


def test_verbatim_child_ranges_can_drop_primitives(O):
    """Documents the reference defect in allocate_children (spatial_sah.rs:375-423): when the references of the right
    child are moved up by fewer slots than the child holds, the verbatim ranges read stale data — primitives vanish.
    The default (fix_child_ranges=True) describes the children where the references actually are."""
    t8 = np.array([[[i, 0, 0], [i + 0.5, 0, 0], [i, 0.5, 0.1]] for i in range(8)], np.float32)
    rc, fixed = O.build_spatial(t8, 1, True)
    rc, verbatim = O.build_spatial(t8, 1, False)
    assert fixed.validate(8) and sorted(fixed.indices[:8].tolist()) == list(range(8))
    assert not verbatim.validate(8)


def test_spatial_tree_traversal_equals_brute_force(O, W):
    tris = W.soup(5000, seed=77, aniso=(8, 1, 1))
    rays = np.concatenate([W.camera_rays(W.soup_camera(64, 64)), W.random_rays(4096, *W.bounds(tris))])
    rc, bvh = O.build_spatial(tris, 1, True)
    assert bvh.validate(len(tris))
    bf = O.brute_force(tris, rays)
    for tree in (bvh, bvh.collapse()):
        hits, _, _ = O.trace(tree, tris, rays)
        assert np.array_equal(hits["t"], bf["t"]) and np.array_equal(hits["prim"], bf["prim"])
    a, c = O.prims_from_triangles(tris)
    rc, binned = O.build(O.BINNED_SAH, a, c, 1)
    assert bvh.sah_cost() < binned.sah_cost() * 1.02  # full-sweep SAH is at least as good as 16-bin SAH
