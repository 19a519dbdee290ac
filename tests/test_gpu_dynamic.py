"""Dynamic scenes: rtbvh_gpu_scene_refit / _refit_device keep a device-resident Bvh + Mbvh + triangle
records in step with moving vertices.  Oracle: Bvh::refit (src/bvh.rs:176-205) on the CPU followed by Mbvh::construct
(src/bvh.rs:381-404) of the refitted tree — node arrays byte-identical, every traversal flavour bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    from rtbvh_b200 import api
    if api.device_count() == 0:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the B200 box")
    return api


def _wobble(tris, frame):
    f = np.float32
    c = tris.mean(axis=1, keepdims=True)
    phase = (c[..., 0:1] * f(7.0) + c[..., 1:2] * f(5.0) + f(frame) * f(0.9)).astype(f)
    d = np.concatenate([np.sin(phase), np.cos(phase * f(1.3)), np.sin(phase * f(0.7) + f(1.0))], axis=-1).astype(f)
    return (tris + f(0.03) * d).astype(f)


@pytest.mark.parametrize("kind", ["sah", "locb"])
def test_scene_refit_matches_cpu_refit_and_recollapse(A, O, W, kind):
    tris0 = W.soup(30_000, seed=W.SEED_SOUP + 11)
    aabbs, centers = O.prims_from_triangles(tris0)
    rc, bvh = O.build(O.BINNED_SAH if kind == "sah" else O.LOCB, aabbs, centers, 1 if kind == "locb" else 3)
    assert rc == 0
    m = bvh.collapse()
    gb, gm = A.Bvh.from_arrays(bvh.nodes, bvh.indices), A.Mbvh.from_arrays(m.nodes, m.indices)
    sc = A.Scene(tris0, bvh=gb, mbvh=gm)
    try:
        cur = bvh
        for frame in range(1, 4):
            tris = _wobble(tris0, frame)
            new_aabbs, _ = O.prims_from_triangles(tris)
            cur = cur.refit(new_aabbs)            # CPU: refit the refitted tree again, like an animation loop does
            cm = cur.collapse()
            if frame == 2:                        # device-vertex flavour on a stream
                import torch
                d = torch.from_numpy(tris.reshape(-1).copy()).cuda()
                sc.refit_device(d, len(tris), 12, torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
            else:
                sc.refit(tris)
            assert sc.read_nodes(A.TREE_BVH).tobytes() == cur.nodes.tobytes(), f"frame {frame}: refitted Bvh differs"
            assert sc.read_nodes(A.TREE_MBVH).tobytes() == cm.nodes.tobytes(), f"frame {frame}: refreshed Mbvh differs"
            rays = np.concatenate([W.camera_rays(W.soup_camera(150, 150)), W.random_rays(20_000, *W.bounds(tris), seed=frame)])
            for tree, otree in ((A.TREE_BVH, cur), (A.TREE_MBVH, cm)):
                assert np.array_equal(sc.intersect(rays, tree), O.trace(otree, tris, rays)[0]), f"frame {frame} tree {tree}"
                assert np.array_equal(sc.occluded(rays, tree), O.trace(otree, tris, rays, mode="any")[0])
            packets = W.pack4(rays[: len(rays) // 4 * 4])
            assert np.array_equal(sc.intersect_packets(packets, A.TREE_MBVH), O.trace_packets(cm, tris, packets)[0])
        # the refitted tree still finds what brute force finds (boxes stay conservative under refit on a tree without
        # the Q3 fallback boxes: LOCB)
        if kind == "locb":
            assert np.array_equal(sc.intersect(rays, A.TREE_MBVH), O.brute_force(tris, rays))
    finally:
        sc.free()


def test_scene_refit_rejects_what_it_cannot_do(A, O, W):
    tris = W.soup(2_000, seed=5)
    aabbs, centers = O.prims_from_triangles(tris)
    rc, bvh = O.build(O.BINNED_SAH, aabbs, centers, 1)
    m = bvh.collapse()
    gm = A.Mbvh.from_arrays(m.nodes, m.indices)
    only_m = A.Scene(tris, bvh=None, mbvh=gm)
    with pytest.raises(A.RtbvhError):
        only_m.refit(tris)                      # no binary tree in the scene
    only_m.free()
    sc = A.Scene(tris, bvh=A.Bvh.from_arrays(bvh.nodes, bvh.indices), mbvh=gm)
    with pytest.raises(A.RtbvhError):
        sc.refit(tris[:-1])                     # triangle count changed
    sc.refit(tris)                              # same positions: boxes may only change by the refit pads
    sc.free()
    # Bvh only (no Mbvh in the scene)
    sb = A.Scene(tris, bvh=A.Bvh.from_arrays(bvh.nodes, bvh.indices))
    moved = _wobble(tris, 1)
    sb.refit(moved)
    want = bvh.refit(O.prims_from_triangles(moved)[0])
    assert sb.read_nodes(A.TREE_BVH).tobytes() == want.nodes.tobytes()
    rays = W.random_rays(5_000, *W.bounds(moved))
    assert np.array_equal(sb.intersect(rays, A.TREE_BVH), O.trace(want, moved, rays)[0])
    sb.free()


@pytest.mark.parametrize("kind", ["sah", "locb"])
def test_scene_build_resident_equals_host_mirrored_build(A, O, W, kind):
    """rtbvh_gpu_scene_build: the trees never leave the device, yet they are the trees create_bvh + create_mbvh produce,
    and tracing them gives what the oracle gets on the oracle-built tree."""
    import torch
    tris = W.soup(40_000, seed=W.SEED_SOUP + 21)
    btype = A.BINNED_SAH if kind == "sah" else A.LOCALLY_ORDERED_CLUSTERED
    host_bvh = A.build_triangles(tris, btype, 2)
    host_m = A.Mbvh.construct(host_bvh)
    sc = A.Scene.build(tris, btype, 2, mbvh=True)
    d_verts = torch.from_numpy(tris.reshape(-1).copy()).cuda()
    sd = A.Scene.build(d_verts, btype, 2, mbvh=True, n_tris=len(tris))
    try:
        for s in (sc, sd):
            assert s.n_nodes == host_bvh.rt.node_count and s.n_mnodes == host_m.rt.node_count
            assert s.read_nodes(A.TREE_BVH).tobytes() == host_bvh.nodes.tobytes()
            assert s.read_nodes(A.TREE_MBVH).tobytes() == host_m.nodes.tobytes()
            assert np.array_equal(s.read_indices(A.TREE_BVH), host_bvh.indices)
        aabbs, centers = O.prims_from_triangles(tris)
        rc, obvh = O.build(O.BINNED_SAH if kind == "sah" else O.LOCB, aabbs, centers, 2)
        om = obvh.collapse()
        rays = np.concatenate([W.camera_rays(W.soup_camera(200, 200)), W.random_rays(30_000, *W.bounds(tris))])
        for tree, otree in ((A.TREE_BVH, obvh), (A.TREE_MBVH, om)):
            assert np.array_equal(sc.intersect(rays, tree), O.trace(otree, tris, rays)[0])
            assert np.array_equal(sd.occluded(rays, tree), O.trace(otree, tris, rays, mode="any")[0])
        # a resident scene refits like any other
        moved = _wobble(tris, 2)
        sc.refit(moved)
        want = obvh.refit(O.prims_from_triangles(moved)[0])
        assert np.array_equal(sc.intersect(rays, A.TREE_MBVH), O.trace(want.collapse(), moved, rays)[0])
    finally:
        sc.free(); sd.free(); host_m.free(); host_bvh.free()
    with pytest.raises(A.RtbvhError):
        A.Scene.build(np.zeros((0, 3, 3), np.float32))
