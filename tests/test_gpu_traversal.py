"""Parity of the CUDA traversal path against the CPU oracle, through the C ABI (include/rtbvh_gpu.h).

Bar (BASELINE.json north_star): traversing a reference-format tree uploaded unchanged returns bit-exact hit
primitive ids (lowest id on equal t) and t within 1e-5 relative.  The kernels reproduce the reference's
arithmetic, so these tests demand BIT-EXACT t as well; TOL_REL documents the contractual tolerance.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL_REL = 1e-5  # contractual; the assertions below are exact (stricter)


@pytest.fixture(scope="module")
def A():
    from rtbvh_b200 import api
    if api.device_count() == 0:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the B200 box")
    return api


def _scene(A, tris, bvh, m):
    return A.Scene(tris, bvh=A.Bvh.from_arrays(bvh.nodes, bvh.indices), mbvh=A.Mbvh.from_arrays(m.nodes, m.indices))


def _check_all_paths(A, O, W, tris, bvh, m, rays, label):
    sc = _scene(A, tris, bvh, m)
    packets = W.pack4(rays[: len(rays) // 4 * 4])
    try:
        for kind, otree in ((A.TREE_BVH, bvh), (A.TREE_MBVH, m)):
            want, _, _ = O.trace(otree, tris, rays)
            got = sc.intersect(rays, kind)
            assert np.array_equal(got["prim"], want["prim"]), f"{label}: ids differ (tree {kind})"
            assert np.array_equal(got["t"], want["t"]), f"{label}: t differs (tree {kind})"
            occ_w, _, _ = O.trace(otree, tris, rays, mode="any")
            assert np.array_equal(sc.occluded(rays, kind), occ_w), f"{label}: any-hit differs (tree {kind})"
            want4, _, _ = O.trace_packets(otree, tris, packets)
            got4 = sc.intersect_packets(packets, kind)
            assert np.array_equal(got4["prim"], want4["prim"]), f"{label}: packet ids differ (tree {kind})"
            assert np.array_equal(got4["t"], want4["t"]), f"{label}: packet t differs (tree {kind})"
            occ4_w, _, _ = O.trace_packets(otree, tris, packets, mode="any")
            assert np.array_equal(sc.occluded_packets(packets, kind), occ4_w), f"{label}: packet any-hit (tree {kind})"
        assert not sc.stack_overflowed()
        # optional ray sorting changes the order of work, never a result
        sc.set_ray_sorting(True)
        for kind, otree in ((A.TREE_BVH, bvh), (A.TREE_MBVH, m)):
            want, _, _ = O.trace(otree, tris, rays)
            assert np.array_equal(sc.intersect(rays, kind), want), f"{label}: sorted launch differs (tree {kind})"
            occ_w, _, _ = O.trace(otree, tris, rays, mode="any")
            assert np.array_equal(sc.occluded(rays, kind), occ_w), f"{label}: sorted any-hit differs (tree {kind})"
        sc.set_ray_sorting(False)
    finally:
        sc.free()


def test_ffi_kat_quad(A, O, W):
    # rtbvh_ffi/src/lib.rs:946-1019: quad at z = 1, ray from the origin along +Z, t = 1e26 in -> t == 1.0
    tris = W.quad()
    aabbs, _ = O.prims_from_triangles(tris, pad=1e-4)
    rc, bvh = O.build(O.BINNED_SAH, aabbs, O.aabb_centers(aabbs), 1)
    m = bvh.collapse()
    sc = _scene(A, tris, bvh, m)
    rays = W.make_rays(np.zeros((1, 3), np.float32), np.array([[0, 0, 1]], np.float32), t_max=np.float32(1e26))
    for kind in (A.TREE_BVH, A.TREE_MBVH):
        h = sc.intersect(rays, kind)
        assert abs(h["t"][0] - 1.0) < np.finfo(np.float32).eps and h["prim"][0] == 0
    sc.free()


@pytest.mark.parametrize("name", ["sah", "locb"])
def test_teapot_benchmark_camera(A, O, W, teapot, teapot_trees, name):
    bvh, m = teapot_trees[name]
    rays = W.camera_rays(W.benchmark_camera(400, 400))
    _check_all_paths(A, O, W, teapot["tris"], bvh, m, rays, f"teapot/{name}/camera")


@pytest.mark.parametrize("name", ["sah", "locb"])
def test_teapot_incoherent(A, O, W, teapot, teapot_trees, name):
    bvh, m = teapot_trees[name]
    rays = W.random_rays(100_000, *W.bounds(teapot["tris"]))
    _check_all_paths(A, O, W, teapot["tris"], bvh, m, rays, f"teapot/{name}/random")


def test_edge_rays(A, O, W, teapot, teapot_trees):
    """Axis-parallel directions (inv_direction = +-inf, NaN slabs -> the SSE operand rule), rays starting on
    box planes, NaN rays, zero-length windows, rays that start inside the model."""
    tris = teapot["tris"]
    lo, hi = W.bounds(tris)
    rng = np.random.default_rng(7)
    n = 4096
    o = (lo + (hi - lo) * rng.random((n, 3))).astype(np.float32)
    d = rng.standard_normal((n, 3)).astype(np.float32)
    axis = rng.integers(0, 3, n)
    d[np.arange(n) % 2 == 0] = 0.0
    d[np.arange(n), axis] = np.where(rng.random(n) < 0.5, -1.0, 1.0)
    # snap some origins exactly onto vertex coordinates (planes of leaf boxes up to the 1e-4 padding)
    v = tris.reshape(-1, 3)
    pick = rng.integers(0, len(v), n)
    snap = np.arange(n) % 3 == 0
    o[snap] = v[pick[snap]]
    rays = W.make_rays(o, d)
    rays["t"][::7] = np.float32(0.5)
    rays["t"][::11] = np.float32(1e-4)
    rays["origin"][5] = np.nan
    rays["direction"][9, 1] = np.nan
    rays["direction"][13] = 0.0
    bvh, m = teapot_trees["sah"]
    _check_all_paths(A, O, W, tris, bvh, m, rays, "teapot/edge")
    bvh, m = teapot_trees["locb"]
    _check_all_paths(A, O, W, tris, bvh, m, rays, "teapot/edge-locb")


def test_empty_and_ragged_batches(A, O, W, teapot, teapot_trees):
    bvh, m = teapot_trees["sah"]
    sc = _scene(A, teapot["tris"], bvh, m)
    assert len(sc.intersect(np.zeros(0, A.RAY_DTYPE))) == 0
    assert len(sc.intersect_packets(np.zeros(0, A.PACKET_DTYPE))) == 0
    rays = W.random_rays(1001, *W.bounds(teapot["tris"]))  # not a multiple of the block size
    want, _, _ = O.trace(m, teapot["tris"], rays)
    assert np.array_equal(sc.intersect(rays), want)
    pk = W.pack4(rays[:1000])[:33]  # 33 packets: a partially filled last warp
    want4, _, _ = O.trace_packets(m, teapot["tris"], pk)
    assert np.array_equal(sc.intersect_packets(pk), want4)
    sc.free()


def test_leaf_sizes_and_single_leaf_root(A, O, W, teapot):
    tris = teapot["tris"][:300]
    aabbs, centers = O.prims_from_triangles(tris)
    rays = W.random_rays(20_000, *W.bounds(tris))
    for leaf in (1, 2, 4, 8, 1000):  # 1000 -> the root is a single leaf holding every primitive
        rc, bvh = O.build(O.BINNED_SAH, aabbs, centers, leaf)
        assert rc == 0
        _check_all_paths(A, O, W, tris, bvh, bvh.collapse(), rays, f"leaf{leaf}")
    rc, bvh = O.build(O.LOCB, aabbs[:2], centers[:2], 1)  # LOCB with <= 2 prims: one root leaf (locb.rs:258-269)
    _check_all_paths(A, O, W, tris[:2], bvh, bvh.collapse(), rays[:4096], "locb2")


def test_soup_100k(A, O, W):
    tris = W.soup(100_000)
    aabbs, centers = O.prims_from_triangles(tris)
    rc, bvh = O.build(O.BINNED_SAH, aabbs, centers, 1)
    m = bvh.collapse()
    rays = np.concatenate([W.camera_rays(W.soup_camera(256, 256), jitter_seed=5, frame=3),
                           W.random_rays(50_000, *W.bounds(tris))])
    _check_all_paths(A, O, W, tris, bvh, m, rays, "soup100k")
    sh = W.shadow_rays(tris, 50_000)
    sc = _scene(A, tris, bvh, m)
    for kind, otree in ((A.TREE_BVH, bvh), (A.TREE_MBVH, m)):
        occ, _, _ = O.trace(otree, tris, sh, mode="any")
        assert np.array_equal(sc.occluded(sh, kind), occ)
    sc.free()


def test_device_resident_api_and_camera_rays(A, O, W, teapot, teapot_trees):
    import torch
    bvh, m = teapot_trees["sah"]
    tris = teapot["tris"]
    sc = _scene(A, tris, bvh, m)
    cam = W.benchmark_camera(512, 512)
    n = 512 * 512
    d_rays = torch.empty(n * 8, dtype=torch.float32, device="cuda")
    d_hits = torch.empty(n * 2, dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    A.generate_camera_rays_device(cam, 0, 512, d_rays, stream=stream)
    sc.intersect_device(d_rays, n, d_hits, A.TREE_MBVH, stream=stream)
    torch.cuda.synchronize()
    rays = d_rays.cpu().numpy().view(A.RAY_DTYPE).reshape(-1)
    ref_rays = W.camera_rays(cam)
    assert np.allclose(rays["direction"], ref_rays["direction"], rtol=0, atol=1e-6)
    hits = d_hits.cpu().numpy().view(A.HIT_DTYPE).reshape(-1)
    want, _, _ = O.trace(m, tris, rays)  # the oracle runs on the very rays the device generated
    assert np.array_equal(hits, want)
    assert not sc.stack_overflowed()
    sc.free()


def test_ray_tiling_and_camera_calls(A, O, W, teapot, teapot_trees):
    """rtbvh_gpu_scene_set_ray_tiling is a work-order hint: 8x8 pixel tiles per warp, every record unchanged and at its
    ray's index (whole 8-row bands are tiled, the tail keeps the linear order).  rtbvh_gpu_intersect_camera_async /
    _occluded_camera_async generate the frames on the device and must deliver exactly the records of
    generate_camera_rays_device + intersect_device, which in turn equal the oracle's on the very same rays."""
    import torch
    bvh, m = teapot_trees["sah"]
    tris = teapot["tris"]
    sc = _scene(A, tris, bvh, m)
    stream = torch.cuda.current_stream().cuda_stream
    try:
        for (w, h, frames) in ((400, 400, 2), (200, 52, 3), (8, 8, 1), (1000, 24, 1)):
            cam = W.benchmark_camera(w, h)
            n = w * h * frames
            d_rays = torch.empty(n * 8, dtype=torch.float32, device="cuda")
            for f in range(frames):
                A.generate_camera_rays_device(cam, 0, h, d_rays[f * w * h * 8:], jitter_seed=77, frame=5 + f, stream=stream)
            torch.cuda.synchronize()
            rays = d_rays.cpu().numpy().view(A.RAY_DTYPE).reshape(-1)
            for tree, otree in ((A.TREE_MBVH, m), (A.TREE_BVH, bvh)):
                want = O.trace(otree, tris, rays)[0]
                want_occ = O.trace(otree, tris, rays, mode="any")[0]
                d_hits = torch.zeros(n * 2, dtype=torch.float32, device="cuda")
                d_occ = torch.zeros(n, dtype=torch.uint8, device="cuda")
                sc.set_ray_tiling(w)
                sc.intersect_device(d_rays, n, d_hits, tree, stream=stream)
                sc.occluded_device(d_rays, n, d_occ, tree, stream=stream)
                torch.cuda.synchronize()
                sc.set_ray_tiling(0)
                assert np.array_equal(d_hits.cpu().numpy().view(A.HIT_DTYPE).reshape(-1), want), (w, h, frames, tree)
                assert np.array_equal(d_occ.cpu().numpy(), want_occ), (w, h, frames, tree)
                # a batch that is not a whole number of bands: the tail is traced in linear order
                sc.set_ray_tiling(w)
                d_hits.zero_()
                sc.intersect_device(d_rays, n - 3, d_hits, tree, stream=stream)
                torch.cuda.synchronize()
                sc.set_ray_tiling(0)
                assert np.array_equal(d_hits.cpu().numpy().view(A.HIT_DTYPE).reshape(-1)[: n - 3], want[: n - 3])
                # frames generated on the device inside the call, records straight to the host
                hh = torch.zeros(n * 2, dtype=torch.float32).pin_memory()
                ho = torch.zeros(n, dtype=torch.uint8).pin_memory()
                t1 = sc.intersect_camera_async(cam, frames, hh.data_ptr(), tree, jitter_seed=77, first_frame=5)
                t2 = sc.intersect_camera_async(cam, frames, ho.data_ptr(), tree, jitter_seed=77, first_frame=5, any_hit=True)
                sc.wait(t2)
                sc.wait(t1)
                assert np.array_equal(hh.numpy().view(A.HIT_DTYPE).reshape(-1), want), (w, h, frames, tree)
                assert np.array_equal(ho.numpy(), want_occ), (w, h, frames, tree)
        with pytest.raises(A.RtbvhError):
            sc.set_ray_tiling(12)  # not a multiple of 8
        assert not sc.stack_overflowed()
    finally:
        sc.free()


def test_full_size_properties_soup_1m(A, O, W):
    """BASELINE config 2 at full geometry size: oracle parity on a sample plus size-independent properties
    (any-hit == closest-hit predicate; a closest hit re-traced with t = t_hit*(1+1e-3) finds the same t)."""
    tris = W.soup(1 << 20)
    aabbs, centers = O.prims_from_triangles(tris)
    rc, bvh = O.build(O.BINNED_SAH, aabbs, centers, 1)
    m = bvh.collapse()
    sc = _scene(A, tris, bvh, m)
    rays = W.camera_rays(W.soup_camera(1000, 1000), jitter_seed=W.SEED_SOUP, frame=0)
    got = sc.intersect(rays, A.TREE_MBVH)
    sample = np.arange(0, len(rays), 16)
    want, _, _ = O.trace(m, tris, rays[sample])
    assert np.array_equal(got[sample], want)
    occ = sc.occluded(rays, A.TREE_MBVH)
    assert np.array_equal(occ.astype(bool), got["prim"] != A.NO_HIT)
    hit = got["prim"] != A.NO_HIT
    again = rays[hit].copy()
    again["t"] = got["t"][hit] * np.float32(1.001)
    got2 = sc.intersect(again, A.TREE_MBVH)
    assert np.array_equal(got2["t"], got["t"][hit]) and np.array_equal(got2["prim"], got["prim"][hit])
    assert not sc.stack_overflowed()
    sc.free()


def _chain_tree(O, levels):
    """A hand-made, maximally unbalanced reference-format tree: inner node k has a leaf child and the next inner node.
    Every box contains the whole scene, so a ray that hits anything keeps one pending entry per level on the stack."""
    tris = np.zeros((levels + 1, 3, 3), np.float32)
    for k in range(levels + 1):
        z = np.float32(1.0 + k)
        tris[k] = [[-1, -1, z], [1, -1, z], [0, 1, z]]
    nodes = np.zeros(2 * levels + 1, O.NODE_DTYPE)
    nodes["min"] = [-2, -2, 0]
    nodes["max"] = [2, 2, levels + 3]
    for k in range(levels):
        inner, leaf, nxt = (0 if k == 0 else 2 * k), 2 * k + 1, 2 * k + 2
        nodes[inner]["count"], nodes[inner]["left_first"] = -1, leaf
        nodes[leaf]["count"], nodes[leaf]["left_first"] = 1, k
    nodes[2 * levels]["count"], nodes[2 * levels]["left_first"] = 1, levels
    # make the leaf the NEAR child for +z rays so the far (inner) child is pushed first ... and popped last
    for k in range(levels):
        nodes[2 * k + 1]["max"] = [2, 2, 1.5 + k]
    return tris, O.Bvh(nodes, np.arange(levels + 1, dtype=np.uint32))


def test_deep_stack_spills_and_overflow_is_reported(A, O, W):
    """The reference's stack has 32 entries (panic / UB beyond, quirk Q10).  Here entries 32..127 spill to local memory
    and deeper rays raise an error instead of corrupting memory."""
    tris, bvh = _chain_tree(O, 100)
    m = bvh.collapse()
    rng = np.random.default_rng(1)
    o = np.stack([rng.uniform(-0.5, 0.5, 2000), rng.uniform(-0.5, 0.5, 2000), np.full(2000, -1.0)], axis=1).astype(np.float32)
    d = np.tile(np.array([[0, 0, 1]], np.float32), (2000, 1)) + rng.normal(0, 0.01, (2000, 3)).astype(np.float32)
    rays = W.make_rays(o, d)
    _, _, c2 = O.trace(bvh, tris, rays, counters=True)
    assert c2["max_stack"] > 32  # the oracle itself needs more than the reference's 32 entries here
    _check_all_paths(A, O, W, tris, bvh, m, rays, "chain100")
    tris, bvh = _chain_tree(O, 300)
    sc = A.Scene(tris, bvh=A.Bvh.from_arrays(bvh.nodes, bvh.indices))
    with pytest.raises(A.RtbvhError):
        sc.intersect(rays, A.TREE_BVH)
    sc.free()


@pytest.mark.parametrize("fix", [True, False])
def test_reference_built_spatial_tree_uploaded_unchanged(A, O, W, fix):
    """Config 5 path: a spatial-split SAH tree built on the CPU (oracle restatement of spatial_sah.rs) is uploaded
    unchanged — prim_indices of length N + 0.75 N with an unused tail, scheduling-order node numbering — and traversed.
    The verbatim tree (fix=False) drops primitives on this scene; parity is against the tree as built either way."""
    tris = W.soup(20_000, seed=W.SEED_SOUP + 5, aniso=(8, 1, 1))
    rc, bvh = O.build_spatial(tris, 1, fix_child_ranges=fix)
    assert rc == 0
    m = bvh.collapse()
    rays = np.concatenate([W.camera_rays(W.soup_camera(200, 200)), W.random_rays(40_000, *W.bounds(tris))])
    _check_all_paths(A, O, W, tris, bvh, m, rays, f"sbvh/fix={fix}")
    # GPU collapse of the (non level-ordered) binary tree equals merge_nodes on the CPU
    gb = A.Bvh.from_arrays(bvh.nodes, bvh.indices)
    gm = A.Mbvh.construct(gb)
    assert gm.nodes.tobytes() == m.nodes.tobytes()
    gm.free()
    if fix:
        bf = O.brute_force(tris, rays)
        sc = A.Scene(tris, bvh=gb)
        assert np.array_equal(sc.intersect(rays, A.TREE_BVH), bf)
        sc.free()


def test_async_submit_wait(A, O, W, teapot, teapot_trees):
    """rtbvh_gpu_intersect_async / rtbvh_gpu_occluded_async / rtbvh_gpu_wait: several batches in flight through the
    shared staging pipeline, pinned and pageable buffers, waits in and out of order — every batch equals the oracle."""
    import torch
    tris = teapot["tris"]
    bvh, m = teapot_trees["sah"]
    sc = _scene(A, tris, bvh, m)
    try:
        batches, want, want_occ = [], [], []
        for k, n in enumerate((300_001, 70_000, 1, 524_288)):
            rays = W.random_rays(n, *W.bounds(tris), seed=0xA51C + k)
            batches.append(rays)
            want.append(O.trace(m, tris, rays)[0])
            want_occ.append(O.trace(m, tris, rays, mode="any")[0])
        h_rays = [torch.from_numpy(r.view(np.float32).reshape(-1).copy()).pin_memory() for r in batches]
        h_hits = [torch.zeros(len(r) * 2, dtype=torch.float32).pin_memory() for r in batches]
        h_occ = [torch.zeros(len(r), dtype=torch.uint8).pin_memory() for r in batches]
        tickets = [sc.intersect_async(h_rays[k].data_ptr(), len(batches[k]), h_hits[k].data_ptr(), A.TREE_MBVH)
                   for k in range(len(batches))]
        tickets_o = [sc.occluded_async(h_rays[k].data_ptr(), len(batches[k]), h_occ[k].data_ptr(), A.TREE_MBVH)
                     for k in range(len(batches))]
        assert tickets == sorted(tickets) and len(set(tickets + tickets_o)) == 2 * len(batches)
        for k in (2, 0, 3, 1):  # batches complete in submission order; waiting out of order is allowed
            sc.wait(tickets[k])
            got = h_hits[k].numpy().view(A.HIT_DTYPE).reshape(-1)
            assert np.array_equal(got, want[k]), f"async batch {k} differs from the oracle"
        sc.wait(0)
        for k in range(len(batches)):
            assert np.array_equal(h_occ[k].numpy(), want_occ[k]), f"async any-hit batch {k} differs"
        sc.wait(tickets[0])  # waiting twice is fine
        with pytest.raises(A.RtbvhError):
            sc.wait(10_000)  # never issued
        # pageable buffers: submit blocks inside the copies, results are the same
        out = np.zeros(len(batches[1]), dtype=A.HIT_DTYPE)
        t = sc.intersect_async(batches[1].ctypes.data, len(batches[1]), out.ctypes.data, A.TREE_MBVH)
        sc.wait(t)
        assert np.array_equal(out, want[1])
        # more submissions than ticket slots without a single wait
        small = W.random_rays(1000, *W.bounds(tris), seed=7)
        hs = torch.from_numpy(small.view(np.float32).reshape(-1).copy()).pin_memory()
        outs = [torch.zeros(2000, dtype=torch.float32).pin_memory() for _ in range(100)]
        last = [sc.intersect_async(hs.data_ptr(), 1000, o.data_ptr(), A.TREE_MBVH) for o in outs]
        sc.wait(last[0])  # recycled slot: already complete
        sc.wait(0)
        w_small = O.trace(m, tris, small)[0]
        assert all(np.array_equal(o.numpy().view(A.HIT_DTYPE).reshape(-1), w_small) for o in outs)
    finally:
        sc.free()


@pytest.mark.parametrize("mode", ["gated", "staged"])
def test_host_pipeline_flavours_in_a_subprocess(A, mode):
    """RTBVH_HOST_MODE=gated (launches that start before their input has arrived; the default wherever allowed) and =staged
    (chunks over four streams; what packets and sorted batches always use).  The variable is read once per process, hence
    the subprocess: blocking, asynchronous, RTRay and split-input calls all return the oracle's records."""
    import subprocess
    import sys
    import os
    code = (
        "import numpy as np, sys; sys.path.insert(0, '.')\n"
        "from rtbvh_b200 import api, workloads as W\n"
        "from oracle import oracle as O\n"
        "tris = W.soup(50_000)\n"
        "aabbs, centers = O.prims_from_triangles(tris)\n"
        "rc, bvh = O.build(O.BINNED_SAH, aabbs, centers, 1); m = bvh.collapse()\n"
        "sc = api.Scene(tris, bvh=None, mbvh=api.Mbvh.from_arrays(m.nodes, m.indices))\n"
        "for n in (1, 100_000, 700_001):\n"
        "    rays = W.random_rays(n, *W.bounds(tris), seed=n)\n"
        "    want = O.trace(m, tris, rays, threads=8)[0]\n"
        "    assert np.array_equal(sc.intersect(rays, api.TREE_MBVH), want), n\n"
        "    assert np.array_equal(sc.occluded(rays, api.TREE_MBVH), O.trace(m, tris, rays, mode='any', threads=8)[0]), n\n"
        "    o = np.ascontiguousarray(rays['origin']); d = np.ascontiguousarray(rays['direction'])\n"
        "    assert np.array_equal(sc.intersect_od(o, d, api.TREE_MBVH), want), n\n"
        "import torch\n"
        "rays = [W.random_rays(3_000_000, *W.bounds(tris), seed=k) for k in range(3)]\n"
        "hr = [torch.from_numpy(r.view(np.float32).reshape(-1).copy()).pin_memory() for r in rays]\n"
        "ho = [torch.from_numpy(np.ascontiguousarray(r['origin']).reshape(-1)).pin_memory() for r in rays]\n"
        "hd = [torch.from_numpy(np.ascontiguousarray(r['direction']).reshape(-1)).pin_memory() for r in rays]\n"
        "hh = [torch.zeros(6_000_000, dtype=torch.float32).pin_memory() for _ in range(6)]\n"
        "tk = [sc.intersect_async(hr[k].data_ptr(), 3_000_000, hh[k].data_ptr(), api.TREE_MBVH) for k in range(3)]\n"
        "tk += [sc.intersect_od_async(ho[k].data_ptr(), hd[k].data_ptr(), 3_000_000, hh[3 + k].data_ptr(), api.TREE_MBVH) for k in range(3)]\n"
        "sc.wait(0)\n"
        "for k in range(3):\n"
        "    want = O.trace(m, tris, rays[k], threads=8)[0]\n"
        "    assert np.array_equal(hh[k].numpy().view(api.HIT_DTYPE).reshape(-1), want), k\n"
        "    assert np.array_equal(hh[3 + k].numpy().view(api.HIT_DTYPE).reshape(-1), want), k\n"
        "print('gated ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RTBVH_HOST_MODE=mode)
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "gated ok" in r.stdout, r.stdout + r.stderr


def test_host_buffer_calls_return_under_synchronous_launches(A):
    """CUDA_LAUNCH_BLOCKING=1 makes every kernel launch synchronous, as ncu and compute-sanitizer do.  The gated host
    pipeline starts a launch before its rays have arrived; all its uploads and watermark writes must therefore be queued
    BEFORE the launch call (round-1 defect: launch first, feed afterwards -> the launch never returned).  Runs smoke() and a
    3 M-ray blocking + asynchronous call from pinned buffers in a subprocess with a hard timeout: a call that always
    returns, like rtbvh_ffi/src/lib.rs:700-731."""
    import subprocess
    import sys
    import os
    code = (
        "import numpy as np, sys; sys.path.insert(0, '.')\n"
        "import __graft_entry__ as G\n"
        "G.smoke()\n"
        "import torch\n"
        "from rtbvh_b200 import api, workloads as W\n"
        "from oracle import oracle as O\n"
        "tris = W.soup(50_000)\n"
        "aabbs, centers = O.prims_from_triangles(tris)\n"
        "rc, bvh = O.build(O.BINNED_SAH, aabbs, centers, 1); m = bvh.collapse()\n"
        "sc = api.Scene(tris, bvh=None, mbvh=api.Mbvh.from_arrays(m.nodes, m.indices))\n"
        "rays = W.random_rays(3_000_000, *W.bounds(tris), seed=5)\n"
        "want = O.trace(m, tris, rays, threads=8)[0]\n"
        "hr = torch.from_numpy(rays.view(np.float32).reshape(-1).copy()).pin_memory()\n"
        "hh = torch.zeros(6_000_000, dtype=torch.float32).pin_memory()\n"
        "sc.intersect_ptr(hr.data_ptr(), 3_000_000, hh.data_ptr(), api.TREE_MBVH)\n"
        "assert np.array_equal(hh.numpy().view(api.HIT_DTYPE).reshape(-1), want)\n"
        "hh.zero_()\n"
        "t = sc.intersect_async(hr.data_ptr(), 3_000_000, hh.data_ptr(), api.TREE_MBVH); sc.wait(t)\n"
        "assert np.array_equal(hh.numpy().view(api.HIT_DTYPE).reshape(-1), want)\n"
        "assert np.array_equal(sc.intersect(rays, api.TREE_MBVH), want)  # pageable input\n"
        "print('blocking ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for extra in ({}, {"RTBVH_HOST_MODE": "gated"}):
        env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1", **extra)
        r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "blocking ok" in r.stdout, r.stdout + r.stderr


def test_split_origin_direction_input(A, O, W, teapot, teapot_trees):
    """rtbvh_gpu_intersect_od / _occluded_od / _od_async / _od_device: origins and directions as packed float3 arrays with a
    common t_min / t_max (the reference FFI's argument shape) give the records of the RTRay calls, bit for bit."""
    import torch
    tris = teapot["tris"]
    bvh, m = teapot_trees["sah"]
    sc = _scene(A, tris, bvh, m)
    try:
        for n, t_max in ((1, 1e34), (100_003, 1e34), (300_000, 9.5)):
            rays = W.random_rays(n, *W.bounds(tris), seed=0x0D + n)
            rays["t"] = np.float32(t_max)
            o = np.ascontiguousarray(rays["origin"]); d = np.ascontiguousarray(rays["direction"])
            for tree, otree in ((A.TREE_BVH, bvh), (A.TREE_MBVH, m)):
                want = O.trace(otree, tris, rays)[0]
                assert np.array_equal(sc.intersect_od(o, d, tree, 1e-4, t_max), want), (n, tree)
                assert np.array_equal(sc.intersect_od(o, d, tree, 1e-4, t_max, any_hit=True), O.trace(otree, tris, rays, mode="any")[0])
            # asynchronous, pinned
            ho, hd = torch.from_numpy(o.reshape(-1).copy()).pin_memory(), torch.from_numpy(d.reshape(-1).copy()).pin_memory()
            hh = torch.zeros(n * 2, dtype=torch.float32).pin_memory()
            t = sc.intersect_od_async(ho.data_ptr(), hd.data_ptr(), n, hh.data_ptr(), A.TREE_MBVH, 1e-4, t_max)
            sc.wait(t)
            assert np.array_equal(hh.numpy().view(A.HIT_DTYPE).reshape(-1), O.trace(m, tris, rays)[0])
            # device resident
            do, dd = ho.cuda(), hd.cuda()
            dh = torch.zeros(n * 2, dtype=torch.float32, device="cuda")
            sc.intersect_od_device(do, dd, n, dh, A.TREE_MBVH, 1e-4, t_max, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert np.array_equal(dh.cpu().numpy().view(A.HIT_DTYPE).reshape(-1), O.trace(m, tris, rays)[0])
        assert len(sc.intersect_od(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))) == 0
    finally:
        sc.free()


@pytest.mark.parametrize("tree_name", ["bvh", "mbvh"])
def test_packet_calls_with_oddly_aligned_device_buffers(A, O, W, teapot, teapot_trees, tree_name):
    """RTRayPacket4 / RTHitPacket4 only promise 4-byte alignment.  The one-lane-per-packet kernels use 16-byte accesses and
    must hand oddly placed buffers to the scalar quad kernel: same results either way (closest and any hit)."""
    import torch
    tris = teapot["tris"]
    bvh, m = teapot_trees["sah"]
    sc = A.Scene(tris, bvh=A.Bvh.from_arrays(bvh.nodes, bvh.indices), mbvh=A.Mbvh.from_arrays(m.nodes, m.indices))
    try:
        rays = np.concatenate([W.camera_rays(W.benchmark_camera(96, 96)), W.random_rays(10_000, *W.bounds(tris), seed=77)])
        packets = W.pack4(rays[: len(rays) // 4 * 4])
        n = len(packets)
        tree, otree = (A.TREE_BVH, bvh) if tree_name == "bvh" else (A.TREE_MBVH, m)
        want = O.trace_packets(otree, tris, packets)[0]
        want_any = O.trace_packets(otree, tris, packets, mode="any")[0]
        flat = torch.from_numpy(packets.view(np.float32).reshape(-1).copy())
        stream = torch.cuda.current_stream().cuda_stream
        for shift in (0, 1):  # 0: 16-byte aligned (lane kernels); 1: shifted by one float (quad kernels)
            d_in = torch.empty(flat.numel() + 4, dtype=torch.float32, device="cuda")
            d_in[shift: shift + flat.numel()] = flat.cuda()
            d_out = torch.zeros(n * 8 + 4, dtype=torch.float32, device="cuda")
            d_occ = torch.zeros(n * 4 + 8, dtype=torch.uint8, device="cuda")
            sc.intersect_packets_device(d_in[shift:], n, d_out[shift:], tree, stream=stream)
            sc.occluded_packets_device(d_in[shift:], n, d_occ[shift:], tree, stream=stream)
            torch.cuda.synchronize()
            got = d_out[shift: shift + n * 8].cpu().numpy().view(A.HIT4_DTYPE).reshape(-1)
            assert np.array_equal(got, want), f"shift {shift}"
            assert np.array_equal(d_occ[shift: shift + n * 4].cpu().numpy().reshape(n, 4), want_any), f"shift {shift}"
        assert not sc.stack_overflowed()
    finally:
        sc.free()


@pytest.mark.parametrize("leaf,kind", [(1, "sah"), (4, "sah"), (1, "locb")])
def test_duplicated_triangles_report_the_lowest_id(A, O, W, leaf, kind):
    """Exactly equal t from duplicated triangles: every flavour (single / packet, closest / any, both trees, sorted launches)
    reports the lowest id like the oracle, whatever order the leaves are reached in; multi-primitive leaves included."""
    tris = W.soup(6_000, seed=0x71E5).astype(np.float32)
    tris[4_000:6_000] = tris[0:2_000]  # 2 000 exact duplicates with higher ids
    aabbs, centers = O.prims_from_triangles(tris)
    rc, bvh = O.build(O.BINNED_SAH if kind == "sah" else O.LOCB, aabbs, centers, leaf)
    assert rc == 0
    m = bvh.collapse()
    rays = np.concatenate([W.camera_rays(W.soup_camera(160, 160)), W.random_rays(30_000, *W.bounds(tris), seed=9)])
    want = O.trace(m, tris, rays)[0]
    hit = want["prim"] != A.NO_HIT
    assert hit.mean() > 0.2 and (want["prim"][hit] < 4_000).all()
    _check_all_paths(A, O, W, tris, bvh, m, rays, f"duplicates leaf={leaf} {kind}")
