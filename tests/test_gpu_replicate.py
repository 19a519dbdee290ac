"""Scene replication through the C ABI (rtbvh_gpu_scene_export / _import / _clone, include/rtbvh_gpu.h; SURVEY.md section 8e:
the tree is built once and replicated, rays are sharded).  The replica must be byte-identical on the device (nodes of both
trees, prim_indices), trace bit-exactly like the oracle, survive the original being freed, and be refittable."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    from rtbvh_b200 import api
    if api.device_count() == 0:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the B200 box")
    return api


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_clone_is_byte_identical_and_independent(A, O, W, teapot, teapot_trees):
    tris = teapot["tris"]
    bvh, m = teapot_trees["sah"]
    src = A.Scene(tris, bvh=A.Bvh.from_arrays(bvh.nodes, bvh.indices), mbvh=A.Mbvh.from_arrays(m.nodes, m.indices))
    dev = A.device_count() - 1  # another GPU when the box has one, else the same device (same code path: cudaMemcpyPeer)
    rep = src.clone(dev)
    for tree in (A.TREE_BVH, A.TREE_MBVH):
        assert rep.read_nodes(tree).tobytes() == src.read_nodes(tree).tobytes()
        assert np.array_equal(rep.read_indices(tree), src.read_indices(tree))
    src.free()  # the replica owns its memory
    A.set_device(dev)
    try:
        rays = W.random_rays(60_000, *W.bounds(tris), seed=0xC10E)
        want, _, _ = O.trace(m, tris, rays)
        assert np.array_equal(rep.intersect(rays, A.TREE_MBVH), want)
        want2, _, _ = O.trace(bvh, tris, rays)
        assert np.array_equal(rep.intersect(rays, A.TREE_BVH), want2)
        # a replica is a full scene: refit it (same vertices -> same boxes as a CPU refit of the tree)
        rep.refit(tris)
        cur = bvh.refit(teapot["aabbs"])
        assert rep.read_nodes(A.TREE_BVH).tobytes() == cur.nodes.tobytes()
        assert rep.read_nodes(A.TREE_MBVH).tobytes() == cur.collapse().nodes.tobytes()
        assert np.array_equal(rep.intersect(rays, A.TREE_MBVH), O.trace(cur.collapse(), tris, rays)[0])
        assert not rep.stack_overflowed()
    finally:
        rep.free()
        A.set_device(0)


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rtbvh_b200 import api, workloads as W
    dev = rank % api.device_count()
    api.set_device(dev)
    tris = W.teapot()
    blob = [None]
    scene = None
    if rank == 0:
        scene = api.Scene.build(tris, api.BINNED_SAH, 1, mbvh=True)  # built ONCE, on rank 0's GPU
        blob[0] = scene.export_bytes()
    dist.broadcast_object_list(blob, src=0)
    if rank != 0:
        scene = api.Scene.import_bytes(blob[0])  # device-to-device copy out of rank 0's allocations
    dist.barrier()  # the exporter keeps its scene alive until every importer is done
    rays = W.random_rays(40_000, *W.bounds(tris), seed=0x1290 + rank)
    np.save(os.path.join(out_dir, f"hits_{rank}.npy"), scene.intersect(rays, api.TREE_MBVH))
    np.save(os.path.join(out_dir, f"occ_{rank}.npy"), scene.occluded(rays, api.TREE_BVH))
    np.save(os.path.join(out_dir, f"mnodes_{rank}.npy"), scene.read_nodes(api.TREE_MBVH).view(np.uint8))
    np.save(os.path.join(out_dir, f"nodes_{rank}.npy"), scene.read_nodes(api.TREE_BVH).view(np.uint8))
    scene.free()
    dist.destroy_process_group()


def test_export_import_between_two_processes(O, W, teapot, teapot_trees, tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    tris = teapot["tris"]
    bvh, m = teapot_trees["sah"]
    assert np.load(tmp_path / "mnodes_1.npy").tobytes() == np.load(tmp_path / "mnodes_0.npy").tobytes()
    assert np.load(tmp_path / "nodes_1.npy").tobytes() == np.load(tmp_path / "nodes_0.npy").tobytes()
    for r in range(world):
        rays = W.random_rays(40_000, *W.bounds(tris), seed=0x1290 + r)
        want, _, _ = O.trace(m, tris, rays)
        assert np.array_equal(np.load(tmp_path / f"hits_{r}.npy"), want), f"rank {r}"
        occ, _, _ = O.trace(bvh, tris, rays, mode="any")
        assert np.array_equal(np.load(tmp_path / f"occ_{r}.npy"), occ), f"rank {r}"


def test_import_rejects_garbage(A):
    with pytest.raises(A.RtbvhError):
        A.Scene.import_bytes(b"\x00" * 512)
