"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous ray sharding, tree broadcast, ragged
hit gather.  The per-rank "device" work is done by the CPU oracle here; on the box the same plumbing drives the
CUDA path (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rays, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle as O
    from rtbvh_b200 import multigpu as MG, workloads as W
    tris = W.teapot()
    tree = None
    if rank == 0:  # only rank 0 builds; the others receive the replica
        aabbs, centers = O.prims_from_triangles(tris)
        rc, bvh = O.build(O.BINNED_SAH, aabbs, centers, 1)
        m = bvh.collapse()
        tree = {"nodes": m.nodes, "indices": m.indices}
    tree = MG.broadcast_arrays(tree, src=0)
    assert tree["nodes"].dtype == O.MNODE_DTYPE
    rays = W.random_rays(n_rays, *W.bounds(tris))           # every rank can regenerate any slice (counter based)
    lo, hi = MG.shard_range(n_rays, rank, world)
    hits, _, _ = O.trace(O.Mbvh(tree["nodes"], tree["indices"]), tris, rays[lo:hi], threads=2)
    local = torch.from_numpy(hits.view(np.uint8).reshape(-1, 8).copy())
    full = MG.all_gather_ragged(local, MG.shard_counts(n_rays, world))
    np.save(os.path.join(out_dir, f"gathered_{rank}.npy"), full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rays", [10_000, 10_001])
def test_sharded_trace_equals_single_process(O, W, teapot, teapot_trees, tmp_path, n_rays):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_rays, str(tmp_path)), nprocs=world, join=True)
    rays = W.random_rays(n_rays, *W.bounds(teapot["tris"]))
    want, _, _ = O.trace(teapot_trees["sah"][1], teapot["tris"], rays)
    for r in range(world):
        got = np.load(tmp_path / f"gathered_{r}.npy").reshape(-1).view(O.HIT_DTYPE)
        assert got.tobytes() == want.tobytes()  # multi-rank result == single-process result, byte for byte


def test_shard_ranges_partition_the_index_space():
    from rtbvh_b200 import multigpu as MG
    for n in (0, 1, 7, 8, 1000, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            r = [MG.shard_range(n, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[g][1] == r[g + 1][0] for g in range(world - 1))
            assert sum(MG.shard_counts(n, world)) == n
            assert max(MG.shard_counts(n, world)) - min(MG.shard_counts(n, world)) <= 1
