"""The multi-GPU gather fused into the traversal kernel (rtbvh_gpu_*_device_scatter + rtbvh_gpu_peer_barrier,
include/rtbvh_gpu.h; SURVEY.md section 8e: rays sharded, tree replicated, hit records gathered).

Two processes (one per rank) exchange cudaIpc handles through a gloo group and trace their shard; each kernel writes
its records straight into BOTH ranks' gather buffers.  The ranks use distinct GPUs when the box has them and share
GPU 0 otherwise (the cudaIpc mapping, the scatter stores and the device barrier are the same code either way).
Every rank's gathered buffer must equal the single-process oracle result byte for byte, for several steps that
alternate between the two gather buffers."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_per_rank, steps, any_hit, out_dir, push=True, tile_w=0):
    import sys
    os.environ["RTBVH_GATHER_PUSH"] = "1" if push else "0"  # chunk-wise push (default) / one store per ray and destination
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from rtbvh_b200 import api, multigpu as MG, workloads as W
    dev = rank % api.device_count()
    torch.cuda.set_device(dev)
    api.set_device(dev)
    tris = W.teapot()
    aabbs, centers = O.prims_from_triangles(tris)
    rc, bvh = O.build(O.BINNED_SAH, aabbs, centers, 1)
    m = bvh.collapse()
    scene = api.Scene(tris, bvh=None, mbvh=api.Mbvh.from_arrays(m.nodes, m.indices))
    scene.set_ray_tiling(tile_w)  # work-order hint: 8x8 tiles inside whole bands of 8 rows of tile_w rays (results unchanged)
    rec = 1 if any_hit else 8
    fg = MG.FusedGather(n_per_rank, rec)
    torch.cuda.set_stream(torch.cuda.Stream())  # not the legacy default stream: the barrier's side stream overlaps it
    stream = torch.cuda.current_stream().cuda_stream
    n_total = world * n_per_rank
    snaps = []
    for k in range(steps):
        rays = W.random_rays(n_total, *W.bounds(tris), seed=0xF00D + k)
        lo = rank * n_per_rank
        d_rays = torch.from_numpy(rays[lo: lo + n_per_rank].view(np.float32).reshape(-1).copy()).cuda()
        local = torch.zeros(n_per_rank * rec, dtype=torch.uint8, device="cuda")
        fg.intersect(scene, d_rays, n_per_rank, k, d_hits=local, stream=stream, any_hit=any_hit)
        # stream-ordered consumer of step k: snapshot the gather buffer behind the step's barrier (side stream)
        fg.wait(k, stream)
        snap = api.device_view(fg.buffer_ptr(k), n_total * rec).clone()
        snaps.append((snap, local))
    torch.cuda.synchronize()
    for k, (snap, local) in enumerate(snaps):
        np.save(os.path.join(out_dir, f"g_{rank}_{k}.npy"), snap.cpu().numpy())
        np.save(os.path.join(out_dir, f"l_{rank}_{k}.npy"), local.cpu().numpy())
    assert not scene.stack_overflowed()
    fg.close()
    scene.free()
    dist.destroy_process_group()


@pytest.mark.parametrize("any_hit,push,n_per_rank,tile_w", [(False, True, 50_000, 0), (True, True, 50_000, 0),
                                                            (False, False, 50_000, 0), (False, True, 50_030, 40),
                                                            (True, True, 50_062, 48)])
def test_fused_gather_equals_single_process(O, W, teapot, teapot_trees, tmp_path, any_hit, push, n_per_rank, tile_w):
    import torch.multiprocessing as mp
    world, steps = 2, 4
    mp.spawn(_worker, args=(world, _free_port(), n_per_rank, steps, any_hit, str(tmp_path), push, tile_w), nprocs=world,
             join=True)
    for k in range(steps):
        rays = W.random_rays(world * n_per_rank, *W.bounds(teapot["tris"]), seed=0xF00D + k)
        want, _, _ = O.trace(teapot_trees["sah"][1], teapot["tris"], rays, mode="any" if any_hit else "closest")
        wb = np.ascontiguousarray(want).view(np.uint8).reshape(-1)
        rec = 1 if any_hit else 8
        for r in range(world):
            got = np.load(tmp_path / f"g_{r}_{k}.npy")
            assert got.tobytes() == wb.tobytes(), f"step {k}: rank {r}'s gather buffer differs from the oracle"
            loc = np.load(tmp_path / f"l_{r}_{k}.npy")
            assert loc.tobytes() == wb[r * n_per_rank * rec: (r + 1) * n_per_rank * rec].tobytes()
