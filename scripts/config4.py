#!/usr/bin/env python
"""BASELINE config 4: 30 M-triangle scene, incoherent any-hit shadow rays, rays sharded across N GPUs.

    python scripts/config4.py                      (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/config4.py

Every rank builds nothing but its own copy of the scene arrays; the tree is built on rank 0 (GPU binned SAH +
collapse), broadcast over NCCL and replicated; each rank traces its own shard of the shadow rays (counter-based
generator, so shards are disjoint) and the occlusion bytes are all_gathered.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from rtbvh_b200 import api, multigpu as MG, workloads as W  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    instances = int(os.environ.get("CFG4_INSTANCES", "30"))
    rays_per_rank = int(os.environ.get("CFG4_RAYS", str(16_000_000)))
    steps = int(os.environ.get("CFG4_STEPS", "10"))
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    api.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t0 = time.time()
    tris = W.instanced_scene(instances)
    t_scene = time.time() - t0
    info = {"triangles": int(len(tris)), "scene_gen_s": t_scene}
    arrays = None
    if rank == 0:
        api.build_triangles(tris[: 1 << 16], api.BINNED_SAH, 1).free()  # module load / pool warm-up on a small input
        bvh = api.build_triangles(tris, api.BINNED_SAH, 1)
        st = api.last_build_stats()
        mbvh = api.Mbvh.construct(bvh)
        cst = api.last_build_stats()
        info.update(build_ms_per_mtri=st["device_ms"] / (len(tris) / 1e6), build_device_ms=st["device_ms"],
                    build_total_ms=st["total_ms"], collapse_device_ms=cst["device_ms"], bvh_nodes=int(bvh.rt.node_count),
                    mbvh_nodes=int(mbvh.rt.node_count))
        arrays = {"mnodes": mbvh.nodes, "indices": mbvh.indices}
    if world > 1:
        arrays = MG.broadcast_arrays(arrays, src=0, device="cuda")
        if rank != 0:
            mbvh = api.Mbvh.from_arrays(arrays["mnodes"], arrays["indices"])
    scene = api.Scene(tris, bvh=None, mbvh=mbvh)
    sort = os.environ.get("CFG4_SORT", "0") == "1"
    scene.set_ray_sorting(sort)
    info["ray_sorting"] = sort
    rays = W.shadow_rays(tris, rays_per_rank, first=rank * rays_per_rank)
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1).copy()).cuda()
    d_occ = torch.empty(rays_per_rank, dtype=torch.uint8, device="cuda")
    g_out = [torch.empty(world * rays_per_rank, dtype=torch.uint8, device="cuda") for _ in range(2)] if world > 1 else None
    stream = torch.cuda.current_stream().cuda_stream
    works = []

    def step(k):
        scene.occluded_device(d_rays, rays_per_rank, d_occ, api.TREE_MBVH, stream=stream)
        if world > 1:
            if len(works) >= 2:
                works[-2].wait()
            works.append(dist.all_gather_into_tensor(g_out[k % 2], d_occ, async_op=True))

    for k in range(3):
        step(k)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        step(3 + k)
    for w in works[-2:]:
        w.wait()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    if scene.stack_overflowed():
        raise RuntimeError("stack overflow")
    if rank == 0:
        from oracle import oracle as O  # checker + CPU baseline on a bounded sample
        threads = max(1, len(os.sched_getaffinity(0)))
        sample = rays[:200_000]
        otree = O.Mbvh(mbvh.nodes.copy(), mbvh.indices.copy())
        want, cms, _ = O.trace(otree, tris, sample, mode="any", threads=threads)
        _, _, cnt = O.trace(otree, tris, sample, mode="any", threads=threads, counters=True)
        got = d_occ[: len(sample)].cpu().numpy()
        nv, nt = cnt["node_visits"] / len(sample), cnt["prim_tests"] / len(sample)
        bpr = 32 + 1 + 128 * nv + 40 * nt
        value = world * steps * rays_per_rank / ms / 1e3
        out = {"config": "config 4: instanced 30M-triangle scene, incoherent any-hit shadow rays, Mbvh (GPU binned SAH)",
               "metric": "Mrays/s any-hit", "value": value, "n_gpus": world, "steps": steps, "rays_per_gpu_per_step": rays_per_rank,
               "ms_per_step": ms / steps, "occluded_fraction": float(got.mean()), "parity_sample_bit_exact": bool(np.array_equal(got, want)),
               "cpu_baseline": {"value": len(sample) / cms / 1e3, "unit": "Mrays/s", "cores": threads, "kind": "port",
                                "sample": "first 200000 shadow rays"},
               "bytes_per_ray": bpr, "node_visits": nv, "tri_tests": nt, "max_stack": cnt["max_stack"],
               "algorithmic_gbs_per_gpu": value / world * 1e6 * bpr / 1e9, **info}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    scene.free()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
