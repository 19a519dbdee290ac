#!/usr/bin/env python
"""BASELINE config 5: a spatial-split SAH *reference* tree uploaded to the GPU unchanged, diffuse-bounce rays,
closest hit, rays sharded across N GPUs.

The tree is built on the CPU by the oracle's restatement of the reference's SpatialSahBuilder (the reference itself
is Rust and cannot run here) — it stands in for "the reference built this tree"; the product only uploads and
traverses it.  `CFG5_TREE=<npz>` loads a tree built beforehand with `python scripts/config5.py --build-tree <npz>`
(the CPU build takes ~85 s per M triangles).  Scene: config-2 soup with offsets stretched x(8,1,1) (long thin
triangles).  Rays: primary camera rays are traced first; every hit spawns one cosine-weighted bounce ray from the hit
point.  Prints one JSON line (rank 0).

    python scripts/config5.py                      (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/config5.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from rtbvh_b200 import workloads as W  # noqa: E402

N_TRIS = int(os.environ.get("CFG5_TRIS", str(1 << 20)))
SEED = W.SEED_SOUP + 5


def scene():
    return W.soup(N_TRIS, seed=SEED, aniso=(8, 1, 1))


def build_tree(path):
    from oracle import oracle as O
    tris = scene()
    t0 = time.time()
    rc, bvh = O.build_spatial(tris, 1, True)
    assert rc == 0 and bvh.validate(len(tris))
    np.savez(path, nodes=bvh.nodes, indices=bvh.indices, build_s=time.time() - t0, stats=np.array(bvh.stats), sah=bvh.sah_cost())
    print("built", len(bvh.nodes), "nodes in", round(time.time() - t0, 1), "s; stats", bvh.stats, "SAH", bvh.sah_cost())


def bounce_rays(torch, d_rays, d_hits, d_tris, seed):
    """One cosine-weighted bounce ray per primary hit (device, torch ops).  Returns [m, 8] float32."""
    rays = d_rays.view(-1, 8)
    hits = d_hits.view(-1, 2)
    prim = hits[:, 1].view(torch.int32)
    ok = prim != -1
    rays, t, prim = rays[ok], hits[ok, 0], prim[ok].long()
    o, d = rays[:, 0:3], rays[:, 4:7]
    p = o + t[:, None] * d
    tri = d_tris[prim]
    n = torch.linalg.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    n = n / n.norm(dim=1, keepdim=True).clamp_min(1e-20)
    n = torch.where((n * d).sum(1, keepdim=True) > 0, -n, n)
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    u = torch.rand((len(p), 2), generator=g, device="cuda")
    r, phi = u[:, 0].sqrt(), u[:, 1] * (2 * np.pi)
    a = torch.where(n[:, 0:1].abs() > 0.9, torch.tensor([0.0, 1.0, 0.0], device="cuda"), torch.tensor([1.0, 0.0, 0.0], device="cuda"))
    tx = torch.linalg.cross(n, a.expand_as(n))
    tx = tx / tx.norm(dim=1, keepdim=True)
    ty = torch.linalg.cross(n, tx)
    nd = tx * (r * phi.cos())[:, None] + ty * (r * phi.sin())[:, None] + n * (1 - u[:, 0]).clamp_min(0).sqrt()[:, None]
    nd = nd / nd.norm(dim=1, keepdim=True)
    out = torch.empty((len(p), 8), dtype=torch.float32, device="cuda")
    out[:, 0:3] = p + n * 1e-4
    out[:, 3] = 1e-4
    out[:, 4:7] = nd
    out[:, 7] = 1e34
    return out.contiguous()


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--build-tree":
        return build_tree(sys.argv[2])
    import torch
    import torch.distributed as dist
    from rtbvh_b200 import api, multigpu as MG

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    frames = int(os.environ.get("CFG5_FRAMES", "16"))
    steps = int(os.environ.get("CFG5_STEPS", "10"))
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    api.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tris = scene()
    info = {"triangles": int(len(tris))}
    arrays = None
    if rank == 0:
        path = os.environ.get("CFG5_TREE")
        if path and os.path.exists(path):
            z = np.load(path)
            nodes, indices = z["nodes"], z["indices"]
            info.update(tree_source=f"oracle SpatialSahBuilder restatement, prebuilt ({float(z['build_s']):.0f} s on CPU)",
                        sbvh_stats=[int(x) for x in z["stats"]], sah=float(z["sah"]))
        else:
            from oracle import oracle as O
            t0 = time.time()
            rc, ob = O.build_spatial(tris, 1, True)
            nodes, indices = ob.nodes, ob.indices
            info.update(tree_source=f"oracle SpatialSahBuilder restatement, built in-run ({time.time() - t0:.0f} s on CPU)",
                        sbvh_stats=list(ob.stats), sah=ob.sah_cost())
        bvh = api.Bvh.from_arrays(nodes, indices)
        mbvh = api.Mbvh.construct(bvh)  # GPU collapse of the reference-format binary tree
        info.update(collapse_device_ms=api.last_build_stats()["device_ms"], bvh_nodes=len(nodes), index_count=len(indices),
                    mbvh_nodes=int(mbvh.rt.node_count))
        arrays = {"mnodes": mbvh.nodes, "mindices": mbvh.indices}
    if world > 1:
        arrays = MG.broadcast_arrays(arrays, src=0, device="cuda")
        if rank != 0:
            mbvh = api.Mbvh.from_arrays(arrays["mnodes"], arrays["mindices"])
    scene_gpu = api.Scene(tris, bvh=None, mbvh=mbvh)
    stream = torch.cuda.current_stream().cuda_stream
    # primary pass: `frames` jittered frames per rank (distinct per rank), then one bounce ray per hit
    cam = W.soup_camera(1000, 1000)
    n_primary = frames * 1_000_000
    d_prim = torch.empty(n_primary * 8, dtype=torch.float32, device="cuda")
    for f in range(frames):
        api.generate_camera_rays_device(cam, 0, 1000, d_prim[f * 8_000_000:], jitter_seed=SEED, frame=rank * frames + f, stream=stream)
    d_phits = torch.empty(n_primary * 2, dtype=torch.float32, device="cuda")
    scene_gpu.intersect_device(d_prim, n_primary, d_phits, api.TREE_MBVH, stream=stream)
    torch.cuda.synchronize()
    d_tris = torch.from_numpy(tris).cuda()
    d_rays = bounce_rays(torch, d_prim, d_phits, d_tris, seed=1234 + rank)
    n = int(d_rays.shape[0])
    # equal shard sizes make the all_gather regular: trim to a common multiple
    nt = torch.tensor([n], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(nt, op=dist.ReduceOp.MIN)
    n = int(nt[0])
    d_rays = d_rays[:n].contiguous()
    del d_prim, d_phits
    d_hits = torch.empty(n * 2, dtype=torch.float32, device="cuda")
    g_out = [torch.empty(world * n * 2, dtype=torch.float32, device="cuda") for _ in range(2)] if world > 1 else None
    results = {}
    for sort in (False, True):
        scene_gpu.set_ray_sorting(sort)
        works = []

        def step(k):
            scene_gpu.intersect_device(d_rays, n, d_hits, api.TREE_MBVH, stream=stream)
            if world > 1:
                if len(works) >= 2:
                    works[-2].wait()
                works.append(dist.all_gather_into_tensor(g_out[k % 2], d_hits, async_op=True))

        for k in range(2):
            step(k)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            step(2 + k)
        for w in works[-2:]:
            w.wait()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        results[sort] = float(t[0])
    if scene_gpu.stack_overflowed():
        raise RuntimeError("stack overflow")
    if rank == 0:
        from oracle import oracle as O  # checker + CPU baseline on a bounded sample
        threads = max(1, len(os.sched_getaffinity(0)))
        sample = d_rays[:200_000].cpu().numpy().view(api.RAY_DTYPE).reshape(-1)
        otree = O.Mbvh(mbvh.nodes.copy(), mbvh.indices.copy())
        want, cms, _ = O.trace(otree, tris, sample, threads=threads)
        _, _, cnt = O.trace(otree, tris, sample, threads=threads, counters=True)
        got = d_hits[: len(sample) * 2].cpu().numpy().view(api.HIT_DTYPE).reshape(-1)
        nv, nt_ = cnt["node_visits"] / len(sample), cnt["prim_tests"] / len(sample)
        bpr = 32 + 8 + 128 * nv + 40 * nt_
        best = min(results.values())
        value = world * steps * n / best / 1e3
        out = {"config": "config 5: reference-built spatial-split SAH tree uploaded unchanged, diffuse bounce rays, closest hit",
               "metric": "Mrays/s closest-hit", "value": value, "n_gpus": world, "steps": steps, "rays_per_gpu_per_step": n,
               "mrays_unsorted": world * steps * n / results[False] / 1e3, "mrays_sorted": world * steps * n / results[True] / 1e3,
               "hit_fraction": float((got["prim"] != api.NO_HIT).mean()), "parity_sample_bit_exact": bool(np.array_equal(got, want)),
               "cpu_baseline": {"value": len(sample) / cms / 1e3, "unit": "Mrays/s", "cores": threads, "kind": "port",
                                "sample": "first 200000 bounce rays"},
               "bytes_per_ray": bpr, "node_visits": nv, "tri_tests": nt_, "max_stack": cnt["max_stack"],
               "algorithmic_gbs_per_gpu": value / world * 1e6 * bpr / 1e9, **info}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    scene_gpu.free()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
