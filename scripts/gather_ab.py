#!/usr/bin/env python
"""Where does the fused gather's time go?  Run under torchrun (N ranks, one GPU each) on config 2's frames:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/gather_ab.py
Per variant: device time per step (CUDA events around the whole loop, max over ranks).
  plain          rtbvh_gpu_intersect_device, no gather, no barrier
  barrier        the same + rtbvh_gpu_peer_barrier per step
  own            scatter into the own gather buffer only, no barrier
  peers          scatter into the peers' buffers only, no barrier
  all            scatter into every buffer, no barrier
  all+barrier    the product path (FusedGather.intersect)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from rtbvh_b200 import api, multigpu as MG, workloads as W  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    api.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    steps, warm, frames = int(os.environ.get("STEPS", 30)), 5, 8
    Wd = Hd = 1000
    n = frames * Wd * Hd
    blob = [None]
    scene = None
    if rank == 0:
        scene = api.Scene.build(W.soup(1 << 20), api.BINNED_SAH, 1, mbvh=True)
        blob[0] = scene.export_bytes()
    dist.broadcast_object_list(blob, src=0)
    if rank != 0:
        scene = api.Scene.import_bytes(blob[0])
    dist.barrier()
    scene.set_ray_tiling(Wd if os.environ.get("TILING", "1") == "1" else 0)
    cam = W.soup_camera(Wd, Hd)
    ring = 6
    stream = torch.cuda.current_stream().cuda_stream
    d_rays = [torch.empty(n * 8, dtype=torch.float32, device="cuda") for _ in range(ring)]
    for b in range(ring):
        for f in range(frames):
            api.generate_camera_rays_device(cam, 0, Hd, d_rays[b][f * Wd * Hd * 8:], jitter_seed=W.SEED_SOUP,
                                            frame=(rank * ring + b) * frames + f, stream=stream)
    d_hits = torch.empty(n * 2, dtype=torch.float32, device="cuda")
    fg = MG.FusedGather(n, 8)
    torch.cuda.synchronize()
    state = {"step": 1000}

    def barrier_step():
        state["step"] += 1
        api.peer_barrier(fg.flag_dests, rank, state["step"], stream)

    def variant(name):
        def step(k):
            b = k % ring
            dests = fg.dests[k % 2]
            if name == "plain" or name == "barrier":
                scene.intersect_device(d_rays[b], n, d_hits, api.TREE_MBVH, stream=stream)
            elif name == "own":
                scene.intersect_device_scatter(d_rays[b], n, [dests[rank]], rank * n, d_hits, api.TREE_MBVH, stream)
            elif name == "peers":
                scene.intersect_device_scatter(d_rays[b], n, [d for r, d in enumerate(dests) if r != rank], rank * n, d_hits,
                                               api.TREE_MBVH, stream)
            else:
                scene.intersect_device_scatter(d_rays[b], n, dests, rank * n, d_hits, api.TREE_MBVH, stream)
            if name in ("barrier", "all+barrier"):
                barrier_step()
        for k in range(warm):
            step(k)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            step(warm + k)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        if rank == 0:
            print(f"N={world} push={os.environ.get('RTBVH_GATHER_PUSH', '1')} {name:12s} {float(t[0]):7.3f} ms/step  "
                  f"{world * n / float(t[0]) / 1e3:8.1f} Mrays/s", flush=True)

    for name in ("plain", "barrier", "own", "peers", "all", "all+barrier", "plain"):
        variant(name)
    fg.close()
    scene.free()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
