#!/usr/bin/env python
"""A/B of traversal-kernel variants on config 2 (1 Mi-triangle soup, Mbvh, jittered primary rays), device-resident rays.

One process per variant (the library is chosen at load time: RTBVH_LIB / RTBVH_TRACE_MODE), a few seconds each:
    python scripts/trace_ab.py [--steps 30] [--check]         # prints one line: Mrays/s, checksum of the hit records
The checksum (sum of prim ids + bit pattern of t over the first batch) must be equal across variants: every variant is
bit-exact or it is not a candidate.  `--packets` measures the packet kernels on the same frames instead.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rtbvh_b200 import api, workloads as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--tris", type=int, default=1 << 20)
    ap.add_argument("--any", action="store_true")
    ap.add_argument("--bvh", action="store_true")
    ap.add_argument("--packets", action="store_true")
    ap.add_argument("--no-tiling", action="store_true")
    ap.add_argument("--name", default=os.environ.get("RTBVH_LIB", "default").split("librtbvh_rs")[-1].strip("_.so") or "default")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    api.set_device(0)
    tris = W.soup(a.tris)
    tree = api.TREE_BVH if a.bvh else api.TREE_MBVH
    scene = api.Scene.build(tris, api.BINNED_SAH, 1, mbvh=True)
    Wd = Hd = 1000
    if not a.no_tiling:
        scene.set_ray_tiling(Wd)
    cam = W.soup_camera(Wd, Hd)
    n = a.frames * Wd * Hd
    ring = 6
    stream = torch.cuda.current_stream().cuda_stream
    d_rays = [torch.empty(n * 8, dtype=torch.float32, device="cuda") for _ in range(ring)]
    for b in range(ring):
        for f in range(a.frames):
            api.generate_camera_rays_device(cam, 0, Hd, d_rays[b][f * Wd * Hd * 8:], jitter_seed=W.SEED_SOUP, frame=b * a.frames + f,
                                            stream=stream)
    torch.cuda.synchronize()
    if a.packets:
        # RayPacket4 of 4 x-adjacent pixels (examples/benchmark.rs:135-141): SoA of 7 x 4 floats per packet
        d_in = []
        for b in range(ring):
            r = d_rays[b].view(n // 4, 4, 8)
            pk = torch.stack([r[:, :, 0], r[:, :, 1], r[:, :, 2], r[:, :, 4], r[:, :, 5], r[:, :, 6], r[:, :, 7]], dim=1).contiguous()
            d_in.append(pk.view(-1))
        d_out = torch.empty(n * 2, dtype=torch.float32, device="cuda")
        d_occ = torch.empty(n, dtype=torch.uint8, device="cuda")

        def step(k):
            if a.any:
                scene.occluded_packets_device(d_in[k % ring], n // 4, d_occ, tree, stream=stream)
            else:
                scene.intersect_packets_device(d_in[k % ring], n // 4, d_out, tree, stream=stream)
    else:
        d_out = torch.empty(n * 2, dtype=torch.float32, device="cuda")
        d_occ = torch.empty(n, dtype=torch.uint8, device="cuda")

        def step(k):
            if a.any:
                scene.occluded_device(d_rays[k % ring], n, d_occ, tree, stream=stream)
            else:
                scene.intersect_device(d_rays[k % ring], n, d_out, tree, stream=stream)
    step(0)
    torch.cuda.synchronize()
    if a.any:
        chk = int(d_occ.sum(dtype=torch.int64))
    else:
        chk = int(d_out.view(torch.int32).to(torch.int64).sum())
    for k in range(a.warmup):
        step(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(a.steps):
        step(a.warmup + k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if scene.stack_overflowed():
        print(a.name, "STACK OVERFLOW")
    kind = ("packet4 " if a.packets else "single ") + ("any" if a.any else "closest") + (" bvh" if a.bvh else " mbvh")
    print(f"{a.name:28s} {os.environ.get('RTBVH_TRACE_MODE', '-'):10s} {kind:22s} {a.steps * n / ms / 1e3:9.1f} Mrays/s  "
          f"{ms / a.steps:7.3f} ms/step  checksum {chk}", flush=True)
    scene.free()


if __name__ == "__main__":
    main()
