#!/usr/bin/env python
"""Throughput of the resident traversal kernel when 8 M rays are traced as k launches on 4 streams (no copies at all):
what the host-buffer pipeline could reach at best for a given chunk size."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from rtbvh_b200 import api, workloads as W  # noqa: E402

tris = W.soup(1 << 20)
scene = api.Scene.build(tris, api.BINNED_SAH, 1, mbvh=True)
n = 8_000_000
cam = W.soup_camera(1000, 1000)
stream = torch.cuda.current_stream().cuda_stream
d_rays = torch.empty(n * 8, dtype=torch.float32, device="cuda")
for f in range(8):
    api.generate_camera_rays_device(cam, 0, 1000, d_rays[f * 8_000_000:], jitter_seed=W.SEED_SOUP, frame=f, stream=stream)
d_hits = torch.empty(n * 2, dtype=torch.float32, device="cuda")
streams = [torch.cuda.Stream() for _ in range(4)]
torch.cuda.synchronize()
for chunks, nstreams in ((1, 1), (4, 1), (4, 4), (16, 1), (16, 4), (64, 4)):
    m = n // chunks
    def step():
        for c in range(chunks):
            st = streams[c % nstreams]
            scene.intersect_device(d_rays[c * m * 8:], m, d_hits[c * m * 2:], api.TREE_MBVH, stream=st.cuda_stream)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for st in streams:
        st.wait_event(a)
    reps = 10
    for _ in range(reps):
        step()
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    b.record()
    b.synchronize()
    print(f"{chunks:3d} launches of {m:8d} rays on {nstreams} stream(s): {n * reps / a.elapsed_time(b) / 1e3:7.0f} Mrays/s", flush=True)
