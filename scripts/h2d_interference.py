#!/usr/bin/env python
"""Does PCIe copy traffic slow the resident traversal kernel down?  Config 2 scene, 8 M resident rays per launch, timed
alone and while another stream uploads 256 MB pinned buffers back to back (and a third downloads 64 MB ones)."""
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
h_in = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
d_in = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
h_out = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
d_out = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()


def run(copies, reps=20):
    for _ in range(3):
        scene.intersect_device(d_rays, n, d_hits, api.TREE_MBVH, stream=stream)
    torch.cuda.synchronize()
    if copies:
        for _ in range(reps + 6):
            with torch.cuda.stream(s_up):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s_down):
                h_out.copy_(d_out, non_blocking=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        scene.intersect_device(d_rays, n, d_hits, api.TREE_MBVH, stream=stream)
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b) / reps
    torch.cuda.synchronize()
    return n / ms / 1e3


print("kernel alone            : %.0f Mrays/s" % run(False))
print("kernel + H2D/D2H traffic: %.0f Mrays/s" % run(True))
print("kernel alone again      : %.0f Mrays/s" % run(False))
