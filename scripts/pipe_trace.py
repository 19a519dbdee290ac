#!/usr/bin/env python
"""Per-chunk timeline of the host-buffer pipeline (RTBVH_PIPE_TRACE=1): one rtbvh_gpu_intersect call over 8 M pinned rays."""
import os
import sys

os.environ["RTBVH_PIPE_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from rtbvh_b200 import api, workloads as W  # noqa: E402

tris = W.soup(1 << 20)
bvh = api.build_triangles(tris, api.BINNED_SAH, 1)
mbvh = api.Mbvh.construct(bvh)
scene = api.Scene(tris, bvh=None, mbvh=mbvh)
n = 8_000_000
cam = W.soup_camera(1000, 1000)
d = torch.empty(n * 8, dtype=torch.float32, device="cuda")
for f in range(8):
    api.generate_camera_rays_device(cam, 0, 1000, d[f * 8_000_000:], jitter_seed=W.SEED_SOUP, frame=f)
h = torch.empty(n * 8, dtype=torch.float32).pin_memory()
h.copy_(d)
o = torch.empty(n * 2, dtype=torch.float32).pin_memory()
for rep in range(3):
    print(f"--- call {rep}", file=sys.stderr, flush=True)
    scene.intersect_ptr(h.data_ptr(), n, o.data_ptr(), api.TREE_MBVH)
