#!/usr/bin/env python
"""Dynamic-scene loop on config 2's geometry: every frame the 1 Mi soup triangles wobble (device-side
vertex update), the scene is refitted in place (rtbvh_gpu_scene_refit_device: boxes -> Bvh refit -> Mbvh slot refresh
-> triangle records) and one 1000x1000 frame of primary rays is traced through the Mbvh.  Reported: refit ms per frame
and per Mtri (CUDA events), traversal Mrays/s on the refitted tree per frame, and the same frame traced through a
tree rebuilt from scratch into a new resident scene (rtbvh_gpu_scene_build_device: binned SAH + collapse + triangle
records, nothing leaves the device) with its wall-clock cost.  One JSON line on stdout."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from rtbvh_b200 import api, workloads as W  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 20
amp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.002
api.set_device(0)
tris0 = W.soup(1 << 20)
n = len(tris0)
bvh = api.build_triangles(tris0, api.BINNED_SAH, 1)
mbvh = api.Mbvh.construct(bvh)
scene = api.Scene(tris0, bvh=bvh, mbvh=mbvh)
stream = torch.cuda.current_stream().cuda_stream
cam = W.soup_camera(1000, 1000)
nr = 1_000_000
d_rays = torch.empty(nr * 8, dtype=torch.float32, device="cuda")
d_hits = torch.empty(nr * 2, dtype=torch.float32, device="cuda")
api.generate_camera_rays_device(cam, 0, 1000, d_rays, jitter_seed=W.SEED_SOUP, frame=0, stream=stream)
base = torch.from_numpy(tris0.reshape(n, 9).copy()).cuda()
cen = base.view(n, 3, 3).mean(dim=1)


def verts_at(frame):
    ph = cen[:, 0:1] * 7.0 + cen[:, 1:2] * 5.0 + frame * 0.35
    d = torch.cat([torch.sin(ph), torch.cos(ph * 1.3), torch.sin(ph * 0.7 + 1.0)], dim=1) * amp * frame
    return (base.view(n, 3, 3) + d[:, None, :]).contiguous().view(-1)


def timed(fn, reps=1):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def trace_rate(sc):
    sc.intersect_device(d_rays, nr, d_hits, api.TREE_MBVH, stream=stream)  # warm
    ms = timed(lambda: sc.intersect_device(d_rays, nr, d_hits, api.TREE_MBVH, stream=stream), reps=5)
    return nr / ms / 1e3


rows = []
base_rate = trace_rate(scene)
for f in range(1, frames + 1):
    v = verts_at(f)
    refit_ms = timed(lambda: scene.refit_device(v, n, 12, stream))
    rate = trace_rate(scene)
    hit_frac = float((d_hits.view(torch.int32)[1::2] != -1).float().mean())
    row = {"frame": f, "refit_ms": refit_ms, "mrays_refit_tree": rate, "hit_frac": hit_frac}
    if f in (1, frames // 2, frames):
        # full rebuild of the moved geometry straight into a new resident scene (binned SAH + collapse + records)
        api.Scene.build(v, api.BINNED_SAH, 1, mbvh=True, n_tris=n).free()  # warm the memory pool for this size
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fs = api.Scene.build(v, api.BINNED_SAH, 1, mbvh=True, n_tris=n)
        torch.cuda.synchronize()
        row["rebuild_resident_wall_ms"] = (time.perf_counter() - t0) * 1e3
        row["rebuild_device_ms"] = api.last_build_stats()["device_ms"]
        row["mrays_rebuilt_tree"] = trace_rate(fs)
        same = torch.empty_like(d_hits)
        fs.intersect_device(d_rays, nr, same, api.TREE_MBVH, stream=stream)
        scene.intersect_device(d_rays, nr, d_hits, api.TREE_MBVH, stream=stream)
        torch.cuda.synchronize()
        # t is tree independent where both trees are conservative; ids may differ only through the reference's Q3 boxes
        row["hits_equal_rebuilt_frac"] = float((same.view(torch.int32) == d_hits.view(torch.int32)).view(-1, 2).all(dim=1).float().mean())
        fs.free()
    rows.append(row)
    print(row, file=sys.stderr, flush=True)
print(json.dumps({"workload": "soup-1Mi-tris wobbling, refit every frame, 1 M primary rays per frame through the Mbvh",
                  "amplitude_per_frame": amp, "static_mrays": base_rate,
                  "refit_ms_median": float(np.median([r["refit_ms"] for r in rows])),
                  "refit_ms_per_mtri": float(np.median([r["refit_ms"] for r in rows])) / (n / 1e6), "frames": rows}))
