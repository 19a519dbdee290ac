#!/usr/bin/env python
"""How many rays would change their answer under ANY internal node format that is conservative (wide nodes that skip binary
levels with exact boxes, quantised boxes, ...)?  SURVEY.md quirk Q3: the reference's binned-SAH fallback leaves some left-child
boxes that do NOT enclose their primitives, and a reference-exact traversal reproduces the resulting misses.  A conservative
format finds those hits again, so its answers differ from the reference's exactly on the rays counted here.

Arbiter: the SAME tree (same topology, same leaves) after rtbvh_gpu_scene_refit with the unchanged vertices — refit rebuilds
every box bottom-up as the union of what is below it (src/bvh.rs:176-205), i.e. the conservative twin of the as-built tree.
Closest-hit records of the two trees are compared on config 2's primary rays (8 M per frame batch) and on incoherent rays;
any-hit flags on config 4's shadow rays (reduced instance count unless --full)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rtbvh_b200 import api, workloads as W  # noqa: E402


def count(scene, tris, d_rays, n, any_hit):
    stream = torch.cuda.current_stream().cuda_stream
    out = {}
    for tree, name in ((api.TREE_MBVH, "mbvh"), (api.TREE_BVH, "bvh")):
        if any_hit:
            a = torch.empty(n, dtype=torch.uint8, device="cuda")
            scene.occluded_device(d_rays, n, a, tree, stream=stream)
        else:
            a = torch.empty(n * 2, dtype=torch.int32, device="cuda")
            scene.intersect_device(d_rays, n, a, tree, stream=stream)
        torch.cuda.synchronize()
        out[name] = a.clone()
    return out


def compare(a, b, any_hit):
    res = {}
    for k in a:
        if any_hit:
            res[k] = {"rays": int(a[k].numel()), "flag_differs": int((a[k] != b[k]).sum()),
                      "occluded_only_in_conservative": int(((a[k] == 0) & (b[k] != 0)).sum())}
        else:
            ra, rb = a[k].view(-1, 2), b[k].view(-1, 2)
            prim = ra[:, 1] != rb[:, 1]
            res[k] = {"rays": int(ra.shape[0]), "prim_differs": int(prim.sum()), "t_bits_differ": int((ra[:, 0] != rb[:, 0]).sum()),
                      "miss_in_reference_hit_in_conservative": int(((ra[:, 1] == -1) & (rb[:, 1] != -1)).sum())}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config4-instances", type=int, default=30)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    api.set_device(0)
    stream = torch.cuda.current_stream().cuda_stream
    report = {}
    # ---- config 2: 1 Mi-triangle soup, binned SAH, leaf size 1
    tris = W.soup(1 << 20)
    scene = api.Scene.build(tris, api.BINNED_SAH, 1, mbvh=True)
    cam = W.soup_camera(1000, 1000)
    n = 8_000_000
    d_rays = torch.empty(n * 8, dtype=torch.float32, device="cuda")
    for f in range(8):
        api.generate_camera_rays_device(cam, 0, 1000, d_rays[f * 8_000_000:], jitter_seed=W.SEED_SOUP, frame=f, stream=stream)
    inc = W.random_rays(4_000_000, *W.bounds(tris), seed=0x93)
    d_inc = torch.from_numpy(inc.view(np.float32).reshape(-1).copy()).cuda()
    before = count(scene, tris, d_rays, n, False)
    before_inc = count(scene, tris, d_inc, len(inc), False)
    scene.refit(tris)  # conservative twin: same topology, boxes rebuilt bottom-up
    after = count(scene, tris, d_rays, n, False)
    after_inc = count(scene, tris, d_inc, len(inc), False)
    report["config2_primary"] = compare(before, after, False)
    report["config2_incoherent"] = compare(before_inc, after_inc, False)
    scene.free()
    # ---- config 4 flavour: instanced scene, shadow rays, any hit
    tris4 = W.instanced_scene(a.config4_instances)
    if tris4 is not None:
        scene = api.Scene.build(tris4, api.BINNED_SAH, 1, mbvh=True)
        rays = W.shadow_rays(tris4, 4_000_000)
        d = torch.from_numpy(rays.view(np.float32).reshape(-1).copy()).cuda()
        b4 = count(scene, tris4, d, len(rays), True)
        scene.refit(tris4)
        a4 = count(scene, tris4, d, len(rays), True)
        report["config4_shadow"] = {"triangles": int(len(tris4)), **compare(b4, a4, True)}
        scene.free()
    print(json.dumps(report), flush=True)


if __name__ == "__main__":
    main()
