#!/usr/bin/env python
"""Measures every traversal variant of SURVEY.md section 8 on configs 1 and 2 (and the builders on bigger scenes),
next to the CPU oracle port on the box's host cores.  Not the driver's bench (that is bench.py): this fills the
coverage table in DESIGN.md / profiles/.  Run on the GPU box:  python scripts/matrix.py <tag> [--big]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import oracle as O  # noqa: E402  (checker / CPU baseline only)
from rtbvh_b200 import api, workloads as W  # noqa: E402

THREADS = max(1, len(os.sched_getaffinity(0)))


def gpu_time(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def traversal_matrix(name, tris, rays_frame, reps_frames, rows):
    aabbs, centers = O.prims_from_triangles(tris)
    bvh = api.build_triangles(tris, api.BINNED_SAH, 1)
    mbvh = api.Mbvh.construct(bvh)
    scene = api.Scene(tris, bvh=bvh, mbvh=mbvh)
    obvh = O.Bvh(bvh.nodes.copy(), bvh.indices.copy())
    ombvh = O.Mbvh(mbvh.nodes.copy(), mbvh.indices.copy())
    n1 = len(rays_frame)
    rays = np.tile(rays_frame, reps_frames)
    packets = W.pack4(rays)
    n = len(rays)
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1).copy()).cuda()
    d_pk = torch.from_numpy(packets.view(np.float32).reshape(-1).copy()).cuda()
    d_hits = torch.empty(n * 2, dtype=torch.float32, device="cuda")
    d_occ = torch.empty(n, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    sample = rays_frame[: min(n1, 400_000)]
    spk = W.pack4(sample[: len(sample) // 4 * 4])
    for tree, tname, otree in ((api.TREE_BVH, "Bvh", obvh), (api.TREE_MBVH, "Mbvh", ombvh)):
        for kind in ("single", "packet4"):
            for mode in ("closest", "any"):
                if kind == "single":
                    if mode == "closest":
                        ms = gpu_time(lambda: scene.intersect_device(d_rays, n, d_hits, tree, stream=stream))
                        got = d_hits[: len(sample) * 2].cpu().numpy().view(api.HIT_DTYPE).reshape(-1)
                        want, cms, cnt = O.trace(otree, tris, sample, threads=THREADS, counters=True)
                        _, cms, _ = O.trace(otree, tris, sample, threads=THREADS)
                        ok = bool(np.array_equal(got, want))
                    else:
                        ms = gpu_time(lambda: scene.occluded_device(d_rays, n, d_occ, tree, stream=stream))
                        got = d_occ[: len(sample)].cpu().numpy()
                        want, cms, cnt = O.trace(otree, tris, sample, mode="any", threads=THREADS, counters=True)
                        _, cms, _ = O.trace(otree, tris, sample, mode="any", threads=THREADS)
                        ok = bool(np.array_equal(got, want))
                    nv, nt = cnt["node_visits"] / len(sample), cnt["prim_tests"] / len(sample)
                    inner = cnt["inner_visits"] / len(sample)
                    bpr = (32 + (8 if mode == "closest" else 1) +
                           (128 * nv if tree == api.TREE_MBVH else 64 * inner) + 40 * nt)
                else:
                    if mode == "closest":
                        ms = gpu_time(lambda: scene.intersect_packets_device(d_pk, n // 4, d_hits, tree, stream=stream))
                        got = d_hits[: len(spk) * 8].cpu().numpy().view(api.HIT4_DTYPE).reshape(-1)
                        want, cms, cnt = O.trace_packets(otree, tris, spk, threads=THREADS, counters=True)
                        _, cms, _ = O.trace_packets(otree, tris, spk, threads=THREADS)
                        ok = bool(np.array_equal(got, want))
                    else:
                        ms = gpu_time(lambda: scene.occluded_packets_device(d_pk, n // 4, d_occ, tree, stream=stream))
                        got = d_occ[: len(spk) * 4].cpu().numpy().reshape(-1, 4)
                        want, cms, cnt = O.trace_packets(otree, tris, spk, mode="any", threads=THREADS, counters=True)
                        _, cms, _ = O.trace_packets(otree, tris, spk, mode="any", threads=THREADS)
                        ok = bool(np.array_equal(got, want))
                    nv, nt = cnt["node_visits"] / len(spk), cnt["prim_tests"] / len(spk)
                    inner = cnt["inner_visits"] / len(spk)
                    bpr = (112 + (32 if mode == "closest" else 4) +
                           (128 * nv if tree == api.TREE_MBVH else 64 * inner) + 40 * nt) / 4
                gpu_mrays = n / ms / 1e3
                cpu_mrays = len(sample) / cms / 1e3
                rows.append(dict(scene=name, tree=tname, rays=kind, query=mode, gpu_mrays=gpu_mrays, cpu_mrays=cpu_mrays,
                                 cpu_cores=THREADS, parity_bit_exact=ok, bytes_per_ray=bpr,
                                 algorithmic_gbs=gpu_mrays * 1e6 * bpr / 1e9, node_visits=nv, tri_tests=nt,
                                 max_stack=cnt["max_stack"]))
                print(json.dumps(rows[-1]), flush=True)
    scene.free()


def build_matrix(name, tris, rows, oracle_too=True):
    mtri = len(tris) / 1e6
    for kind, kname, okind in ((api.BINNED_SAH, "binned_sah", O.BINNED_SAH), (api.LOCALLY_ORDERED_CLUSTERED, "locb", O.LOCB)):
        api.build_triangles(tris, kind, 1).free()
        dev = []
        b = None
        for _ in range(3):
            if b is not None:
                b.free()
            b = api.build_triangles(tris, kind, 1)
            st = api.last_build_stats()
            dev.append(st["device_ms"])
        m = api.Mbvh.construct(b)
        cst = api.last_build_stats()
        sah = O.Bvh(b.nodes.copy(), b.indices.copy()).sah_cost()
        row = dict(scene=name, builder=kname, tris=len(tris), gpu_ms_per_mtri=float(np.median(dev)) / mtri,
                   gpu_collapse_ms=cst["device_ms"], nodes=int(b.rt.node_count), mnodes=int(m.rt.node_count), sah=sah,
                   locb_iterations=st["iterations"])
        if oracle_too:
            aabbs, centers = O.prims_from_triangles(tris)
            t0 = time.time()
            rc, ob = O.build(okind, aabbs, centers, 1, parallel=True)
            row["cpu_ms_per_mtri"] = (time.time() - t0) * 1e3 / mtri
            row["cpu_sah"] = ob.sah_cost()
            row["cpu_collapse_ms"] = ob.collapse().collapse_ms
            row["cpu_cores_note"] = "oracle port: serial builder (LOCB: parallel Morton sort only, like the reference)"
        rows.append(row)
        print(json.dumps(row), flush=True)
        m.free()
        b.free()


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "matrix"
    big = "--big" in sys.argv
    api.set_device(0)
    rows = []
    teapot = W.teapot()
    traversal_matrix("teapot/benchmark-camera", teapot, W.camera_rays(W.benchmark_camera(1000, 1000)), 8, rows)
    traversal_matrix("teapot/incoherent", teapot, W.random_rays(1_000_000, *W.bounds(teapot)), 8, rows)
    soup = W.soup(1 << 20)
    traversal_matrix("soup1Mi/primary", soup, W.camera_rays(W.soup_camera(1000, 1000), jitter_seed=W.SEED_SOUP), 8, rows)
    traversal_matrix("soup1Mi/incoherent", soup, W.random_rays(1_000_000, *W.bounds(soup)), 8, rows)
    build_matrix("teapot", teapot, rows)
    build_matrix("soup1Mi", soup, rows)
    if big:
        build_matrix("heightfield10M", W.heightfield(2237, 2237), rows, oracle_too=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"{tag}_matrix.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
