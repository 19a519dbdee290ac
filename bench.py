#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config, one JSON line.

Workload (configs[1]): synthetic 1 Mi-triangle soup (rtbvh_b200/workloads.soup, seed 0x50A90002), binned-SAH
Bvh collapsed to an Mbvh, primary rays closest-hit: 1000x1000 pinhole frames from (0.5, 0.5, -1.5), 50 degree
fov, hashed sub-pixel jitter per (frame, pixel).  One STEP = `--frames-per-step` frames (default 8 = 8 M rays,
256 MB of rays: larger than the 126 MB L2, so no flush is needed between steps).  With the default
--steps 125 the timed region traces exactly the config's 1 B rays.

  value  : Mrays/s, rays already resident in HBM, one traversal kernel launch per step, CUDA events on the
           launching stream, max over ranks.
  e2e    : Mrays/s through the host-buffer C-ABI calls: pinned host rays -> H2D -> kernel -> D2H hit records inside
           the timed region every step.  value = submit/wait flavour with split origin / direction arrays
           (rtbvh_gpu_intersect_od_async + rtbvh_gpu_wait, two steps in flight, 24 B per ray);
           rtray_async_value = the same with 32-byte RTRay records; sync_call_value = the blocking rtbvh_gpu_intersect.
  roofline.achieved : algorithmic bytes per ray (32 + 8 + 128*n_m + 40*n_p; n_m, n_p = node visits / triangle
           tests per ray counted by the instrumented CPU oracle on a sample of the same rays, SURVEY.md
           section 8d) x rays per launch / mean launch duration; peak = MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline : the CPU oracle (a C++ port of the reference's loop, OpenMP over chunks of 1000 rays like
           examples/benchmark.rs:25) on all host cores, on a bounded sample of the same rays.

`--impl reference` times that CPU port alone (the reference itself is Rust and cannot be built in this image).
`--config {1,3,4,5}` runs the other BASELINE.json configs (bench_configs.py) with the same JSON shape; config 1 (teapot,
examples/benchmark.rs) is the one with published reference numbers and fills `vs_baseline`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rtbvh_b200 import workloads as W  # noqa: E402

METRIC = "Mrays/s closest-hit (Mbvh, binned SAH, 1Mi-triangle soup, primary rays)"
WIDTH = HEIGHT = 1000
N_TRIS = 1 << 20


from bench_common import ClockSampler, bind_to_gpu_numa_node, host_threads, log, measured_peak_gbs, ncu_traffic, nvlink_kib  # noqa: E402,F401


def build_scene_host():
    """Triangles + tree for config 2.  Returns (tris, bvh arrays, mbvh arrays, info)."""
    tris = W.soup(N_TRIS)
    return tris


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle): used as cpu_baseline on rank 0 and as the whole of --impl reference
# ------------------------------------------------------------------------------------------------
def oracle_tree(tris):
    from oracle import oracle as O
    aabbs, centers = O.prims_from_triangles(tris)
    t0 = time.time()
    rc, bvh = O.build(O.BINNED_SAH, aabbs, centers, 1)
    assert rc == 0
    build_s = time.time() - t0
    m = bvh.collapse()
    return O, bvh, m, build_s


def oracle_threaded_build_ms(O, tris):
    """The CPU builder with the reference's threaded scheduling (subtrees > 1024 primitives on other threads, src/utils.rs:
    189-289) on every host core: the honest CPU figure next to the GPU builder's ms per Mtri.  None on any failure."""
    try:
        aabbs, centers = O.prims_from_triangles(tris)
        rc, b = O.build(O.BINNED_SAH, aabbs, centers, 1, parallel=True)
        return b.build_ms if rc == 0 else None
    except Exception:
        return None


def cpu_sample_rate(O, m, tris, rays, target_s=10.0):
    """Times the oracle on a bounded sample; returns (Mrays/s, n_sample, counters per ray)."""
    threads = host_threads()
    probe = rays[: min(len(rays), 200_000)]
    _, ms, _ = O.trace(m, tris, probe, threads=threads)
    rate = len(probe) / max(ms, 1e-3) * 1e3
    n = int(min(len(rays), max(len(probe), rate * target_s)))
    _, ms, _ = O.trace(m, tris, rays[:n], threads=threads)
    _, _, cnt = O.trace(m, tris, rays[: min(n, 1_000_000)], threads=threads, counters=True)
    nc = min(n, 1_000_000)
    return n / ms / 1e3, n, {k: cnt[k] / nc for k in ("node_visits", "prim_tests")}, cnt["max_stack"], threads


def cpu_single_thread_rate(O, m, tris, rays, n=400_000):
    """The same loop on ONE thread (SURVEY.md section 8d asks for it next to the all-core figure); None on any failure."""
    try:
        sample = rays[: min(len(rays), n)]
        _, ms, _ = O.trace(m, tris, sample, threads=1)
        return len(sample) / max(ms, 1e-3) / 1e3
    except Exception:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tris = build_scene_host()
    O, bvh, m, build_s = oracle_tree(tris)
    cam = W.soup_camera(WIDTH, HEIGHT)
    rows = 250  # each step: a bounded sample of the frame (250 rows x 1000 px = 250 k rays)
    threads = host_threads()

    def step(k):
        rays = W.camera_rays(cam, y0=(k * rows) % HEIGHT, y1=(k * rows) % HEIGHT + rows, jitter_seed=W.SEED_SOUP,
                             frame=k * rows // HEIGHT)
        _, ms, _ = O.trace(m, tris, rays, threads=threads)
        return len(rays), ms

    for k in range(args.warmup):
        step(k)
    n_tot, ms_tot = 0, 0.0
    for k in range(args.steps):
        n, ms = step(args.warmup + k)
        n_tot += n
        ms_tot += ms
    v = n_tot / ms_tot / 1e3
    sample = f"{rows}x{WIDTH} jittered camera rays per step ({n_tot} rays total), Mbvh single-ray closest hit"
    one = cpu_single_thread_rate(O, m, tris, W.camera_rays(cam, y0=0, y1=400, jitter_seed=W.SEED_SOUP, frame=0))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "soup-1Mi-tris binned-SAH Mbvh primary rays closest-hit (BASELINE configs[1])",
                   "note": "reference is Rust and cannot be built in this image (no rustc/cargo): this is the C++ "
                           "oracle port of its loop, OpenMP dynamic chunks of 1000 rays like benchmark.rs:25"},
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample,
                         "single_thread_value": one},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "oracle_build_ms_per_mtri": build_s * 1e3 / (N_TRIS / 1e6),
        "oracle_build_ms_per_mtri_all_cores": (lambda v: v / (N_TRIS / 1e6) if v else None)(oracle_threaded_build_ms(O, tris)),
        "oracle_build_note": "binned SAH, 1 Mi triangles: one thread (deterministic numbering) and the reference's threaded "
                             f"scheduling on {threads} threads (subtrees > 1024 primitives run on other threads)",
    }))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from rtbvh_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line (the JSON): libraries that print there (NCCL's version banner) go to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if api.device_count() == 0:
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    api.set_device(local)
    # everything below runs on ONE non-default stream: the fused gather's step barrier lives on a side stream, and a side stream
    # only overlaps work that is not on the legacy default stream (which synchronises with every blocking stream)
    torch.cuda.set_stream(torch.cuda.Stream())
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    all_cpus = set(os.sched_getaffinity(0))
    numa = bind_to_gpu_numa_node(local) if os.environ.get("RTBVH_BENCH_NUMA", "1") == "1" else "off"

    tris = build_scene_host()
    info = {"host_numa": numa}
    if world == 1:
        bvh, mbvh, info = build_trees(api, tris, info, rank)
        scene = api.Scene(tris, bvh=None, mbvh=mbvh)
    else:
        # the tree is built once (rank 0) and replicated through the C ABI: rank 0 exports its resident
        # scene (cudaIpc handles, 512 bytes over the process group), every other rank copies it device to device over NVLink
        # (rtbvh_gpu_scene_export / rtbvh_gpu_scene_import) — no host hop, no collective
        scene, blob = None, [None]
        if rank == 0:
            bvh, mbvh, info = build_trees(api, tris, info, rank)
            scene = api.Scene(tris, bvh=None, mbvh=mbvh)
            try:
                blob[0] = scene.export_bytes()
            except api.RtbvhError as e:
                log(f"[bench] scene export failed: {e}")
        dist.broadcast_object_list(blob, src=0)
        ok = 1
        t_rep = time.perf_counter()
        if rank != 0:
            try:
                if blob[0] is None:
                    raise api.RtbvhError(1, "no export")
                scene = api.Scene.import_bytes(blob[0])
            except api.RtbvhError as e:
                log(f"[bench] scene import failed on rank {rank}: {e}")
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # also the barrier that keeps rank 0's scene alive until all imports are done
        rep_ms = (time.perf_counter() - t_rep) * 1e3
        if int(flag[0]) == 1:
            info["replication"] = ("tree built on rank 0; rtbvh_gpu_scene_export -> rtbvh_gpu_scene_import on every other rank: "
                                   f"device-to-device copies over NVLink, {rep_ms:.1f} ms for all ranks")
        else:  # no cudaIpc / peer access on this box: replicate through the process group instead
            from rtbvh_b200 import multigpu as MG
            if rank != 0 and scene is not None:
                scene.free()
            arrays = {"mnodes": mbvh.nodes, "indices": mbvh.indices} if rank == 0 else None
            arrays = MG.broadcast_arrays(arrays, src=0, device="cuda")
            if rank != 0:
                mbvh = api.Mbvh.from_arrays(arrays["mnodes"], arrays["indices"])
                scene = api.Scene(tris, bvh=None, mbvh=mbvh)
            info["replication"] = "tree built on rank 0, broadcast over NCCL (cudaIpc import unavailable), one replica per GPU"
    sort_rays = os.environ.get("RTBVH_BENCH_SORT", "0") == "1"  # experiment knob; primary rays are coherent already
    scene.set_ray_sorting(sort_rays)
    info["ray_sorting"] = sort_rays
    # work-order hint for image-ordered batches: 8x8 pixel tiles per warp instead of 64-pixel row segments (results unchanged)
    tiling = os.environ.get("RTBVH_BENCH_TILING", "1") == "1"
    scene.set_ray_tiling(WIDTH if tiling else 0)
    info["ray_tiling"] = f"rtbvh_gpu_scene_set_ray_tiling({WIDTH})" if tiling else "off"

    fps = args.frames_per_step
    rays_per_step = fps * WIDTH * HEIGHT
    cam = W.soup_camera(WIDTH, HEIGHT)
    free_b, _ = torch.cuda.mem_get_info()
    ring = int(max(2, min(args.steps + args.warmup, (free_b * 0.6) // (rays_per_step * 40))))
    if os.environ.get("RTBVH_BENCH_RING"):  # experiment knob: fewer distinct ray buffers (still larger than L2 together)
        ring = max(2, min(ring, int(os.environ["RTBVH_BENCH_RING"])))
    stream = torch.cuda.current_stream().cuda_stream
    d_rays = [torch.empty(rays_per_step * 8, dtype=torch.float32, device="cuda") for _ in range(ring)]
    d_hits = [torch.empty(rays_per_step * 2, dtype=torch.float32, device="cuda") for _ in range(ring)]
    # weak scaling: rank r traces its own frames (global step index = rank * steps + k)
    for b in range(ring):
        g = (rank * (args.steps + args.warmup) + b) * fps
        for f in range(fps):
            off = f * WIDTH * HEIGHT * 8
            api.generate_camera_rays_device(cam, 0, HEIGHT, d_rays[b][off:], jitter_seed=W.SEED_SOUP, frame=g + f,
                                            stream=stream)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # hit records of every rank are gathered over NVLink (all_gather, 8 B per ray), asynchronously so that the
    # transfer of step k overlaps the traversal of step k+1
    gather = "none" if (world == 1 or args.no_gather) else args.gather
    g_out = ([torch.empty(world * rays_per_step * 2, dtype=torch.float32, device="cuda") for _ in range(2)]
             if gather == "nccl" else None)
    works = []
    fused = None
    if gather == "fused":
        # the gather is part of the traversal kernel: every record is stored into all ranks' gather buffers (P2P over
        # NVLink) as its ray finishes; a one-block device barrier closes each step (rtbvh_b200/multigpu.py FusedGather)
        from rtbvh_b200 import multigpu as MG
        try:
            fused = MG.FusedGather(rays_per_step, 8)
            ok = 1
        except Exception as e:  # no cudaIpc / peer access between these ranks: the NCCL gather is the other GPU path
            log(f"[bench] fused gather unavailable on rank {rank}: {e}")
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag[0]) == 0:
            if fused is not None:
                fused.close()
            fused, gather = None, "nccl"
            g_out = [torch.empty(world * rays_per_step * 2, dtype=torch.float32, device="cuda") for _ in range(2)]
            info["gather_note"] = "cudaIpc peer mapping failed on this box: fell back from the fused gather to NCCL all_gather"

    def step(k):
        b = k % ring
        if fused is not None:
            fused.intersect(scene, d_rays[b], rays_per_step, k, d_hits=d_hits[b], stream=stream)
            return
        scene.intersect_device(d_rays[b], rays_per_step, d_hits[b], api.TREE_MBVH, stream=stream)
        if gather == "nccl":
            if len(works) >= 2:
                works[-2].wait()  # the output buffer about to be reused
            works.append(dist.all_gather_into_tensor(g_out[k % 2], d_hits[b], async_op=True))

    for k in range(args.warmup):
        step(k)
    # rank-0-only host work (sub-processes: tens of milliseconds) goes BEFORE the barrier: with a cross-rank step barrier inside
    # the timed region, a rank that starts its loop late makes every other rank's first step wait inside ITS timed region
    sampler = ClockSampler(local)
    nvl0 = nvlink_kib(local) if (rank == 0 and world > 1) else None
    if rank == 0 and os.environ.get("RTBVH_BENCH_NOSAMPLER") != "1":
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)] if os.environ.get("RTBVH_BENCH_STEPTIMES") == "1" else None
    ev0.record()
    for k in range(args.steps):
        step(args.warmup + k)
        if step_ev is not None:
            step_ev[k].record()
    for w in works[-2:]:
        w.wait()
    if fused is not None:  # the last step's gather is complete when its barrier has passed: inside the timed region
        fused.wait(args.warmup + args.steps - 1, stream)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if step_ev is not None:
        ts = [ev0.elapsed_time(e) for e in step_ev]
        log(f"[bench] rank {rank} step end times (ms): " + " ".join(f"{t:.2f}" for t in ts))
    clocks = sampler.stop() if rank == 0 else None
    if nvl0 is not None:
        nvl1 = nvlink_kib(local)
        if nvl1 is not None:
            # the fused gather sends every record of this rank to each of the world-1 peers and receives theirs
            expect = args.steps * rays_per_step * 8 * (world - 1) if gather != "none" else 0
            info["nvlink_gpu0"] = {"tx_bytes": (nvl1[0] - nvl0[0]) * 1024, "rx_bytes": (nvl1[1] - nvl0[1]) * 1024,
                                   "expected_payload_bytes_each_way": expect,
                                   "how": "nvidia-smi nvlink -gt d around the timed region (all links of GPU 0; counters include "
                                          "protocol overhead and the step barriers' flag traffic)"}
    if scene.stack_overflowed():
        raise RuntimeError("traversal stack overflow")
    if fused is not None:
        # the last timed step's gather buffer must hold every rank's records: compare against an NCCL all_gather
        k_last = args.warmup + args.steps - 1
        ref = torch.empty(world * rays_per_step * 2, dtype=torch.float32, device="cuda")
        dist.all_gather_into_tensor(ref, d_hits[k_last % ring])
        got = api.device_view(fused.buffer_ptr(k_last), world * rays_per_step * 8)
        info["fused_gather_equals_all_gather"] = bool(torch.equal(got, ref.view(torch.uint8)))
        if not info["fused_gather_equals_all_gather"]:
            raise RuntimeError("fused gather buffer differs from the NCCL all_gather of the same records")

    # ---- the same frames as RayPacket4 (four x-adjacent pixels, examples/benchmark.rs:135-141): reported next to the headline
    try:
        n_pk = min(3, ring)
        d_pk = []
        for b in range(n_pk):
            r = d_rays[b].view(rays_per_step // 4, 4, 8)
            d_pk.append(torch.stack([r[:, :, 0], r[:, :, 1], r[:, :, 2], r[:, :, 4], r[:, :, 5], r[:, :, 6], r[:, :, 7]],
                                    dim=1).contiguous().view(-1))
        d_pk_hits = torch.empty(rays_per_step * 2, dtype=torch.float32, device="cuda")
        pk_steps = max(3, min(args.steps, 12))
        for k in range(3):
            scene.intersect_packets_device(d_pk[k % n_pk], rays_per_step // 4, d_pk_hits, api.TREE_MBVH, stream=stream)
        barrier()
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for k in range(pk_steps):
            scene.intersect_packets_device(d_pk[k % n_pk], rays_per_step // 4, d_pk_hits, api.TREE_MBVH, stream=stream)
        pe1.record()
        torch.cuda.synchronize()
        info["packet4"] = {"value": world * pk_steps * rays_per_step / pe0.elapsed_time(pe1) / 1e3, "unit": "Mrays/s",
                           "steps": pk_steps, "kernel": "trace_mbvh_packet_lane_kernel<closest> (one lane per RayPacket4)",
                           "note": "rank-local device time (not max over ranks); packets of 4 x-adjacent pixels of the same frames"}
        del d_pk, d_pk_hits
    except Exception as e:  # an extra: never costs the bench line
        info["packet4"] = {"error": f"{type(e).__name__}: {e}"}

    # ---- e2e: host buffers through rtbvh_gpu_intersect -------------------------------------------
    n_host = min(3, ring)
    h_rays = [torch.empty(rays_per_step * 8, dtype=torch.float32).pin_memory() for _ in range(n_host)]
    h_hits = [torch.empty(rays_per_step * 2, dtype=torch.float32).pin_memory() for _ in range(n_host)]
    for b in range(n_host):
        h_rays[b].copy_(d_rays[b])
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for k in range(min(2, args.warmup)):
        scene.intersect_ptr(h_rays[k % n_host].data_ptr(), rays_per_step, h_hits[k % n_host].data_ptr(), api.TREE_MBVH)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        scene.intersect_ptr(h_rays[k % n_host].data_ptr(), rays_per_step, h_hits[k % n_host].data_ptr(), api.TREE_MBVH)
    torch.cuda.synchronize()
    e2e_sync_ms = (time.perf_counter() - t0) * 1e3
    # the same steps through the submit / wait flavour of the call, double-buffered the way a renderer streams frames:
    # step k is submitted while step k-1 is in flight; before a host buffer pair is reused its step has been waited for
    barrier()
    tickets = []
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        if k >= 2:
            scene.wait(tickets[k - 2])
        tickets.append(scene.intersect_async(h_rays[k % n_host].data_ptr(), rays_per_step, h_hits[k % n_host].data_ptr(),
                                             api.TREE_MBVH))
    scene.wait(0)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    same_async = bool(torch.equal(h_hits[(e2e_steps - 1) % n_host].view(torch.int32),
                                  d_hits[(e2e_steps - 1) % n_host].cpu().view(torch.int32)))
    e2e_rtray_ms = e2e_ms
    # the same steps with the rays in the reference FFI's argument shape: origins[3n] + directions[3n] (24 B per ray across
    # PCIe instead of 32; t_min = 1e-4 and t = 1e34 are the constants Ray::new sets, src/ray.rs:166-182)
    h_o = [torch.empty(rays_per_step * 3, dtype=torch.float32).pin_memory() for _ in range(n_host)]
    h_d = [torch.empty(rays_per_step * 3, dtype=torch.float32).pin_memory() for _ in range(n_host)]
    for b in range(n_host):
        r8 = h_rays[b].view(rays_per_step, 8)
        h_o[b].view(rays_per_step, 3).copy_(r8[:, 0:3])
        h_d[b].view(rays_per_step, 3).copy_(r8[:, 4:7])
        h_hits[b].zero_()
    for k in range(min(2, args.warmup)):
        scene.intersect_od_ptr(h_o[k % n_host].data_ptr(), h_d[k % n_host].data_ptr(), rays_per_step, h_hits[k % n_host].data_ptr(),
                               api.TREE_MBVH, 1e-4, 1e34)
    barrier()
    tickets = []
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        if k >= 2:
            scene.wait(tickets[k - 2])
        tickets.append(scene.intersect_od_async(h_o[k % n_host].data_ptr(), h_d[k % n_host].data_ptr(), rays_per_step,
                                                h_hits[k % n_host].data_ptr(), api.TREE_MBVH, 1e-4, 1e34))
    scene.wait(0)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    same_od = bool(torch.equal(h_hits[(e2e_steps - 1) % n_host].view(torch.int32),
                               d_hits[(e2e_steps - 1) % n_host].cpu().view(torch.int32)))
    # the host-buffer path must agree with the resident path on the same rays
    same = bool(torch.equal(h_hits[0].view(torch.int32), d_hits[0].cpu().view(torch.int32)))
    # the same steps with the primary rays generated on the device (rtbvh_gpu_intersect_camera_async): only the records cross
    # PCIe.  Step k re-creates ring slot k's frames (same seed, same frame numbers), so its records must equal that slot's.
    def cam_submit(k):
        b = k % n_host
        g = (rank * (args.steps + args.warmup) + b) * fps
        return scene.intersect_camera_async(cam, fps, h_hits[b].data_ptr(), api.TREE_MBVH, jitter_seed=W.SEED_SOUP, first_frame=g)
    for b in range(n_host):
        h_hits[b].zero_()
    scene.wait(cam_submit(0))
    barrier()
    tickets = []
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        if k >= 2:
            scene.wait(tickets[k - 2])
        tickets.append(cam_submit(k))
    scene.wait(0)
    e2e_cam_ms = (time.perf_counter() - t0) * 1e3
    same_cam = bool(torch.equal(h_hits[(e2e_steps - 1) % n_host].view(torch.int32),
                                d_hits[(e2e_steps - 1) % n_host].cpu().view(torch.int32)))

    t = torch.tensor([ms, e2e_ms, e2e_sync_ms, e2e_rtray_ms, e2e_cam_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, e2e_sync_ms, e2e_rtray_ms, e2e_cam_ms = (float(x) for x in t)
    os.sched_setaffinity(0, all_cpus)  # the CPU baseline below uses every host core again

    if rank == 0:
        total_rays = world * args.steps * rays_per_step
        value = total_rays / ms / 1e3
        e2e = world * e2e_steps * rays_per_step / e2e_ms / 1e3
        # ---- CPU baseline + algorithmic bytes from the instrumented oracle on a sample of step 0's rays
        cpu = None
        bytes_per_ray, nm, npr = None, None, None
        if not args.no_cpu:
            from oracle import oracle as O
            # bounded sample: the rays of up to 16 timed steps (128 M rays, ~10 s of CPU work on 16 cores)
            n_batches = min(ring, 16) if world == 1 else 1
            host_rays = np.concatenate([d_rays[b].cpu().numpy().view(api.RAY_DTYPE).reshape(-1) for b in range(n_batches)])
            otree = O.Mbvh(mbvh.nodes.copy(), mbvh.indices.copy())
            if world == 1:
                rate, n_s, per_ray, max_stack, threads = cpu_sample_rate(O, otree, tris, host_rays)
            else:
                # N > 1: the CPU baseline is an N = 1 figure (the other ranks would idle in the barrier while rank 0 times it);
                # only the oracle's work counters are taken here, on 1 M rays, for the roofline's algorithmic bytes
                threads = host_threads()
                nc = min(len(host_rays), 1_000_000)
                _, _, cnt = O.trace(otree, tris, host_rays[:nc], threads=threads, counters=True)
                rate, n_s, per_ray, max_stack = None, nc, {k: cnt[k] / nc for k in ("node_visits", "prim_tests")}, cnt["max_stack"]
            nm, npr = per_ray["node_visits"], per_ray["prim_tests"]
            bytes_per_ray = 32 + 8 + 128 * nm + 40 * npr
            if rate is not None:
                cpu = {"value": rate, "unit": "Mrays/s", "cores": threads, "kind": "port",
                       "sample": f"first {n_s} rays of the timed workload (the very rays the GPU steps trace), Mbvh single-ray "
                                 f"closest hit, OpenMP dynamic chunks of 1000",
                       "single_thread_value": cpu_single_thread_rate(O, otree, tris, host_rays)}
                tb = oracle_threaded_build_ms(O, tris)
                cpu["binned_sah_build_ms_per_mtri"] = tb / (N_TRIS / 1e6) if tb else None
                cpu["binned_sah_build_note"] = (f"the port's binned-SAH builder on this scene with the reference's threaded scheduling "
                                                f"(subtrees > 1024 primitives on other threads, src/utils.rs:189-289), {threads} threads")
            # parity spot check inside the bench: oracle vs GPU on the sample
            want, _, _ = O.trace(otree, tris, host_rays[:200_000], threads=threads)
            got = d_hits[0][: 200_000 * 2].cpu().numpy().view(api.HIT_DTYPE).reshape(-1)
            info["parity_sample_bit_exact"] = bool(np.array_equal(want, got))
            info["oracle_max_stack"] = int(max_stack)
        peak, peak_src = measured_peak_gbs()
        roof = None
        if bytes_per_ray is not None:
            launch_ms = ms / args.steps
            achieved = bytes_per_ray * rays_per_step / (launch_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic(), "peak_source": peak_src,
                    "bytes_per_ray": bytes_per_ray, "node_visits_per_ray": nm, "tri_tests_per_ray": npr,
                    "kernel": "trace_single_persistent_kernel<MBVH, closest>", "launch_ms": launch_ms}
        out = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "soup-1Mi-tris binned-SAH Mbvh primary rays closest-hit (BASELINE configs[1])",
                       "rays_per_step": rays_per_step, "frames_per_step": fps, "ray_ring_batches": ring,
                       "l2": "inputs larger than L2 (256 MB rays per step, distinct buffers)",
                       "trace_mode": os.environ.get("RTBVH_TRACE_MODE", "persistent"),
                       "sharding": ("rays sharded by frame range, tree replicated; hit records all_gathered over NVLink by "
                                    f"NCCL ({world * rays_per_step * 8} B per step per GPU, overlapped)" if gather == "nccl" else
                                    "rays sharded by frame range, tree replicated; gather fused into the traversal kernel: "
                                    f"P2P stores into every rank's gather buffer ({world * rays_per_step * 8} B per step per "
                                    "GPU) + one-block device barrier per step, no NCCL on the data path" if gather == "fused"
                                    else "single GPU" if world == 1 else "rays sharded, tree replicated, no gather"), **info},
            "clocks": clocks,
            # own kernels per rank inside the timed region: one traversal launch per step (+ one step-barrier kernel with the fused gather)
            "gpu_launches": args.steps * (2 if gather == "fused" else 1),
            "e2e": {"value": e2e, "unit": "Mrays/s", "h2d_bytes_per_step": rays_per_step * 24,
                    "d2h_bytes_per_step": rays_per_step * 8, "steps": e2e_steps,
                    "host_equals_resident": same and same_async and same_od and same_cam,
                    "camera_value": world * e2e_steps * rays_per_step / e2e_cam_ms / 1e3,
                    "camera_call": "rtbvh_gpu_intersect_camera_async + rtbvh_gpu_wait: primary rays generated on the device from the "
                                   "camera (64 B of parameters per step), host hit records out every step, two steps in flight; "
                                   f"h2d 0, d2h {rays_per_step * 8} B per step",
                    "call": "rtbvh_gpu_intersect_od_async + rtbvh_gpu_wait: pinned host origins[3n] + directions[3n] in (the "
                            "reference FFI's argument shape, 24 B per ray), host hit records out every step, two steps in "
                            "flight (double-buffered)",
                    "rtray_async_value": world * e2e_steps * rays_per_step / e2e_rtray_ms / 1e3,
                    "rtray_async": "rtbvh_gpu_intersect_async + rtbvh_gpu_wait with 32-byte RTRay records (256 MB H2D per step)",
                    "sync_call_value": world * e2e_steps * rays_per_step / e2e_sync_ms / 1e3,
                    "sync_call": "rtbvh_gpu_intersect with RTRay records (blocking: the pipeline fills and drains inside every call)"},
            "roofline": roof, "cpu_baseline": cpu,
        }
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    if fused is not None:
        fused.close()
    scene.free()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def build_trees(api, tris, info, rank):
    """Tree for the bench: built on the GPU (rtbvh_gpu_create_bvh_triangles -> create_mbvh), the product path.  (Reference-built
    trees uploaded unchanged are measured by scripts/matrix.py and scripts/config5.py, not here.)"""
    mtri = len(tris) / 1e6
    api.build_triangles(tris, api.BINNED_SAH, 1).free()  # warm-up: module load, memory pool growth
    dev, tot, bvh = [], [], None
    for _ in range(3):
        if bvh is not None:
            bvh.free()
        bvh = api.build_triangles(tris, api.BINNED_SAH, 1)
        st = api.last_build_stats()
        dev.append(st["device_ms"])
        tot.append(st["total_ms"])
    # collapse through create_mbvh right behind create_bvh, like a caller of the reference's ABI: median of 3 (the first call
    # also pins the host mirror's block, which later calls recycle — same warm-up rule as for the build above)
    cdev, ctot, mbvh = [], [], None
    for _ in range(4):
        if mbvh is not None:
            mbvh.free()
        mbvh = api.Mbvh.construct(bvh)
        cst = api.last_build_stats()
        cdev.append(cst["device_ms"])
        ctot.append(cst["total_ms"])
    cst = {"device_ms": float(np.median(cdev[1:])), "total_ms": float(np.median(ctot[1:])), "first_call_total_ms": ctot[0]}
    # the same build straight into a device-resident scene (no host mirror): wall clock of the whole call from host
    # vertices, i.e. H2D of 36 MB of vertices + prims + binned SAH + collapse + triangle records
    api.Scene.build(tris, api.BINNED_SAH, 1, mbvh=True).free()
    res_wall, res_dev = [], []
    for _ in range(3):
        t0 = time.perf_counter()
        rs = api.Scene.build(tris, api.BINNED_SAH, 1, mbvh=True)
        res_wall.append((time.perf_counter() - t0) * 1e3)
        res_dev.append(api.last_build_stats()["device_ms"])
        rs.free()
    # the build's own roofline: 148 + 60 * D-bar algorithmic bytes per triangle, D-bar = mean
    # leaf depth (per primitive) of the tree just built, over the measured device time of one build
    try:
        ds = W.leaf_depth_stats(bvh.nodes)
        bpt = W.binned_sah_bytes_per_tri(ds["mean_leaf_depth_per_prim"])
        peak, peak_src = measured_peak_gbs()
        ach = bpt * len(tris) / (float(np.median(dev)) * 1e-3) / 1e9
        build_roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                          "peak_source": peak_src, "bytes_per_tri": bpt,
                          "mean_leaf_depth": ds["mean_leaf_depth_per_prim"], "max_depth": ds["max_depth"],
                          "kernels": "all kernels of one binned-SAH build (level loop + small-subtree kernel)"}
    except Exception as e:  # a statistics failure must not cost the bench line
        build_roofline = {"error": f"{type(e).__name__}: {e}"}
    info.update(tree="gpu-built: rtbvh_gpu_create_bvh_triangles(BinnedSAH) + create_mbvh",
                build={"binned_sah_ms_per_mtri": float(np.median(dev)) / mtri,
                       "roofline": build_roofline,
                       "binned_sah_ms_per_mtri_incl_h2d_d2h": float(np.median(tot)) / mtri,
                       "binned_sah_device_ms_runs": dev, "collapse_device_ms": cst["device_ms"],
                       "collapse_ms_incl_h2d_d2h": cst["total_ms"], "collapse_first_call_ms_incl_h2d_d2h": cst["first_call_total_ms"],
                       "bvh_nodes": int(bvh.rt.node_count),
                       "mbvh_nodes": int(mbvh.rt.node_count),
                       "resident_scene_build_wall_ms": float(np.median(res_wall)),
                       "resident_scene_build_device_ms": float(np.median(res_dev)),
                       "timing": "CUDA events around the builder kernels, triangles resident -> tree resident; median of 3"})
    return bvh, mbvh, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 125 (config 2: the config's 1 B rays), 40 / 10 / 20 / 20 for configs 1 / 3 / 4 / 5")
    ap.add_argument("--warmup", type=int, default=None, help="default 5 (config 3: 3)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json config (1-based): 2 = the config the metric is quoted on (default); 1, 3, 4, 5 live in "
                         "bench_configs.py with the same JSON shape")
    ap.add_argument("--frames-per-step", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=40)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / roofline sample (profiling runs)")
    ap.add_argument("--no-gather", action="store_true", help="N > 1: do not gather the hit records")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"],
                    help="N > 1: fused = P2P stores from inside the traversal kernel; nccl = all_gather on a side stream")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = {1: 40, 2: 125, 3: 10, 4: 20, 5: 20}[args.config]
    if args.warmup is None:
        args.warmup = 3 if args.config == 3 else 5
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself one rank per GPU (the driver uses torchrun directly)
        import socket
        sock = socket.socket()
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
        sock.close()
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port", str(port),
                                   os.path.abspath(__file__)] + sys.argv[1:])
    if args.config != 2:
        import bench_configs
        bench_configs.run(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
