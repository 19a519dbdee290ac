"""bench_configs.py — BASELINE.json configs 1, 3, 4 and 5 behind `bench.py --config N`, with the JSON shape of the default
(config 2) line: metric / value / unit / e2e / roofline / cpu_baseline / clocks / gpu_launches.

  1  teapot.obj (6 320 triangles), binned-SAH Bvh + Mbvh, the camera of examples/benchmark.rs, single rays and RayPacket4.
     The only config with published numbers (reference README.md:31-41, Ryzen 5950X): `vs_baseline` = Mbvh single rays over
     125.45 Mrays/s; every flavour and the build carry their own ratio in config.flavours / config.build.
  3  10 M-triangle heightfield mesh: locally-ordered-clustering build + Mbvh collapse; metric = build ms per Mtri (lower is
     better), SAH cost reported.
  4  30 M-triangle instanced scene, incoherent any-hit shadow rays, rays sharded over the ranks (tree replicated).
  5  spatial-split SAH tree built by the CPU restatement of the reference builder, uploaded unchanged, one diffuse bounce ray
     per primary hit, closest hit, rays sharded over the ranks.

`--impl reference` times the CPU oracle port of the same path on all host cores on a bounded sample per step (the Rust
reference cannot be built in this image).  oracle/ is imported only as the checker, for `cpu_baseline`, for the reference arm
and — config 5 only — as the stand-in for "the reference built this tree" (the product never builds spatial-split trees)."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

from bench_common import ClockSampler, host_threads, log, measured_peak_gbs
from rtbvh_b200 import workloads as W

PUBLISHED = {  # /root/reference/README.md:31-41 (binned-SAH teapot rows), AMD Ryzen 9 5950X, 32 threads
    "bvh_single": 77.38225, "bvh_packet4": 283.20187, "mbvh_single": 125.44958, "mbvh_packet4": 368.95465,
    "binned_sah_build_ms": 3.81, "mbvh_collapse_ms": 1.156,
}


class Ctx:
    """Process-group plumbing shared by the GPU arms: one process per GPU, NCCL only for barriers / max-over-ranks."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        from rtbvh_b200 import api
        self.torch, self.dist, self.api = torch, dist, api
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        sys.stdout.flush()
        self.real_stdout = os.dup(1)  # stdout carries exactly ONE line (the JSON)
        os.dup2(2, 1)
        if api.device_count() == 0:
            raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(self.local)
        api.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.stream = torch.cuda.current_stream().cuda_stream

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def min_over_ranks(self, v):
        t = self.torch.tensor([v], dtype=self.torch.int64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return int(t[0])

    def timed(self, step, steps, warmup):
        """warmup untimed steps, then exactly `steps` steps between CUDA events on the launching stream, barrier +
        synchronize on both sides, max over ranks.  Returns (ms, clocks)."""
        torch = self.torch
        for k in range(warmup):
            step(k)
        sampler = ClockSampler(self.local)
        if self.rank == 0:
            sampler.start()  # before the barrier: the sub-process start must not skew rank 0's loop against the other ranks'
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            step(warmup + k)
        e1.record()
        self.barrier()
        ms = self.max_over_ranks(e0.elapsed_time(e1))[0]
        return ms, (sampler.stop() if self.rank == 0 else None)

    def emit(self, out):
        if self.rank == 0:
            sys.stdout.flush()
            os.dup2(self.real_stdout, 1)
            print(json.dumps(out), flush=True)
            os.dup2(2, 1)

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def emit_reference(args, metric, unit, value, ms_per_step, hib, workload, sample, extra=None, cores=None):
    out = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": hib, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload,
                      "note": "reference is Rust and cannot be built in this image (no rustc/cargo): this is the C++ oracle "
                              "port of its path on the host cores"},
           "cpu_baseline": {"value": value, "unit": unit, "cores": cores or host_threads(), "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if extra:
        out.update(extra)
    print(json.dumps(out), flush=True)


def traversal_roofline(bytes_per_ray, rays_per_launch, launch_ms, kernel, nv, nt, traffic=None):
    peak, src = measured_peak_gbs()
    ach = bytes_per_ray * rays_per_launch / (launch_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
            "peak_source": src, "bytes_per_ray": bytes_per_ray, "node_visits_per_ray": nv, "tri_tests_per_ray": nt,
            "kernel": kernel, "launch_ms": launch_ms}


def traffic_for(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of this config's dominant kernel, from the committed ncu
    capture (profiles/ncu_traffic.json, key per config); None when no capture of this workload exists."""
    try:
        return float(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")))[key])
    except Exception:
        return None


def host_e2e_async(ctx, scene, submit, n_bufs, steps):
    """Host-buffer flavour: `submit(k)` enqueues step k from pinned host buffers (returns a ticket), two steps in flight."""
    ctx.barrier()
    tickets = []
    t0 = time.perf_counter()
    for k in range(steps):
        if k >= 2:
            scene.wait(tickets[k - 2])
        tickets.append(submit(k % n_bufs))
    scene.wait(0)
    return (time.perf_counter() - t0) * 1e3


# ======================================================================================================================
# config 1: teapot, examples/benchmark.rs
# ======================================================================================================================
C1_METRIC = "Mrays/s closest-hit (teapot 6320 tris, binned-SAH Mbvh, examples/benchmark.rs camera, single rays)"
C1_WORKLOAD = "teapot.obj binned-SAH Bvh + Mbvh, 1000x1000 benchmark camera frames, single + packet4 (BASELINE configs[0])"


def config1_reference(args):
    from oracle import oracle as O
    tris = W.teapot()
    aabbs, centers = O.prims_from_triangles(tris)
    t0 = time.perf_counter()
    rc, bvh = O.build(O.BINNED_SAH, aabbs, centers, 1)
    build_ms = (time.perf_counter() - t0) * 1e3
    m = bvh.collapse()
    rays = W.camera_rays(W.benchmark_camera(1000, 1000))  # one frame per step; the reference traces 100 identical frames
    threads = host_threads()
    for _ in range(args.warmup):
        O.trace(m, tris, rays[:200_000], threads=threads)
    ms_tot = 0.0
    for _ in range(args.steps):
        ms_tot += O.trace(m, tris, rays, threads=threads)[1]
    v = args.steps * len(rays) / ms_tot / 1e3
    flav = {}
    pk = W.pack4(rays)
    for name, fn in (("bvh_single", lambda: O.trace(bvh, tris, rays, threads=threads)[1]),
                     ("bvh_packet4", lambda: O.trace_packets(bvh, tris, pk, threads=threads)[1]),
                     ("mbvh_packet4", lambda: O.trace_packets(m, tris, pk, threads=threads)[1])):
        flav[name] = len(rays) / fn() / 1e3
    flav["mbvh_single"] = v
    emit_reference(args, C1_METRIC, "Mrays/s", v, ms_tot / args.steps, True, C1_WORKLOAD,
                   f"one 1000x1000 benchmark-camera frame per step ({args.steps} steps), Mbvh single-ray closest hit, OpenMP "
                   f"dynamic chunks of 1000", extra={"flavours_mrays": flav, "oracle_build_ms": build_ms,
                                                     "oracle_collapse_ms": m.collapse_ms, "published_5950x": PUBLISHED})


def config1_gpu(args):
    ctx = Ctx()
    torch, api = ctx.torch, ctx.api
    tris = W.teapot()
    # ---- build through the drop-in ABI (host triangles in, host-mirrored trees out) and resident ----------------------
    api.build_triangles(tris, api.BINNED_SAH, 1).free()
    dev, tot, bvh = [], [], None
    for _ in range(5):
        if bvh is not None:
            bvh.free()
        bvh = api.build_triangles(tris, api.BINNED_SAH, 1)
        st = api.last_build_stats()
        dev.append(st["device_ms"])
        tot.append(st["total_ms"])
    mbvh = api.Mbvh.construct(bvh)
    cst = api.last_build_stats()
    scene = api.Scene(tris, bvh=bvh, mbvh=mbvh)
    build = {"binned_sah_device_ms": float(np.median(dev)), "binned_sah_ms_incl_h2d_d2h": float(np.median(tot)),
             "collapse_device_ms": cst["device_ms"], "collapse_ms_incl_h2d_d2h": cst["total_ms"],
             "published_binned_sah_ms_5950x": PUBLISHED["binned_sah_build_ms"],
             "vs_baseline_build": PUBLISHED["binned_sah_build_ms"] / float(np.median(tot)),
             "note": "vs_baseline_build = README's 3.81 ms over the whole create_bvh-style call incl. copies (higher = faster here)",
             "bvh_nodes": int(bvh.rt.node_count), "mbvh_nodes": int(mbvh.rt.node_count)}
    # ---- rays: the benchmark's frame, `fps` identical frames per step in distinct buffers (ring > L2) ------------------
    fps = args.frames_per_step
    frame = W.camera_rays(W.benchmark_camera(1000, 1000))
    n = fps * len(frame)
    ring = 3
    h_frame = torch.from_numpy(frame.view(np.float32).reshape(-1).copy())
    d_rays = [h_frame.repeat(fps).cuda() for _ in range(ring)]
    pk = W.pack4(frame)
    h_pk = torch.from_numpy(pk.view(np.float32).reshape(-1).copy())
    d_pk = [h_pk.repeat(fps).cuda() for _ in range(ring)]
    d_hits = [torch.empty(n * 2, dtype=torch.float32, device="cuda") for _ in range(ring)]
    flavours = {
        "mbvh_single": lambda k: scene.intersect_device(d_rays[k % ring], n, d_hits[k % ring], api.TREE_MBVH, stream=ctx.stream),
        "mbvh_packet4": lambda k: scene.intersect_packets_device(d_pk[k % ring], n // 4, d_hits[k % ring], api.TREE_MBVH,
                                                                 stream=ctx.stream),
        "bvh_single": lambda k: scene.intersect_device(d_rays[k % ring], n, d_hits[k % ring], api.TREE_BVH, stream=ctx.stream),
        "bvh_packet4": lambda k: scene.intersect_packets_device(d_pk[k % ring], n // 4, d_hits[k % ring], api.TREE_BVH,
                                                                stream=ctx.stream),
    }
    flav = {}
    for name in ("bvh_single", "bvh_packet4", "mbvh_packet4"):
        ms_f, _ = ctx.timed(flavours[name], max(3, args.steps // 4), 3)
        v = ctx.world * max(3, args.steps // 4) * n / ms_f / 1e3
        flav[name] = {"value": v, "published_5950x": PUBLISHED[name], "vs_baseline": v / PUBLISHED[name]}
    ms, clocks = ctx.timed(flavours["mbvh_single"], args.steps, args.warmup)
    value = ctx.world * args.steps * n / ms / 1e3
    flav["mbvh_single"] = {"value": value, "published_5950x": PUBLISHED["mbvh_single"], "vs_baseline": value / PUBLISHED["mbvh_single"]}
    # ---- e2e: pinned host origins + directions in, host records out, two steps in flight --------------------------------
    r8 = h_frame.view(-1, 8).repeat(fps, 1)
    h_o = [r8[:, 0:3].contiguous().view(-1).pin_memory() for _ in range(2)]
    h_d = [r8[:, 4:7].contiguous().view(-1).pin_memory() for _ in range(2)]
    h_h = [torch.zeros(n * 2, dtype=torch.float32).pin_memory() for _ in range(2)]
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    sub = lambda b: scene.intersect_od_async(h_o[b].data_ptr(), h_d[b].data_ptr(), n, h_h[b].data_ptr(), api.TREE_MBVH, 1e-4, 1e34)
    host_e2e_async(ctx, scene, sub, 2, 2)
    e2e_ms = ctx.max_over_ranks(host_e2e_async(ctx, scene, sub, 2, e2e_steps))[0]
    same = bool(torch.equal(h_h[0].view(torch.int32), d_hits[0].cpu().view(torch.int32)))
    if scene.stack_overflowed():
        raise RuntimeError("traversal stack overflow")
    if ctx.rank == 0:
        cpu, roof, info = None, None, {}
        if not args.no_cpu:
            from oracle import oracle as O
            threads = host_threads()
            otree = O.Mbvh(mbvh.nodes.copy(), mbvh.indices.copy())
            want, cms, _ = O.trace(otree, tris, frame, threads=threads)
            _, _, cnt = O.trace(otree, tris, frame, threads=threads, counters=True)
            got = d_hits[0][: len(frame) * 2].cpu().numpy().view(api.HIT_DTYPE).reshape(-1)
            info["parity_sample_bit_exact"] = bool(np.array_equal(want, got))
            nv, nt = cnt["node_visits"] / len(frame), cnt["prim_tests"] / len(frame)
            cpu = {"value": len(frame) / cms / 1e3, "unit": "Mrays/s", "cores": threads, "kind": "port",
                   "sample": "one benchmark-camera frame (1 M rays), Mbvh single-ray closest hit, OpenMP dynamic chunks of 1000"}
            roof = traversal_roofline(32 + 8 + 128 * nv + 40 * nt, n, ms / args.steps, "trace_single_persistent_kernel<MBVH, closest>",
                                      nv, nt, traffic_for("config1_dram_bytes_per_launch"))
        ctx.emit({"metric": C1_METRIC, "value": value, "unit": "Mrays/s", "n_gpus": ctx.world, "steps": args.steps,
                  "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                  "vs_baseline": value / PUBLISHED["mbvh_single"], "dtype": "f32", "data": "synthetic",
                  "config": {"workload": C1_WORKLOAD, "rays_per_step": n, "frames_per_step": fps,
                             "l2": f"{ring} distinct ray buffers of {n * 32 >> 20} MB each (larger than L2 together with the records)",
                             "vs_baseline_source": "reference README.md:38-39 (Mbvh binned SAH, 32 threads, Ryzen 9 5950X): 125.45 Mrays/s",
                             "flavours": flav, "build": build, "sharding": "tree replicated, every rank traces its own frames, no gather",
                             **info},
                  "clocks": clocks, "gpu_launches": args.steps,
                  "e2e": {"value": ctx.world * e2e_steps * n / e2e_ms / 1e3, "unit": "Mrays/s", "h2d_bytes_per_step": n * 24,
                          "d2h_bytes_per_step": n * 8, "steps": e2e_steps, "host_equals_resident": same,
                          "call": "rtbvh_gpu_intersect_od_async + rtbvh_gpu_wait (pinned origins[3n] + directions[3n] in, hit records out)"},
                  "roofline": roof, "cpu_baseline": cpu})
    scene.free()
    ctx.close()


# ======================================================================================================================
# config 3: 10 M-triangle mesh, LOCB build + Mbvh collapse
# ======================================================================================================================
C3_METRIC = "LOCB build + Mbvh collapse, ms per Mtri (10M-triangle heightfield mesh)"
C3_WORKLOAD = "heightfield 2237x2237x2 = 10 008 338 triangles, locally-ordered-clustering Bvh + Mbvh collapse (BASELINE configs[2])"


def c3_mesh(side=2237):
    return W.heightfield(side, side)


def config3_reference(args):
    from oracle import oracle as O
    side = 500  # bounded sample: a 500 000-triangle heightfield of the same generator per step (the LOCB port is serial but for the sort)
    tris = c3_mesh(side)
    aabbs, centers = O.prims_from_triangles(tris)
    mtri = len(tris) / 1e6

    def step():
        t0 = time.perf_counter()
        rc, b = O.build(O.LOCB, aabbs, centers, 1, parallel=True)
        m = b.collapse()
        return (time.perf_counter() - t0) * 1e3, b

    for _ in range(min(args.warmup, 1)):
        step()
    tot, b = 0.0, None
    for _ in range(args.steps):
        ms, b = step()
        tot += ms
    v = tot / args.steps / mtri
    emit_reference(args, C3_METRIC, "ms/Mtri", v, tot / args.steps, False, C3_WORKLOAD,
                   f"{len(tris)}-triangle heightfield of the same generator per step (1/20 of the config), LOCB + collapse",
                   extra={"sah_sample": b.sah_cost(), "kappa_sample": b.kappa})


def config3_gpu(args):
    ctx = Ctx()
    torch, api = ctx.torch, ctx.api
    tris = c3_mesh()
    n_tris = len(tris)
    mtri = n_tris / 1e6
    h_verts = torch.from_numpy(tris.reshape(-1)).pin_memory()
    d_verts = h_verts.cuda()
    kind = api.LOCALLY_ORDERED_CLUSTERED
    holder = {}

    def step(k):
        if "s" in holder:
            holder["s"].free()
        holder["s"] = api.Scene.build(d_verts, kind, 1, mbvh=True, n_tris=n_tris, vertex_stride=12)

    ms, clocks = ctx.timed(step, args.steps, args.warmup)
    st = api.last_build_stats()
    value = ms / args.steps / mtri
    # e2e: the same build from HOST vertices through rtbvh_gpu_scene_build (H2D inside), tree sizes read back
    e2e_steps = max(1, min(args.steps, args.e2e_steps, 5))
    hv = h_verts.numpy().reshape(-1, 3, 3)

    def host_step():
        s = api.Scene.build(hv, kind, 1, mbvh=True)
        nn = s.n_nodes
        s.free()
        return nn

    host_step()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_step()
    e2e_ms = ctx.max_over_ranks((time.perf_counter() - t0) * 1e3)[0]
    # the reference's own ABI (create_bvh-shaped call with host mirrors + create_mbvh)
    # (first call: the pooled page-locked host mirrors are pinned — once per size class and process; second call: recycled)
    b = api.build_triangles(tris, kind, 1)
    abi_first = api.last_build_stats()
    m = api.Mbvh.construct(b)
    abi_c_first = api.last_build_stats()
    m.free()
    b.free()
    b = api.build_triangles(tris, kind, 1)
    abi = api.last_build_stats()
    m = api.Mbvh.construct(b)
    abi_c = api.last_build_stats()
    if ctx.rank == 0:
        from oracle import oracle as O
        info = {"bvh_nodes": int(b.rt.node_count), "mbvh_nodes": int(m.rt.node_count), "locb_iterations": int(st["iterations"]),
                "sah": O.Bvh(b.nodes, b.indices).sah_cost(),
                "reference_abi": {"first_call_create_bvh_ms_per_mtri_incl_h2d_d2h": abi_first["total_ms"] / mtri,
                                  "first_call_create_mbvh_ms_incl_h2d_d2h": abi_c_first["total_ms"],
                                  "create_bvh_ms_per_mtri_incl_h2d_d2h": abi["total_ms"] / mtri, "create_bvh_device_ms_per_mtri": abi["device_ms"] / mtri,
                                  "create_mbvh_ms_incl_h2d_d2h": abi_c["total_ms"], "create_mbvh_device_ms": abi_c["device_ms"]}}
        cpu, roof = None, None
        if not args.no_cpu:
            side = 707  # ~1 M triangles: bounded CPU sample of the same generator
            st_tris = c3_mesh(side)
            aabbs, centers = O.prims_from_triangles(st_tris)
            t0 = time.perf_counter()
            rc, ob = O.build(O.LOCB, aabbs, centers, 1, parallel=True)
            om = ob.collapse()
            cms = (time.perf_counter() - t0) * 1e3
            gb = api.build_triangles(st_tris, kind, 1)
            gm = api.Mbvh.construct(gb)
            info["parity_sample_byte_identical"] = bool(np.array_equal(gb.nodes.view(np.uint8), ob.nodes.view(np.uint8)) and
                                                        np.array_equal(gb.indices, ob.indices) and
                                                        np.array_equal(gm.nodes.view(np.uint8), om.nodes.view(np.uint8)))
            info["sah_sample"] = {"gpu": O.Bvh(gb.nodes, gb.indices).sah_cost(), "oracle": ob.sah_cost()}
            cpu = {"value": cms / (len(st_tris) / 1e6), "unit": "ms/Mtri", "cores": host_threads(), "kind": "port",
                   "sample": f"{len(st_tris)}-triangle heightfield of the same generator (1/10 of the config), LOCB + collapse; the port "
                             f"is serial except the Morton sort, like the reference"}
            kappa = ob.kappa
            bpt = 196 + 92 * kappa + 128
            peak, src = measured_peak_gbs()
            ach = bpt * n_tris / (ms / args.steps * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic_for("config3_dram_bytes_per_build"),
                    "peak_source": src, "bytes_per_tri": bpt, "kappa": kappa,
                    "kernel": "all kernels of one LOCB build + collapse (morton, radix sort, locb_nn/flag/scan/write per iteration, collapse)",
                    "launch_ms": ms / args.steps,
                    "note": "kappa (sum of cluster counts over the iterations / N) from the 1 M-triangle sample of the same generator"}
        ctx.emit({"metric": C3_METRIC, "value": value, "unit": "ms/Mtri", "n_gpus": ctx.world, "steps": args.steps,
                  "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": False, "scaling": "weak",
                  "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                  "config": {"workload": C3_WORKLOAD, "triangles": n_tris, "step": "one rtbvh_gpu_scene_build_device (LOCB + collapse + "
                             "triangle records) from device vertices into a device-resident scene", "l2": "360 MB of vertices + 1.3 GB of "
                             "nodes per build: far larger than L2", "sharding": "replicas only: every rank builds its own copy", **info},
                  "clocks": clocks, "gpu_launches": int(args.steps * (12 + 5 * max(1, int(st["iterations"])))),
                  "e2e": {"value": e2e_ms / e2e_steps / mtri, "unit": "ms/Mtri", "h2d_bytes_per_step": n_tris * 36, "d2h_bytes_per_step": 8,
                          "steps": e2e_steps, "call": "rtbvh_gpu_scene_build from pinned host vertices (H2D inside), tree sizes read back"},
                  "roofline": roof, "cpu_baseline": cpu})
    if "s" in holder:
        holder["s"].free()
    ctx.close()


# ======================================================================================================================
# config 4: 30 M-triangle scene, incoherent any-hit shadow rays
# ======================================================================================================================
C4_METRIC = "Mrays/s any-hit (30M-triangle instanced scene, binned-SAH Mbvh, incoherent shadow rays)"
C4_WORKLOAD = "30 instanced displaced spheres + room = 30 033 370 triangles, incoherent shadow rays, any hit (BASELINE configs[3])"


def config4_reference(args):
    from oracle import oracle as O
    inst = 2  # bounded sample: the same generator at 2 instances (2 M triangles): the CPU build of 30 M takes ~90 s per run
    tris = W.instanced_scene(inst)
    aabbs, centers = O.prims_from_triangles(tris)
    rc, bvh = O.build(O.BINNED_SAH, aabbs, centers, 1)
    m = bvh.collapse()
    threads = host_threads()
    n = 200_000

    def step(k):
        rays = W.shadow_rays(tris, n, first=k * n)
        return O.trace(m, tris, rays, mode="any", threads=threads)[1]

    for k in range(args.warmup):
        step(k)
    tot = sum(step(args.warmup + k) for k in range(args.steps))
    v = args.steps * n / tot / 1e3
    emit_reference(args, C4_METRIC, "Mrays/s", v, tot / args.steps, True, C4_WORKLOAD,
                   f"{n} shadow rays per step on the same generator at {inst} instances ({len(tris)} triangles), Mbvh any hit")


def config4_gpu(args):
    ctx = Ctx()
    torch, api = ctx.torch, ctx.api
    instances = int(os.environ.get("CFG4_INSTANCES", "30"))
    n = int(os.environ.get("CFG4_RAYS", str(16_000_000)))
    t0 = time.time()
    tris = W.instanced_scene(instances)
    info = {"triangles": int(len(tris)), "scene_gen_s": time.time() - t0}
    # every rank builds its own replica on its GPU (deterministic builder: replicas are identical; 60 ms instead of a 4 GB broadcast)
    api.Scene.build(tris[: 1 << 16], api.BINNED_SAH, 1, mbvh=True).free()
    t0 = time.perf_counter()
    scene = api.Scene.build(tris, api.BINNED_SAH, 1, mbvh=True)
    info["scene_build_wall_ms"] = (time.perf_counter() - t0) * 1e3
    info["build_device_ms_per_mtri"] = api.last_build_stats()["device_ms"] / (len(tris) / 1e6)
    sort = os.environ.get("CFG4_SORT", "1") == "1"
    scene.set_ray_sorting(sort)
    info["ray_sorting"] = sort
    rays = W.shadow_rays(tris, n, first=ctx.rank * n)
    h_rays = torch.from_numpy(rays.view(np.float32).reshape(-1).copy()).pin_memory()
    d_rays = h_rays.cuda()
    d_occ = torch.empty(n, dtype=torch.uint8, device="cuda")
    g_out = [torch.empty(ctx.world * n, dtype=torch.uint8, device="cuda") for _ in range(2)] if ctx.world > 1 else None
    works = []

    def step(k):
        scene.occluded_device(d_rays, n, d_occ, api.TREE_MBVH, stream=ctx.stream)
        if ctx.world > 1:  # occlusion bytes of every rank gathered over NVLink, overlapped with the next step
            if len(works) >= 2:
                works[-2].wait()
            works.append(ctx.dist.all_gather_into_tensor(g_out[k % 2], d_occ, async_op=True))

    def step_and_drain(k):
        step(k)
        if k == args.warmup + args.steps - 1:
            for w in works[-2:]:
                w.wait()

    ms, clocks = ctx.timed(step_and_drain, args.steps, args.warmup)
    value = ctx.world * args.steps * n / ms / 1e3
    h_occ = [torch.zeros(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
    e2e_steps = max(1, min(args.steps, args.e2e_steps, 10))
    sub = lambda b: scene.occluded_async(h_rays.data_ptr(), n, h_occ[b].data_ptr(), api.TREE_MBVH)
    host_e2e_async(ctx, scene, sub, 2, 1)
    e2e_ms = ctx.max_over_ranks(host_e2e_async(ctx, scene, sub, 2, e2e_steps))[0]
    same = bool(torch.equal(h_occ[0], d_occ.cpu()))
    if scene.stack_overflowed():
        raise RuntimeError("traversal stack overflow")
    if ctx.rank == 0:
        cpu, roof = None, None
        if not args.no_cpu:
            from oracle import oracle as O
            threads = host_threads()
            sample = rays[:200_000]
            otree = O.Mbvh(scene.read_nodes(api.TREE_MBVH), scene.read_indices(api.TREE_MBVH))
            want, cms, _ = O.trace(otree, tris, sample, mode="any", threads=threads)
            _, _, cnt = O.trace(otree, tris, sample, mode="any", threads=threads, counters=True)
            got = d_occ[: len(sample)].cpu().numpy()
            info["parity_sample_bit_exact"] = bool(np.array_equal(got, want))
            info["occluded_fraction"] = float(got.mean())
            nv, nt = cnt["node_visits"] / len(sample), cnt["prim_tests"] / len(sample)
            cpu = {"value": len(sample) / cms / 1e3, "unit": "Mrays/s", "cores": threads, "kind": "port",
                   "sample": "first 200000 shadow rays of rank 0's shard on the GPU-built tree (read back), Mbvh any hit"}
            roof = traversal_roofline(32 + 1 + 128 * nv + 40 * nt, n, ms / args.steps,
                                      "trace_single_persistent_kernel<MBVH, any> (+ ray_keys + radix sort when sorting)", nv, nt,
                                      traffic_for("config4_dram_bytes_per_launch"))
        ctx.emit({"metric": C4_METRIC, "value": value, "unit": "Mrays/s", "n_gpus": ctx.world, "steps": args.steps,
                  "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                  "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                  "config": {"workload": C4_WORKLOAD, "rays_per_step": n, "l2": "tree + records 5.7 GB, 512 MB of rays per step: far larger than L2",
                             "sharding": "tree replicated (every rank builds the same tree on its GPU), shadow rays sharded by index range, "
                                         "occlusion bytes all_gathered over NVLink (NCCL, overlapped)" if ctx.world > 1 else "single GPU", **info},
                  "clocks": clocks, "gpu_launches": args.steps * (3 if sort else 1),
                  "e2e": {"value": ctx.world * e2e_steps * n / e2e_ms / 1e3, "unit": "Mrays/s", "h2d_bytes_per_step": n * 32,
                          "d2h_bytes_per_step": n, "steps": e2e_steps, "host_equals_resident": same,
                          "call": "rtbvh_gpu_occluded_async + rtbvh_gpu_wait (pinned RTRay records in, occlusion bytes out)"},
                  "roofline": roof, "cpu_baseline": cpu})
    scene.free()
    ctx.close()


# ======================================================================================================================
# config 5: reference-built spatial-split SAH tree uploaded unchanged, diffuse bounce rays
# ======================================================================================================================
C5_METRIC = "Mrays/s closest-hit (spatial-split SAH reference tree uploaded unchanged, Mbvh, diffuse bounce rays)"
C5_SEED = W.SEED_SOUP + 5


def c5_tris(n):
    return W.soup(n, seed=C5_SEED, aniso=(8, 1, 1))


def c5_workload(n):
    return (f"soup of {n} long thin triangles (offsets x(8,1,1)), tree from the CPU restatement of SpatialSahBuilder uploaded "
            f"unchanged + GPU collapse, one cosine-weighted bounce ray per primary hit (BASELINE configs[4])")


def bounce_rays(torch, d_rays, d_hits, d_tris, seed):
    """One cosine-weighted bounce ray per primary hit (device, torch ops: workload generation, not the measured path)."""
    rays = d_rays.view(-1, 8)
    hits = d_hits.view(-1, 2)
    prim = hits[:, 1].view(torch.int32)
    ok = prim != -1
    rays, t, prim = rays[ok], hits[ok, 0], prim[ok].long()
    o, d = rays[:, 0:3], rays[:, 4:7]
    p = o + t[:, None] * d
    tri = d_tris[prim]
    nrm = torch.linalg.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    nrm = nrm / nrm.norm(dim=1, keepdim=True).clamp_min(1e-20)
    nrm = torch.where((nrm * d).sum(1, keepdim=True) > 0, -nrm, nrm)
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    u = torch.rand((len(p), 2), generator=g, device="cuda")
    r, phi = u[:, 0].sqrt(), u[:, 1] * (2 * np.pi)
    a = torch.where(nrm[:, 0:1].abs() > 0.9, torch.tensor([0.0, 1.0, 0.0], device="cuda"), torch.tensor([1.0, 0.0, 0.0], device="cuda"))
    tx = torch.linalg.cross(nrm, a.expand_as(nrm))
    tx = tx / tx.norm(dim=1, keepdim=True)
    ty = torch.linalg.cross(nrm, tx)
    nd = tx * (r * phi.cos())[:, None] + ty * (r * phi.sin())[:, None] + nrm * (1 - u[:, 0]).clamp_min(0).sqrt()[:, None]
    nd = nd / nd.norm(dim=1, keepdim=True)
    out = torch.empty((len(p), 8), dtype=torch.float32, device="cuda")
    out[:, 0:3] = p + nrm * 1e-4
    out[:, 3] = 1e-4
    out[:, 4:7] = nd
    out[:, 7] = 1e34
    return out.contiguous()


def config5_reference(args):
    from oracle import oracle as O
    n_tris = int(os.environ.get("CFG5_REF_TRIS", str(1 << 17)))  # bounded: the spatial-split CPU build takes ~85 s per Mtri
    tris = c5_tris(n_tris)
    rc, bvh = O.build_spatial(tris, 1, True)
    m = bvh.collapse()
    threads = host_threads()
    lo, hi = W.bounds(tris)
    n = 100_000

    def step(k):
        rays = W.random_rays(n, lo, hi, seed=C5_SEED, first=k * n)  # incoherent stand-in for the bounce set (no GPU in this arm)
        return O.trace(m, tris, rays, threads=threads)[1]

    for k in range(args.warmup):
        step(k)
    tot = sum(step(args.warmup + k) for k in range(args.steps))
    v = args.steps * n / tot / 1e3
    emit_reference(args, C5_METRIC, "Mrays/s", v, tot / args.steps, True, c5_workload(n_tris),
                   f"{n} incoherent rays per step on a {n_tris}-triangle scene of the same generator (spatial-split tree by the same "
                   f"restatement), Mbvh closest hit")


def config5_gpu(args):
    ctx = Ctx()
    torch, api = ctx.torch, ctx.api
    from rtbvh_b200 import multigpu as MG
    n_tris = int(os.environ.get("CFG5_TRIS", str(1 << 20)))
    frames = int(os.environ.get("CFG5_FRAMES", "16"))
    tris = c5_tris(n_tris)
    info = {"triangles": n_tris}
    arrays = None
    if ctx.rank == 0:
        path = os.environ.get("CFG5_TREE")
        if path and os.path.exists(path):
            z = np.load(path)
            nodes, indices = z["nodes"], z["indices"]
            info.update(tree_source=f"oracle SpatialSahBuilder restatement, prebuilt ({float(z['build_s']):.0f} s on CPU)")
        else:
            from oracle import oracle as O  # stands in for "the reference built this tree"; the product only uploads it
            t0 = time.time()
            rc, ob = O.build_spatial(tris, 1, True)
            nodes, indices = ob.nodes, ob.indices
            info.update(tree_source=f"oracle SpatialSahBuilder restatement, built in-run ({time.time() - t0:.0f} s on one CPU core)",
                        sbvh_stats=list(ob.stats), sah=ob.sah_cost())
        bvh = api.Bvh.from_arrays(nodes, indices)
        mbvh = api.Mbvh.construct(bvh)  # GPU collapse of the reference-format binary tree
        info.update(collapse_device_ms=api.last_build_stats()["device_ms"], bvh_nodes=len(nodes), index_count=len(indices),
                    mbvh_nodes=int(mbvh.rt.node_count))
        arrays = {"mnodes": mbvh.nodes, "mindices": mbvh.indices}
    if ctx.world > 1:
        arrays = MG.broadcast_arrays(arrays, src=0, device="cuda")
        if ctx.rank != 0:
            mbvh = api.Mbvh.from_arrays(arrays["mnodes"], arrays["mindices"])
    scene = api.Scene(tris, bvh=None, mbvh=mbvh)
    cam = W.soup_camera(1000, 1000)
    n_primary = frames * 1_000_000
    d_prim = torch.empty(n_primary * 8, dtype=torch.float32, device="cuda")
    for f in range(frames):
        api.generate_camera_rays_device(cam, 0, 1000, d_prim[f * 8_000_000:], jitter_seed=C5_SEED, frame=ctx.rank * frames + f,
                                        stream=ctx.stream)
    d_phits = torch.empty(n_primary * 2, dtype=torch.float32, device="cuda")
    scene.intersect_device(d_prim, n_primary, d_phits, api.TREE_MBVH, stream=ctx.stream)
    torch.cuda.synchronize()
    d_tris = torch.from_numpy(tris).cuda()
    d_rays = bounce_rays(torch, d_prim, d_phits, d_tris, seed=1234 + ctx.rank)
    n = ctx.min_over_ranks(int(d_rays.shape[0]))  # equal shards
    d_rays = d_rays[:n].contiguous().view(-1)
    del d_prim, d_phits, d_tris
    d_hits = torch.empty(n * 2, dtype=torch.float32, device="cuda")
    g_out = [torch.empty(ctx.world * n * 2, dtype=torch.float32, device="cuda") for _ in range(2)] if ctx.world > 1 else None
    sort = os.environ.get("CFG5_SORT", "1") == "1"
    scene.set_ray_sorting(sort)
    info["ray_sorting"] = sort
    works = []

    def step(k):
        scene.intersect_device(d_rays, n, d_hits, api.TREE_MBVH, stream=ctx.stream)
        if ctx.world > 1:
            if len(works) >= 2:
                works[-2].wait()
            works.append(ctx.dist.all_gather_into_tensor(g_out[k % 2], d_hits, async_op=True))
        if k == args.warmup + args.steps - 1:
            for w in works[-2:]:
                w.wait()

    ms, clocks = ctx.timed(step, args.steps, args.warmup)
    value = ctx.world * args.steps * n / ms / 1e3
    h_rays = torch.empty(n * 8, dtype=torch.float32).pin_memory()
    h_rays.copy_(d_rays)
    h_hits = [torch.zeros(n * 2, dtype=torch.float32).pin_memory() for _ in range(2)]
    e2e_steps = max(1, min(args.steps, args.e2e_steps, 10))
    sub = lambda b: scene.intersect_async(h_rays.data_ptr(), n, h_hits[b].data_ptr(), api.TREE_MBVH)
    host_e2e_async(ctx, scene, sub, 2, 1)
    e2e_ms = ctx.max_over_ranks(host_e2e_async(ctx, scene, sub, 2, e2e_steps))[0]
    same = bool(torch.equal(h_hits[0].view(torch.int32), d_hits.cpu().view(torch.int32)))
    if scene.stack_overflowed():
        raise RuntimeError("traversal stack overflow")
    if ctx.rank == 0:
        cpu, roof = None, None
        if not args.no_cpu:
            from oracle import oracle as O
            threads = host_threads()
            sample = d_rays[: 200_000 * 8].cpu().numpy().view(api.RAY_DTYPE).reshape(-1)
            otree = O.Mbvh(mbvh.nodes.copy(), mbvh.indices.copy())
            want, cms, _ = O.trace(otree, tris, sample, threads=threads)
            _, _, cnt = O.trace(otree, tris, sample, threads=threads, counters=True)
            got = d_hits[: len(sample) * 2].cpu().numpy().view(api.HIT_DTYPE).reshape(-1)
            info["parity_sample_bit_exact"] = bool(np.array_equal(got, want))
            info["hit_fraction"] = float((got["prim"] != api.NO_HIT).mean())
            nv, nt = cnt["node_visits"] / len(sample), cnt["prim_tests"] / len(sample)
            cpu = {"value": len(sample) / cms / 1e3, "unit": "Mrays/s", "cores": threads, "kind": "port",
                   "sample": "first 200000 bounce rays of rank 0's shard, Mbvh closest hit"}
            roof = traversal_roofline(32 + 8 + 128 * nv + 40 * nt, n, ms / args.steps,
                                      "trace_single_persistent_kernel<MBVH, closest> (+ ray_keys + radix sort when sorting)", nv, nt,
                                      traffic_for("config5_dram_bytes_per_launch"))
        ctx.emit({"metric": C5_METRIC, "value": value, "unit": "Mrays/s", "n_gpus": ctx.world, "steps": args.steps,
                  "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                  "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                  "config": {"workload": c5_workload(n_tris), "rays_per_step": n,
                             "l2": f"{n * 32 >> 20} MB of rays + {n * 8 >> 20} MB of records per step: larger than L2",
                             "sharding": "tree uploaded on rank 0's host, replicated (NCCL broadcast of the arrays), bounce rays sharded by "
                                         "frame range, hit records all_gathered over NVLink (NCCL, overlapped)" if ctx.world > 1 else "single GPU",
                             **info},
                  "clocks": clocks, "gpu_launches": args.steps * (3 if sort else 1),
                  "e2e": {"value": ctx.world * e2e_steps * n / e2e_ms / 1e3, "unit": "Mrays/s", "h2d_bytes_per_step": n * 32,
                          "d2h_bytes_per_step": n * 8, "steps": e2e_steps, "host_equals_resident": same,
                          "call": "rtbvh_gpu_intersect_async + rtbvh_gpu_wait (pinned RTRay records in, hit records out)"},
                  "roofline": roof, "cpu_baseline": cpu})
    scene.free()
    ctx.close()


RUNNERS = {1: (config1_gpu, config1_reference), 3: (config3_gpu, config3_reference), 4: (config4_gpu, config4_reference),
           5: (config5_gpu, config5_reference)}


def run(args):
    gpu, ref = RUNNERS[args.config]
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            ref(args)
        return
    gpu(args)
