#!/usr/bin/env python
"""Concurrent pinned H2D / D2H ceiling of the box: every rank copies 256 MB up and 64 MB down (the bytes of one e2e bench
step) back to back on two streams, all ranks at once.  Run under torchrun with N = 1, 2, 4, 8; rank 0 prints one JSON line:
per-rank and aggregate GB/s, and the e2e Mrays/s ceiling they imply for 32-byte rays + 8-byte records (and 24 + 8, 0 + 8)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench_common import bind_to_gpu_numa_node  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    numa = bind_to_gpu_numa_node(local) if os.environ.get("RTBVH_BENCH_NUMA", "1") == "1" else "off"
    up, down, reps = 256 << 20, 64 << 20, 20
    h_up = torch.empty(up, dtype=torch.uint8).pin_memory()
    h_dn = torch.empty(down, dtype=torch.uint8).pin_memory()
    d_up = torch.empty(up, dtype=torch.uint8, device="cuda")
    d_dn = torch.empty(down, dtype=torch.uint8, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def run(do_up, do_dn):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s_up.wait_event(e0)
        s_dn.wait_event(e0)
        for _ in range(reps):
            if do_up:
                with torch.cuda.stream(s_up):
                    d_up.copy_(h_up, non_blocking=True)
            if do_dn:
                with torch.cuda.stream(s_dn):
                    h_dn.copy_(d_dn, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s_up)
        torch.cuda.current_stream().wait_stream(s_dn)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps  # ms per (up, down) pair

    out = {}
    for name, (u, d) in {"h2d_only": (True, False), "d2h_only": (False, True), "both": (True, True)}.items():
        run(u, d)
        ms = run(u, d)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        gb = ((up if u else 0) + (down if d else 0)) / 1e9
        out[name] = {"ms_per_step_max_over_ranks": ms, "gbs_per_rank": gb / (ms * 1e-3), "gbs_aggregate": world * gb / (ms * 1e-3)}
    both = out["both"]["ms_per_step_max_over_ranks"] * 1e-3
    rays = 8_000_000
    out["e2e_ceiling_mrays"] = {
        "rtray_32B_in_8B_out": world * rays / both / 1e6,
        "camera_0B_in_8B_out": world * rays / (out["d2h_only"]["ms_per_step_max_over_ranks"] * 1e-3) / 1e6,
        "note": "8 M rays per step per rank: 256 MB up + 64 MB down (RTRay flavour); the split origin/direction flavour moves 192 MB up",
    }
    if rank == 0:
        print(json.dumps({"n_gpus": world, "host_numa_rank0": numa, **out}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
