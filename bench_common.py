"""Helpers shared by bench.py (config 2, the driver's default) and bench_configs.py (configs 1, 3, 4, 5)."""
from __future__ import annotations

import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the traversal kernel per launch, from the committed
    `ncu --set full` capture of this same workload (profiles/ncu_traffic.json); None if absent."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return float(json.load(open(p))["dram_bytes_per_launch"])
    except Exception:
        return None


def nvlink_kib(index: int):
    """Sum of the NVLink data counters of GPU `index` over all links: (tx KiB, rx KiB), or None when nvidia-smi cannot
    report them (`nvidia-smi nvlink -gt d`)."""
    import re
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True,
                             timeout=20).stdout
        tx = [int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out)]
        rx = [int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out)]
        if not tx and not rx:
            return None
        return sum(tx), sum(rx)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """All host cores this process may use (torchrun sets OMP_NUM_THREADS=1, which must not shrink the CPU arm)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def bind_to_gpu_numa_node(local: int) -> str:
    """Pins this process (and with it the first-touch placement of the pinned host buffers it allocates afterwards) to the CPUs
    of the NUMA node its GPU hangs off.  torchrun sets no affinity: without this, a rank's staging buffers may sit on the other
    socket and every H2D / D2H crosses the inter-socket link.  Returns a short description for the JSON line."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return f"gpu {bus}: no NUMA node reported"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return f"gpu {bus}: NUMA node {node} has no allowed CPU"
        os.sched_setaffinity(0, allowed)
        return f"gpu {bus}: bound to NUMA node {node} ({len(allowed)} CPUs)"
    except Exception as e:  # never fatal: the bench runs unbound
        return f"not bound ({type(e).__name__}: {e})"
