#!/usr/bin/env python
"""ONE binned-SAH build of the 1 Mi-triangle soup (no warm-up, no second run) — the target of `ncu --set full` captures of
the builder kernels (scripts/gpu_ncu_build.sh): every kernel instance ncu sees belongs to that single build."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtbvh_b200 import api, workloads as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
tris = W.soup(n)
b = api.build_triangles(tris, api.BINNED_SAH, 1)
print("soup", len(tris), "binned", api.last_build_stats(), b.rt.node_count, flush=True)
