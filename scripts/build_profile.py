#!/usr/bin/env python
"""Builds one scene on the GPU (binned SAH + collapse, optionally LOCB) — target of `ncu --metrics gpu__time_duration.sum`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtbvh_b200 import api, workloads as W  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "soup"
tris = {"soup": lambda: W.soup(1 << 20), "field": lambda: W.heightfield(2237, 2237), "scene30m": lambda: W.instanced_scene(30)}[scene]()
api.build_triangles(tris[: 1 << 16], api.BINNED_SAH, 1).free()
for rep in range(2):
    b = api.build_triangles(tris, api.BINNED_SAH, 1)
    print(scene, len(tris), "binned", api.last_build_stats(), b.rt.node_count, flush=True)
    if rep == 0:
        b.free()
m = api.Mbvh.construct(b)
print("collapse", api.last_build_stats(), m.rt.node_count, flush=True)
if "--locb" in sys.argv:
    l = api.build_triangles(tris, api.LOCALLY_ORDERED_CLUSTERED, 1)
    print("locb", api.last_build_stats(), l.rt.node_count, flush=True)
