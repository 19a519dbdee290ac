#!/usr/bin/env python
"""Binned-SAH / LOCB build times on the three geometry sizes of BASELINE.json (1 Mi soup, 10 M height field, 30 M
instanced scene): device ms (CUDA events, median of 3 after a warm-up build) per Mtri, plus collapse.  One JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtbvh_b200 import api, workloads as W  # noqa: E402

which = sys.argv[1:] or ["soup", "field", "scene30m"]
gen = {"soup": lambda: W.soup(1 << 20), "field": lambda: W.heightfield(2237, 2237), "scene30m": lambda: W.instanced_scene(30)}
out = {}
for name in which:
    tris = gen[name]()
    mtri = len(tris) / 1e6
    row = {"triangles": len(tris)}
    for kind, label in ((api.BINNED_SAH, "binned_sah"), (api.LOCALLY_ORDERED_CLUSTERED, "locb")):
        api.build_triangles(tris, kind, 1).free()
        dev, b = [], None
        for _ in range(3):
            if b is not None:
                b.free()
            b = api.build_triangles(tris, kind, 1)
            dev.append(api.last_build_stats()["device_ms"])
        m = api.Mbvh.construct(b)
        cst = api.last_build_stats()
        row[label] = {"device_ms": float(np.median(dev)), "ms_per_mtri": float(np.median(dev)) / mtri, "runs": dev,
                      "nodes": int(b.rt.node_count), "collapse_device_ms": cst["device_ms"], "mbvh_nodes": int(m.rt.node_count)}
        m.free()
        b.free()
        print(name, label, row[label], file=sys.stderr, flush=True)
    out[name] = row
print(json.dumps(out))
