#!/usr/bin/env python
"""Builder A/B for one library (RTBVH_LIB): sha256 of nodes + indices and the builder's device time on four scenes.
Every variant of the builder must print the same digests (the trees are deterministic); compare the device_ms columns."""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtbvh_b200 import api, workloads as W  # noqa: E402

name = os.environ.get("RTBVH_LIB", "default").split("librtbvh_rs")[-1].strip("_.so") or "default"
scenes = {"teapot": W.teapot(), "soup64k": W.soup(1 << 16), "soup1m": W.soup(1 << 20), "soup4m": W.soup(1 << 22)}
for sname, tris in scenes.items():
    for kind, kname in ((api.BINNED_SAH, "sah"), (api.LOCALLY_ORDERED_CLUSTERED, "locb")):
        if kind == api.LOCALLY_ORDERED_CLUSTERED and sname != "soup1m":
            continue
        ms = []
        for rep in range(5):
            b = api.build_triangles(tris, kind, 1)
            ms.append(api.last_build_stats()["device_ms"])
            h = hashlib.sha256(np.ascontiguousarray(b.nodes).tobytes() + np.ascontiguousarray(b.indices).tobytes()).hexdigest()[:12]
            b.free()
        print(f"{name:12s} {sname:8s} {kname:4s} sha {h}  device_ms median {np.median(ms[1:]):7.3f}  min {min(ms[1:]):7.3f}", flush=True)
