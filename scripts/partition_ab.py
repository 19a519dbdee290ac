#!/usr/bin/env python
"""A/B of the binned-SAH partition pass: CUB scan-by-key (RTBVH_SAH_PARTITION=cub) vs the two block kernels (default).
The mode is read once per process, so every mode runs in its own child; the children print a digest of the trees
(nodes + indices must be byte-identical: a stable partition has one result) and the builder's device time."""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    from rtbvh_b200 import api, workloads as W
    out = {}
    scenes = {"teapot": W.teapot(), "soup64k": W.soup(1 << 16), "soup1m": W.soup(1 << 20), "soup3m": W.soup(3 << 20)}
    dup = W.soup(1 << 14).copy()
    dup[1000:9000] = dup[1000]  # 8 000 identical triangles: unsplittable ranges (leaf / fallback rules)
    scenes["dups"] = dup
    for name, tris in scenes.items():
        for leaf in ((1, 4) if name in ("soup64k", "dups") else (1,)):
            ms = []
            for rep in range(4 if len(tris) >= (1 << 20) else 1):
                b = api.build_triangles(tris, api.BINNED_SAH, leaf)
                ms.append(api.last_build_stats()["device_ms"])
                h = hashlib.sha256(np.ascontiguousarray(b.nodes).tobytes() + np.ascontiguousarray(b.indices).tobytes()).hexdigest()
                nodes = int(b.rt.node_count)
                b.free()
            out[f"{name}/leaf{leaf}"] = {"sha": h, "nodes": nodes, "device_ms": ms}
    print(json.dumps(out))


def main():
    res = {}
    for mode in ("cub", "block"):
        env = dict(os.environ)
        env["RTBVH_SAH_PARTITION"] = mode
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True, timeout=300)
        if r.returncode != 0:
            print(mode, "FAILED", r.stderr[-2000:])
            sys.exit(1)
        res[mode] = json.loads(r.stdout.strip().splitlines()[-1])
    same = {k: res["cub"][k]["sha"] == res["block"][k]["sha"] and res["cub"][k]["nodes"] == res["block"][k]["nodes"] for k in res["cub"]}
    summary = {"identical": same, "all_identical": all(same.values()),
               "device_ms": {k: {m: [round(x, 3) for x in res[m][k]["device_ms"]] for m in res} for k in res["cub"]}}
    print(json.dumps(summary))
    sys.exit(0 if summary["all_identical"] else 2)


if __name__ == "__main__":
    child() if "--child" in sys.argv else main()
