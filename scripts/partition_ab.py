#!/usr/bin/env python
"""A/B of the binned-SAH level-loop variants: legacy (CUB scan-by-key partition, separate split / warp-task / scan / emit
launches) vs the block partition, the merged split + warp-task launch and the fused scan + emit kernel (the defaults).
The knobs are read once per process, so every mode runs in its own child; the children print a digest of the trees
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
    scenes = {"teapot": W.teapot(), "soup64k": W.soup(1 << 16), "soup1m": W.soup(1 << 20)}
    if os.environ.get("AB_QUICK") != "1":
        scenes["soup3m"] = W.soup(3 << 20)
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


MODES = {  # environment of every child; "all" = the library's defaults
    "legacy": {"RTBVH_SAH_PARTITION": "cub", "RTBVH_SAH_MERGE": "0", "RTBVH_SAH_SCANEMIT": "0"},
    "part": {"RTBVH_SAH_MERGE": "0", "RTBVH_SAH_SCANEMIT": "0"},
    "part+merge": {"RTBVH_SAH_SCANEMIT": "0"},
    "part+scanemit": {"RTBVH_SAH_MERGE": "0"},
    "all": {},
}
if os.environ.get("AB_LS") == "1":  # the experimental level-synchronous small-subtree kernel (never verified on a GPU yet)
    MODES["all+ls"] = {"RTBVH_SAH_SMALL": "ls"}


def main():
    res = {}
    only = [m for m in os.environ.get("AB_ONLY", "").split(",") if m]
    modes = {m: e for m, e in MODES.items() if not only or m in only or m == "legacy"}
    for mode, extra in modes.items():
        env = {k: v for k, v in os.environ.items() if not k.startswith("RTBVH_SAH_")}
        env.update(extra)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True, timeout=300)
        if r.returncode != 0:
            print(json.dumps({"mode": mode, "failed": r.stderr[-1500:]}))
            continue
        res[mode] = json.loads(r.stdout.strip().splitlines()[-1])
    ref = res["legacy"]
    same = {m: {k: res[m][k]["sha"] == ref[k]["sha"] and res[m][k]["nodes"] == ref[k]["nodes"] for k in ref} for m in res}
    summary = {"identical_to_legacy": {m: all(v.values()) for m, v in same.items()},
               "differing": {m: [k for k, ok in v.items() if not ok] for m, v in same.items() if not all(v.values())},
               "device_ms_median": {k: {m: round(sorted(res[m][k]["device_ms"])[len(res[m][k]["device_ms"]) // 2], 3) for m in res}
                                    for k in ref}}
    print(json.dumps(summary))
    sys.exit(0 if all(summary["identical_to_legacy"].values()) and len(res) == len(modes) else 2)


if __name__ == "__main__":
    child() if "--child" in sys.argv else main()
