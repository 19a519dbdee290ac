"""Regenerates the fixtures under tests/golden/.  Run in the build container (needs /root/reference):

    python tests/golden/make_fixtures.py

teapot_tris.npy     [6320,3,3] float32 — the triangles of /root/reference/objects/teapot.obj in file order
                    (the reference loads it with l3d and takes vertices.chunks_exact(3), src/lib.rs:232-240).
                    Only `v` and triangular `f` records exist in that file.
oracle_golden.npz   regression vectors of the CPU oracle on the teapot (tree hashes, SAH, hits of a fixed ray
                    set): written by this script from oracle/liboracle.so, checked by tests/test_oracle_golden.py
                    so that an accidental change of the oracle is caught on CPU.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def load_obj(path):
    vs, fs = [], []
    with open(path) as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                vs.append([float(x) for x in p[1:4]])
            elif p[0] == "f":
                idx = [int(tok.split("/")[0]) for tok in p[1:]]
                assert len(idx) == 3
                fs.append([i - 1 if i > 0 else len(vs) + i for i in idx])
    v = np.asarray(vs, dtype=np.float32)
    return v[np.asarray(fs, dtype=np.int64)]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    tris = load_obj("/root/reference/objects/teapot.obj")
    assert tris.shape == (6320, 3, 3)
    np.save(os.path.join(HERE, "teapot_tris.npy"), tris)

    from oracle import oracle as O
    from rtbvh_b200 import workloads as W
    aabbs, centers = O.prims_from_triangles(tris)
    out = {}
    cam = W.benchmark_camera(256, 256)
    rays = np.concatenate([W.camera_rays(cam), W.random_rays(16384, *W.bounds(tris))])
    for name, kind in (("sah", O.BINNED_SAH), ("locb", O.LOCB)):
        rc, bvh = O.build(kind, aabbs, centers, 1)
        assert rc == 0
        m = bvh.collapse()
        out[name + "_nodes_sha"] = sha(bvh.nodes)
        out[name + "_indices_sha"] = sha(bvh.indices)
        out[name + "_mnodes_sha"] = sha(m.nodes)
        out[name + "_node_count"] = len(bvh.nodes)
        out[name + "_mnode_count"] = len(m.nodes)
        out[name + "_sah"] = bvh.sah_cost()
        h2, _, _ = O.trace(bvh, tris, rays)
        h4, _, _ = O.trace(m, tris, rays)
        out[name + "_bvh_hits"] = h2
        out[name + "_mbvh_hits"] = h4
        # the flavours the first version did not pin: any hit, packets of four (closest + any), on a smaller ray set
        sub = rays[::5][: 16384 // 4 * 4]
        packets = W.pack4(sub)
        for tag, tree in (("bvh", bvh), ("mbvh", m)):
            out[f"{name}_{tag}_any"] = O.trace(tree, tris, sub, mode="any")[0]
            out[f"{name}_{tag}_packet_hits"] = O.trace_packets(tree, tris, packets)[0]
            out[f"{name}_{tag}_packet_any"] = O.trace_packets(tree, tris, packets, mode="any")[0]
    # refit (src/bvh.rs:176-205, topology kept) and the spatial-split restatement (fixed child ranges)
    rc, bvh = O.build(O.BINNED_SAH, aabbs, centers, 1)
    moved = aabbs.copy()
    moved["min"] += np.float32(0.25)
    moved["max"] += np.float32(0.5)
    out["refit_nodes_sha"] = sha(bvh.refit(moved).nodes)
    sp = O.build_spatial(tris[:2000], 1)
    sp = sp[1] if isinstance(sp, tuple) else sp
    out["spatial_nodes_sha"] = sha(sp.nodes)
    out["spatial_indices_sha"] = sha(sp.indices)
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **out)
    print({k: v for k, v in out.items() if not isinstance(v, np.ndarray)})


if __name__ == "__main__":
    main()
