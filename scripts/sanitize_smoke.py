#!/usr/bin/env python
"""Small end-to-end exercise of every kernel family for `compute-sanitizer --tool memcheck` (and racecheck / synccheck):
single rays and packets on both trees, any hit, tiling, the chunk-wise gather push into the own buffer, camera rays, resident
build + refit, clone.  Results are compared with the oracle so that a silent corruption also fails the run."""
import os
import sys

import numpy as np

os.environ.setdefault("RTBVH_GATHER_PUSH", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import oracle as O  # noqa: E402
from rtbvh_b200 import api, workloads as W  # noqa: E402

api.set_device(0)
torch.cuda.set_device(0)
tris = W.teapot()
aabbs, centers = O.prims_from_triangles(tris)
rc, bvh = O.build(O.BINNED_SAH, aabbs, centers, 1)
m = bvh.collapse()
sc = api.Scene(tris, bvh=api.Bvh.from_arrays(bvh.nodes, bvh.indices), mbvh=api.Mbvh.from_arrays(m.nodes, m.indices))
rays = np.concatenate([W.camera_rays(W.benchmark_camera(64, 64)), W.random_rays(5_000, *W.bounds(tris), seed=3)])
for tree, ot in ((api.TREE_MBVH, m), (api.TREE_BVH, bvh)):
    assert np.array_equal(sc.intersect(rays, tree), O.trace(ot, tris, rays)[0])
    assert np.array_equal(sc.occluded(rays, tree), O.trace(ot, tris, rays, mode="any")[0])
    pk = W.pack4(rays[: len(rays) // 4 * 4])
    assert np.array_equal(sc.intersect_packets(pk, tree), O.trace_packets(ot, tris, pk)[0])
    assert np.array_equal(sc.occluded_packets(pk, tree), O.trace_packets(ot, tris, pk, mode="any")[0])
# device-resident, tiled, gather push into the own buffer
stream = torch.cuda.current_stream().cuda_stream
sc.set_ray_tiling(64)
cam = W.benchmark_camera(64, 64)
n = 64 * 64 * 3 + 37
d_rays = torch.empty(n * 8, dtype=torch.float32, device="cuda")
host = np.concatenate([W.camera_rays(cam)] * 3 + [W.random_rays(37, *W.bounds(tris), seed=5)])
d_rays.copy_(torch.from_numpy(host.view(np.float32).reshape(-1).copy()))
buf = api.PeerBuffer((n + 2) * 8)
d_hits = torch.zeros(n * 2, dtype=torch.float32, device="cuda")
sc.intersect_device_scatter(d_rays, n, [buf.ptr.value], 2, d_hits, api.TREE_MBVH, stream)
torch.cuda.synchronize()
want = O.trace(m, tris, host)[0]
assert np.array_equal(d_hits.cpu().numpy().view(api.HIT_DTYPE).reshape(-1), want)
got = api.device_view(buf.ptr.value, (n + 2) * 8).cpu().numpy()[16:].view(api.HIT_DTYPE).reshape(-1)
assert np.array_equal(got, want)
d_occ = torch.zeros(n, dtype=torch.uint8, device="cuda")
buf2 = api.PeerBuffer(n + 2)
sc.occluded_device_scatter(d_rays, n, [buf2.ptr.value], 2, d_occ, api.TREE_MBVH, stream)
torch.cuda.synchronize()
want_occ = O.trace(m, tris, host, mode="any")[0]
assert np.array_equal(d_occ.cpu().numpy(), want_occ)
assert np.array_equal(api.device_view(buf2.ptr.value, n + 2).cpu().numpy()[2:], want_occ)
buf.free()
buf2.free()
# resident build, refit, clone
rs = api.Scene.build(tris, api.BINNED_SAH, 1, mbvh=True)
rs.refit(tris)
cl = rs.clone(0)
a, b = rs.intersect(rays, api.TREE_MBVH), cl.intersect(rays, api.TREE_MBVH)
assert np.array_equal(a, b)
ls = api.Scene.build(tris, api.LOCALLY_ORDERED_CLUSTERED, 1, mbvh=True)
assert (ls.intersect(rays, api.TREE_MBVH)["prim"] != api.NO_HIT).any()
for s in (rs, cl, ls, sc):
    s.free()
print("sanitize_smoke ok")
