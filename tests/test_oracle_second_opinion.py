"""A second, independent restatement of the reference's single-ray traversal — written from the Rust sources, in plain
Python with numpy float32 scalars (one IEEE rounding per operation), sharing no code with oracle/rtbvh_oracle.hpp — checked
against the C++ oracle bit for bit on teapot trees.  The oracle cannot be pinned by running the reference (no Rust toolchain
in this image); two restatements that were transcribed separately and agree on every ray narrow what "parity unpinned" leaves
open to misreadings that BOTH share.

Followed sources (file:line of /root/reference):
  Ray::new                         src/ray.rs:166-182
  Aabb::intersect                  src/aabb.rs:146-181
  BvhNode::sort_nodes              src/bvh_node.rs:150-177
  BvhIndexIterator::next           src/iter_indices.rs:69-106 (ctor :32-46: NaN origin / direction -> nothing is visited)
  MbvhNode::intersect              src/mbvh_node.rs:177-240 (glam Vec4 min / max = SSE: the SECOND operand when either is NaN)
  MbvhIndexIterator::next          src/iter_indices.rs:267-312 (ctor :222-240)
  SpatialTriangle::intersect       src/builders/spatial_sah.rs:131-163 (glam Vec3 cross / dot, scalar)
  caller loop                      examples/benchmark.rs:25-31; any hit = the FFI callback returning true (rtbvh_ffi/src/lib.rs:572-576)
  RayPacket4::new                  src/ray.rs:64-103
  Aabb::intersect4                 src/aabb.rs:218-244
  BvhNode::sort_nodes4             src/bvh_node.rs:180-211
  BvhPacketIndexIterator           src/iter_indices.rs:121-209 (a NaN in any lane rejects the packet)
  MbvhNode::intersect4             src/mbvh_node.rs:243-295
  MbvhPacketIndexIterator::next    src/iter_indices.rs:370-414
  SpatialTriangle::intersect4      src/builders/spatial_sah.rs:165-244 (eps 1e-6, t >= t_min; caller passes t_min = 1e-4)
plus the north star's id rule: among candidates with exactly equal accepted t the lowest primitive id is reported."""
import numpy as np
import pytest

F = np.float32
NO_HIT = 0xFFFFFFFF


def _ray_new(o, d, t_min, t):
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = [F(1.0) / d[0], F(1.0) / d[1], F(1.0) / d[2]]
    signs = [1 if d[k] < F(0.0) else 0 for k in range(3)]
    return {"o": o, "d": d, "inv": inv, "signs": signs, "t_min": t_min, "t": t, "prim": NO_HIT}


def _tri_intersect(ray, v0, v1, v2, prim):
    e1 = [v1[k] - v0[k] for k in range(3)]
    e2 = [v2[k] - v0[k] for k in range(3)]
    d = ray["d"]
    cross = lambda a, b: [a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]]
    dot = lambda a, b: (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]
    h = cross(d, e2)
    a = dot(e1, h)
    if a > F(-1e-5) and a < F(1e-5):
        return
    f = F(1.0) / a
    s = [ray["o"][k] - v0[k] for k in range(3)]
    u = f * dot(s, h)
    if not (F(0.0) <= u <= F(1.0)):
        return
    q = cross(s, e1)
    v = f * dot(d, q)
    if v < F(0.0) or (u + v) > F(1.0):
        return
    t = f * dot(e2, q)
    if t > ray["t_min"] and t < ray["t"]:
        ray["t"] = t
        ray["prim"] = prim
    elif t > ray["t_min"] and ray["prim"] != NO_HIT and t == ray["t"] and prim < ray["prim"]:
        ray["prim"] = prim  # north star: ties broken by the lowest id


def _aabb_intersect(node, ray):
    box = (node["min"], node["max"])
    o, inv, sg = ray["o"], ray["inv"], ray["signs"]
    ray_min = (box[sg[0]][0] - o[0]) * inv[0]
    ray_max = (box[1 - sg[0]][0] - o[0]) * inv[0]
    y_min = (box[sg[1]][1] - o[1]) * inv[1]
    y_max = (box[1 - sg[1]][1] - o[1]) * inv[1]
    if ray_min > y_max or y_min > ray_max:
        return None
    if y_min > ray_min:
        ray_min = y_min
    if y_max < ray_max:
        ray_max = y_max
    z_min = (box[sg[2]][2] - o[2]) * inv[2]
    z_max = (box[1 - sg[2]][2] - o[2]) * inv[2]
    if ray_min > z_max or z_min > ray_max:
        return None
    if z_max < ray_max:
        ray_max = z_max
    return ray_max if ray_max > ray["t_min"] else None


def _trace_bvh(nodes, indices, tris, ray, tri=None):
    tri = tri or _tri_intersect
    if len(nodes) == 0 or any(np.isnan(x) for x in ray["o"]) or any(np.isnan(x) for x in ray["d"]):
        return
    stack = [0]
    while stack:
        node = nodes[stack.pop()]
        count, left_first = int(node["count"]), int(node["left_first"])
        if count > -1:
            for i in range(count):
                p = int(indices[left_first + i])
                tri(ray, tris[p][0], tris[p][1], tris[p][2], p)
        elif left_first > -1:
            left = _aabb_intersect(nodes[left_first], ray)
            right = _aabb_intersect(nodes[left_first + 1], ray)
            if left is not None and right is not None:
                if left < right:
                    stack += [left_first, left_first + 1]
                else:
                    stack += [left_first + 1, left_first]
            elif left is not None:
                stack.append(left_first)
            elif right is not None:
                stack.append(left_first + 1)


def _sse_min(a, b):  # _mm_min_ps(a, b): a < b ? a : b
    return a if a < b else b


def _sse_max(a, b):  # _mm_max_ps(a, b): a > b ? a : b
    return a if a > b else b


def _mnode_intersect(n, ray):
    o, inv = ray["o"], ray["inv"]
    t_min, result = [None] * 4, [False] * 4
    with np.errstate(invalid="ignore", over="ignore"):
        for s in range(4):
            tx0, tx1 = (n["min_x"][s] - o[0]) * inv[0], (n["max_x"][s] - o[0]) * inv[0]
            ty0, ty1 = (n["min_y"][s] - o[1]) * inv[1], (n["max_y"][s] - o[1]) * inv[1]
            tz0, tz1 = (n["min_z"][s] - o[2]) * inv[2], (n["max_z"][s] - o[2]) * inv[2]
            tmn = _sse_max(_sse_min(tx0, tx1), _sse_max(_sse_min(ty0, ty1), _sse_min(tz0, tz1)))
            tmx = _sse_min(_sse_max(tx0, tx1), _sse_min(_sse_max(ty0, ty1), _sse_max(tz0, tz1)))
            t_min[s] = tmn
            result[s] = bool(tmx >= tmn) and bool(tmn < ray["t"])
    ids = [0, 1, 2, 3]

    def cswap(i, j, keys_too=True):
        if t_min[i] > t_min[j]:
            if keys_too:
                t_min[i], t_min[j] = t_min[j], t_min[i]
            ids[i], ids[j] = ids[j], ids[i]
    cswap(0, 1)
    cswap(2, 3)
    cswap(0, 2)
    cswap(1, 3)
    cswap(2, 3, keys_too=False)  # the reference's last comparator swaps the ids only
    return ids, result


def _trace_mbvh(mnodes, indices, tris, ray, tri=None):
    tri = tri or _tri_intersect
    if len(mnodes) == 0:
        return
    current, stack = 0, []
    hit = _mnode_intersect(mnodes[0], ray)
    while True:
        node = mnodes[current]
        ids, result = hit
        for i in range(4):
            sid = ids[3 - i]
            if result[sid]:
                count, left_first = int(node["counts"][sid]), int(node["children"][sid])
                if count > -1:
                    for j in range(count):
                        p = int(indices[left_first + j])
                        tri(ray, tris[p][0], tris[p][1], tris[p][2], p)
                elif left_first > -1:
                    stack.append(left_first)
        if not stack:
            return
        current = stack.pop()
        hit = _mnode_intersect(mnodes[current], ray)


class _Stop(Exception):
    pass


def _tri_any(ray, v0, v1, v2, prim):
    t0 = ray["t"]
    _tri_intersect(ray, v0, v1, v2, prim)
    if ray["t"] < t0:  # SpatialTriangle::intersect returned true -> the callback breaks the loop
        raise _Stop


# ---- packets of four rays: lists of four float32 per component ------------------------------------------------------
def _packet_new(pk):
    P = {k: [F(x) for x in pk[k]] for k in ("origin_x", "origin_y", "origin_z", "direction_x", "direction_y", "direction_z", "t")}
    with np.errstate(divide="ignore", invalid="ignore"):
        for ax in "xyz":
            P["inv_" + ax] = [F(1.0) / d for d in P["direction_" + ax]]
    P["prim"] = [NO_HIT] * 4
    return P


def _aabb_intersect4(node, P):
    lo, hi = node["min"], node["max"]
    t_near, any_lane = [None] * 4, False
    for i in range(4):
        t1 = [(lo[k] - P["origin_" + ax][i]) * P["inv_" + ax][i] for k, ax in enumerate("xyz")]
        t2 = [(hi[k] - P["origin_" + ax][i]) * P["inv_" + ax][i] for k, ax in enumerate("xyz")]
        mn = [_sse_min(t1[k], t2[k]) for k in range(3)]
        mx = [_sse_max(t1[k], t2[k]) for k in range(3)]
        t_min = _sse_max(mn[0], _sse_max(mn[1], mn[2]))
        t_max = _sse_min(mx[0], _sse_min(mx[1], mx[2]))
        t_near[i] = t_min
        any_lane = any_lane or (bool(t_max > F(0.0)) and bool(t_max > t_min) and bool(t_min < P["t"][i]))
    return t_near if any_lane else None


def _tri_intersect4(P, v0, v1, v2, prim, t_min=F(1e-4)):
    e1 = [v1[k] - v0[k] for k in range(3)]
    e2 = [v2[k] - v0[k] for k in range(3)]
    for i in range(4):
        d = [P["direction_x"][i], P["direction_y"][i], P["direction_z"][i]]
        o = [P["origin_x"][i], P["origin_y"][i], P["origin_z"][i]]
        h = [d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0]]
        a = (e1[0] * h[0] + e1[1] * h[1]) + e1[2] * h[2]
        ok = bool(a <= F(-1e-6)) or bool(a >= F(1e-6))
        f = F(1.0) / a
        sv = [o[k] - v0[k] for k in range(3)]
        u = f * ((sv[0] * h[0] + sv[1] * h[1]) + sv[2] * h[2])
        ok = ok and bool(u >= F(0.0)) and bool(u <= F(1.0))
        q = [sv[1] * e1[2] - sv[2] * e1[1], sv[2] * e1[0] - sv[0] * e1[2], sv[0] * e1[1] - sv[1] * e1[0]]
        v = f * ((d[0] * q[0] + d[1] * q[1]) + d[2] * q[2])
        ok = ok and bool(v >= F(0.0)) and bool((u + v) <= F(1.0))
        t = f * ((e2[0] * q[0] + e2[1] * q[1]) + e2[2] * q[2])
        ok = ok and bool(t >= t_min)
        if ok and t < P["t"][i]:
            P["t"][i] = t
            P["prim"][i] = prim
        elif ok and P["prim"][i] != NO_HIT and t == P["t"][i] and prim < P["prim"][i]:
            P["prim"][i] = prim  # north star: ties broken by the lowest id


def _trace_bvh_packet(nodes, indices, tris, P):
    comps = [P[k] for k in ("origin_x", "origin_y", "origin_z", "direction_x", "direction_y", "direction_z")]
    if len(nodes) == 0 or any(np.isnan(x) for c in comps for x in c):
        return
    stack = [0]
    while stack:
        node = nodes[stack.pop()]
        count, left_first = int(node["count"]), int(node["left_first"])
        if count > -1:
            for i in range(count):
                p = int(indices[left_first + i])
                _tri_intersect4(P, tris[p][0], tris[p][1], tris[p][2], p)
        elif left_first > -1:
            left = _aabb_intersect4(nodes[left_first], P)
            right = _aabb_intersect4(nodes[left_first + 1], P)
            if left is not None and right is not None:
                if any(bool(left[i] < right[i]) for i in range(4)):
                    stack += [left_first, left_first + 1]
                else:
                    stack += [left_first + 1, left_first]
            elif left is not None:
                stack.append(left_first)
            elif right is not None:
                stack.append(left_first + 1)


def _mnode_intersect4(n, P):
    result = [False] * 4
    for i in range(4):
        for s in range(4):
            t1 = (n["min_x"][s] - P["origin_x"][i]) * P["inv_x"][i]
            t2 = (n["max_x"][s] - P["origin_x"][i]) * P["inv_x"][i]
            t_min, t_max = _sse_min(t1, t2), _sse_max(t1, t2)
            t1 = (n["min_y"][s] - P["origin_y"][i]) * P["inv_y"][i]
            t2 = (n["max_y"][s] - P["origin_y"][i]) * P["inv_y"][i]
            t_min, t_max = _sse_max(t_min, _sse_min(t1, t2)), _sse_min(t_max, _sse_max(t1, t2))
            t1 = (n["min_z"][s] - P["origin_z"][i]) * P["inv_z"][i]
            t2 = (n["max_z"][s] - P["origin_z"][i]) * P["inv_z"][i]
            t_min, t_max = _sse_max(t_min, _sse_min(t1, t2)), _sse_min(t_max, _sse_max(t1, t2))
            result[s] = result[s] or (bool(t_max > t_min) and bool(t_min < P["t"][i]))
    return result


def _trace_mbvh_packet(mnodes, indices, tris, P):
    if len(mnodes) == 0:
        return
    current, stack = 0, []
    result = _mnode_intersect4(mnodes[0], P)
    while True:
        node = mnodes[current]
        for i in range(4):
            sid = 3 - i  # ids = [0, 1, 2, 3]: no ordering for packets
            if result[sid]:
                count, left_first = int(node["counts"][sid]), int(node["children"][sid])
                if count > -1:
                    for j in range(count):
                        p = int(indices[left_first + j])
                        _tri_intersect4(P, tris[p][0], tris[p][1], tris[p][2], p)
                elif left_first > -1:
                    stack.append(left_first)
        if not stack:
            return
        current = stack.pop()
        result = _mnode_intersect4(mnodes[current], P)


def _rays(W, tris):
    cam = W.camera_rays(W.benchmark_camera(40, 40))
    rnd = W.random_rays(900, *W.bounds(tris), seed=0x2ED0)
    edge = rnd[:60].copy()
    edge["direction"][:20, 0] = 0.0      # axis-parallel: +-inf inverse directions, NaN slab products
    edge["direction"][20:40, 1] = -0.0
    edge["direction"][40:50] = 0.0
    edge["origin"][50:55, 2] = np.nan
    edge["direction"][55:60, 1] = np.nan
    return np.concatenate([cam, rnd, edge])


@pytest.mark.parametrize("builder", ["sah", "locb"])
def test_python_restatement_agrees_with_the_cpp_oracle(O, W, teapot, teapot_trees, builder):
    tris = teapot["tris"].astype(np.float32)
    bvh, m = teapot_trees[builder]
    rays = _rays(W, tris)
    want_b = O.trace(bvh, tris, rays, threads=1)[0]
    want_m = O.trace(m, tris, rays, threads=1)[0]
    T = [[[F(c) for c in v] for v in t] for t in tris]
    for kind, nodes, idx, want in (("bvh", bvh.nodes, bvh.indices, want_b), ("mbvh", m.nodes, m.indices, want_m)):
        got = np.zeros(len(rays), dtype=want.dtype)
        for k, r in enumerate(rays):
            ray = _ray_new([F(x) for x in r["origin"]], [F(x) for x in r["direction"]], F(r["t_min"]), F(r["t"]))
            with np.errstate(all="ignore"):
                (_trace_bvh if kind == "bvh" else _trace_mbvh)(nodes, idx, T, ray)
            got["t"][k], got["prim"][k] = ray["t"], ray["prim"]
        same_t = got["t"].view(np.uint32) == want["t"].view(np.uint32)
        same_p = got["prim"] == want["prim"]
        bad = np.nonzero(~(same_t & same_p))[0]
        assert len(bad) == 0, f"{builder}/{kind}: {len(bad)} of {len(rays)} rays differ, first {bad[:5]}: {got[bad[:3]]} vs {want[bad[:3]]}"


@pytest.mark.parametrize("builder", ["sah", "locb"])
def test_python_restatement_agrees_on_any_hit_and_packets(O, W, teapot, teapot_trees, builder):
    tris = teapot["tris"].astype(np.float32)
    bvh, m = teapot_trees[builder]
    rays = _rays(W, tris)[:1600]
    T = [[[F(c) for c in v] for v in t] for t in tris]
    # any hit: the callback returns true on the first accepted candidate
    for kind, nodes, idx, tree in (("bvh", bvh.nodes, bvh.indices, bvh), ("mbvh", m.nodes, m.indices, m)):
        want = O.trace(tree, tris, rays, mode="any", threads=1)[0]
        got = np.zeros(len(rays), dtype=np.uint8)
        for k, r in enumerate(rays):
            ray = _ray_new([F(x) for x in r["origin"]], [F(x) for x in r["direction"]], F(r["t_min"]), F(r["t"]))
            try:
                with np.errstate(all="ignore"):
                    (_trace_bvh if kind == "bvh" else _trace_mbvh)(nodes, idx, T, ray, tri=_tri_any)
            except _Stop:
                got[k] = 1
        assert np.array_equal(got, want), f"{builder}/{kind} any hit: {(got != want).sum()} rays differ"
    # packets of four consecutive rays, closest hit
    packets = W.pack4(rays[: len(rays) // 4 * 4])
    for kind, nodes, idx, tree in (("bvh", bvh.nodes, bvh.indices, bvh), ("mbvh", m.nodes, m.indices, m)):
        want = O.trace_packets(tree, tris, packets, threads=1)[0]
        got = np.zeros(len(packets), dtype=want.dtype)
        for k, pk in enumerate(packets):
            P = _packet_new(pk)
            with np.errstate(all="ignore"):
                (_trace_bvh_packet if kind == "bvh" else _trace_mbvh_packet)(nodes, idx, T, P)
            got["t"][k], got["prim"][k] = P["t"], P["prim"]
        same = (got["t"].view(np.uint32) == want["t"].view(np.uint32)).all(axis=1) & (got["prim"] == want["prim"]).all(axis=1)
        bad = np.nonzero(~same)[0]
        assert len(bad) == 0, f"{builder}/{kind} packets: {len(bad)} of {len(packets)} differ, first {bad[:5]}: {got[bad[:2]]} vs {want[bad[:2]]}"


@pytest.mark.parametrize("leaf,kind", [(1, "sah"), (3, "sah"), (8, "sah"), (1, "locb")])
def test_second_opinion_on_duplicated_triangles_and_wide_leaves(O, W, leaf, kind):
    """Exactly equal t from duplicated triangles (the lowest id must be reported, whatever order the leaves come in) and
    leaves with several primitives: the two restatements must still agree bit for bit."""
    tris = W.soup(240, seed=0x71E5).astype(np.float32)
    tris[160:240] = tris[0:80]  # 80 exact duplicates with higher ids
    aabbs, centers = O.prims_from_triangles(tris)
    rc, bvh = O.build(O.BINNED_SAH if kind == "sah" else O.LOCB, aabbs, centers, leaf)
    assert rc == 0
    m = bvh.collapse()
    rays = np.concatenate([W.camera_rays(W.soup_camera(24, 24)), W.random_rays(400, *W.bounds(tris), seed=5)])
    T = [[[F(c) for c in v] for v in t] for t in tris]
    brute = O.brute_force(tris, rays)
    for name, nodes, idx, tree in (("bvh", bvh.nodes, bvh.indices, bvh), ("mbvh", m.nodes, m.indices, m)):
        want = O.trace(tree, tris, rays, threads=1)[0]
        got = np.zeros(len(rays), dtype=want.dtype)
        for k, r in enumerate(rays):
            ray = _ray_new([F(x) for x in r["origin"]], [F(x) for x in r["direction"]], F(r["t_min"]), F(r["t"]))
            with np.errstate(all="ignore"):
                (_trace_bvh if name == "bvh" else _trace_mbvh)(nodes, idx, T, ray)
            got["t"][k], got["prim"][k] = ray["t"], ray["prim"]
        assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32)) and np.array_equal(got["prim"], want["prim"]), name
        hit = want["prim"] != NO_HIT
        assert hit.any() and (want["prim"][hit] < 160).all(), "a duplicate's higher id was reported"
        if kind == "locb":  # conservative boxes: the walk finds what brute force finds
            assert np.array_equal(want, brute), name
    packets = W.pack4(rays[: len(rays) // 4 * 4])
    for name, nodes, idx, tree in (("bvh", bvh.nodes, bvh.indices, bvh), ("mbvh", m.nodes, m.indices, m)):
        want = O.trace_packets(tree, tris, packets, threads=1)[0]
        for k, pk in enumerate(packets):
            P = _packet_new(pk)
            with np.errstate(all="ignore"):
                (_trace_bvh_packet if name == "bvh" else _trace_mbvh_packet)(nodes, idx, T, P)
            assert np.array_equal(np.array(P["t"], dtype=F).view(np.uint32), want["t"][k].view(np.uint32)), (name, k)
            assert np.array_equal(np.array(P["prim"], dtype=np.uint32), want["prim"][k]), (name, k)
