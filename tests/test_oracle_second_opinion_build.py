"""A second, independent restatement of the reference's binned-SAH builder — written from src/builders/binned_sah.rs in
plain Python / numpy float32, sharing no code with oracle/rtbvh_oracle.hpp — must produce the very bytes the C++ oracle
produces (nodes incl. numbering, prim_indices incl. the order the swap partition leaves), and the same after Mbvh collapse
is applied by the oracle to both.  See tests/test_oracle_second_opinion.py for why.

Followed sources (file:line of /root/reference):
  BinnedSahBuilder::build               src/builders/binned_sah.rs:346-399 (root = Aabb::union_of_list, src/aabb.rs:125-131)
  BinnedSahBuildTask::run / find_split  src/builders/binned_sah.rs:80-128, :132-282 (incl. the fallback and its left-box quirk)
  partition (swap with slice[count])    src/utils.rs:76-96
  AtomicNodeStack::allocate             src/builders/mod.rs:59-76 (child pairs from a counter that starts at 1)
  TaskSpawner::run_task                 src/utils.rs:243-288, one thread: the child with more work is run first
  Aabb helpers                          src/aabb.rs:252-273 (grow_bb), :313-322 (offset_by), :343-346 (half_area), :354-363
"""
import numpy as np
import pytest

F = np.float32
BINS, MAX_DEPTH, TRAVERSAL_COST = 16, 64, F(1.0)


def _half_area(mn, mx):
    d = mx - mn
    return (d[0] + d[1]) * d[2] + d[0] * d[1]


def _empty():
    return np.full(3, 1e34, dtype=F), np.full(3, -1e34, dtype=F)


def build_binned_sah(aabb_min, aabb_max, centers, max_leaf_size=1):
    n = len(centers)
    nodes_min = np.zeros((2 * n - 1, 3), dtype=F)
    nodes_max = np.zeros((2 * n - 1, 3), dtype=F)
    count = np.zeros(2 * n - 1, dtype=np.int32)
    left_first = np.zeros(2 * n - 1, dtype=np.int32)
    for k in range(2 * n - 1):  # BvhNode::new(): empty box, count -1, left_first -1
        nodes_min[k], nodes_max[k] = _empty()
    count[:] = -1
    left_first[:] = -1
    idx = np.arange(n, dtype=np.uint32)
    node_counter = 1
    delta = F(0.0001)
    # root = union_of_list(aabbs).with_offset(1e-4)
    nodes_min[0] = aabb_min.min(axis=0) - delta
    nodes_max[0] = aabb_max.max(axis=0) + delta
    stack = [(0, 0, n, 0)]  # (node, begin, end, depth)
    while stack:
        node, begin, end, depth = stack.pop()
        nodes_min[node] = nodes_min[node] - delta
        nodes_max[node] = nodes_max[node] + delta

        def make_leaf():
            nodes_min[node] = nodes_min[node] - delta
            nodes_max[node] = nodes_max[node] + delta
            left_first[node] = begin
            count[node] = end - begin
        work = end - begin
        if work <= 1 or depth >= MAX_DEPTH:
            make_leaf()
            continue
        mn, mx = nodes_min[node].copy(), nodes_max[node].copy()
        with np.errstate(divide="ignore", invalid="ignore"):
            center_to_bin = (F(1.0) / (mx - mn)) * F(BINS)
        bin_offset = (-mn) * center_to_bin
        ids = idx[begin:end]
        with np.errstate(invalid="ignore"):
            raw = centers[ids] * center_to_bin + bin_offset          # one rounding per operation (float32 arrays)
        raw = np.where(raw > 0, raw, F(0.0))                          # f32::max(x, 0.0); NaN -> 0
        bins_of = np.minimum(BINS - 1, raw.astype(np.int64))          # `as usize`, then min(bin_count - 1, .)
        bmin = np.full((3, BINS, 3), 1e34, dtype=F)
        bmax = np.full((3, BINS, 3), -1e34, dtype=F)
        bcnt = np.zeros((3, BINS), dtype=np.int64)
        for ax in range(3):
            for b in np.unique(bins_of[:, ax]):
                sel = ids[bins_of[:, ax] == b]
                bmin[ax, b] = aabb_min[sel].min(axis=0)
                bmax[ax, b] = aabb_max[sel].max(axis=0)
                bcnt[ax, b] = len(sel)
        best = []
        for ax in range(3):  # find_split
            right_cost = np.full(BINS, np.finfo(F).max, dtype=F)
            cmn, cmx = _empty()
            cnt = 0
            for i in range(BINS - 1, 0, -1):
                cmn, cmx = np.minimum(cmn, bmin[ax, i]), np.maximum(cmx, bmax[ax, i])
                cnt += int(bcnt[ax, i])
                with np.errstate(over="ignore", invalid="ignore"):
                    right_cost[i] = _half_area(cmn, cmx) * F(cnt)
            cmn, cmx = _empty()
            cnt = 0
            best_cost, best_count = np.finfo(F).max, BINS
            for i in range(BINS - 1):
                cmn, cmx = np.minimum(cmn, bmin[ax, i]), np.maximum(cmx, bmax[ax, i])
                cnt += int(bcnt[ax, i])
                with np.errstate(over="ignore", invalid="ignore"):
                    cost = _half_area(cmn, cmx) * F(cnt) + right_cost[i + 1]
                if cost < best_cost:
                    best_cost, best_count = cost, i + 1
            best.append((best_cost, best_count))
        best_axis = 0
        if best[0][0] > best[1][0]:
            best_axis = 1
        if best[best_axis][0] > best[2][0]:
            best_axis = 2
        split_index = best[best_axis][1]
        max_split_cost = _half_area(mn, mx) * (F(work) - TRAVERSAL_COST)
        if best[best_axis][1] == BINS or best[best_axis][0] >= max_split_cost:
            if work > max_leaf_size:
                ext = mx - mn
                a = 0
                if ext[1] > ext[0]:
                    a = 1
                if ext[2] > ext[a]:
                    a = 2
                best_axis = a
                c = 0
                for i in range(BINS - 1):
                    c += int(bcnt[best_axis, i])
                    if c >= (work * 2 // 5 + 1):
                        split_index = i + 1
                        break
            else:
                make_leaf()
                continue
        # utils::partition: elements that pass the check are swapped to the front in encounter order
        goes_left = bins_of[:, best_axis] < split_index
        cnt_left = 0
        seg = idx[begin:end]
        for i in range(work):
            # the check looks at the element CURRENTLY at position i (earlier swaps may have moved a right-goer there)
            if goes_left[i]:
                seg[i], seg[cnt_left] = seg[cnt_left], seg[i]
                goes_left[i], goes_left[cnt_left] = goes_left[cnt_left], goes_left[i]
                cnt_left += 1
        begin_right = begin + cnt_left
        if begin < begin_right < end:
            left = node_counter
            node_counter += 2
            left_first[node] = left
            count[node] = -1
            lmn, lmx = _empty()
            for i in range(best[best_axis][1]):  # quirk: the SAH split count of the final axis, not split_index
                lmn, lmx = np.minimum(lmn, bmin[best_axis, i]), np.maximum(lmx, bmax[best_axis, i])
            rmn, rmx = _empty()
            for i in range(split_index, BINS):
                rmn, rmx = np.minimum(rmn, bmin[best_axis, i]), np.maximum(rmx, bmax[best_axis, i])
            nodes_min[left], nodes_max[left] = lmn, lmx
            nodes_min[left + 1], nodes_max[left + 1] = rmn, rmx
            a_task = (left, begin, begin_right, depth + 1)
            b_task = (left + 1, begin_right, end, depth + 1)
            if (a_task[2] - a_task[1]) < (b_task[2] - b_task[1]):
                a_task, b_task = b_task, a_task
            stack.append(b_task)
            stack.append(a_task)  # popped next: the child with more work runs first
            continue
        make_leaf()
    return nodes_min[:node_counter], nodes_max[:node_counter], count[:node_counter], left_first[:node_counter], idx


@pytest.mark.parametrize("scene,leaf", [("teapot", 1), ("teapot", 4), ("soup900", 1), ("dups", 2)])
def test_python_builder_produces_the_oracles_bytes(O, W, teapot, scene, leaf):
    if scene == "teapot":
        tris = teapot["tris"]
    elif scene == "soup900":
        tris = W.soup(900, seed=0xB11D)
    else:  # identical triangles: unsplittable ranges exercise the fallback, its left-box quirk and the leaf rules
        tris = W.soup(600, seed=0xD0B1).copy()
        tris[100:400] = tris[100]
    aabbs, centers = O.prims_from_triangles(tris)
    rc, want = O.build(O.BINNED_SAH, aabbs, centers, leaf)
    assert rc == 0
    mn, mx, cnt, lf, idx = build_binned_sah(np.ascontiguousarray(aabbs["min"], dtype=F), np.ascontiguousarray(aabbs["max"], dtype=F),
                                            np.ascontiguousarray(centers, dtype=F).reshape(-1, 3), leaf)
    assert len(cnt) == len(want.nodes)
    assert np.array_equal(idx, want.indices), "prim_indices differ"
    assert np.array_equal(cnt, want.nodes["count"]) and np.array_equal(lf, want.nodes["left_first"]), "topology / numbering differs"
    assert mn.view(np.uint32).tobytes() == np.ascontiguousarray(want.nodes["min"]).view(np.uint32).tobytes(), "min corners differ"
    assert mx.view(np.uint32).tobytes() == np.ascontiguousarray(want.nodes["max"]).view(np.uint32).tobytes(), "max corners differ"


# ---- locally-ordered clustering ---------------------------------------------------------------------------------------
# Followed sources: LocallyOrderedClusteringBuilder::build / cluster  src/builders/locb.rs:48-328 (search radius 14, first
# strict minimum scanning j = i-14 .. i-1 then i+1 .. i+14, merge iff mutual, the lower index keeps the parent slot),
# MortonEncoder::{new, encode, get_sorted_indices} + morton_split  src/morton.rs:10-102 (stable sort by code),
# prefix_sum (inclusive)  src/utils.rs:42-58.
def _morton_split(v):
    v = v & 0x3FF
    out = 0
    for b in range(10):
        out |= ((v >> b) & 1) << (3 * b)
    return out


def build_locb(aabb_min, aabb_max, centers):
    n = len(centers)
    delta = F(0.0001)
    wmin, wmax = aabb_min.min(axis=0) - delta, aabb_max.max(axis=0) + delta
    world_to_grid = F(1024.0) * (F(1.0) / (wmax - wmin))
    grid_offset = (-wmin) * world_to_grid
    grid = centers * world_to_grid + grid_offset
    g = np.clip(np.nan_to_num(grid, nan=0.0).astype(np.float64).astype(np.int64), 0, 1023)  # `as i32` truncates; then min / max
    codes = np.array([_morton_split(int(x)) | (_morton_split(int(y)) << 1) | (_morton_split(int(z)) << 2) for x, y, z in g],
                     dtype=np.uint32)
    prim_indices = np.argsort(codes, kind="stable").astype(np.uint32)
    nc = 2 * n - 1
    bmin, bmax = np.zeros((nc, 3), dtype=F), np.zeros((nc, 3), dtype=F)
    count, left_first = np.full(nc, -1, dtype=np.int32), np.zeros(nc, dtype=np.int32)
    begin, end, previous_end = nc - n, nc, nc
    bmin[begin:end] = aabb_min[prim_indices] - delta
    bmax[begin:end] = aabb_max[prim_indices] + delta
    count[begin:end] = 1
    left_first[begin:end] = np.arange(n, dtype=np.int32)
    cur = [bmin, bmax, count, left_first]
    nxt = [a.copy() for a in cur]
    R = 14
    fmax = np.finfo(F).max
    iterations = 0
    while end - begin > 1:
        iterations += 1
        imn, imx, icnt, ilf = cur
        m = end - begin
        lo, hi = imn[begin:end], imx[begin:end]
        # candidate matrix in scan order: columns 0..13 = j = i-14 .. i-1, columns 14..27 = j = i+1 .. i+14
        dist = np.full((m, 2 * R), fmax, dtype=F)
        for d in range(1, R + 1):
            if d >= m:
                break
            umn, umx = np.minimum(lo[:-d], lo[d:]), np.maximum(hi[:-d], hi[d:])
            dd = umx - umn
            ha = (dd[:, 0] + dd[:, 1]) * dd[:, 2] + dd[:, 0] * dd[:, 1]
            dist[d:, R - d] = ha        # backward neighbour j = i - d of i
            dist[:-d, R + d - 1] = ha   # forward neighbour j = i + d of i
        col = np.argmin(dist, axis=1)   # first minimum in scan order == first strict minimum
        off = np.where(col < R, col - R, col - R + 1)
        nb = np.arange(m) + off
        i_all = np.arange(m)
        mutual = nb[nb] == i_all
        merged = np.cumsum((mutual & (i_all < nb)).astype(np.int64))
        merged_count = int(merged[-1])
        unmerged_count = m - merged_count
        children_count = 2 * merged_count
        children_begin = end - children_count
        unmerged_begin = end - (children_count + unmerged_count)
        omn, omx, ocnt, olf = nxt
        for i in range(m):
            j = int(nb[i])
            if mutual[i]:
                if i < j:
                    dst = unmerged_begin + j - int(merged[j])
                    first_child = children_begin + (int(merged[i]) - 1) * 2
                    omn[dst] = np.minimum(lo[j], lo[i])
                    omx[dst] = np.maximum(hi[j], hi[i])
                    ocnt[dst], olf[dst] = -1, first_child
                    for a, b in zip(nxt, cur):
                        a[first_child] = b[begin + i]
                        a[first_child + 1] = b[begin + j]
            else:
                dst = unmerged_begin + i - int(merged[i])
                for a, b in zip(nxt, cur):
                    a[dst] = b[begin + i]
        for a, b in zip(nxt, cur):
            a[end:previous_end] = b[end:previous_end]
        cur, nxt = nxt, cur
        previous_end = end
        begin, end = unmerged_begin, children_begin
    return cur, prim_indices, iterations


@pytest.mark.parametrize("scene", ["teapot", "soup900"])
def test_python_locb_produces_the_oracles_bytes(O, W, teapot, scene):
    tris = teapot["tris"] if scene == "teapot" else W.soup(900, seed=0xB11D)
    aabbs, centers = O.prims_from_triangles(tris)
    rc, want = O.build(O.LOCB, aabbs, centers, 1)
    assert rc == 0
    (mn, mx, cnt, lf), idx, iters = build_locb(np.ascontiguousarray(aabbs["min"], dtype=F), np.ascontiguousarray(aabbs["max"], dtype=F),
                                               np.ascontiguousarray(centers, dtype=F).reshape(-1, 3))
    assert np.array_equal(idx, want.indices), "Morton order differs"
    assert np.array_equal(cnt, want.nodes["count"]) and np.array_equal(lf, want.nodes["left_first"]), "topology differs"
    assert mn.view(np.uint32).tobytes() == np.ascontiguousarray(want.nodes["min"]).view(np.uint32).tobytes()
    assert mx.view(np.uint32).tobytes() == np.ascontiguousarray(want.nodes["max"]).view(np.uint32).tobytes()
    if scene == "teapot":
        assert iters == 37  # the survey's own count


# ---- collapse to the 4-wide layout --------------------------------------------------------------------------------------
# Followed sources: Mbvh::construct  src/bvh.rs:381-404; MbvhNode::merge_nodes  src/mbvh_node.rs:297-411 (all four slots start
# with the PARENT's box; a grandchild's box overwrites it, a direct leaf child keeps it; m-nodes numbered by pool_ptr in
# recursion order); MbvhNode::new  src/mbvh_node.rs:55-78.
def collapse(nodes):
    import sys
    n = len(nodes)
    M = {k: np.full((n, 4), v, dtype=F) for k, v in (("min_x", 1e34), ("min_y", 1e34), ("min_z", 1e34), ("max_x", -1e34),
                                                     ("max_y", -1e34), ("max_z", -1e34))}
    children, counts = np.full((n, 4), -1, dtype=np.int32), np.full((n, 4), -1, dtype=np.int32)
    pool = [1]
    is_leaf = lambda k: int(nodes["count"][k]) >= 0
    lf_of = lambda k: int(nodes["left_first"][k])

    def set_bounds(m, slot, k):
        for a, ax in enumerate("xyz"):
            M["min_" + ax][m, slot] = nodes["min"][k][a]
            M["max_" + ax][m, slot] = nodes["max"][k][a]

    def merge(m, cur):
        for i in range(4):
            set_bounds(m, i, cur)
        sel, leafs = [-1] * 4, [-1] * 4
        if lf_of(cur) >= 0:
            for side, base in ((lf_of(cur), 0), (lf_of(cur) + 1, 2)):
                if side < n and lf_of(side) >= 0:
                    if is_leaf(side):
                        sel[base], leafs[base] = lf_of(side), int(nodes["count"][side])
                    else:
                        for off in (0, 1):
                            g = lf_of(side) + off
                            if is_leaf(g):
                                sel[base + off], leafs[base + off] = lf_of(g), int(nodes["count"][g])
                            else:
                                sel[base + off] = g
                            set_bounds(m, base + off, g)
        for i in range(4):
            node, cnt = sel[i], leafs[i]
            if node >= 0 and cnt >= 0:
                children[m, i], counts[m, i] = node, cnt
                continue
            if node == -1:
                continue
            if is_leaf(node):
                children[m, i], counts[m, i] = lf_of(node), int(nodes["count"][node])
                set_bounds(m, i, node)
            else:
                new_m = pool[0]
                pool[0] += 1
                children[m, i] = new_m
                set_bounds(m, i, node)
                merge(new_m, node)

    sys.setrecursionlimit(10000)
    merge(0, 0)
    k = pool[0]
    return {**{a: b[:k] for a, b in M.items()}, "children": children[:k], "counts": counts[:k]}


@pytest.mark.parametrize("builder", ["sah", "locb"])
def test_python_collapse_produces_the_oracles_bytes(O, teapot, teapot_trees, builder):
    bvh, want = teapot_trees[builder]
    got = collapse(bvh.nodes)
    assert len(got["children"]) == len(want.nodes)
    for key in ("min_x", "max_x", "min_y", "max_y", "min_z", "max_z"):
        assert got[key].view(np.uint32).tobytes() == np.ascontiguousarray(want.nodes[key]).view(np.uint32).tobytes(), key
    assert np.array_equal(got["children"], want.nodes["children"]) and np.array_equal(got["counts"], want.nodes["counts"])


# ---- refit ------------------------------------------------------------------------------------------------------------
# Followed source: Bvh::refit  src/bvh.rs:176-205 — reverse sweep (children sit behind their parent), leaf = union of its
# primitives' new boxes, inner = union of its two children, each padded by 1e-4.  The reference then assigns the fresh Aabb to
# node.bounds, which also zeroes count / left_first (Aabb::new has extras = 0) and leaves a destroyed tree; like the oracle and
# the GPU this restatement keeps the topology fields (the one deliberate deviation).
def refit(nodes, indices, new_min, new_max):
    mn, mx = np.array(nodes["min"], dtype=F), np.array(nodes["max"], dtype=F)
    delta = F(0.0001)
    for i in range(len(nodes) - 1, -1, -1):
        bmn, bmx = _empty()
        lf, cnt = int(nodes["left_first"][i]), int(nodes["count"][i])
        if lf >= 0:
            if cnt >= 0:
                for k in range(cnt):
                    p = int(indices[lf + k])
                    bmn, bmx = np.minimum(bmn, new_min[p]), np.maximum(bmx, new_max[p])
            else:
                bmn, bmx = np.minimum(bmn, mn[lf]), np.maximum(bmx, mx[lf])
                bmn, bmx = np.minimum(bmn, mn[lf + 1]), np.maximum(bmx, mx[lf + 1])
            bmn, bmx = bmn - delta, bmx + delta
        mn[i], mx[i] = bmn, bmx
    return mn, mx


@pytest.mark.parametrize("builder", ["sah", "locb"])
def test_python_refit_produces_the_oracles_bytes(O, W, teapot, teapot_trees, builder):
    bvh, _ = teapot_trees[builder]
    moved = teapot["tris"].copy()
    moved[:, :, 1] += (0.05 * np.sin(7.0 * moved[:, :, 0])).astype(np.float32)  # the same triangles, wobbling
    new_aabbs, _ = O.prims_from_triangles(moved)
    want = bvh.refit(new_aabbs)
    mn, mx = refit(bvh.nodes, bvh.indices, np.ascontiguousarray(new_aabbs["min"], dtype=F), np.ascontiguousarray(new_aabbs["max"], dtype=F))
    assert mn.view(np.uint32).tobytes() == np.ascontiguousarray(want.nodes["min"]).view(np.uint32).tobytes()
    assert mx.view(np.uint32).tobytes() == np.ascontiguousarray(want.nodes["max"]).view(np.uint32).tobytes()
    assert np.array_equal(want.nodes["count"], bvh.nodes["count"]) and np.array_equal(want.nodes["left_first"], bvh.nodes["left_first"])
