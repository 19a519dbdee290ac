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
