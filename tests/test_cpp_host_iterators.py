"""The host side of include/rtbvh.hpp on the CPU: the `&T` iterators (traverse_iter / traverse_iter_packet over
Bvh::from_raw / Mbvh::from_raw trees) driving SpatialTriangle::intersect / intersect4, i.e. the loops of
examples/benchmark.rs:25-31 and :55-61 written against the C++ mirror — compared bit for bit (t) with the CPU oracle's own
walk of the same reference-format trees.  No GPU call is made by the program."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_iterators.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_iterators")


@pytest.fixture(scope="module")
def exe():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    lib = os.path.join(ROOT, "rtbvh_b200")
    subprocess.run([cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
                    "-L", lib, "-lrtbvh_rs", f"-Wl,-rpath,{lib}"], check=True)
    return EXE


@pytest.mark.parametrize("name", ["sah", "locb"])
def test_host_iterator_loops_equal_the_oracle(exe, tmp_path, O, W, teapot, teapot_trees, name):
    from rtbvh_b200 import api as A
    tris = teapot["tris"]
    bvh, m = teapot_trees[name]
    cam = W.camera_rays(W.benchmark_camera(96, 96))
    rnd = W.random_rays(6000, *W.bounds(tris))
    rays = np.concatenate([cam, rnd])
    rays = rays[: len(rays) // 4 * 4]
    packets = np.concatenate([W.pack4(cam[: len(cam) // 4 * 4]), W.pack4(rnd[: len(rnd) // 4 * 4])])
    np.ascontiguousarray(bvh.nodes).tofile(tmp_path / "bvh_nodes.bin")
    np.ascontiguousarray(m.nodes).tofile(tmp_path / "mbvh_nodes.bin")
    np.ascontiguousarray(bvh.indices, dtype=np.uint32).tofile(tmp_path / "indices.bin")
    np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 9).tofile(tmp_path / "tris.bin")
    np.ascontiguousarray(rays, dtype=A.RAY_DTYPE).tofile(tmp_path / "rays.bin")
    np.ascontiguousarray(packets, dtype=A.PACKET_DTYPE).tofile(tmp_path / "packets.bin")
    assert np.array_equal(bvh.indices, m.indices)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    for tree, tag in ((bvh, "bvh"), (m, "mbvh")):
        want, _, _ = O.trace(tree, tris, rays)
        got = np.fromfile(tmp_path / f"out_{tag}_single.bin", dtype=A.HIT_DTYPE)
        assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32)), f"{name}/{tag}: single-ray t differs"
        assert np.array_equal(got["prim"] == A.NO_HIT, want["prim"] == A.NO_HIT)
        # ids: the reference loop keeps the FIRST primitive that reaches the final t, the oracle reports the lowest id
        # among exactly equal t (north-star tie rule) - they may differ only on exact ties
        differ = got["prim"] != want["prim"]
        assert differ.mean() < 0.01
        wantp, _, _ = O.trace_packets(tree, tris, packets)
        gotp = np.fromfile(tmp_path / f"out_{tag}_packet.bin", dtype=A.HIT4_DTYPE)
        assert np.array_equal(gotp["t"].view(np.uint32), wantp["t"].view(np.uint32)), f"{name}/{tag}: packet t differs"
        assert np.array_equal(gotp["prim"] == A.NO_HIT, wantp["prim"] == A.NO_HIT)
        assert (gotp["prim"] != wantp["prim"]).mean() < 0.01
    assert (want["prim"] != A.NO_HIT).mean() > 0.3  # the sample does hit the teapot
