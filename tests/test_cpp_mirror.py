"""Compiles and runs tests/cpp/test_reference_api.cpp: the reference's own tests written against the C++ mirror of its
API (include/rtbvh.hpp), linked against librtbvh_rs.so.  The compile step alone runs without a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_reference_api")


def _compile():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    lib = os.path.join(ROOT, "rtbvh_b200")
    subprocess.run([cxx, "-std=c++17", "-O1", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "test_reference_api.cpp"), "-o", EXE, "-L", lib, "-lrtbvh_rs",
                    f"-Wl,-rpath,{lib}"], check=True)


def test_cpp_mirror_compiles_and_links():
    _compile()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_reference_tests_through_cpp_mirror():
    srcs = [os.path.join(ROOT, "tests", "cpp", "test_reference_api.cpp"), os.path.join(ROOT, "include", "rtbvh.hpp"),
            os.path.join(ROOT, "include", "rtbvh_iter.hpp"), os.path.join(ROOT, "include", "rtbvh_gpu.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(f) > os.path.getmtime(EXE) for f in srcs):
        _compile()  # a binary older than its sources would test yesterday's code
    r = subprocess.run([EXE, os.path.join(ROOT, "tests", "golden", "teapot_tris.npy")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok:")
