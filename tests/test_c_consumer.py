"""A plain C11 translation unit against include/rtbvh.h + include/rtbvh_gpu.h (tests/c/ffi_consumer.c): the headers
must be valid C under -Wall -Wextra -Wpedantic -Werror, the POD sizes of rtbvh_ffi's `same_size` test hold as
_Static_asserts, and the program links against librtbvh_rs.so.  Without a GPU it checks that the library refuses to
build (no CPU fallback); on the B200 it runs rtbvh_ffi's `create_delete` and `intersect` tests through the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "ffi_consumer.c")
EXE = os.path.join(ROOT, "tests", "c", "ffi_consumer")


def _compile():
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    lib = os.path.join(ROOT, "rtbvh_b200")
    subprocess.run([cc, "-std=c11", "-Wall", "-Wextra", "-Wpedantic", "-Werror", "-O1", "-ffp-contract=off", "-I",
                    os.path.join(ROOT, "include"), SRC, "-o", EXE, "-L", lib, "-lrtbvh_rs", f"-Wl,-rpath,{lib}", "-lm"],
                   check=True)


def _fresh():
    deps = [SRC, os.path.join(ROOT, "include", "rtbvh.h"), os.path.join(ROOT, "include", "rtbvh_gpu.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(f) > os.path.getmtime(EXE) for f in deps):
        _compile()


def test_headers_are_valid_c_and_no_device_means_error():
    _compile()
    from rtbvh_b200 import api
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present: the GPU flavour of this test runs the program")
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok: no CUDA device")


@pytest.mark.gpu
def test_reference_ffi_tests_from_plain_c():
    _fresh()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok: create_delete")


CXX_SRC = os.path.join(ROOT, "tests", "cpp", "ffi_cxx_caller.cpp")
CXX_EXE = os.path.join(ROOT, "tests", "cpp", "ffi_cxx_caller")


def _compile_cxx():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    lib = os.path.join(ROOT, "rtbvh_b200")
    subprocess.run([cxx, "-std=c++17", "-Wall", "-Wextra", "-Werror", "-O1", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                    CXX_SRC, "-o", CXX_EXE, "-L", lib, "-lrtbvh_rs", f"-Wl,-rpath,{lib}"], check=True)


def test_reference_cxx_header_spellings_compile():
    """rtbvh_ffi also generates a C++ header (namespace rtbvh, guard RTBVH_HPP): a caller written against it compiles
    unchanged against include/rtbvh.hpp; without a GPU it sees the builders refuse."""
    _compile_cxx()
    from rtbvh_b200 import api
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present: the GPU flavour runs the program")
    r = subprocess.run([CXX_EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok: no CUDA device")


@pytest.mark.gpu
def test_reference_cxx_header_caller_runs():
    deps = [CXX_SRC, os.path.join(ROOT, "include", "rtbvh.hpp"), os.path.join(ROOT, "include", "rtbvh.h")]
    if not os.path.exists(CXX_EXE) or any(os.path.getmtime(f) > os.path.getmtime(CXX_EXE) for f in deps):
        _compile_cxx()
    r = subprocess.run([CXX_EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok: create / collapse")
