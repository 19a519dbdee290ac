import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure only)."""
    from oracle import oracle
    oracle.build_lib()
    return oracle


@pytest.fixture(scope="session")
def W():
    from rtbvh_b200 import workloads
    return workloads


@pytest.fixture(scope="session")
def teapot(O, W):
    tris = W.teapot()
    aabbs, centers = O.prims_from_triangles(tris)
    return dict(tris=tris, aabbs=aabbs, centers=centers)


@pytest.fixture(scope="session")
def teapot_trees(O, teapot):
    out = {}
    for name, kind in (("sah", O.BINNED_SAH), ("locb", O.LOCB)):
        rc, bvh = O.build(kind, teapot["aabbs"], teapot["centers"], 1)
        assert rc == 0
        out[name] = (bvh, bvh.collapse())
    return out
