import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracing-optimized_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Build the product libraries and the CPU oracle once (cheap when up to date)."""
    import __graft_entry__ as ge
    ge.build()


@pytest.fixture(scope="session")
def crt(built):
    import crt_b200
    return crt_b200


@pytest.fixture(scope="session")
def oracle(built):
    import oracle as o
    return o


@pytest.fixture(scope="session")
def small_scene(crt):
    """The golden-vector scene: staircase at detail 0.1, 32x32 textures, 5 triangles per leaf."""
    return crt.Scene.staircase(0.1, 32, 5)


@pytest.fixture(scope="session")
def medium_scene(crt):
    return crt.Scene.staircase(0.25, 64, 5)
