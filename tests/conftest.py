import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib_built():
    """libvipe_ba.so, built on demand (nvcc cross-compiles sm_100a without a GPU)."""
    from vipe_b200 import build

    return build.build()


@pytest.fixture(scope="session")
def problems():
    from vipe_b200.synthetic import make_problem

    cache = {}

    def get(name, **kw):
        key = (name, tuple(sorted(kw.items())))
        if key not in cache:
            cache[key] = make_problem(name, **kw)
        return cache[key]

    return get
