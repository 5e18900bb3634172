import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def engine_ctx():
    """One engine context per test session (GPU tests only)."""
    from feddlib_b200 import build as fb_build
    if not os.path.exists(fb_build.SO):
        fb_build.build()
    from feddlib_b200 import Context
    return Context(0)


@pytest.fixture(autouse=True)
def _default_scatter_mode(request):
    """Every GPU test starts and ends in the default scatter mode (gather): what a test covers must not depend on
    the mode the previous test left in the session-scoped context."""
    if request.node.get_closest_marker("gpu") is None:
        yield
        return
    ctx = request.getfixturevalue("engine_ctx")
    ctx.set_scatter_mode("gather")
    yield
    ctx.set_scatter_mode("gather")
