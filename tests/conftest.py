import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def ddf():
    """The product package; the CUDA library must be present (no fallback)."""
    import ddf.jl_b200 as pkg
    pkg.load()
    return pkg


@pytest.fixture(scope="session")
def ctx(ddf):
    c = ddf.Context(0)
    ddf.set_default_context(c)
    return c
