import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ctx():
    """One CUDA context of the product library for the GPU tests."""
    from realisticaudioraytracing2d_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):      # normally built by __graft_entry__.build(); nvcc is on the box too
        from realisticaudioraytracing2d_b200 import build as _b
        _b.build()
    c = _capi.Context(0)
    yield c
    c.destroy()
