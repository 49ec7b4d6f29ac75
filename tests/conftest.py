import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as oc
    oc.build()
    return oc


@pytest.fixture(scope="session")
def ctx():
    """The product context: libp2b.so on cuda:0.  Fails loudly when the library or the GPU is missing."""
    from phase2_bn254_b200 import lib
    c = lib.Context(0)
    yield c
    c.close()
