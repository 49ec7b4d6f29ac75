import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """A clean checkout has no built artefacts (*.so are git-ignored): build the C-ABI library and the oracle once."""
    lib_path = os.path.join(ROOT, "phase2_bn254_b200", "libp2b.so")
    if not os.path.exists(lib_path):
        import subprocess
        subprocess.check_call([sys.executable, "-c", "import __graft_entry__ as g; g.build()"], cwd=ROOT)


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
