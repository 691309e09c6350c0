import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle import loader
    return loader.oracle()


@pytest.fixture(scope="session")
def ref_scalar():
    """The reference compiled from source, scalar flavour (skips if oracle/_ref was not built)."""
    from oracle import loader
    if not loader.ref_available("scalar"):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return loader.ref("scalar", tpp=1)


@pytest.fixture(scope="session")
def vpb():
    """libvpic_b200.so bound to cuda:0 (GPU tests only)."""
    from old_vpic_b200 import lib
    L = lib.load()
    L.vpb_init(0)
    return L
