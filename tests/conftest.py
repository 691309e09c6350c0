import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# Tests that compare two realisations of a many-step run (the float sums of the deposits are ordered differently on every
# GPU run, so one particle's cell crossing or wall hit can fall a step apart): their bounds carry head-room, but under
# `pytest -x` a failure there must not hide the per-call parity tests, so they are collected last.
_LAST = ("test_trecon_part_deck_as_shipped", "test_trecon_part_deck_at_a_scaled_configs2_shape",
         "test_reference_deck_with_walls_sheet_and_hydro_dump", "test_reference_deck_with_absorbing_walls",
         "test_reference_wall_decks_on_two_ranks")


def pytest_collection_modifyitems(config, items):
    items.sort(key=lambda it: 1 if it.name.split("[")[0] in _LAST else 0)


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
