"""The library's host-staged transport (old_vpic_b200/csrc/vpb_mp_transport.hpp: exchanges through the HOST PROGRAM'S
message layer, used when ranks share a GPU) on CPU ranks: tests/mp_transport_harness.cpp includes that header, links the
reference's own mp_dmp (oracle/_ref/hybrid/libvpic_ref_scalar.a) over oracle/mpi_shim and replays the library's exchange
pattern -- sends by face 0..5, receives by face 3,4,5,0,1,2, two faces towards the same peer when an axis has two
ranks, empty messages, MB-sized messages, mp_allsum_d -- checking every word that arrives."""
import os
import subprocess

import pytest

from test_ref_multirank import run_ranks

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HOST = os.path.join(ROOT, "oracle", "_ref", "hybrid", "libvpic_ref_scalar.a")

pytestmark = pytest.mark.skipif(not os.path.exists(HOST), reason="oracle/_ref/hybrid not built")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("mpt") / "harness")
    cmd = ["g++", "-std=c++11", "-O1", "-w", "-I" + os.path.join(ROOT, "old_vpic_b200", "csrc"),
           os.path.join(HERE, "mp_transport_harness.cpp"), HOST, "-ldl", "-lm", "-lpthread", "-rdynamic", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


@pytest.mark.parametrize("topo", [(2, 1, 1), (2, 2, 1), (3, 1, 2), (1, 1, 4)])
def test_exchange_pattern_over_the_reference_mp_layer(harness, topo):
    world = topo[0] * topo[1] * topo[2]
    run_ranks(world, {"VPIC_SHIM_SLOT_MB": "8"}, argv=[harness] + [str(t) for t in topo], marker="MP_TRANSPORT_OK rank=%d")
