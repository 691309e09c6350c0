"""The reference itself on W>1 ranks (separate processes over oracle/mpi_shim's shared-memory transport) against
the oracle's multi-rank pieces, bit for bit -- see tests/ref_mpi_worker.py.  CPU only; skipped where oracle/_ref
was not built (it is built in the dev container by __graft_entry__.build())."""
import os
import subprocess
import sys
import tempfile
import time

import pytest

from helpers import loader

HERE = os.path.dirname(os.path.abspath(__file__))

pytestmark = pytest.mark.skipif(not loader.ref_available("scalar"), reason="oracle/_ref not built")


def run_ranks(world, env_extra, timeout=240, argv=None, cwd=None, marker="REF_MPI_OK rank=%d"):
    argv = argv or [sys.executable, os.path.join(HERE, "ref_mpi_worker.py")]
    with tempfile.NamedTemporaryFile(prefix="vpic_shim_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as shm:
        procs = []
        for r in range(world):
            env = dict(os.environ, VPIC_SHIM_NPROC=str(world), VPIC_SHIM_RANK=str(r), VPIC_SHIM_SHM=shm.name,
                       VPIC_SHIM_SLOT_MB="1")
            env.update(env_extra)
            procs.append(subprocess.Popen(argv, env=env, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        t0, failed = time.time(), False
        while any(p.poll() is None for p in procs):
            if any(p.poll() not in (None, 0) for p in procs) or time.time() - t0 > timeout:
                failed = True       # a rank died: its peers would spin in the shim for ever
                for p in procs:
                    if p.poll() is None:
                        p.kill()
                break
            time.sleep(0.05)
        outs = [p.communicate()[0] for p in procs]
    assert not failed and all(p.returncode == 0 for p in procs), "\n".join(o[-3000:] for o in outs)
    for r, o in enumerate(outs):
        assert marker is None or marker % r in o, o[-3000:]
    return outs


@pytest.mark.parametrize("topo,kind,gn", [((2, 1, 1), "periodic", (8, 6, 4)), ((1, 1, 2), "absorbing", (5, 4, 6)),
                                           ((1, 2, 1), "metal", (4, 8, 1)), ((2, 2, 1), "periodic", (8, 6, 3)),
                                           ((2, 2, 2), "sheet", (8, 6, 6))])
def test_reference_on_ranks_matches_oracle(topo, kind, gn):
    run_ranks(topo[0] * topo[1] * topo[2],
              {"REFW_TOPO": ",".join(map(str, topo)), "REFW_KIND": kind, "REFW_GN": ",".join(map(str, gn))})


def read_energies(path):
    import numpy as np
    return np.array([[float(x) for x in line.split()] for line in open(path) if line.strip() and not line.startswith("%")])


@pytest.mark.parametrize("world", [2, 4])
def test_reference_deck_on_ranks(world, tmp_path):
    """A whole reference program (main.cxx + vpic_simulation + oracle/decks/thermal_small.cxx, hot path in the scalar
    flavour) on `world` processes: every rank draws the same particles and keeps its x-slab, so the energies must
    be those of the committed one-rank run up to the order of the float sums (measured: 8e-7)."""
    import numpy as np
    exe = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "thermal_small.op")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/thermal_small.op not built")
    run_ranks(world, {"VPIC_SHIM_SLOT_MB": "2"}, argv=[exe, "-tpp=1"], cwd=str(tmp_path), marker=None)
    got = read_energies(tmp_path / "energies")
    want = read_energies(os.path.join(HERE, "golden", "deck_thermal_small_energies.txt"))
    assert got.shape == want.shape == (21, 9)
    rel = np.abs(got[:, 1:] - want[:, 1:]) / np.abs(want[:, 1:]).max(axis=0)
    assert rel.max() < 1e-5, rel.max(axis=0)


def test_reference_deck_grows_tight_arrays(tmp_path):
    """VPB_DECK_MAXNP=16450 leaves each rank 66 particles of head-room: boundary_p has to grow the arrays
    (boundary_p.c:416-447).  Here the pure reference does it; tests/test_gpu_deck.py asks the same of the library."""
    import numpy as np
    exe = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "thermal_small.op")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/thermal_small.op not built")
    outs = run_ranks(2, {"VPIC_SHIM_SLOT_MB": "2", "VPB_DECK_MAXNP": "16450"}, argv=[exe, "-tpp=1"], cwd=str(tmp_path), marker=None)
    assert sum(o.count("Resizing local") for o in outs) >= 1
    got = read_energies(tmp_path / "energies")
    want = read_energies(os.path.join(HERE, "golden", "deck_thermal_small_energies.txt"))
    rel = np.abs(got[:, 1:] - want[:, 1:]) / np.abs(want[:, 1:]).max(axis=0)
    assert rel.max() < 1e-5, rel.max(axis=0)


@pytest.mark.parametrize("deck", ["absorb_small", "sheet_small"])
def test_reference_wall_decks_on_two_ranks(deck, tmp_path):
    """The decks with walls, split along x over two reference ranks: six absorbing faces with Higdon fields and
    absorbed particles (absorb_small), and the trecon-part geometry with conducting reflecting z walls, a force-free
    sheet and a hydro dump (sheet_small).  Energies of the one-rank run, particle counts exactly."""
    import numpy as np
    exe = os.path.join(os.path.dirname(HERE), "oracle", "_ref", deck + ".op")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/%s.op not built" % deck)
    run_ranks(2, {"VPIC_SHIM_SLOT_MB": "2"}, argv=[exe, "-tpp=1"], cwd=str(tmp_path), marker=None)
    got = read_energies(tmp_path / "energies")
    want = read_energies(os.path.join(HERE, "golden", "deck_%s_energies.txt" % deck))
    scale = np.abs(want[:, 1:]).max(axis=0)
    scale[:6] = want[:, 1:7].sum(axis=1).max()
    assert (np.abs(got[:, 1:] - want[:, 1:]) / scale).max() < 1e-5
    if deck == "absorb_small":
        tot = {}
        for r in range(2):
            for line in open(tmp_path / ("counts.%d" % r)):
                k, v = line.split()
                tot[k] = tot.get(k, 0) + int(v)
        want_counts = {k: int(v) for k, v in (line.split() for line in open(os.path.join(HERE, "golden", "deck_absorb_small_counts.txt")))}
        assert tot == want_counts
