"""Pins the ORCHESTRATION of the history oracle to the reference's own main loop.

tests/test_gpu_history.py::cpu_history is this repository's restatement of vpic_simulation::advance()
(src/vpic/advance.cxx:13-244) -- the call order the C++ driver (csrc/vpb_step.cu) and sim.py follow, and what every
GPU history test is compared against.  Here the REAL loop runs: oracle/decks/pin_history.cxx on the reference alone
(main.cxx, initialize(), advance(), scalar hot path) writes the state advance() starts from and, after every step, the
eight numbers of dump_energies as raw doubles.  cpu_history, fed that state and the oracle kernels, must reproduce the
20-step history: the kernels are bit-exact per call (tests/test_oracle_vs_ref.py), every species holds a multiple of 16
particles (all through pipeline 0: float sums in array order), so any difference beyond the fp64 summation order of
the energy diagnostics means the restated call order is not the reference's."""
import os
import subprocess

import numpy as np
import pytest

from helpers import abi, host_grid, loader
from test_gpu_history import SORT, STEPS, cpu_history, oracle_kernels

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "pin_history.op")

pytestmark = pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/pin_history.op not built")


def read_state(path):
    raw = open(path, "rb").read()
    nsp, nv = np.frombuffer(raw, np.int32, 2)
    off = 8
    f = abi.aligned_zeros(int(nv), abi.field_dtype)
    f.view(np.uint8)[:] = np.frombuffer(raw, np.uint8, int(nv) * 80, off)
    off += int(nv) * 80
    species = []
    for _ in range(int(nsp)):
        n = int(np.frombuffer(raw, np.int32, 1, off)[0])
        q_m = float(np.frombuffer(raw, np.float32, 1, off + 4)[0])
        p = abi.aligned_zeros(n, abi.particle_dtype)
        p.view(np.uint8)[:] = np.frombuffer(raw, np.uint8, n * 48, off + 8)
        off += 8 + n * 48
        species.append({"p": p, "q_m": q_m})
    assert off == len(raw)
    return f, species


@pytest.mark.parametrize("clean,sync", [(5, 0), (5, 4), (0, 3)])
def test_cpu_history_is_the_reference_main_loop(orc, tmp_path, clean, sync):
    """clean: both divergence-cleaning intervals (with the err>0 guards of advance.cxx:164-172,185-193); sync:
    sync_shared_interval (advance.cxx:199-208)."""
    assert (STEPS, SORT) == (20, 5)          # what the deck was written for
    env = dict(os.environ, VPB_PIN_SYNC=str(sync), VPB_PIN_CLEAN=str(clean))
    r = subprocess.run([EXE, "-tpp=1"], cwd=tmp_path, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    want = np.fromfile(tmp_path / "hist.bin", np.float64).reshape(-1, 8)
    assert want.shape == (STEPS + 1, 8)
    f0, species = read_state(tmp_path / "state0.bin")
    assert [len(s["p"]) % 16 for s in species] == [0, 0] and [s["q_m"] for s in species] == [1.0, -1.0]   # ion first: list order
    g = host_grid((12, 10, 8), "periodic")
    errors = []
    got = cpu_history(oracle_kernels(orc), g, species, STEPS, clean, clean, f_init=f0, sync_shared=sync, errors=errors)
    assert len(errors) == len([k for k in range(STEPS) if (clean and k % clean == 0) or (sync and k % sync == 0)])
    if sync:     # one rank: both copies of a periodic face see the same arithmetic, the desynchronisation error is exactly 0
        assert all(e["desync"] == 0 for e in errors if "desync" in e)
    assert got.shape == (STEPS, 8)
    rel = np.abs(got - want[1:]) / np.abs(want[1:]).max(axis=0)
    assert rel.max() < 1e-12, rel.max(axis=0)
    assert want[1:, :6].sum() > 0 and np.all(want[:, 6:] > 0)
