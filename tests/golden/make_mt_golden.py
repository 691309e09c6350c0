#!/usr/bin/env python
"""Regenerates tests/golden/ref_mt_stream.npz and ref_thermal_c1_load.npz from the REFERENCE ALONE (development container:
oracle/_ref built from /root/reference):

  ref_mt_stream.npz        util/mtrand/mtrand.c itself through oracle/_ref/libvpic_ref_scalar.so: for seeds 0, 7 and
                           0xfffffffe the first 2000 words of mt_urand, then 500 mt_drand, then 4000 mt_drandn (continuing
                           one stream per seed); and for seed 7 the 55 tail-layer deviates among the first 250000 mt_drandn
                           with their positions
  ref_thermal_c1_load.npz  the two particle arrays vpic_simulation::initialize() leaves after the load loop of
                           oracle/decks/thermal_c1.cxx (seed_rand(7); 6^3 cells x 5 per cell), dumped by the deck itself

tests/test_gpu_mt.py compares the device stream and the device load with them; tests/test_oracle_mt.py the oracle.
usage: python tests/golden/make_mt_golden.py"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import loader  # noqa: E402
from old_vpic_b200 import abi  # noqa: E402


def main():
    R = loader.ref("scalar")
    R.new_mt_rng.restype, R.new_mt_rng.argtypes = C.c_void_p, [C.c_uint]
    for name in ("mt_urand_fill", "mt_drand_fill", "mt_drandn_fill"):
        getattr(R, name).restype, getattr(R, name).argtypes = None, [C.c_void_p, C.c_void_p, C.c_size_t]
    out = {}
    for seed in (0, 7, 0xfffffffe):
        r = C.c_void_p(R.new_mt_rng(seed))
        w, u, n = np.zeros(2000, np.uint32), np.zeros(500), np.zeros(4000)
        R.mt_urand_fill(r, w.ctypes.data, len(w))
        R.mt_drand_fill(r, u.ctypes.data, len(u))
        R.mt_drandn_fill(r, n.ctypes.data, len(n))
        out["words_%d" % seed], out["drand_%d" % seed], out["drandn_%d" % seed] = w, u, n
    r = C.c_void_p(R.new_mt_rng(7))
    big = np.zeros(250000)
    R.mt_drandn_fill(r, big.ctypes.data, len(big))
    where = np.flatnonzero(np.abs(big) > 3.6554204190269413)
    out["tail_where_7"], out["tail_value_7"] = where.astype(np.int64), big[where]
    np.savez_compressed(os.path.join(HERE, "ref_mt_stream.npz"), **out)
    print("ref_mt_stream.npz:", {k: v.shape for k, v in out.items()})

    cells, ppc = 6, 5
    with tempfile.TemporaryDirectory() as work:
        dump = os.path.join(work, "load.bin")
        env = dict(os.environ, VPB_DECK_CELLS=str(cells), VPB_DECK_PPC=str(ppc), VPB_DECK_STEPS="1", VPB_DECK_DUMP_LOAD=dump, VPB_DECK_ENERGIES="0")
        subprocess.run([os.path.join(loader.REF_DIR, "thermal_c1.op"), "-tpp=1"], cwd=work, env=env, check=True, capture_output=True, timeout=300)
        raw = np.fromfile(dump, np.uint8)
    n = cells ** 3 * ppc
    assert tuple(raw[:8].view(np.int32)) == (n, n)
    e = raw[8:8 + 48 * n].view(abi.particle_dtype)
    i = raw[8 + 48 * n:8 + 96 * n].view(abi.particle_dtype)
    np.savez_compressed(os.path.join(HERE, "ref_thermal_c1_load.npz"), cells=cells, ppc=ppc, seed=7, vth=0.1, electron=e, ion=i)
    print("ref_thermal_c1_load.npz:", n, "particles per species")


if __name__ == "__main__":
    main()
