#!/usr/bin/env python
"""Regenerates the deck fixtures of tests/golden/ from the REFERENCE ALONE (run in the development container, where
oracle/build_ref.sh + oracle/build_hybrid.sh have compiled /root/reference into oracle/_ref):

  deck_thermal_small_energies.txt, deck_sheet_small_energies.txt, deck_sheet_small_ehydro.{hdr,npz},
  deck_absorb_small_{energies,counts}.txt   oracle/decks/*.cxx on oracle/_ref/<deck>.op, one rank, -tpp=1
  deck_turbulence_energies.txt              decks/trecon-part/turbulence.cxx AS SHIPPED (config.h: 16x16x1 cells,
                                            topology 2x2x1, 2500 steps) on oracle/_ref/turbulence.op, four ranks over
                                            oracle/mpi_shim's shared-memory transport
  deck_turbulence_c2s_energies.txt          the same deck source with the config.h knobs of a scaled-down BASELINE
                                            configs[2] (128 x 1 x 64 cells, one rank, 200 steps; the directory of
                                            symlinks + generated config.h that oracle/build_hybrid.sh makes) on
                                            oracle/_ref/turbulence_c2s.op

The hot path of these executables is the reference's scalar flavour, the one libvpic_b200 reproduces bit for bit per
call.  usage: python tests/golden/make_deck_golden.py [--check]   (--check: compare instead of overwrite)"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_one(deck, work):
    subprocess.run([os.path.join(REF, deck + ".op"), "-tpp=1"], cwd=work, check=True, capture_output=True, timeout=900)


def run_ranks(exe, world, work):
    from test_ref_multirank import run_ranks as rr
    rr(world, {"VPIC_SHIM_SLOT_MB": "2"}, timeout=900, argv=[exe, "-tpp=1"], cwd=work, marker=None)


def emit(src, name, check):
    dst = os.path.join(HERE, name)
    if check:
        same = open(src, "rb").read() == open(dst, "rb").read()
        print(("same     " if same else "DIFFERS  ") + name)
        return same
    shutil.copyfile(src, dst)
    print("wrote    " + name)
    return True


def main():
    check = "--check" in sys.argv
    ok = True
    with tempfile.TemporaryDirectory() as t:
        for deck in ("thermal_small", "sheet_small", "absorb_small"):
            w = os.path.join(t, deck)
            os.mkdir(w)
            run_one(deck, w)
            ok &= emit(os.path.join(w, "energies"), "deck_%s_energies.txt" % deck, check)
            if deck == "absorb_small":
                ok &= emit(os.path.join(w, "counts"), "deck_absorb_small_counts.txt", check)
            if deck == "sheet_small":      # dump_hydro: 123-byte header, then hydro_t[(nx+2)(ny+2)(nz+2)]
                raw = open(os.path.join(w, "ehydro.0"), "rb").read()
                nvox = 1404
                hb = len(raw) - nvox * 64
                open(os.path.join(w, "hdr"), "wb").write(raw[:hb])
                ok &= emit(os.path.join(w, "hdr"), "deck_sheet_small_ehydro.hdr", check)
                hyd = np.frombuffer(raw[hb:], dtype=np.float32).reshape(-1, 16)
                if check:
                    z = np.load(os.path.join(HERE, "deck_sheet_small_ehydro.npz"))
                    same = int(z["header_bytes"]) == hb and np.array_equal(z["hydro"].view(np.uint32), hyd.view(np.uint32))
                    print(("same     " if same else "DIFFERS  ") + "deck_sheet_small_ehydro.npz")
                    ok &= same
                else:
                    np.savez_compressed(os.path.join(HERE, "deck_sheet_small_ehydro.npz"), hydro=hyd, header_bytes=np.int64(hb))
                    print("wrote    deck_sheet_small_ehydro.npz")
        w = os.path.join(t, "turbulence")
        os.mkdir(w)
        run_ranks(os.path.join(REF, "turbulence.op"), 4, w)
        ok &= emit(os.path.join(w, "rundata", "energies"), "deck_turbulence_energies.txt", check)
        w = os.path.join(t, "turbulence_c2s")
        os.mkdir(w)
        run_one("turbulence_c2s", w)
        ok &= emit(os.path.join(w, "rundata", "energies"), "deck_turbulence_c2s_energies.txt", check)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
