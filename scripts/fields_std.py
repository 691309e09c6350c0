#!/usr/bin/env python
"""The standard (material-table) advance_e alone on an n^3 periodic field-only box, two materials so that the kernel reads
every voxel's material ids (the `standard` leg of bench.py's fields_c2): a few steps, timed per launch.  For ncu captures:
    ncu --set full -k regex:advance_e_kernel -c 2 ... python scripts/fields_std.py 512
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from old_vpic_b200 import grid as G, lib  # noqa: E402
from old_vpic_b200.sim import NativeSimulation  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    L = lib.load()
    L.vpb_init(0)
    g = G.make_grid((n, n, n), "periodic", field_only=True)
    sim = NativeSimulation(g, n_mat=2, vacuum=False, L=L)
    L.vpb_load_plane_wave(sim.dom, sim.field_ptr, 8, 1.0)
    for _ in range(2):
        sim.advance()
    L.vpb_sync()
    L.vpb_prof_enable(1)
    for _ in range(steps):
        sim.advance()
    tot, cnt = C.c_double(0), C.c_int(0)
    L.vpb_prof_collect(3, C.byref(tot), C.byref(cnt), 0)
    ms = tot.value / max(cnt.value, 1)
    print("advance_e standard %d^3: %.3f ms per launch, %.1f GB/s algorithmic (84 B/cell), %.1f GB/s layout (112 B/cell)"
          % (n, ms, 84.0 * n ** 3 / ms / 1e6, 112.0 * n ** 3 / ms / 1e6))
    sim.free()


if __name__ == "__main__":
    main()
