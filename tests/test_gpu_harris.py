"""GPU: the geometry of decks/trecon-part (BASELINE configs[2], scaled down): a 2D nx x 1 x nz box, periodic in x and
y, conducting walls that reflect particles at z = 0 and z = Lz (turbulence.cxx:262-270), a force-free current sheet
Bx = b0 tanh(z/L), By = b0 sech(z/L) (turbulence.cxx:441-460) and a hot pair plasma with the electrons drifting along
the sheet.  The C++ time-step driver must reproduce the CPU oracle's energy history from the same particles and
fields, reflect every particle that reaches a wall (no movers, none lost) and keep the wall conditions (tangential E
and normal B zero on the walls).

Tolerance.  This plasma is hot (vth 0.3 c) in a small box with walls, so now and then ONE particle's in-cell test
(advance_p.cxx:124-125) falls the other way in the two runs -- their currents are added in different orders, their
fields differ in the last bit -- and that particle deposits into the neighbouring cell for a step (SURVEY.md 8a,
"exactness hot spots"; seen here at step 5 of the second case, scripts/dbg_harris4.py).  It shifts the three small
field-energy columns (ex, ez, cbz: 1e-2 of the field energy) by 1e-3 of THEIR size and nothing else.  Field columns are
therefore compared on the scale of the total field energy, kinetic columns on their own: 1e-4 over 20 steps."""
import os

import numpy as np
import pytest

from helpers import abi, host_grid, random_particles
from test_gpu_history import SORT, STEPS, cpu_history, oracle_kernels
from old_vpic_b200.sim import NativeSimulation

pytestmark = pytest.mark.gpu


def trecon_grid(nx, nz):
    g = host_grid((nx, 1, nz), "periodic")
    for s in (-1, 1):
        g.set_fbc(abi.boundary(0, 0, s), abi.PEC_FIELDS)
        g.set_pbc(abi.boundary(0, 0, s), abi.REFLECT_PARTICLES)
    return g


def sheet_fields(g, b0=0.8, width=2.0):
    nx, ny, nz = g.n
    f = abi.aligned_zeros(g.nv, abi.field_dtype)
    z = (np.arange(g.nv) // ((nx + 2) * (ny + 2))).astype(np.float64)
    zc = (z - 0.5) - 0.5 * nz                       # cell centres, sheet in the middle of the box
    f["cbx"] = (b0 * np.tanh(zc / width)).astype(np.float32)
    f["cby"] = (b0 / np.cosh(zc / width)).astype(np.float32)
    return f


def sheet_species(g, ppc, seed, drift=0.2, width=2.0):
    rng = np.random.default_rng(seed)
    n = g.n[0] * g.n[1] * g.n[2] * ppc
    e = random_particles(rng, g, n, vth=0.3, sort=True, q=-1.0 / ppc)
    i = random_particles(rng, g, n, vth=0.3, sort=True, q=+1.0 / ppc)
    i["dx"], i["dy"], i["dz"], i["i"] = e["dx"], e["dy"], e["dz"], e["i"]
    nx, ny, nz = g.n
    zc = (e["i"] // ((nx + 2) * (ny + 2))) - 0.5 - 0.5 * nz
    e["uy"] += (drift / np.cosh(zc / width) ** 2).astype(np.float32)     # the electrons carry the sheet current
    return [{"p": e, "q_m": -1.0}, {"p": i, "q_m": 1.0}]


@pytest.mark.parametrize("nx,nz,clean", [(32, 16, 0), (24, 20, 5)])
def test_trecon_geometry_history(vpb, orc, nx, nz, clean):
    g = trecon_grid(nx, nz)
    ppc = 16
    f0 = sheet_fields(g)
    h_cpu = cpu_history(oracle_kernels(orc), g, sheet_species(g, ppc, 9), STEPS, clean, clean, f_init=f0)
    sim = NativeSimulation(g, L=vpb)
    sim.set_intervals(clean, clean)
    inputs = sheet_species(g, ppc, 9)
    for k, sp in enumerate(inputs):
        s = sim.define_species("s%d" % k, sp["q_m"], len(sp["p"]) + 64, sort_interval=SORT)
        sim.set_particles(s, sp["p"])
    sim.set_fields(f0)
    hist = []
    for _ in range(STEPS):
        sim.advance()
        hist.append(sim.energies())
    h_gpu = np.array(hist)
    scale = np.abs(h_cpu).max(axis=0)
    scale[:6] = h_cpu[:, :6].sum(axis=1).max()
    assert (np.abs(h_gpu - h_cpu) / scale).max() < 1e-4, (np.abs(h_gpu - h_cpu) / scale).max(axis=0)
    tot_c, tot_g = h_cpu.sum(axis=1), h_gpu.sum(axis=1)
    assert np.max(np.abs(tot_g - tot_c) / tot_c) < 2e-5
    assert h_cpu[-1, 3] > 0 and h_cpu[-1, 4] > 0          # Bx and By energy of the sheet
    # every particle is still there (reflection, no absorption), inside its cell, inside the box
    for sp, inp in zip(sim.species, inputs):
        out = sim.get_particles(sp)
        assert np.array_equal(np.sort(out["tag"]), np.sort(inp["p"]["tag"]))
        iz = out["i"] // ((nx + 2) * 3)
        assert iz.min() >= 1 and iz.max() <= nz
    # wall conditions after 20 steps: tangential E on the z walls (node planes 1 and nz+1) and normal B vanish
    f = sim.get_fields().reshape(nz + 2, 3, nx + 2)
    for plane in (1, nz + 1):
        assert not np.any(f["ex"][plane, 1, 1:nx + 1]) and not np.any(f["ey"][plane, 1, 1:nx + 1])
        assert not np.any(f["cbz"][plane, 1, 1:nx + 1])
    sim.free()


def test_trecon_cells_and_field_strength_history(vpb, orc):
    """The same geometry with the deck's OWN cell shape and field strength (turbulence.cxx:86-160): cells of 0.488 x 1.95 x
    0.488 c/wpe, dt = 0.99 Courant, wce/wpe = 10 (b0 = 10), sheet half-thickness 6, vth = 0.6 c -- anisotropic cells and
    a gyro-phase of 3.4 rad per step, which none of the unit-cube histories exercise."""
    nx, nz, ppc = 32, 24, 16
    cell = (1000.0 / 2048, 500.0 / 256, 500.0 / 1024)
    g = host_grid((nx, 1, nz), "periodic", L=(cell[0] * nx, cell[1], cell[2] * nz))
    for sgn in (-1, 1):
        g.set_fbc(abi.boundary(0, 0, sgn), abi.PEC_FIELDS)
        g.set_pbc(abi.boundary(0, 0, sgn), abi.REFLECT_PARTICLES)
    g.set_units(0.99 / np.sqrt(1.0 / cell[0] ** 2 + 1.0 / cell[2] ** 2), 1.0, 1.0, 0.0)
    f0 = abi.aligned_zeros(g.nv, abi.field_dtype)
    zc = ((np.arange(g.nv) // ((nx + 2) * 3)) - 0.5 - 0.5 * nz) * cell[2]
    f0["cbx"] = (10.0 * np.tanh(zc / 6.0)).astype(np.float32)
    f0["cby"] = (10.0 / np.cosh(zc / 6.0)).astype(np.float32)

    def species():
        rng = np.random.default_rng(19)
        n = nx * nz * ppc
        vol = cell[0] * cell[1] * cell[2]
        e = random_particles(rng, g, n, vth=0.6, sort=True, q=-vol / ppc)
        i = random_particles(rng, g, n, vth=0.6, sort=True, q=+vol / ppc)
        i["dx"], i["dy"], i["dz"], i["i"] = e["dx"], e["dy"], e["dz"], e["i"]
        return [{"p": e, "q_m": -1.0}, {"p": i, "q_m": 1.0}]

    h_cpu = cpu_history(oracle_kernels(orc), g, species(), STEPS, 5, 5, f_init=f0)
    sim = NativeSimulation(g, L=vpb)
    sim.set_intervals(5, 5)
    inputs = species()
    for k, sp in enumerate(inputs):
        s = sim.define_species("s%d" % k, sp["q_m"], len(sp["p"]) + 64, sort_interval=SORT)
        sim.set_particles(s, sp["p"])
    sim.set_fields(f0)
    h_gpu = []
    for _ in range(STEPS):
        sim.advance()
        h_gpu.append(sim.energies())
    h_gpu = np.array(h_gpu)
    scale = np.abs(h_cpu).max(axis=0)
    scale[:6] = h_cpu[:, :6].sum(axis=1).max()
    assert (np.abs(h_gpu - h_cpu) / scale).max() < 1e-4, (np.abs(h_gpu - h_cpu) / scale).max(axis=0)
    for sp, inp in zip(sim.species, inputs):
        assert np.array_equal(np.sort(sim.get_particles(sp)["tag"]), np.sort(inp["p"]["tag"]))
    sim.free()
