"""GPU counterpart of tests/test_oracle_vs_ref_fuzz.py: random shapes, cell sizes and constants, every face with its
own field and particle boundary condition; CUDA through the C ABI against the oracle (which that file pins to the
reference on the same kind of configurations).  Stencils, ghost fills and synchronisations bit-exact; float sums made
with atomics within 2e-5; boundary_p survivors as a set.
First run on hardware in round 2 (profiles/r2a_gpu_pytest_all.txt)."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import (abi, assert_bits_equal, host_grid, max_rel, random_fields, random_interpolator, random_particles,
                     vacuum_coefficients)
from old_vpic_b200 import lib
from old_vpic_b200.abi import ptr

pytestmark = pytest.mark.gpu

FBC = [abi.PEC_FIELDS, abi.PMC_FIELDS, abi.SYMMETRIC_FIELDS, abi.ABSORB_FIELDS]
PBC = [abi.REFLECT_PARTICLES, abi.ABSORB_PARTICLES]
AXES = ((1, 0, 0), (0, 1, 0), (0, 0, 1))
TOL = 2e-5


def random_grid(rng):
    n = tuple(int(v) for v in rng.choice([1, 2, 3, 4, 5, 7, 12], size=3))
    if n == (1, 1, 1):
        n = (3, 1, 2)
    cell = rng.uniform(0.4, 1.6, 3)
    g = host_grid(n, "periodic", L=tuple(float(c * m) for c, m in zip(cell, n)))
    s = g.struct
    cvac, eps0, damp = float(rng.uniform(0.7, 1.2)), float(rng.uniform(0.6, 1.9)), float(rng.choice([0.0, 0.03]))
    dims = [d for d, m in ((s.dx, n[0]), (s.dy, n[1]), (s.dz, n[2])) if m > 1]
    g.set_units(float(rng.uniform(0.6, 0.95)) / (cvac * np.sqrt(sum(1.0 / d ** 2 for d in dims))), cvac, eps0, damp)
    desc = []
    for ax, (i, j, k) in enumerate(AXES):
        if n[ax] == 1 or rng.random() < 0.3:
            desc.append("periodic")
            continue
        for sgn in (-1, 1):
            b = abi.boundary(sgn * i, sgn * j, sgn * k)
            g.set_fbc(b, int(rng.choice(FBC)))
            g.set_pbc(b, int(rng.choice(PBC)))
        desc.append("walls")
    return g, n, desc


@pytest.mark.parametrize("planes", [0, 1])
@pytest.mark.parametrize("seed", range(30))
def test_random_configuration(vpb, orc, seed, planes):
    rng = np.random.default_rng(2000 + seed)
    g, n, desc = random_grid(rng)
    what = "n=%s %s" % (n, desc)
    M = lib.field_methods(vpb, 0)
    f = random_fields(rng, g, n_mat=3)
    m = vacuum_coefficients(3, rng)
    vpb.vpb_register_material_coefficients(ptr(m), 3)
    f_o, f_g = f.copy(), f.copy()

    def both(name, orc_call, gpu_call):
        orc_call(); gpu_call()
        assert_bits_equal(f_g, f_o, "%s %s" % (name, what))

    both("advance_b", lambda: orc.orc_advance_b(ptr(f_o), g.ref(), 0.5, 1), lambda: M.advance_b(ptr(f_g), g.ref(), 0.5))
    both("advance_e", lambda: orc.orc_advance_e(ptr(f_o), ptr(m), g.ref(), 0), lambda: M.advance_e(ptr(f_g), ptr(m), g.ref()))
    both("advance_b", lambda: orc.orc_advance_b(ptr(f_o), g.ref(), 0.5, 1), lambda: M.advance_b(ptr(f_g), g.ref(), 0.5))
    both("synchronize_jf", lambda: orc.orc_synchronize_jf(ptr(f_o), g.ref()), lambda: M.synchronize_jf(ptr(f_g), g.ref()))
    both("synchronize_rho", lambda: orc.orc_synchronize_rho(ptr(f_o), g.ref()), lambda: M.synchronize_rho(ptr(f_g), g.ref()))
    both("synchronize_tang_e_norm_b", lambda: orc.orc_synchronize_tang_e_norm_b(ptr(f_o), g.ref()),
         lambda: M.synchronize_tang_e_norm_b(ptr(f_g), g.ref()))
    both("compute_div_e_err", lambda: orc.orc_compute_div_e_err(ptr(f_o), ptr(m), g.ref()), lambda: M.compute_div_e_err(ptr(f_g), ptr(m), g.ref()))
    both("clean_div_e", lambda: orc.orc_clean_div_e(ptr(f_o), ptr(m), g.ref()), lambda: M.clean_div_e(ptr(f_g), ptr(m), g.ref()))
    both("compute_div_b_err", lambda: orc.orc_compute_div_b_err(ptr(f_o), g.ref()), lambda: M.compute_div_b_err(ptr(f_g), g.ref()))
    both("clean_div_b", lambda: orc.orc_clean_div_b(ptr(f_o), g.ref()), lambda: M.clean_div_b(ptr(f_g), g.ref()))
    both("compute_rhob", lambda: orc.orc_compute_rhob(ptr(f_o), ptr(m), g.ref()), lambda: M.compute_rhob(ptr(f_g), ptr(m), g.ref()))
    both("compute_curl_b", lambda: orc.orc_compute_curl_b(ptr(f_o), ptr(m), g.ref()), lambda: M.compute_curl_b(ptr(f_g), ptr(m), g.ref()))

    np_ = 16 * int(rng.integers(5, 200)) + int(rng.integers(0, 16))
    p = random_particles(rng, g, np_, vth=0.7, sort=True, edge_frac=0.03)
    p["q"] = rng.uniform(0.5, 1.5, np_).astype(np.float32)
    fi = random_interpolator(rng, g, amp=0.3)
    vpb.vpb_set_tuning(b"dropin.particle_planes", planes)
    try:
        p_o, p_g = p.copy(), p.copy()
        a_o = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
        a_g = a_o.copy()
        pm_o = abi.aligned_zeros(np_, abi.mover_dtype)
        pm_g = pm_o.copy()
        nm_o = orc.orc_advance_p(ptr(p_o), np_, -0.8, ptr(pm_o), np_, ptr(a_o), ptr(fi), g.ref())
        nm_g = vpb.advance_p(ptr(p_g), np_, -0.8, ptr(pm_g), np_, ptr(a_g), ptr(fi), g.ref())
        assert nm_g == nm_o, what
        assert_bits_equal(p_g, p_o, "particles " + what)
        assert_bits_equal(pm_g[:nm_g], pm_o[:nm_o], "movers " + what)
        assert max_rel(a_g.view(np.float32).reshape(-1, 12), a_o.view(np.float32).reshape(-1, 12)) < TOL
        # boundary_p: absorbing faces remove the movers advance_p left (reflecting ones were handled by move_p)
        sp = abi.SpeciesStruct()
        sp.id, sp.np, sp.max_np, sp.p = 0, np_, np_, p_g.ctypes.data
        sp.nm, sp.max_nm, sp.pm = nm_g, np_, pm_g.ctypes.data
        sp.q_m = -0.8
        f_o2, f_g2 = f_o.copy(), f_o.copy()
        vpb.boundary_p(C.byref(sp), ptr(f_g2), ptr(a_g), g.ref(), None)
        out = [abi.aligned_zeros(nm_o + 1, abi.injector_dtype) for _ in range(6)]
        outp = (C.c_void_p * 6)(*[o.ctypes.data for o in out])
        n_out = (C.c_int * 6)()
        np_o = orc.orc_boundary_p_pack(ptr(p_o), np_, ptr(pm_o), nm_o, 0, ptr(f_o2), g.ref(), 0, 1, outp, n_out)
        assert sum(n_out) == 0 and sp.np == np_o and sp.nm == 0, what
        og, oo = np.argsort(p_g["tag"][:sp.np]), np.argsort(p_o["tag"][:np_o])
        assert_bits_equal(p_g[:sp.np][og], p_o[:np_o][oo], "survivors (as a set) " + what)
        scale = max(float(np.abs(f_o2["rhob"]).max()), 1e-30)
        assert float(np.abs(f_g2["rhob"] - f_o2["rhob"]).max()) <= TOL * scale, what
    finally:
        vpb.vpb_set_tuning(b"dropin.particle_planes", 0)

    h = abi.aligned_zeros(g.nv, abi.hydro_dtype)
    h.view(np.float32)[:] = rng.standard_normal(h.view(np.float32).shape).astype(np.float32)
    h_o, h_g = h.copy(), h.copy()
    orc.orc_synchronize_hydro(ptr(h_o), g.ref(), 0, 1)
    vpb.synchronize_hydro(ptr(h_g), g.ref())
    assert_bits_equal(h_g, h_o, "synchronize_hydro " + what)
