"""GPU parity on grids whose cells are NOT unit cubes and whose constants are not 1 (dx != dy != dz, cvac = 0.9,
eps0 = 1.7, damp = 0.02) -- the CUDA counterpart of tests/test_oracle_vs_ref_aniso.py (which pins the oracle to the
reference on the same grids).  Every other GPU parity test runs with dx = dy = dz = cvac = eps0 = 1, where a swapped
rdx/rdy, a missing eps0 or a c/dt mix-up cannot show.  Same bars as the isotropic tests.

First run on hardware in round 2 (profiles/r2a_gpu_pytest_all.txt)."""
import os

import numpy as np
import pytest

from helpers import (abi, assert_bits_equal, host_grid, max_rel, random_fields, random_interpolator, random_particles,
                     vacuum_coefficients)
from old_vpic_b200 import lib
from old_vpic_b200.abi import ptr

pytestmark = pytest.mark.gpu

SHAPES = [(6, 5, 4), (8, 1, 6), (1, 1, 16), (20, 12, 9)]
CELL = (0.7, 1.3, 0.45)
TOL = 2e-5


def aniso_grid(n, kind, fbc=None, damp=0.02):
    g = host_grid(n, kind, L=tuple(c * m for c, m in zip(CELL, n)))
    s = g.struct
    dims = [d for d, m in ((s.dx, n[0]), (s.dy, n[1]), (s.dz, n[2])) if m > 1]
    g.set_units(0.93 / (0.9 * np.sqrt(sum(1.0 / d ** 2 for d in dims))), 0.9, 1.7, damp)
    if fbc is not None:
        for ax, (i, j, k) in enumerate(((1, 0, 0), (0, 1, 0), (0, 0, 1))):
            if g.n[ax] > 1:
                for sgn in (-1, 1):
                    g.set_fbc(abi.boundary(sgn * i, sgn * j, sgn * k), fbc)
    return g


def acc_floats(a):
    return a.view(np.float32).reshape(-1, 12)


@pytest.mark.parametrize("planes", [0, 1])
@pytest.mark.parametrize("kind", ["periodic", "metal", "absorbing"])
@pytest.mark.parametrize("n", SHAPES)
def test_particles_aniso(vpb, orc, planes, kind, n):
    g = aniso_grid(n, kind)
    rng = np.random.default_rng(71)
    np_ = 16 * 300 + 5
    p = random_particles(rng, g, np_, vth=0.6, sort=True, edge_frac=0.02)
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
        assert nm_g == nm_o
        assert_bits_equal(p_g, p_o, "particles")
        assert_bits_equal(pm_g[:nm_g], pm_o[:nm_o], "movers")
        assert max_rel(acc_floats(a_g), acc_floats(a_o)) < TOL
        for which in ("center_p", "uncenter_p"):
            q_o, q_g = p.copy(), p.copy()
            getattr(orc, "orc_" + which)(ptr(q_o), np_, 0.7, ptr(fi), g.ref())
            getattr(vpb, which)(ptr(q_g), np_, 0.7, ptr(fi), g.ref())
            assert_bits_equal(q_g, q_o, which)
        e_o = orc.orc_energy_p(ptr(p), np_, -0.8, ptr(fi), g.ref())
        assert vpb.energy_p(ptr(p), np_, -0.8, ptr(fi), g.ref()) == pytest.approx(e_o, rel=1e-12)
        f = random_fields(rng, g)
        f_o, f_g = f.copy(), f.copy()
        orc.orc_accumulate_rho_p(ptr(f_o), ptr(p), np_, g.ref())
        vpb.accumulate_rho_p(ptr(f_g), ptr(p), np_, g.ref())
        assert max_rel(f_g["rhof"], f_o["rhof"]) < TOL
        one = p[7:8].copy()
        f_o, f_g = f.copy(), f.copy()
        orc.orc_accumulate_rhob(ptr(f_o), ptr(one), g.ref())
        vpb.accumulate_rhob(ptr(f_g), ptr(one), g.ref())
        assert_bits_equal(f_g, f_o, "accumulate_rhob of one particle")
        h_o = abi.aligned_zeros(g.nv, abi.hydro_dtype)
        h_g = h_o.copy()
        orc.orc_accumulate_hydro_p(ptr(h_o), ptr(p), np_, -0.8, ptr(fi), g.ref())
        vpb.accumulate_hydro_p(ptr(h_g), ptr(p), np_, -0.8, ptr(fi), g.ref())
        for name in ("jx", "jy", "jz", "rho", "px", "py", "pz", "ke", "txx", "tyy", "tzz", "tyz", "tzx", "txy"):
            assert float(np.max(np.abs(h_g[name] - h_o[name]))) <= TOL * float(np.max(np.abs(h_o[name]))), name
        # same input on both sides: the boundary operations are bit-exact
        h_in = abi.aligned_zeros(g.nv, abi.hydro_dtype)
        h_in.view(np.float32)[:] = rng.standard_normal(h_in.view(np.float32).shape).astype(np.float32)
        h_o, h_g = h_in.copy(), h_in.copy()
        orc.orc_synchronize_hydro(ptr(h_o), g.ref(), 0, 1)
        vpb.synchronize_hydro(ptr(h_g), g.ref())
        assert_bits_equal(h_g, h_o, "synchronize_hydro")
    finally:
        vpb.vpb_set_tuning(b"dropin.particle_planes", 0)


@pytest.mark.parametrize("n", SHAPES)
def test_species_field_coupling_aniso(vpb, orc, n):
    g = aniso_grid(n, "periodic")
    rng = np.random.default_rng(73)
    f = random_fields(rng, g)
    fi_o = abi.aligned_zeros(g.nv, abi.interpolator_dtype)
    fi_g = fi_o.copy()
    orc.orc_load_interpolator(ptr(fi_o), ptr(f), g.ref())
    vpb.load_interpolator(ptr(fi_g), ptr(f), g.ref())
    assert_bits_equal(fi_g, fi_o, "interpolator")
    a = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    for k in ("jx", "jy", "jz"):
        a[k] = rng.standard_normal((g.nv, 4))
    f_o, f_g = f.copy(), f.copy()
    orc.orc_unload_accumulator(ptr(f_o), ptr(a), g.ref())
    vpb.unload_accumulator(ptr(f_g), ptr(a), g.ref())
    assert_bits_equal(f_g, f_o, "jf after unload")


@pytest.mark.parametrize("fbc", [None, abi.PEC_FIELDS, abi.SYMMETRIC_FIELDS, abi.PMC_FIELDS, abi.ABSORB_FIELDS])
@pytest.mark.parametrize("n", SHAPES)
def test_fields_aniso(vpb, orc, fbc, n):
    g = aniso_grid(n, "periodic", fbc)
    M = lib.field_methods(vpb, 0)
    rng = np.random.default_rng(72)
    f = random_fields(rng, g, n_mat=3)
    m = vacuum_coefficients(3, rng)
    vpb.vpb_register_material_coefficients(ptr(m), 3)
    f_o, f_g = f.copy(), f.copy()

    def both(name, orc_call, gpu_call):
        orc_call(); gpu_call()
        assert_bits_equal(f_g, f_o, name)

    for frac in (0.5, 1.0):
        both("advance_b", lambda: orc.orc_advance_b(ptr(f_o), g.ref(), frac, 1), lambda: M.advance_b(ptr(f_g), g.ref(), frac))
    for _ in range(2):
        both("advance_e", lambda: orc.orc_advance_e(ptr(f_o), ptr(m), g.ref(), 0), lambda: M.advance_e(ptr(f_g), ptr(m), g.ref()))
    both("synchronize_jf", lambda: orc.orc_synchronize_jf(ptr(f_o), g.ref()), lambda: M.synchronize_jf(ptr(f_g), g.ref()))
    both("synchronize_rho", lambda: orc.orc_synchronize_rho(ptr(f_o), g.ref()), lambda: M.synchronize_rho(ptr(f_g), g.ref()))
    e = [0.0, 0.0]
    both("synchronize_tang_e_norm_b", lambda: e.__setitem__(0, orc.orc_synchronize_tang_e_norm_b(ptr(f_o), g.ref())),
         lambda: e.__setitem__(1, M.synchronize_tang_e_norm_b(ptr(f_g), g.ref())))
    assert e[1] == pytest.approx(e[0], rel=1e-12)
    both("compute_div_e_err", lambda: orc.orc_compute_div_e_err(ptr(f_o), ptr(m), g.ref()), lambda: M.compute_div_e_err(ptr(f_g), ptr(m), g.ref()))
    out = np.zeros(2)
    orc.orc_rms_div_e_err_local(ptr(out), ptr(f_o), g.ref())
    assert M.compute_rms_div_e_err(ptr(f_g), g.ref()) == pytest.approx(g.struct.eps0 * np.sqrt(out[0] / out[1]), rel=1e-12)
    both("clean_div_e", lambda: orc.orc_clean_div_e(ptr(f_o), ptr(m), g.ref()), lambda: M.clean_div_e(ptr(f_g), ptr(m), g.ref()))
    both("compute_div_b_err", lambda: orc.orc_compute_div_b_err(ptr(f_o), g.ref()), lambda: M.compute_div_b_err(ptr(f_g), g.ref()))
    orc.orc_rms_div_b_err_local(ptr(out), ptr(f_o), g.ref())
    assert M.compute_rms_div_b_err(ptr(f_g), g.ref()) == pytest.approx(g.struct.eps0 * np.sqrt(out[0] / out[1]), rel=1e-12)
    both("clean_div_b", lambda: orc.orc_clean_div_b(ptr(f_o), g.ref()), lambda: M.clean_div_b(ptr(f_g), g.ref()))
    both("compute_rhob", lambda: orc.orc_compute_rhob(ptr(f_o), ptr(m), g.ref()), lambda: M.compute_rhob(ptr(f_g), ptr(m), g.ref()))
    both("compute_curl_b", lambda: orc.orc_compute_curl_b(ptr(f_o), ptr(m), g.ref()), lambda: M.compute_curl_b(ptr(f_g), ptr(m), g.ref()))
    en_o, en_g = np.zeros(6), np.zeros(6)
    orc.orc_energy_f(ptr(en_o), ptr(f_o), ptr(m), g.ref()); M.energy_f(ptr(en_g), ptr(f_g), ptr(m), g.ref())
    np.testing.assert_allclose(en_g, en_o, rtol=1e-12)
    g0 = aniso_grid(n, "periodic", fbc, damp=0.0)
    V = lib.field_methods(vpb, 1)
    f0 = random_fields(rng, g0, n_mat=1)
    f_o, f_g = f0.copy(), f0.copy()
    orc.orc_advance_e(ptr(f_o), None, g0.ref(), 1)
    V.advance_e(ptr(f_g), None, g0.ref())
    assert_bits_equal(f_g, f_o, "vfa_advance_e")
