"""Oracle vs the compiled reference on grids whose cells are NOT unit cubes and whose constants are not 1
(dx != dy != dz, cvac = 0.9, eps0 = 1.7, damp = 0.02): every other parity test uses dx = dy = dz = cvac = eps0 = 1,
where a swapped rdx/rdy, a missing eps0 or a c/dt mix-up is invisible.  Same bar: bit-identical."""
import ctypes as C

import numpy as np
import pytest

from helpers import (RefGrid, abi, assert_bits_equal, loader, random_fields, random_interpolator, random_particles,
                     vacuum_coefficients)
from old_vpic_b200.abi import ptr
from test_oracle_vs_ref import _accumulators, _set_bcs

SHAPES = [(6, 5, 4), (8, 1, 6), (1, 1, 16)]
CELL = (0.7, 1.3, 0.45)


def aniso_grid(L, n, kind, fbc=None):
    g = RefGrid(L, n, kind, Lbox=tuple(c * m for c, m in zip(CELL, n)))
    s = g.struct
    s.cvac, s.eps0, s.damp = 0.9, 1.7, 0.02
    dims = [d for d, m in ((s.dx, n[0]), (s.dy, n[1]), (s.dz, n[2])) if m > 1]
    s.dt = 0.93 / (s.cvac * np.sqrt(sum(1.0 / d ** 2 for d in dims)))
    if fbc is not None:
        _set_bcs(L, g, fbc)
    assert len({round(s.dx, 6), round(s.dy, 6), round(s.dz, 6)}) == 3
    return g


@pytest.mark.parametrize("kind", ["periodic", "metal", "absorbing"])
@pytest.mark.parametrize("n", SHAPES)
def test_particles_aniso(orc, ref_scalar, kind, n):
    L = ref_scalar
    g = aniso_grid(L, n, kind)
    rng = np.random.default_rng(71)
    np_ = 16 * 300
    p = random_particles(rng, g, np_, vth=0.6, sort=True, edge_frac=0.02)
    p["q"] = rng.uniform(0.5, 1.5, np_).astype(np.float32)
    fi = random_interpolator(rng, g, amp=0.3)
    # advance_p + move_p
    p_r, p_o = p.copy(), p.copy()
    a_r, _ = _accumulators(L, g)
    a_o = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    pm_r = abi.aligned_zeros(np_, abi.mover_dtype)
    pm_o = pm_r.copy()
    nm_r = L.advance_p(ptr(p_r), np_, -0.8, ptr(pm_r), np_, ptr(a_r), ptr(fi), g.ref())
    L.reduce_accumulators(ptr(a_r), g.ref())
    nm_o = orc.orc_advance_p(ptr(p_o), np_, -0.8, ptr(pm_o), np_, ptr(a_o), ptr(fi), g.ref())
    assert nm_o == nm_r
    assert_bits_equal(p_o, p_r, "particles")
    assert_bits_equal(pm_o[:nm_o], pm_r[:nm_r], "movers")
    assert_bits_equal(a_o, a_r[:g.nv], "accumulators")
    assert int((p_o["i"] != p["i"]).sum()) > 0
    # center / uncenter / energy
    for which in ("center_p", "uncenter_p"):
        q_r, q_o = p.copy(), p.copy()
        getattr(L, which)(ptr(q_r), np_, 0.7, ptr(fi), g.ref())
        getattr(orc, "orc_" + which)(ptr(q_o), np_, 0.7, ptr(fi), g.ref())
        assert_bits_equal(q_o, q_r, which)
    assert L.energy_p(ptr(p), np_, -0.8, ptr(fi), g.ref()) == orc.orc_energy_p(ptr(p), np_, -0.8, ptr(fi), g.ref())
    # rho_p, rhob, unload_accumulator, load_interpolator
    f = random_fields(rng, g)
    f_r, f_o = f.copy(), f.copy()
    L.accumulate_rho_p(ptr(f_r), ptr(p), np_, g.ref()); orc.orc_accumulate_rho_p(ptr(f_o), ptr(p), np_, g.ref())
    for k in range(0, np_, 97):
        one = p[k:k + 1].copy()
        L.accumulate_rhob(ptr(f_r), ptr(one), g.ref()); orc.orc_accumulate_rhob(ptr(f_o), ptr(one), g.ref())
    L.unload_accumulator(ptr(f_r), ptr(a_r), g.ref()); orc.orc_unload_accumulator(ptr(f_o), ptr(a_o), g.ref())
    assert_bits_equal(f_o, f_r, "rhof, rhob, jf")
    fi_r = abi.aligned_zeros(g.nv, abi.interpolator_dtype)
    fi_o = fi_r.copy()
    L.load_interpolator(ptr(fi_r), ptr(f_r), g.ref()); orc.orc_load_interpolator(ptr(fi_o), ptr(f_o), g.ref())
    assert_bits_equal(fi_o, fi_r, "interpolator")
    # hydro moments
    h_r = abi.aligned_zeros(g.nv, abi.hydro_dtype)
    h_o = h_r.copy()
    L.accumulate_hydro_p(ptr(h_r), ptr(p), np_, -0.8, ptr(fi), g.ref())
    orc.orc_accumulate_hydro_p(ptr(h_o), ptr(p), np_, -0.8, ptr(fi), g.ref())
    L.synchronize_hydro(ptr(h_r), g.ref()); orc.orc_synchronize_hydro(ptr(h_o), g.ref(), 0, 1)
    assert_bits_equal(h_o, h_r, "hydro")


@pytest.mark.parametrize("fbc", [None, abi.PEC_FIELDS, abi.SYMMETRIC_FIELDS, abi.PMC_FIELDS, abi.ABSORB_FIELDS])
@pytest.mark.parametrize("n", SHAPES)
def test_fields_aniso(orc, ref_scalar, fbc, n):
    L = ref_scalar
    g = aniso_grid(L, n, "periodic", fbc)
    M = loader.ref_methods(L, 0)
    rng = np.random.default_rng(72)
    f = random_fields(rng, g, n_mat=3)
    m = vacuum_coefficients(3, rng)
    f_r, f_o = f.copy(), f.copy()

    def both(name, ref_call, orc_call):
        ref_call(); orc_call()
        assert_bits_equal(f_o, f_r, name)

    for frac in (0.5, 1.0):
        both("advance_b", lambda: M.advance_b(ptr(f_r), g.ref(), frac), lambda: orc.orc_advance_b(ptr(f_o), g.ref(), frac, 1))
    for _ in range(2):
        both("advance_e", lambda: M.advance_e(ptr(f_r), ptr(m), g.ref()), lambda: orc.orc_advance_e(ptr(f_o), ptr(m), g.ref(), 0))
    both("synchronize_jf", lambda: M.synchronize_jf(ptr(f_r), g.ref()), lambda: orc.orc_synchronize_jf(ptr(f_o), g.ref()))
    both("synchronize_rho", lambda: M.synchronize_rho(ptr(f_r), g.ref()), lambda: orc.orc_synchronize_rho(ptr(f_o), g.ref()))
    e = [0.0, 0.0]
    both("synchronize_tang_e_norm_b", lambda: e.__setitem__(0, M.synchronize_tang_e_norm_b(ptr(f_r), g.ref())),
         lambda: e.__setitem__(1, orc.orc_synchronize_tang_e_norm_b(ptr(f_o), g.ref())))
    assert e[1] == pytest.approx(e[0], rel=1e-13)
    both("compute_div_e_err", lambda: M.compute_div_e_err(ptr(f_r), ptr(m), g.ref()), lambda: orc.orc_compute_div_e_err(ptr(f_o), ptr(m), g.ref()))
    out = np.zeros(2)
    orc.orc_rms_div_e_err_local(ptr(out), ptr(f_o), g.ref())
    assert g.struct.eps0 * np.sqrt(out[0] / out[1]) == pytest.approx(M.compute_rms_div_e_err(ptr(f_r), g.ref()), rel=1e-12)
    both("clean_div_e", lambda: M.clean_div_e(ptr(f_r), ptr(m), g.ref()), lambda: orc.orc_clean_div_e(ptr(f_o), ptr(m), g.ref()))
    both("compute_div_b_err", lambda: M.compute_div_b_err(ptr(f_r), g.ref()), lambda: orc.orc_compute_div_b_err(ptr(f_o), g.ref()))
    orc.orc_rms_div_b_err_local(ptr(out), ptr(f_o), g.ref())
    assert g.struct.eps0 * np.sqrt(out[0] / out[1]) == pytest.approx(M.compute_rms_div_b_err(ptr(f_r), g.ref()), rel=1e-12)
    both("clean_div_b", lambda: M.clean_div_b(ptr(f_r), g.ref()), lambda: orc.orc_clean_div_b(ptr(f_o), g.ref()))
    both("compute_rhob", lambda: M.compute_rhob(ptr(f_r), ptr(m), g.ref()), lambda: orc.orc_compute_rhob(ptr(f_o), ptr(m), g.ref()))
    both("compute_curl_b", lambda: M.compute_curl_b(ptr(f_r), ptr(m), g.ref()), lambda: orc.orc_compute_curl_b(ptr(f_o), ptr(m), g.ref()))
    en_r, en_o = np.zeros(6), np.zeros(6)
    M.energy_f(ptr(en_r), ptr(f_r), ptr(m), g.ref()); orc.orc_energy_f(ptr(en_o), ptr(f_o), ptr(m), g.ref())
    np.testing.assert_allclose(en_o, en_r, rtol=1e-13)
    # vacuum advance (damp must be 0)
    g.struct.damp = 0.0
    V = loader.ref_methods(L, 1)
    f0 = random_fields(rng, g, n_mat=1)
    f_r, f_o = f0.copy(), f0.copy()
    V.advance_e(ptr(f_r), None, g.ref()); orc.orc_advance_e(ptr(f_o), None, g.ref(), 1)
    assert_bits_equal(f_o, f_r, "vfa_advance_e")


def test_material_coefficients_aniso(orc, ref_scalar):
    from helpers import MATERIAL_TABLE, material_list
    L = ref_scalar
    g = aniso_grid(L, (4, 3, 2), "periodic")
    M = loader.ref_methods(L, 0)
    head = C.c_void_p(None)
    for name, eps, mu, sig, zeta in MATERIAL_TABLE:
        L.new_material(name.encode(), *[float(v) for v in eps + mu + sig + zeta], C.byref(head))
    addr = M.new_material_coefficients(g.ref(), head)
    n = len(MATERIAL_TABLE)
    ref_m = np.ctypeslib.as_array(C.cast(addr, C.POINTER(C.c_uint8)), shape=(64 * n,)).view(abi.material_coefficient_dtype).copy()
    mine, keep = material_list()
    out = abi.aligned_zeros(n, abi.material_coefficient_dtype)
    orc.orc_material_coefficients(ptr(out), C.byref(mine), g.ref())
    for k in [k for k in abi.material_coefficient_dtype.names if not k.startswith("pad")]:
        assert np.array_equal(out[k].view(np.uint32), ref_m[k].view(np.uint32)), k
