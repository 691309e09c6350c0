"""Randomised configurations the hand-written parity cases do not reach: every face with ITS OWN field and particle
boundary condition (pec / pmc / symmetric / absorbing fields, reflecting / absorbing particles, or periodic in pairs),
random small shapes (degenerate axes included), random cell sizes and constants.  Oracle vs the compiled reference,
bit for bit, on the field solve, the divergence cleaning, the particle push with movers, boundary_p and the hydro
synchronisation.  Seeds are fixed: the cases are the same every run."""
import ctypes as C

import numpy as np
import pytest

from helpers import (RefGrid, abi, assert_bits_equal, host_grid, loader, random_fields, random_interpolator, random_particles,
                     vacuum_coefficients)
from old_vpic_b200.abi import ptr
from test_oracle_vs_ref import _accumulators

FBC = [abi.PEC_FIELDS, abi.PMC_FIELDS, abi.SYMMETRIC_FIELDS, abi.ABSORB_FIELDS]
PBC = [abi.REFLECT_PARTICLES, abi.ABSORB_PARTICLES]
AXES = ((1, 0, 0), (0, 1, 0), (0, 0, 1))


def random_grid(L, rng):
    n = tuple(int(v) for v in rng.choice([1, 2, 3, 4, 5, 7], size=3))
    if n == (1, 1, 1):
        n = (3, 1, 2)
    cell = rng.uniform(0.4, 1.6, 3)
    g = RefGrid(L, n, "periodic", Lbox=tuple(float(c * m) for c, m in zip(cell, n)))
    s = g.struct
    s.cvac, s.eps0, s.damp = float(rng.uniform(0.7, 1.2)), float(rng.uniform(0.6, 1.9)), float(rng.choice([0.0, 0.03]))
    dims = [d for d, m in ((s.dx, n[0]), (s.dy, n[1]), (s.dz, n[2])) if m > 1]
    s.dt = float(rng.uniform(0.6, 0.95)) / (s.cvac * np.sqrt(sum(1.0 / d ** 2 for d in dims)))
    desc = []
    for ax, (i, j, k) in enumerate(AXES):
        if n[ax] == 1 or rng.random() < 0.3:
            desc.append("periodic")
            continue                                    # keep this axis periodic (both faces joined to the rank itself)
        for sgn in (-1, 1):
            b = abi.boundary(sgn * i, sgn * j, sgn * k)
            L.set_fbc(g.ref(), b, int(rng.choice(FBC)))
            L.set_pbc(g.ref(), b, int(rng.choice(PBC)))
        desc.append("walls")
    return g, n, desc


@pytest.mark.parametrize("seed", range(40))
def test_random_configuration(orc, ref_scalar, seed):
    L = ref_scalar
    rng = np.random.default_rng(1000 + seed)
    g, n, desc = random_grid(L, rng)
    M = loader.ref_methods(L, 0)
    f = random_fields(rng, g, n_mat=3)
    m = vacuum_coefficients(3, rng)
    f_r, f_o = f.copy(), f.copy()

    def both(name, ref_call, orc_call):
        ref_call(); orc_call()
        assert_bits_equal(f_o, f_r, "%s n=%s %s" % (name, n, desc))

    both("advance_b", lambda: M.advance_b(ptr(f_r), g.ref(), 0.5), lambda: orc.orc_advance_b(ptr(f_o), g.ref(), 0.5, 1))
    both("advance_e", lambda: M.advance_e(ptr(f_r), ptr(m), g.ref()), lambda: orc.orc_advance_e(ptr(f_o), ptr(m), g.ref(), 0))
    both("advance_b", lambda: M.advance_b(ptr(f_r), g.ref(), 0.5), lambda: orc.orc_advance_b(ptr(f_o), g.ref(), 0.5, 1))
    both("synchronize_jf", lambda: M.synchronize_jf(ptr(f_r), g.ref()), lambda: orc.orc_synchronize_jf(ptr(f_o), g.ref()))
    both("synchronize_rho", lambda: M.synchronize_rho(ptr(f_r), g.ref()), lambda: orc.orc_synchronize_rho(ptr(f_o), g.ref()))
    both("synchronize_tang_e_norm_b", lambda: M.synchronize_tang_e_norm_b(ptr(f_r), g.ref()),
         lambda: orc.orc_synchronize_tang_e_norm_b(ptr(f_o), g.ref()))
    both("compute_div_e_err", lambda: M.compute_div_e_err(ptr(f_r), ptr(m), g.ref()), lambda: orc.orc_compute_div_e_err(ptr(f_o), ptr(m), g.ref()))
    both("clean_div_e", lambda: M.clean_div_e(ptr(f_r), ptr(m), g.ref()), lambda: orc.orc_clean_div_e(ptr(f_o), ptr(m), g.ref()))
    both("compute_div_b_err", lambda: M.compute_div_b_err(ptr(f_r), g.ref()), lambda: orc.orc_compute_div_b_err(ptr(f_o), g.ref()))
    both("clean_div_b", lambda: M.clean_div_b(ptr(f_r), g.ref()), lambda: orc.orc_clean_div_b(ptr(f_o), g.ref()))
    both("compute_rhob", lambda: M.compute_rhob(ptr(f_r), ptr(m), g.ref()), lambda: orc.orc_compute_rhob(ptr(f_o), ptr(m), g.ref()))
    both("compute_curl_b", lambda: M.compute_curl_b(ptr(f_r), ptr(m), g.ref()), lambda: orc.orc_compute_curl_b(ptr(f_o), ptr(m), g.ref()))

    # particles: push with movers, then boundary_p (absorbing faces remove, reflecting faces were handled by move_p)
    np_ = 16 * int(rng.integers(5, 60))
    p = random_particles(rng, g, np_, vth=0.7, sort=True, edge_frac=0.03)
    p["q"] = rng.uniform(0.5, 1.5, np_).astype(np.float32)
    fi = random_interpolator(rng, g, amp=0.3)
    p_r, p_o = p.copy(), p.copy()
    a_r, _ = _accumulators(L, g)
    a_o = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    pm_r = abi.aligned_zeros(np_, abi.mover_dtype)
    pm_o = pm_r.copy()
    nm_r = L.advance_p(ptr(p_r), np_, -0.8, ptr(pm_r), np_, ptr(a_r), ptr(fi), g.ref())
    L.reduce_accumulators(ptr(a_r), g.ref())
    nm_o = orc.orc_advance_p(ptr(p_o), np_, -0.8, ptr(pm_o), np_, ptr(a_o), ptr(fi), g.ref())
    assert nm_o == nm_r, (n, desc)
    assert_bits_equal(p_o, p_r, "particles n=%s %s" % (n, desc))
    assert_bits_equal(pm_o[:nm_o], pm_r[:nm_r], "movers")
    assert_bits_equal(a_o, a_r[:g.nv], "accumulators")
    sp = abi.SpeciesStruct()
    sp.id, sp.np, sp.max_np, sp.p = 0, np_, np_, p_r.ctypes.data
    sp.nm, sp.max_nm, sp.pm = nm_r, np_, pm_r.ctypes.data
    sp.q_m = -0.8
    L.boundary_p(C.byref(sp), ptr(f_r), ptr(a_r), g.ref(), None)
    out = [abi.aligned_zeros(nm_o + 1, abi.injector_dtype) for _ in range(6)]
    outp = (C.c_void_p * 6)(*[o.ctypes.data for o in out])
    n_out = (C.c_int * 6)()
    np_o = orc.orc_boundary_p_pack(ptr(p_o), np_, ptr(pm_o), nm_o, 0, ptr(f_o), g.ref(), 0, 1, outp, n_out)
    assert sum(n_out) == 0 and np_o == sp.np and sp.nm == 0
    assert_bits_equal(p_r[:sp.np], p_o[:np_o], "survivors")
    assert_bits_equal(f_r, f_o, "rhob after boundary_p")

    h = abi.aligned_zeros(g.nv, abi.hydro_dtype)
    h.view(np.float32)[:] = rng.standard_normal(h.view(np.float32).shape).astype(np.float32)
    h_r, h_o = h.copy(), h.copy()
    L.synchronize_hydro(ptr(h_r), g.ref())
    orc.orc_synchronize_hydro(ptr(h_o), g.ref(), 0, 1)
    assert_bits_equal(h_o, h_r, "synchronize_hydro")


@pytest.mark.parametrize("seed", range(12))
def test_host_grid_mirror_with_mixed_boundaries(ref_scalar, seed):
    """old_vpic_b200.grid (what the GPU tests and bench.py build their grid_t with): set_fbc / set_pbc per face against
    the reference's ops.c on the same random choices -- bc[] and neighbor[] identical."""
    L = ref_scalar
    rng = np.random.default_rng(3000 + seed)
    n = tuple(int(v) for v in rng.choice([1, 2, 3, 5], size=3))
    cell = rng.uniform(0.4, 1.6, 3)
    Lbox = tuple(float(c * m) for c, m in zip(cell, n))
    r = RefGrid(L, n, "periodic", Lbox=Lbox)
    h = host_grid(n, "periodic", L=Lbox, dt=r.struct.dt)
    for ax, (i, j, k) in enumerate(AXES):
        if n[ax] == 1 or rng.random() < 0.3:
            continue
        for sgn in (-1, 1):
            b = abi.boundary(sgn * i, sgn * j, sgn * k)
            fbc, pbc = int(rng.choice(FBC)), int(rng.choice(PBC))
            L.set_fbc(r.ref(), b, fbc); h.set_fbc(b, fbc)
            L.set_pbc(r.ref(), b, pbc); h.set_pbc(b, pbc)
    assert list(r.struct.bc) == list(h.struct.bc)
    assert np.array_equal(r.neighbor, h.neighbor)
    for name in ("dx", "dy", "dz", "rdx", "rdy", "rdz", "x0", "x1", "y1", "z1", "rangel", "rangeh"):
        assert getattr(r.struct, name) == getattr(h.struct, name), name
