"""Pins the CPU oracle (oracle/*.c) against the reference itself, compiled from
source into oracle/_ref (scalar flavour): identical seeded inputs, bit-identical
outputs.  This is what makes the oracle trustworthy as the checker for the CUDA
path (the reference ships no golden vectors)."""
import ctypes as C

import numpy as np
import pytest

from helpers import (RefGrid, abi, assert_bits_equal, hot, loader, random_fields, random_interpolator,
                     random_particles, vacuum_coefficients)
from old_vpic_b200.abi import ptr

KINDS = ["periodic", "metal", "absorbing"]
SHAPES = [(6, 5, 4), (8, 1, 6), (1, 1, 16)]


def _accumulators(L, g):
    """(1+n_pipeline) replicas as the reference sizes them (sf_interface.c:65-72)."""
    stride = (g.nv + 1) // 2 * 2
    return abi.aligned_zeros((1 + L.refh_n_pipeline()) * stride, abi.accumulator_dtype), stride


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("n", SHAPES)
def test_advance_p(orc, ref_scalar, kind, n):
    L = ref_scalar
    g = RefGrid(L, n, kind)
    rng = np.random.default_rng(11)
    np_ = 16 * 400  # whole bundles of 16: everything goes through pipeline 0 (advance_p.cxx:41)
    p = random_particles(rng, g, np_, vth=0.6, sort=True, edge_frac=0.02)
    fi = random_interpolator(rng, g, amp=0.3)
    q_m, max_nm = -1.0, np_
    # reference
    p_r = p.copy()
    a_r, stride = _accumulators(L, g)
    pm_r = abi.aligned_zeros(max_nm, abi.mover_dtype)
    nm_r = L.advance_p(ptr(p_r), np_, q_m, ptr(pm_r), max_nm, ptr(a_r), ptr(fi), g.ref())
    L.reduce_accumulators(ptr(a_r), g.ref())
    # oracle
    p_o = p.copy()
    a_o = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    pm_o = abi.aligned_zeros(max_nm, abi.mover_dtype)
    nm_o = orc.orc_advance_p(ptr(p_o), np_, q_m, ptr(pm_o), max_nm, ptr(a_o), ptr(fi), g.ref())
    assert nm_o == nm_r
    if kind == "absorbing":
        assert nm_r > 0, "test should exercise unresolved movers"
    assert_bits_equal(p_o, p_r, "particles")
    assert_bits_equal(pm_o[:nm_o], pm_r[:nm_r], "movers")
    assert_bits_equal(a_o, a_r[:g.nv], "accumulators")
    moved = int((p_o["i"] != p["i"]).sum())
    assert moved > 0, "test should exercise cell crossings"


def test_move_p_direct(orc, ref_scalar):
    L = ref_scalar
    g = RefGrid(L, (5, 4, 3), "metal")
    rng = np.random.default_rng(5)
    p = random_particles(rng, g, 64, vth=0.5)
    a_r = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    a_o = a_r.copy()
    p_r, p_o = p.copy(), p.copy()
    for k in range(64):
        m = np.zeros(1, abi.mover_dtype)
        m["dispx"], m["dispy"], m["dispz"] = rng.uniform(-1.5, 1.5, 3)
        m["i"] = k
        m_r, m_o = m.copy(), m.copy()
        r = L.move_p(ptr(p_r), ptr(m_r), ptr(a_r), g.ref())
        o = orc.orc_move_p(ptr(p_o), ptr(m_o), ptr(a_o), g.ref())
        assert r == o
        assert_bits_equal(m_o, m_r, "mover %d" % k)
    assert_bits_equal(p_o, p_r, "particles")
    assert_bits_equal(a_o, a_r, "accumulators")


@pytest.mark.parametrize("which", ["center_p", "uncenter_p"])
def test_center_uncenter(orc, ref_scalar, which):
    L = ref_scalar
    g = RefGrid(L, (6, 5, 4))
    rng = np.random.default_rng(3)
    p = random_particles(rng, g, 16 * 50 + 7, vth=0.4)
    fi = random_interpolator(rng, g, amp=0.3)
    p_r, p_o = p.copy(), p.copy()
    getattr(L, which)(ptr(p_r), len(p), 0.7, ptr(fi), g.ref())
    getattr(orc, "orc_" + which)(ptr(p_o), len(p), 0.7, ptr(fi), g.ref())
    assert_bits_equal(p_o, p_r, which)
    assert not np.array_equal(p_o["ux"], p["ux"])


def test_energy_p(orc, ref_scalar):
    L = ref_scalar
    g = RefGrid(L, (6, 5, 4))
    rng = np.random.default_rng(4)
    p = random_particles(rng, g, 16 * 64, vth=0.4)
    fi = random_interpolator(rng, g, amp=0.3)
    e_r = L.energy_p(ptr(p), len(p), -1.0, ptr(fi), g.ref())
    e_o = orc.orc_energy_p(ptr(p), len(p), -1.0, ptr(fi), g.ref())
    assert e_r == e_o  # one pipeline, same summation order


@pytest.mark.parametrize("kind", KINDS)
def test_rho_p_and_rhob(orc, ref_scalar, kind):
    L = ref_scalar
    g = RefGrid(L, (5, 4, 3), kind)
    rng = np.random.default_rng(6)
    p = random_particles(rng, g, 500, vth=0.1)
    f = random_fields(rng, g)
    f_r, f_o = f.copy(), f.copy()
    L.accumulate_rho_p(ptr(f_r), ptr(p), len(p), g.ref())
    orc.orc_accumulate_rho_p(ptr(f_o), ptr(p), len(p), g.ref())
    for k in range(0, 500, 7):
        one = p[k:k + 1].copy()
        L.accumulate_rhob(ptr(f_r), ptr(one), g.ref())
        orc.orc_accumulate_rhob(ptr(f_o), ptr(one), g.ref())
    assert_bits_equal(f_o, f_r, "rhof/rhob")


def test_sort_p(orc, ref_scalar):
    L = ref_scalar
    g = RefGrid(L, (6, 5, 4))
    rng = np.random.default_rng(8)
    np_ = 3000
    p = random_particles(rng, g, np_, sort=False)
    sp_list = C.c_void_p(None)
    sp = L.new_species(b"e", -1.0, np_ + 10, 100, 20, 1, C.byref(sp_list))
    S = abi.SpeciesStruct.from_address(sp)
    C.memmove(S.p, p.ctypes.data, p.nbytes)
    S.np = np_
    L.sort_p(C.c_void_p(sp), g.ref())
    S = abi.SpeciesStruct.from_address(sp)
    p_r = np.ctypeslib.as_array(C.cast(S.p, C.POINTER(C.c_uint8)), shape=(np_ * 48,)).view(abi.particle_dtype).copy()
    part_r = np.ctypeslib.as_array(C.cast(S.partition, C.POINTER(C.c_int32)), shape=(g.nv + 1,)).copy()
    p_o = abi.aligned_zeros(np_, abi.particle_dtype)
    part_o = np.zeros(g.nv + 1, np.int32)
    orc.orc_sort_p(ptr(p), ptr(p_o), np_, ptr(part_o), g.ref())
    assert_bits_equal(p_o, p_r, "sorted particles")
    assert np.array_equal(part_o, part_r)
    assert np.all(np.diff(p_o["i"]) >= 0)


@pytest.mark.parametrize("n", SHAPES)
def test_load_interpolator_unload_accumulator(orc, ref_scalar, n):
    L = ref_scalar
    g = RefGrid(L, n)
    rng = np.random.default_rng(9)
    f = random_fields(rng, g)
    fi_r = abi.aligned_zeros(g.nv, abi.interpolator_dtype)
    fi_o = fi_r.copy()
    L.load_interpolator(ptr(fi_r), ptr(f), g.ref())
    orc.orc_load_interpolator(ptr(fi_o), ptr(f), g.ref())
    assert_bits_equal(fi_o, fi_r, "interpolator")
    a, stride = _accumulators(L, g)
    a["jx"][:g.nv] = rng.standard_normal((g.nv, 4))
    a["jy"][:g.nv] = rng.standard_normal((g.nv, 4))
    a["jz"][:g.nv] = rng.standard_normal((g.nv, 4))
    f_r, f_o = f.copy(), f.copy()
    L.unload_accumulator(ptr(f_r), ptr(a), g.ref())
    orc.orc_unload_accumulator(ptr(f_o), ptr(a), g.ref())
    assert_bits_equal(f_o, f_r, "jf after unload")


def _set_bcs(L, g, fbc):
    """Give every non-degenerate face the local field bc `fbc`."""
    for ax, (i, j, k) in enumerate(((1, 0, 0), (0, 1, 0), (0, 0, 1))):
        if g.n[ax] > 1:
            for s in (-1, 1):
                L.set_fbc(g.ref(), abi.boundary(s * i, s * j, s * k), fbc)


FBCS = [None, abi.PEC_FIELDS, abi.SYMMETRIC_FIELDS, abi.PMC_FIELDS, abi.ABSORB_FIELDS]


@pytest.mark.parametrize("fbc", FBCS)
@pytest.mark.parametrize("n", SHAPES)
def test_field_advance(orc, ref_scalar, fbc, n):
    """advance_b, advance_e (standard, 3 materials, TCA damping) and the vacuum variant."""
    L = ref_scalar
    g = RefGrid(L, n, "periodic", damp=0.01)
    if fbc is not None:
        _set_bcs(L, g, fbc)
    M = loader.ref_methods(L, 0)
    rng = np.random.default_rng(10)
    f = random_fields(rng, g, n_mat=3)
    m = vacuum_coefficients(3, rng)
    f_r, f_o = f.copy(), f.copy()
    for frac in (0.5, 1.0):
        M.advance_b(ptr(f_r), g.ref(), frac)
        orc.orc_advance_b(ptr(f_o), g.ref(), frac, 1)
        assert_bits_equal(f_o, f_r, "advance_b")
    for it in range(2):
        M.advance_e(ptr(f_r), ptr(m), g.ref())
        orc.orc_advance_e(ptr(f_o), ptr(m), g.ref(), 0)
        assert_bits_equal(f_o, f_r, "advance_e")
    # vacuum field advance needs damp==0 and trivial materials (vfa.c:60-75)
    g.struct.damp = 0.0
    V = loader.ref_methods(L, 1)
    f0 = random_fields(rng, g, n_mat=1)
    f_r, f_o = f0.copy(), f0.copy()
    V.advance_e(ptr(f_r), None, g.ref())
    orc.orc_advance_e(ptr(f_o), None, g.ref(), 1)
    assert_bits_equal(f_o, f_r, "vfa_advance_e")


@pytest.mark.parametrize("fbc", FBCS)
@pytest.mark.parametrize("n", SHAPES)
def test_sync_and_div_clean(orc, ref_scalar, fbc, n):
    L = ref_scalar
    g = RefGrid(L, n, "periodic")
    if fbc is not None:
        _set_bcs(L, g, fbc)
    M = loader.ref_methods(L, 0)
    rng = np.random.default_rng(12)
    f = random_fields(rng, g, n_mat=3)
    m = vacuum_coefficients(3, rng)
    f_r, f_o = f.copy(), f.copy()
    steps = [
        ("clear_jf", lambda: M.clear_jf(ptr(f_r), g.ref()), lambda: orc.orc_clear_jf(ptr(f_o), g.ref())),
        ("clear_rhof", lambda: M.clear_rhof(ptr(f_r), g.ref()), lambda: orc.orc_clear_rhof(ptr(f_o), g.ref())),
    ]
    for name, a, b in steps:
        a(); b()
        assert_bits_equal(f_o, f_r, name)
    # refill what was cleared so the synchronisations have something to do
    for k in ("jfx", "jfy", "jfz", "rhof"):
        v = rng.standard_normal(g.nv).astype(np.float32)
        f_r[k] = v
        f_o[k] = v
    M.synchronize_jf(ptr(f_r), g.ref()); orc.orc_synchronize_jf(ptr(f_o), g.ref())
    assert_bits_equal(f_o, f_r, "synchronize_jf")
    M.synchronize_rho(ptr(f_r), g.ref()); orc.orc_synchronize_rho(ptr(f_o), g.ref())
    assert_bits_equal(f_o, f_r, "synchronize_rho")
    e_r = M.synchronize_tang_e_norm_b(ptr(f_r), g.ref())
    e_o = orc.orc_synchronize_tang_e_norm_b(ptr(f_o), g.ref())
    assert_bits_equal(f_o, f_r, "synchronize_tang_e_norm_b")
    assert e_o == pytest.approx(e_r, rel=1e-13)
    M.compute_div_e_err(ptr(f_r), ptr(m), g.ref()); orc.orc_compute_div_e_err(ptr(f_o), ptr(m), g.ref())
    assert_bits_equal(f_o, f_r, "compute_div_e_err")
    out = np.zeros(2)
    orc.orc_rms_div_e_err_local(ptr(out), ptr(f_o), g.ref())
    assert g.struct.eps0 * np.sqrt(out[0] / out[1]) == pytest.approx(M.compute_rms_div_e_err(ptr(f_r), g.ref()), rel=1e-12)
    M.clean_div_e(ptr(f_r), ptr(m), g.ref()); orc.orc_clean_div_e(ptr(f_o), ptr(m), g.ref())
    assert_bits_equal(f_o, f_r, "clean_div_e")
    M.compute_div_b_err(ptr(f_r), g.ref()); orc.orc_compute_div_b_err(ptr(f_o), g.ref())
    assert_bits_equal(f_o, f_r, "compute_div_b_err")
    orc.orc_rms_div_b_err_local(ptr(out), ptr(f_o), g.ref())
    assert g.struct.eps0 * np.sqrt(out[0] / out[1]) == pytest.approx(M.compute_rms_div_b_err(ptr(f_r), g.ref()), rel=1e-12)
    M.clean_div_b(ptr(f_r), g.ref()); orc.orc_clean_div_b(ptr(f_o), g.ref())
    assert_bits_equal(f_o, f_r, "clean_div_b")
    M.compute_rhob(ptr(f_r), ptr(m), g.ref()); orc.orc_compute_rhob(ptr(f_o), ptr(m), g.ref())
    assert_bits_equal(f_o, f_r, "compute_rhob")
    M.compute_curl_b(ptr(f_r), ptr(m), g.ref()); orc.orc_compute_curl_b(ptr(f_o), ptr(m), g.ref())
    assert_bits_equal(f_o, f_r, "compute_curl_b")
    en_r, en_o = np.zeros(6), np.zeros(6)
    M.energy_f(ptr(en_r), ptr(f_r), ptr(m), g.ref()); orc.orc_energy_f(ptr(en_o), ptr(f_o), ptr(m), g.ref())
    np.testing.assert_allclose(en_o, en_r, rtol=1e-13)


# ---------------------------------------------------------------------------------------------------------
# hydro moments (hydro_p.c, sf_interface/hydro.c)
# ---------------------------------------------------------------------------------------------------------
def _hydro_fields(a):
    return np.ascontiguousarray(a).view(np.float32).reshape(-1, 16)[:, :14]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("n", SHAPES)
def test_accumulate_hydro_p(orc, ref_scalar, kind, n):
    """Serial in the reference (one loop over the particles): the oracle reproduces every bit, including the two
    expressions the reference evaluates in double (hydro_p.c:86,93)."""
    L = ref_scalar
    g = RefGrid(L, n, kind)
    rng = np.random.default_rng(41)
    p = random_particles(rng, g, 5000, vth=0.8, sort=False)
    p["q"] = rng.uniform(0.5, 1.5, len(p)).astype(np.float32)
    fi = random_interpolator(rng, g, amp=0.4)
    for q_m in (-1.0, 0.25):
        h_r = abi.aligned_zeros(g.nv, abi.hydro_dtype)
        h_o = h_r.copy()
        L.accumulate_hydro_p(ptr(h_r), ptr(p), len(p), q_m, ptr(fi), g.ref())
        orc.orc_accumulate_hydro_p(ptr(h_o), ptr(p), len(p), q_m, ptr(fi), g.ref())
        assert np.any(_hydro_fields(h_r) != 0)
        assert_bits_equal(h_o, h_r, "hydro moments q_m=%g" % q_m)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("n", SHAPES)
def test_synchronize_hydro(orc, ref_scalar, kind, n):
    L = ref_scalar
    g = RefGrid(L, n, kind)
    rng = np.random.default_rng(42)
    h = abi.aligned_zeros(g.nv, abi.hydro_dtype)
    h.view(np.float32)[:] = rng.standard_normal(h.view(np.float32).shape).astype(np.float32)
    h_r, h_o = h.copy(), h.copy()
    L.local_adjust_hydro(ptr(h_r), g.ref())
    orc.orc_local_adjust_hydro(ptr(h_o), g.ref(), 1)
    assert_bits_equal(h_o, h_r, "local_adjust_hydro")
    h_r, h_o = h.copy(), h.copy()
    L.synchronize_hydro(ptr(h_r), g.ref())
    orc.orc_synchronize_hydro(ptr(h_o), g.ref(), 0, 1)
    assert_bits_equal(h_o, h_r, "synchronize_hydro")
    if kind == "periodic":
        assert np.any(_hydro_fields(h_o) != _hydro_fields(h))


def test_material_coefficients(orc, ref_scalar):
    """new_material_coefficients (sfa.c:127-168) through the reference's vtable, from a list made by the reference's
    own new_material, against the oracle fed a list built in Python."""
    from helpers import MATERIAL_TABLE, material_list
    L = ref_scalar
    g = RefGrid(L, (4, 3, 2), "periodic")
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
    names = [k for k in abi.material_coefficient_dtype.names if not k.startswith("pad")]
    for k in names:
        assert np.array_equal(out[k].view(np.uint32), ref_m[k].view(np.uint32)), k
    assert out["nonconductive"].tolist() == [1.0, 0.0, 0.0]


@pytest.mark.parametrize("n,np_", [((6, 5, 4), 4000), ((1, 1, 16), 500), ((8, 1, 6), 3001)])
def test_boundary_p_absorbing(orc, ref_scalar, n, np_):
    """boundary_p.c:77-505 on one rank with absorbing walls: every mover advance_p left is removed, its charge goes to
    rhob (accumulate_rhob, :9-71) and the holes are back-filled from the tail.  The oracle's serial loop is the
    reference's, so survivors (ORDER included), counts and rhob are bit-identical."""
    L = ref_scalar
    g = RefGrid(L, n, "absorbing")
    rng = np.random.default_rng(31)
    p = random_particles(rng, g, np_, vth=0.7, edge_frac=0.02)
    fi = random_interpolator(rng, g, amp=0.2)
    f = random_fields(rng, g)
    a = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    pm = abi.aligned_zeros(np_, abi.mover_dtype)
    nm = orc.orc_advance_p(ptr(p), np_, -1.0, ptr(pm), np_, ptr(a), ptr(fi), g.ref())
    assert nm > 0
    # oracle
    p_o, f_o = p.copy(), f.copy()
    out = [abi.aligned_zeros(nm + 1, abi.injector_dtype) for _ in range(6)]
    outp = (C.c_void_p * 6)(*[o.ctypes.data for o in out])
    n_out = (C.c_int * 6)()
    np_o = orc.orc_boundary_p_pack(ptr(p_o), np_, ptr(pm), nm, 0, ptr(f_o), g.ref(), 0, 1, outp, n_out)
    assert sum(n_out) == 0 and np_o == np_ - nm
    # reference, on a species_t of its own layout
    p_r, f_r, pm_r = p.copy(), f.copy(), pm.copy()
    a_r, _ = _accumulators(L, g)
    sp = abi.SpeciesStruct()
    sp.id, sp.np, sp.max_np, sp.p = 0, np_, np_, p_r.ctypes.data
    sp.nm, sp.max_nm, sp.pm = nm, np_, pm_r.ctypes.data
    sp.q_m = -1.0
    L.boundary_p(C.byref(sp), ptr(f_r), ptr(a_r), g.ref(), None)
    assert sp.np == np_o and sp.nm == 0
    assert_bits_equal(p_r[:sp.np], p_o[:np_o], "survivors, in the reference's order")
    assert_bits_equal(f_r, f_o, "fields (rhob)")
