"""GPU parity: the field-advance vtable of libvpic_b200.so (standard and vacuum)
and the species<->field coupling kernels against the CPU oracle, for every local
field boundary condition and for periodic (self-exchanged) faces.

Bar: stencil kernels are BIT-EXACT (same expression order, no FMA).  fp64
reductions (energies, rms, desync error) differ in summation order: rel 1e-12.
"""
import ctypes as C

import numpy as np
import pytest

from helpers import abi, assert_bits_equal, host_grid, random_fields, vacuum_coefficients
from old_vpic_b200 import lib
from old_vpic_b200.abi import ptr

pytestmark = pytest.mark.gpu

SHAPES = [(6, 5, 4), (8, 1, 6), (1, 1, 16), (33, 17, 9)]
FBCS = [None, abi.PEC_FIELDS, abi.SYMMETRIC_FIELDS, abi.PMC_FIELDS, abi.ABSORB_FIELDS]


def _grid(n, fbc, damp=0.0):
    g = host_grid(n, "periodic", damp=damp)
    if fbc is not None:
        for ax, (i, j, k) in enumerate(((1, 0, 0), (0, 1, 0), (0, 0, 1))):
            if g.n[ax] > 1:
                for s in (-1, 1):
                    g.set_fbc(abi.boundary(s * i, s * j, s * k), fbc)
    return g


@pytest.mark.parametrize("n", SHAPES)
def test_load_interpolator_unload_accumulator(vpb, orc, n):
    g = host_grid(n)
    rng = np.random.default_rng(9)
    f = random_fields(rng, g)
    fi_o = abi.aligned_zeros(g.nv, abi.interpolator_dtype)
    fi_o["ex"] = 7.0   # ghosts and _pad must survive
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
    a_g = a.copy()
    vpb.clear_accumulators(ptr(a_g), g.ref())
    assert not np.any(a_g.view(np.uint8))


@pytest.mark.parametrize("march", [1, 0, 2])
@pytest.mark.parametrize("fbc", FBCS)
@pytest.mark.parametrize("n", SHAPES)
def test_field_advance(vpb, orc, fbc, n, march):
    """march = tuning fields.march_z: advance_e with every thread marching up z (1, the default: the material-table
    kernel; 2: also the one-material and vacuum kernels) or one voxel per thread (0) -- the same bits either way"""
    vpb.vpb_set_tuning(b"fields.march_z", march)
    try:
        _field_advance(vpb, orc, fbc, n)
    finally:
        vpb.vpb_set_tuning(b"fields.march_z", 1)


def _field_advance(vpb, orc, fbc, n):
    g = _grid(n, fbc, damp=0.01)
    M = lib.field_methods(vpb, 0)
    rng = np.random.default_rng(10)
    f = random_fields(rng, g, n_mat=3)
    m = vacuum_coefficients(3, rng)
    vpb.vpb_register_material_coefficients(ptr(m), 3)
    f_o, f_g = f.copy(), f.copy()
    for frac in (0.5, 1.0):
        orc.orc_advance_b(ptr(f_o), g.ref(), frac, 1)
        M.advance_b(ptr(f_g), g.ref(), frac)
        assert_bits_equal(f_g, f_o, "advance_b")
    for it in range(2):
        orc.orc_advance_e(ptr(f_o), ptr(m), g.ref(), 0)
        M.advance_e(ptr(f_g), ptr(m), g.ref())
        assert_bits_equal(f_g, f_o, "advance_e")
    # single-material fast path
    m1 = vacuum_coefficients(1)
    vpb.vpb_register_material_coefficients(ptr(m1), 1)
    f1 = random_fields(rng, g, n_mat=1)
    f_o, f_g = f1.copy(), f1.copy()
    orc.orc_advance_e(ptr(f_o), ptr(m1), g.ref(), 0)
    M.advance_e(ptr(f_g), ptr(m1), g.ref())
    assert_bits_equal(f_g, f_o, "advance_e (one material)")
    # vacuum vtable
    g0 = _grid(n, fbc, damp=0.0)
    V = lib.field_methods(vpb, 1)
    f_o, f_g = f1.copy(), f1.copy()
    orc.orc_advance_e(ptr(f_o), None, g0.ref(), 1)
    V.advance_e(ptr(f_g), None, g0.ref())
    assert_bits_equal(f_g, f_o, "vfa_advance_e")


@pytest.mark.parametrize("fbc", FBCS)
@pytest.mark.parametrize("n", SHAPES)
def test_sync_and_div_clean(vpb, orc, fbc, n):
    g = _grid(n, fbc)
    M = lib.field_methods(vpb, 0)
    rng = np.random.default_rng(12)
    f = random_fields(rng, g, n_mat=3)
    m = vacuum_coefficients(3, rng)
    vpb.vpb_register_material_coefficients(ptr(m), 3)
    f_o, f_g = f.copy(), f.copy()
    orc.orc_clear_jf(ptr(f_o), g.ref()); M.clear_jf(ptr(f_g), g.ref())
    assert_bits_equal(f_g, f_o, "clear_jf")
    orc.orc_clear_rhof(ptr(f_o), g.ref()); M.clear_rhof(ptr(f_g), g.ref())
    assert_bits_equal(f_g, f_o, "clear_rhof")
    for k in ("jfx", "jfy", "jfz", "rhof"):
        v = rng.standard_normal(g.nv).astype(np.float32)
        f_o[k] = v
        f_g[k] = v
    orc.orc_synchronize_jf(ptr(f_o), g.ref()); M.synchronize_jf(ptr(f_g), g.ref())
    assert_bits_equal(f_g, f_o, "synchronize_jf")
    orc.orc_synchronize_rho(ptr(f_o), g.ref()); M.synchronize_rho(ptr(f_g), g.ref())
    assert_bits_equal(f_g, f_o, "synchronize_rho")
    e_o = orc.orc_synchronize_tang_e_norm_b(ptr(f_o), g.ref())
    e_g = M.synchronize_tang_e_norm_b(ptr(f_g), g.ref())
    assert_bits_equal(f_g, f_o, "synchronize_tang_e_norm_b")
    assert e_g == pytest.approx(e_o, rel=1e-12, abs=1e-300)
    orc.orc_compute_div_e_err(ptr(f_o), ptr(m), g.ref()); M.compute_div_e_err(ptr(f_g), ptr(m), g.ref())
    assert_bits_equal(f_g, f_o, "compute_div_e_err")
    out = np.zeros(2)
    orc.orc_rms_div_e_err_local(ptr(out), ptr(f_o), g.ref())
    assert M.compute_rms_div_e_err(ptr(f_g), g.ref()) == pytest.approx(g.struct.eps0 * np.sqrt(out[0] / out[1]), rel=1e-12)
    orc.orc_clean_div_e(ptr(f_o), ptr(m), g.ref()); M.clean_div_e(ptr(f_g), ptr(m), g.ref())
    assert_bits_equal(f_g, f_o, "clean_div_e")
    orc.orc_compute_div_b_err(ptr(f_o), g.ref()); M.compute_div_b_err(ptr(f_g), g.ref())
    assert_bits_equal(f_g, f_o, "compute_div_b_err")
    orc.orc_rms_div_b_err_local(ptr(out), ptr(f_o), g.ref())
    assert M.compute_rms_div_b_err(ptr(f_g), g.ref()) == pytest.approx(g.struct.eps0 * np.sqrt(out[0] / out[1]), rel=1e-12)
    orc.orc_clean_div_b(ptr(f_o), g.ref()); M.clean_div_b(ptr(f_g), g.ref())
    assert_bits_equal(f_g, f_o, "clean_div_b")
    orc.orc_compute_rhob(ptr(f_o), ptr(m), g.ref()); M.compute_rhob(ptr(f_g), ptr(m), g.ref())
    assert_bits_equal(f_g, f_o, "compute_rhob")
    orc.orc_compute_curl_b(ptr(f_o), ptr(m), g.ref()); M.compute_curl_b(ptr(f_g), ptr(m), g.ref())
    assert_bits_equal(f_g, f_o, "compute_curl_b")
    en_o, en_g = np.zeros(6), np.zeros(6)
    orc.orc_energy_f(ptr(en_o), ptr(f_o), ptr(m), g.ref()); M.energy_f(ptr(en_g), ptr(f_g), ptr(m), g.ref())
    np.testing.assert_allclose(en_g, en_o, rtol=1e-12)


def test_vacuum_plane_wave_properties(vpb):
    """Size-independent checks on a larger periodic box (no oracle): a plane wave advanced with the vacuum
    solver keeps div B at rounding level and conserves field energy to O(dt^2) over a period."""
    n = (64, 8, 8)
    g = host_grid(n, "periodic")
    V = lib.field_methods(vpb, 1)
    f = abi.aligned_zeros(g.nv, abi.field_dtype)
    F = f.reshape(g.shape)
    x = np.arange(g.shape[2], dtype=np.float64)
    k = 2 * np.pi * 2 / n[0]
    F["ey"][:] = np.sin(k * (x - 1))[None, None, :]
    F["cbz"][:] = np.sin(k * (x - 0.5))[None, None, :]
    en0, en = np.zeros(6), np.zeros(6)
    V.energy_f(ptr(en0), ptr(f), None, g.ref())
    for _ in range(40):
        V.advance_b(ptr(f), g.ref(), 0.5)
        V.advance_e(ptr(f), None, g.ref())
        V.advance_b(ptr(f), g.ref(), 0.5)
    V.energy_f(ptr(en), ptr(f), None, g.ref())
    assert abs(en.sum() - en0.sum()) / en0.sum() < 2e-3
    V.compute_div_b_err(ptr(f), g.ref())
    assert V.compute_rms_div_b_err(ptr(f), g.ref()) < 1e-6


def test_field_only_grid_and_plane_wave_loader(vpb, orc):
    """A grid without a neighbor table (field-only, bench.py's configs[1] leg) runs the same kernels: the device
    plane-wave loader reproduces the host formula bit for bit and 10 device steps match the oracle's."""
    from old_vpic_b200 import grid as gridmod
    from old_vpic_b200.sim import NativeSimulation
    n = (32, 6, 5)
    g = gridmod.make_grid(n, "periodic", field_only=True)
    assert g.neighbor is None and not g.struct.neighbor
    sim = NativeSimulation(g, n_mat=1, vacuum=True, L=vpb)
    vpb.vpb_load_plane_wave(sim.dom, sim.field_ptr, 2, 0.5)
    f_g = sim.get_fields()
    x = np.arange(g.shape[2], dtype=np.float64)
    k = 2 * np.pi * 2 / n[0]
    f_o = abi.aligned_zeros(g.nv, abi.field_dtype)
    F = f_o.reshape(g.shape)
    F["ey"][:] = (np.float32(0.5) * np.cos(k * (x - 1)).astype(np.float32))[None, None, :]
    F["cbz"][:] = (np.float32(0.5) * np.cos(k * (x - 0.5)).astype(np.float32))[None, None, :]
    assert np.abs(f_g["ey"] - f_o["ey"]).max() < 1e-6 and np.abs(f_g["cbz"] - f_o["cbz"]).max() < 1e-6
    f_o = f_g.copy()
    go = host_grid(n, "periodic")       # the oracle walks a full grid_t
    en0 = sum(sim.energies()[:6])
    for _ in range(10):
        sim.advance()
        orc.orc_clear_jf(ptr(f_o), go.ref())
        orc.orc_synchronize_jf(ptr(f_o), go.ref())
        orc.orc_advance_b(ptr(f_o), go.ref(), 0.5, 1)
        orc.orc_advance_e(ptr(f_o), None, go.ref(), 1)
        orc.orc_advance_b(ptr(f_o), go.ref(), 0.5, 1)
    assert_bits_equal(sim.get_fields(), f_o, "field-only advance")
    assert abs(sum(sim.energies()[:6]) - en0) / en0 < 2e-3
    sim.free()


def test_new_field_advance(vpb, orc):
    """new_field_advance / delete_field_advance (field_advance.c:3-28): the bound field array is zero-filled and the
    material coefficients are the reference's (sfa.c:127-168), bit for bit."""
    import ctypes as C
    from helpers import MATERIAL_TABLE, material_list
    g = host_grid((5, 4, 3), "periodic")
    head, keep = material_list()
    n = len(MATERIAL_TABLE)
    want = abi.aligned_zeros(n, abi.material_coefficient_dtype)
    orc.orc_material_coefficients(ptr(want), C.byref(head), g.ref())
    for which in (0, 2):
        fa_addr = vpb.new_field_advance(g.ref(), C.byref(head), vpb.vpb_field_advance_table(which))
        fa = (C.c_void_p * 3).from_address(fa_addr)          # f, m, g
        assert fa[2] == C.addressof(g.struct)
        f = np.ctypeslib.as_array(C.cast(fa[0], C.POINTER(C.c_uint8)), shape=(g.nv * 80,))
        assert not np.any(f)
        m = np.ctypeslib.as_array(C.cast(fa[1], C.POINTER(C.c_uint8)), shape=(64 * n,)).view(abi.material_coefficient_dtype)
        for k in abi.material_coefficient_dtype.names:
            if not k.startswith("pad"):
                assert np.array_equal(m[k].view(np.uint32), want[k].view(np.uint32)), k
        # the bound method table is the library's: one field step through it runs
        tab = abi.FieldAdvanceMethods.from_address(fa_addr + 24)
        adv_b = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_float)(tab.advance_b)
        adv_b(fa[0], g.ref(), 0.5)
        vpb.delete_field_advance(fa_addr)
