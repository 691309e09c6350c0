"""Golden fixtures = outputs of the reference itself (tests/golden/make_golden.py, run where the
reference is present).  The CPU oracle must reproduce them bit for bit everywhere; the CUDA path
must too (-m gpu), except float sums whose order differs (accumulators, rho: 2e-5; fp64: 1e-12)."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import GOLDEN, abi, assert_bits_equal, host_grid, max_rel
from old_vpic_b200.abi import ptr

KINDS = ["periodic", "metal", "absorbing"]
SEQ_ORACLE = {
    "synchronize_jf": lambda O, f, m, g: O.orc_synchronize_jf(f, g),
    "advance_b": lambda O, f, m, g: O.orc_advance_b(f, g, 0.5, 1),
    "advance_e": lambda O, f, m, g: O.orc_advance_e(f, m, g, 0),
    "synchronize_rho": lambda O, f, m, g: O.orc_synchronize_rho(f, g),
    "compute_div_e_err": lambda O, f, m, g: O.orc_compute_div_e_err(f, m, g),
    "clean_div_e": lambda O, f, m, g: O.orc_clean_div_e(f, m, g),
    "compute_div_b_err": lambda O, f, m, g: O.orc_compute_div_b_err(f, g),
    "clean_div_b": lambda O, f, m, g: O.orc_clean_div_b(f, g),
    "compute_curl_b": lambda O, f, m, g: O.orc_compute_curl_b(f, m, g),
    "compute_rhob": lambda O, f, m, g: O.orc_compute_rhob(f, m, g),
}


def load(kind):
    z = np.load(os.path.join(GOLDEN, "ref_%s.npz" % kind))
    g = host_grid(tuple(int(v) for v in z["grid_n"]), kind, dt=float(z["grid_dt"]), damp=float(z["grid_damp"]))
    assert list(g.struct.bc) == list(z["grid_bc"])
    assert np.array_equal(g.neighbor, z["grid_neighbor"])     # host grid mirror == reference's grid
    return z, g


def al(a):
    b = abi.aligned_empty(len(a), a.dtype)
    b[:] = a
    return b


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_reproduces_reference(orc, kind):
    z, g = load(kind)
    p, fi = al(z["adv_p_in"]), al(z["adv_fi"])
    np_ = len(p)
    a = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    pm = abi.aligned_zeros(np_, abi.mover_dtype)
    nm = orc.orc_advance_p(ptr(p), np_, -1.0, ptr(pm), np_, ptr(a), ptr(fi), g.ref())
    assert nm == int(z["adv_nm"])
    assert_bits_equal(p, z["adv_p_out"], "advance_p particles")
    assert_bits_equal(pm[:nm], z["adv_pm_out"], "movers")
    assert_bits_equal(a, z["adv_a_out"], "accumulators")
    q = al(z["adv_p_in"])
    orc.orc_center_p(ptr(q), np_, 0.7, ptr(fi), g.ref())
    assert_bits_equal(q, z["center_out"], "center_p")
    orc.orc_uncenter_p(ptr(q), np_, 0.7, ptr(fi), g.ref())
    assert_bits_equal(q, z["uncenter_out"], "uncenter_p")
    assert orc.orc_energy_p(ptr(al(z["adv_p_in"])), np_, -1.0, ptr(fi), g.ref()) == float(z["energy_p"])
    f, m = al(z["f_in"]), al(z["m"])
    fi2 = abi.aligned_zeros(g.nv, abi.interpolator_dtype)
    orc.orc_load_interpolator(ptr(fi2), ptr(f), g.ref())
    assert_bits_equal(fi2, z["load_interp_out"], "load_interpolator")
    orc.orc_unload_accumulator(ptr(f), ptr(a), g.ref())
    assert_bits_equal(f, z["unload_out"], "unload_accumulator")
    orc.orc_accumulate_rho_p(ptr(f), ptr(al(z["adv_p_out"])), np_, g.ref())
    assert_bits_equal(f, z["rho_p_out"], "accumulate_rho_p")
    for name in z["seq_order"]:
        SEQ_ORACLE[str(name)](orc, ptr(f), ptr(m), g.ref())
        assert_bits_equal(f, z["seq_" + str(name)], str(name))
    err = orc.orc_synchronize_tang_e_norm_b(ptr(f), g.ref())
    assert_bits_equal(f, z["seq_synchronize_tang_e_norm_b"], "synchronize_tang_e_norm_b")
    assert err == pytest.approx(float(z["sync_teb_err"]), rel=1e-13, abs=1e-300)
    en = np.zeros(6)
    orc.orc_energy_f(ptr(en), ptr(f), ptr(m), g.ref())
    np.testing.assert_allclose(en, z["energy_f"], rtol=1e-13)
    out = np.zeros(2)
    orc.orc_rms_div_e_err_local(ptr(out), ptr(f), g.ref())
    assert g.struct.eps0 * np.sqrt(out[0] / out[1]) == pytest.approx(float(z["rms_div_e"]), rel=1e-12)
    orc.orc_rms_div_b_err_local(ptr(out), ptr(f), g.ref())
    assert g.struct.eps0 * np.sqrt(out[0] / out[1]) == pytest.approx(float(z["rms_div_b"]), rel=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_cuda_reproduces_reference(vpb, kind):
    from old_vpic_b200 import lib
    z, g = load(kind)
    M = lib.field_methods(vpb, 0)
    p, fi = al(z["adv_p_in"]), al(z["adv_fi"])
    np_ = len(p)
    a = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    pm = abi.aligned_zeros(np_, abi.mover_dtype)
    nm = vpb.advance_p(ptr(p), np_, -1.0, ptr(pm), np_, ptr(a), ptr(fi), g.ref())
    assert nm == int(z["adv_nm"])
    assert_bits_equal(p, z["adv_p_out"], "advance_p particles")
    assert_bits_equal(pm[:nm], z["adv_pm_out"], "movers")
    assert max_rel(a.view(np.float32), z["adv_a_out"].view(np.float32)) < 2e-5
    q = al(z["adv_p_in"])
    vpb.center_p(ptr(q), np_, 0.7, ptr(fi), g.ref())
    assert_bits_equal(q, z["center_out"], "center_p")
    vpb.uncenter_p(ptr(q), np_, 0.7, ptr(fi), g.ref())
    assert_bits_equal(q, z["uncenter_out"], "uncenter_p")
    assert vpb.energy_p(ptr(al(z["adv_p_in"])), np_, -1.0, ptr(fi), g.ref()) == pytest.approx(float(z["energy_p"]), rel=1e-12)
    f, m = al(z["f_in"]), al(z["m"])
    vpb.vpb_register_material_coefficients(ptr(m), len(m))
    fi2 = abi.aligned_zeros(g.nv, abi.interpolator_dtype)
    vpb.load_interpolator(ptr(fi2), ptr(f), g.ref())
    assert_bits_equal(fi2, z["load_interp_out"], "load_interpolator")
    a_ref = al(z["adv_a_out"])
    vpb.unload_accumulator(ptr(f), ptr(a_ref), g.ref())
    assert_bits_equal(f, z["unload_out"], "unload_accumulator")
    vpb.accumulate_rho_p(ptr(f), ptr(al(z["adv_p_out"])), np_, g.ref())
    assert max_rel(f["rhof"], z["rho_p_out"]["rhof"]) < 2e-5
    f = al(z["rho_p_out"])      # continue from the reference's state so the stencils can be compared bit for bit
    calls = {
        "synchronize_jf": lambda: M.synchronize_jf(ptr(f), g.ref()), "advance_b": lambda: M.advance_b(ptr(f), g.ref(), 0.5),
        "advance_e": lambda: M.advance_e(ptr(f), ptr(m), g.ref()), "synchronize_rho": lambda: M.synchronize_rho(ptr(f), g.ref()),
        "compute_div_e_err": lambda: M.compute_div_e_err(ptr(f), ptr(m), g.ref()),
        "clean_div_e": lambda: M.clean_div_e(ptr(f), ptr(m), g.ref()), "compute_div_b_err": lambda: M.compute_div_b_err(ptr(f), g.ref()),
        "clean_div_b": lambda: M.clean_div_b(ptr(f), g.ref()), "compute_curl_b": lambda: M.compute_curl_b(ptr(f), ptr(m), g.ref()),
        "compute_rhob": lambda: M.compute_rhob(ptr(f), ptr(m), g.ref()),
    }
    for name in z["seq_order"]:
        calls[str(name)]()
        assert_bits_equal(f, z["seq_" + str(name)], str(name))
    err = M.synchronize_tang_e_norm_b(ptr(f), g.ref())
    assert_bits_equal(f, z["seq_synchronize_tang_e_norm_b"], "synchronize_tang_e_norm_b")
    assert err == pytest.approx(float(z["sync_teb_err"]), rel=1e-12, abs=1e-300)
    en = np.zeros(6)
    M.energy_f(ptr(en), ptr(f), ptr(m), g.ref())
    np.testing.assert_allclose(en, z["energy_f"], rtol=1e-12)
    assert M.compute_rms_div_e_err(ptr(f), g.ref()) == pytest.approx(float(z["rms_div_e"]), rel=1e-12)
    assert M.compute_rms_div_b_err(ptr(f), g.ref()) == pytest.approx(float(z["rms_div_b"]), rel=1e-12)


# ---------------------------------------------------------------------------------------------------------
# hydro moments (tests/golden/ref_hydro_*.npz, make_golden.py::make_hydro)
# ---------------------------------------------------------------------------------------------------------
HYDRO_QM = (("e", -1.0), ("i", 0.25))


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_reproduces_reference_hydro(orc, kind):
    z, g = load(kind)
    zh = np.load(os.path.join(GOLDEN, "ref_hydro_%s.npz" % kind))
    p, fi = al(z["adv_p_in"]), al(z["adv_fi"])
    for tag, q_m in HYDRO_QM:
        h = abi.aligned_zeros(g.nv, abi.hydro_dtype)
        orc.orc_accumulate_hydro_p(ptr(h), ptr(p), len(p), q_m, ptr(fi), g.ref())
        assert_bits_equal(h, zh["hydro_%s_accumulated" % tag], "accumulate_hydro_p " + tag)
        orc.orc_synchronize_hydro(ptr(h), g.ref(), 0, 1)
        assert_bits_equal(h, zh["hydro_%s_synchronized" % tag], "synchronize_hydro " + tag)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_cuda_reproduces_reference_hydro(vpb, kind):
    z, g = load(kind)
    zh = np.load(os.path.join(GOLDEN, "ref_hydro_%s.npz" % kind))
    p, fi = al(z["adv_p_in"]), al(z["adv_fi"])
    for tag, q_m in HYDRO_QM:
        want = zh["hydro_%s_accumulated" % tag]
        h = abi.aligned_zeros(g.nv, abi.hydro_dtype)
        vpb.accumulate_hydro_p(ptr(h), ptr(p), len(p), q_m, ptr(fi), g.ref())
        for n in want.dtype.names:
            if n != "_pad":
                assert max_rel(h[n], want[n]) < 2e-5, n
        h = al(want)      # continue from the reference's sums: the face operations are bit-exact
        vpb.synchronize_hydro(ptr(h), g.ref())
        assert_bits_equal(h, zh["hydro_%s_synchronized" % tag], "synchronize_hydro " + tag)
