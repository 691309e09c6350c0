"""TEST INFRASTRUCTURE -- ctypes access to the CPU oracle and (when built) the
reference compiled from source.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libvpic_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

_vp, _i, _f, _d = C.c_void_p, C.c_int, C.c_float, C.c_double


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", HERE, "libvpic_oracle.so"])


def _sig(lib, name, res, args):
    fn = getattr(lib, name)
    fn.restype, fn.argtypes = res, args
    return fn


_oracle = None


def oracle():
    """The C restatement (oracle_particles.c, oracle_fields.c)."""
    global _oracle
    if _oracle is not None:
        return _oracle
    if not os.path.exists(ORACLE_SO):
        build_oracle()
    L = C.CDLL(ORACLE_SO)
    _sig(L, "orc_advance_p", _i, [_vp, _i, _f, _vp, _i, _vp, _vp, _vp])
    _sig(L, "orc_move_p", _i, [_vp, _vp, _vp, _vp])
    _sig(L, "orc_center_p", None, [_vp, _i, _f, _vp, _vp])
    _sig(L, "orc_uncenter_p", None, [_vp, _i, _f, _vp, _vp])
    _sig(L, "orc_energy_p", _d, [_vp, _i, _f, _vp, _vp])
    _sig(L, "orc_accumulate_rho_p", None, [_vp, _vp, _i, _vp])
    _sig(L, "orc_accumulate_rhob", None, [_vp, _vp, _vp])
    _sig(L, "orc_clear_hydro", None, [_vp, _vp])
    _sig(L, "orc_accumulate_hydro_p", None, [_vp, _vp, _i, _f, _vp, _vp])
    _sig(L, "orc_local_adjust_hydro", None, [_vp, _vp, _i])
    _sig(L, "orc_synchronize_hydro", None, [_vp, _vp, _i, _i])
    _sig(L, "orc_hydro_face_floats", _i, [_i, _vp])
    _sig(L, "orc_hydro_face_pack", None, [_i, _vp, _vp, _vp])
    _sig(L, "orc_hydro_face_unpack", None, [_i, _vp, _vp, _vp])
    _sig(L, "orc_sort_p", None, [_vp, _vp, _i, _vp, _vp])
    _sig(L, "orc_boundary_p_pack", _i, [_vp, _i, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp])
    _sig(L, "orc_boundary_p_inject", _i, [_vp, _vp, _vp, _i, _vp, _i, _i, _vp, _vp])
    _sig(L, "orc_load_interpolator", None, [_vp, _vp, _vp])
    _sig(L, "orc_clear_accumulators", None, [_vp, _vp])
    _sig(L, "orc_unload_accumulator", None, [_vp, _vp, _vp])
    _sig(L, "orc_advance_b", None, [_vp, _vp, _f, _i])
    _sig(L, "orc_advance_e_update", None, [_vp, _vp, _vp, _i])
    for n in ("tang_b", "norm_e", "div_b"):
        _sig(L, "orc_local_ghost_" + n, None, [_vp, _vp, _i])
    for n in ("tang_e", "norm_b", "div_e", "jf", "rhof", "rhob"):
        _sig(L, "orc_local_adjust_" + n, None, [_vp, _vp, _i])
    _sig(L, "orc_face_message_floats", _i, [_i, _i, _vp])
    _sig(L, "orc_face_pack", _i, [_i, _i, _vp, _vp, _vp])
    _sig(L, "orc_face_unpack", _d, [_i, _i, _vp, _vp, _vp])
    _sig(L, "orc_advance_e", None, [_vp, _vp, _vp, _i])
    _sig(L, "orc_synchronize_jf", None, [_vp, _vp])
    _sig(L, "orc_synchronize_rho", None, [_vp, _vp])
    _sig(L, "orc_synchronize_tang_e_norm_b", _d, [_vp, _vp])
    _sig(L, "orc_compute_div_e_err", None, [_vp, _vp, _vp])
    _sig(L, "orc_clean_div_e", None, [_vp, _vp, _vp])
    _sig(L, "orc_compute_div_b_err", None, [_vp, _vp])
    _sig(L, "orc_clean_div_b", None, [_vp, _vp])
    _sig(L, "orc_compute_rhob", None, [_vp, _vp, _vp])
    _sig(L, "orc_compute_curl_b", None, [_vp, _vp, _vp])
    _sig(L, "orc_div_e_err_update", None, [_vp, _vp, _vp, _i])
    _sig(L, "orc_clean_div_e_update", None, [_vp, _vp, _vp])
    _sig(L, "orc_clean_div_b_update", None, [_vp, _vp])
    _sig(L, "orc_curl_b_update", None, [_vp, _vp, _vp])
    _sig(L, "orc_clear_jf", None, [_vp, _vp])
    _sig(L, "orc_clear_rhof", None, [_vp, _vp])
    _sig(L, "orc_energy_f", None, [_vp, _vp, _vp, _vp])
    _sig(L, "orc_rms_div_e_err_local", None, [_vp, _vp, _vp])
    _sig(L, "orc_rms_div_b_err_local", None, [_vp, _vp, _vp])
    _sig(L, "orc_material_coefficients", None, [_vp, _vp, _vp])
    _oracle = L
    return L


GHOST_TANG_B, GHOST_NORM_E, GHOST_DIV_B, SYNC_JF, SYNC_RHO, SYNC_TEB = range(6)

_refs = {}


def ref_available(flavour="scalar"):
    return os.path.exists(os.path.join(REF_DIR, "libvpic_ref_%s.so" % flavour))


def ref(flavour="scalar", tpp=1):
    """The reference itself (oracle/_ref, built by oracle/build_ref.sh).

    flavour 'scalar': every particle/voxel through the reference's scalar C
    pipelines (bit-exact oracle).  flavour 'sse': as shipped (V4/SSE + pthreads),
    the CPU baseline.  Booted once per process with `tpp` pipelines.
    """
    if flavour in _refs:
        return _refs[flavour]
    path = os.path.join(REF_DIR, "libvpic_ref_%s.so" % flavour)
    if not os.path.exists(path):
        raise FileNotFoundError(path + " (run oracle/build_ref.sh where /root/reference exists)")
    L = C.CDLL(path)
    _sig(L, "refh_boot", _i, [_i])
    _sig(L, "refh_n_pipeline", _i, [])
    _sig(L, "refh_vtable", _vp, [_i])
    _sig(L, "refh_is_v4", _i, [])
    _sig(L, "refh_layout", _i, [_vp, _i])
    L.refh_boot(tpp)
    _sig(L, "new_grid", _vp, [])
    _sig(L, "partition_periodic_box", None, [_vp] + [_d] * 6 + [_i] * 6)
    _sig(L, "partition_absorbing_box", None, [_vp] + [_d] * 6 + [_i] * 7)
    _sig(L, "partition_metal_box", None, [_vp] + [_d] * 6 + [_i] * 6)
    _sig(L, "set_fbc", None, [_vp, _i, _i])
    _sig(L, "set_pbc", None, [_vp, _i, _i])
    _sig(L, "advance_p", _i, [_vp, _i, _f, _vp, _i, _vp, _vp, _vp])
    _sig(L, "move_p", _i, [_vp, _vp, _vp, _vp])
    _sig(L, "center_p", None, [_vp, _i, _f, _vp, _vp])
    _sig(L, "uncenter_p", None, [_vp, _i, _f, _vp, _vp])
    _sig(L, "energy_p", _d, [_vp, _i, _f, _vp, _vp])
    _sig(L, "accumulate_rho_p", None, [_vp, _vp, _i, _vp])
    _sig(L, "accumulate_rhob", None, [_vp, _vp, _vp])
    _sig(L, "accumulate_hydro_p", None, [_vp, _vp, _i, _f, _vp, _vp])
    _sig(L, "synchronize_hydro", None, [_vp, _vp])
    _sig(L, "local_adjust_hydro", None, [_vp, _vp])
    _sig(L, "new_hydro", _vp, [_vp])
    _sig(L, "clear_hydro", None, [_vp, _vp])
    _sig(L, "sort_p", None, [_vp, _vp])
    _sig(L, "boundary_p", None, [_vp, _vp, _vp, _vp, _vp])
    _sig(L, "new_species", _vp, [C.c_char_p, _f, _i, _i, _i, _i, _vp])
    _sig(L, "new_accumulators", _vp, [_vp])
    _sig(L, "load_interpolator", None, [_vp, _vp, _vp])
    _sig(L, "clear_accumulators", None, [_vp, _vp])
    _sig(L, "reduce_accumulators", None, [_vp, _vp])
    _sig(L, "unload_accumulator", None, [_vp, _vp, _vp])
    _sig(L, "new_material", C.c_uint16, [C.c_char_p] + [_f] * 12 + [_vp])
    _refs[flavour] = L
    return L


def ref_methods(L, which=0):
    """The reference's field-advance vtable as callables (field_advance.h:185-302).
    which: 0 standard, 1 vacuum, 2 standard_v4, 3 vacuum_v4."""
    from old_vpic_b200.abi import FieldAdvanceMethods
    tab = FieldAdvanceMethods.from_address(L.refh_vtable(which))
    sigs = {
        "new_field": (_vp, [_vp]), "delete_field": (None, [_vp]),
        "new_material_coefficients": (_vp, [_vp, _vp]), "delete_material_coefficients": (None, [_vp]),
        "advance_b": (None, [_vp, _vp, _f]), "advance_e": (None, [_vp, _vp, _vp]),
        "energy_f": (None, [_vp, _vp, _vp, _vp]), "clear_jf": (None, [_vp, _vp]),
        "synchronize_jf": (None, [_vp, _vp]), "clear_rhof": (None, [_vp, _vp]), "synchronize_rho": (None, [_vp, _vp]),
        "compute_rhob": (None, [_vp, _vp, _vp]), "compute_curl_b": (None, [_vp, _vp, _vp]),
        "synchronize_tang_e_norm_b": (_d, [_vp, _vp]), "compute_div_e_err": (None, [_vp, _vp, _vp]),
        "compute_rms_div_e_err": (_d, [_vp, _vp]), "clean_div_e": (None, [_vp, _vp, _vp]),
        "compute_div_b_err": (None, [_vp, _vp]), "compute_rms_div_b_err": (_d, [_vp, _vp]),
        "clean_div_b": (None, [_vp, _vp]),
    }

    class M:
        pass

    m = M()
    for n, (res, args) in sigs.items():
        setattr(m, n, C.CFUNCTYPE(res, *args)(getattr(tab, n)))
    return m
