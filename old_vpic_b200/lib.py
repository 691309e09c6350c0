"""ctypes binding of libvpic_b200.so -- the host-side mirror of the reference's C API
for the hot path (same names, argument order and meaning; see include/vpic_b200.h).

The library is CUDA-only: load() raises if the shared object is missing, and every
entry point exits loudly if no sm_100 device is usable.  There is no CPU fallback
and nothing here imports the oracle.
"""
import ctypes as C
import os

from . import abi

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libvpic_b200.so")

_vp, _i, _f, _d, _sz, _l = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t, C.c_long

# name -> (restype, argtypes); mirrors include/vpic_b200.h
SIGNATURES = {
    # (A) reference-named entry points
    "advance_p": (_i, [_vp, _i, _f, _vp, _i, _vp, _vp, _vp]),
    "move_p": (_i, [_vp, _vp, _vp, _vp]),
    "sort_p": (None, [_vp, _vp]),
    "center_p": (None, [_vp, _i, _f, _vp, _vp]),
    "uncenter_p": (None, [_vp, _i, _f, _vp, _vp]),
    "energy_p": (_d, [_vp, _i, _f, _vp, _vp]),
    "accumulate_rho_p": (None, [_vp, _vp, _i, _vp]),
    "accumulate_rhob": (None, [_vp, _vp, _vp]),
    "boundary_p": (None, [_vp, _vp, _vp, _vp, _vp]),
    "new_hydro": (_vp, [_vp]),
    "delete_hydro": (None, [_vp]),
    "clear_hydro": (None, [_vp, _vp]),
    "accumulate_hydro_p": (None, [_vp, _vp, _i, _f, _vp, _vp]),
    "synchronize_hydro": (None, [_vp, _vp]),
    "local_adjust_hydro": (None, [_vp, _vp]),
    "new_interpolator": (_vp, [_vp]),
    "delete_interpolator": (None, [_vp]),
    "new_accumulators": (_vp, [_vp]),
    "delete_accumulators": (None, [_vp]),
    "load_interpolator": (None, [_vp, _vp, _vp]),
    "clear_accumulators": (None, [_vp, _vp]),
    "reduce_accumulators": (None, [_vp, _vp]),
    "unload_accumulator": (None, [_vp, _vp, _vp]),
    "new_field_advance": (_vp, [_vp, _vp, _vp]),
    "delete_field_advance": (None, [_vp]),
    "util_malloc_aligned": (None, [C.c_char_p, _vp, _sz, _sz]),
    "util_free_aligned": (None, [_vp]),
    # (B) device-resident layer
    "vpb_init": (_i, [_i]),
    "vpb_shutdown": (None, []),
    "vpb_device_sm_count": (_i, []),
    "vpb_l2_fetch_granularity": (_i, []),
    "vpb_dev_alloc": (_vp, [_sz]),
    "vpb_dev_free": (None, [_vp]),
    "vpb_malloc_managed": (_vp, [_sz]),
    "vpb_host_alloc_pinned": (_vp, [_sz]),
    "vpb_host_free_pinned": (None, [_vp]),
    "vpb_h2d": (None, [_vp, _vp, _sz]),
    "vpb_d2h": (None, [_vp, _vp, _sz]),
    "vpb_d2d": (None, [_vp, _vp, _sz]),
    "vpb_memset": (None, [_vp, _i, _sz]),
    "vpb_sync": (None, []),
    "vpb_stream": (_vp, []),
    "vpb_timer_start": (None, [_i]),
    "vpb_timer_stop": (None, [_i]),
    "vpb_timer_ms": (_f, [_i]),
    "vpb_launch_count": (_l, [_i]),
    "vpb_prof_enable": (None, [_i]),
    "vpb_prof_collect": (None, [_i, _vp, _vp, _i]),
    "vpb_prof_list": (_i, [_i, _vp, _i]),
    "vpb_trace_enable": (None, [_i]),
    "vpb_trace_reset": (None, []),
    "vpb_trace_report": (None, []),
    "vpb_trace_get": (C.c_double, [C.c_char_p, _vp]),
    "vpb_load_thermal": (None, [_vp, _vp, _i, _f, _f, C.c_ulonglong, _l]),
    "vpb_copy_positions": (None, [_vp, _vp, _l]),
    "vpb_mt_create": (_vp, [C.c_uint]),
    "vpb_mt_destroy": (None, [_vp]),
    "vpb_mt_set_state": (None, [_vp, _vp]),
    "vpb_mt_get_state": (None, [_vp, _vp]),
    "vpb_mt_words": (None, [_vp, _vp, _l]),
    "vpb_mt_draw": (None, [_vp, C.c_char_p, _l, _vp]),
    "vpb_mt_ziggurat_table": (None, [_vp, _vp, _vp]),
    "vpb_inject_from_draws": (_i, [_vp, _vp, _i, _i, _vp, _i, _l, _vp, _vp, _vp, _vp, C.c_double, _l, _l]),
    "vpb_load_pairs_mt": (_l, [_vp, _vp, _l, _vp, _vp, C.c_double, C.c_double, C.c_double, C.c_double, _vp, _i, _vp, _i, _vp, _i, _l, _l]),
    "vpb_load_plane_wave": (None, [_vp, _vp, _i, _f]),
    "vpb_domain_set_particle_layout": (None, [_vp, _l]),
    "vpb_domain_particle_layout": (_l, [_vp]),
    "vpb_particle_convert": (None, [_vp, _vp, _vp, _l, _i]),
    "vpb_copy_positions_dom": (None, [_vp, _vp, _vp, _l]),
    "vpb_domain_set_field_layout": (None, [_vp, _i]),
    "vpb_domain_field_layout": (_i, [_vp]),
    "vpb_field_bytes": (C.c_size_t, [_vp]),
    "vpb_domain_set_interpolator_layout": (None, [_vp, _i]),
    "vpb_interpolator_bytes": (C.c_size_t, [_vp]),
    "vpb_field_convert": (None, [_vp, _vp, _vp, _i]),
    "vpb_comm_unique_id": (None, [_vp]),
    "vpb_comm_init": (None, [_i, _i, _vp]),
    "vpb_comm_autoboot": (_i, [_vp]),
    "vpb_boundary_set_grow_hook": (None, [_vp, _vp]),
    "vpb_boundary_set_local_injectors": (None, [_vp, _i]),
    "vpb_sim_set_callbacks": (None, [_vp, _vp]),
    "vpb_comm_finalize": (None, []),
    "vpb_comm_rank": (_i, []),
    "vpb_comm_nproc": (_i, []),
    "vpb_comm_allsum_d": (None, [_vp, _i]),
    "vpb_set_world": (None, [_i]),
    "vpb_register_material_coefficients": (None, [_vp, _i]),
    "vpb_grid_changed": (None, [_vp]),
    "vpb_domain_of_grid": (_vp, [_vp]),
    "vpb_staging_release": (None, []),
    "vpb_staging_bytes": (None, [_vp, _vp]),
    "vpb_field_advance_table": (_vp, [_i]),
    "vpb_set_tuning": (None, [C.c_char_p, _i]),
    "vpb_get_tuning": (_i, [C.c_char_p]),
    "vpb_domain_create": (_vp, [_vp, _i, _i]),
    "vpb_domain_destroy": (None, [_vp]),
    "vpb_domain_nvoxel": (_l, [_vp]),
    "vpb_advance_p": (None, [_vp, _vp, _i, _f, _vp, _i, _vp, _vp, _vp]),
    "vpb_advance_p_ordered": (None, [_vp, _vp, _i, _f, _vp, _i, _vp, _vp, _vp, _vp]),
    "vpb_advance_p_ignored": (_i, []),
    "vpb_center_p": (None, [_vp, _vp, _i, _f, _vp]),
    "vpb_uncenter_p": (None, [_vp, _vp, _i, _f, _vp]),
    "vpb_energy_p": (None, [_vp, _vp, _i, _f, _vp, _vp]),
    "vpb_accumulate_rho_p": (None, [_vp, _vp, _vp, _i]),
    "vpb_move_p_one": (None, [_vp, _vp, _vp, _vp, _vp]),
    "vpb_accumulate_rhob_one": (None, [_vp, _vp, _vp]),
    "vpb_boundary_p": (None, [_vp, _vp, _i, _vp, _vp]),
    "vpb_boundary_p_round": (None, [_vp, _vp, _i, _vp, _vp, _i]),
    "vpb_sort_p": (None, [_vp, _vp, _vp, _i, _vp]),
    "vpb_sort_p_planes": (None, [_vp, _vp, _vp, _i, _vp]),
    "vpb_sort_p_planes_ahead": (None, [_vp, _vp, _vp, _i, _vp, _i]),
    "vpb_energy_spectrum": (None, [_vp, _vp, _i, C.c_double, _i, _vp, C.c_double, C.c_double, _i, _vp]),
    "vpb_tracer_records": (None, [_vp, _vp, _i, _vp, _f, _f, _f, _i, _vp]),
    "vpb_deck_energy_spectrum": (None, [_vp, _i, C.c_double, _i, _vp, C.c_double, C.c_double, _i, _vp, _vp]),
    "vpb_deck_tracer_records": (None, [_vp, _i, _vp, _vp, _i, _vp]),
    "vpb_sort_p_planes_grouped": (None, [_vp, _vp, _vp, _i, _vp, _i]),
    "vpb_sort_group_order": (_l, [_i, _i, _i, _vp, _vp, _vp]),
    "vpb_sort_group_keys": (_l, [_vp]),
    "vpb_clear_hydro": (None, [_vp, _vp]),
    "vpb_accumulate_hydro_p": (None, [_vp, _vp, _vp, _i, _f, _vp]),
    "vpb_local_adjust_hydro": (None, [_vp, _vp]),
    "vpb_synchronize_hydro": (None, [_vp, _vp]),
    "vpb_load_interpolator": (None, [_vp, _vp, _vp]),
    "vpb_clear_accumulators": (None, [_vp, _vp]),
    "vpb_unload_accumulator": (None, [_vp, _vp, _vp]),
    "vpb_advance_b": (None, [_vp, _vp, _f]),
    "vpb_advance_e": (None, [_vp, _vp, _vp, _i, _i]),
    "vpb_clear_jf": (None, [_vp, _vp]),
    "vpb_clear_rhof": (None, [_vp, _vp]),
    "vpb_synchronize_jf": (None, [_vp, _vp]),
    "vpb_synchronize_rho": (None, [_vp, _vp]),
    "vpb_energy_f": (None, [_vp, _vp, _vp, _i, _vp]),
    "vpb_compute_div_e_err": (None, [_vp, _vp, _vp, _i]),
    "vpb_compute_rms_div_e_err": (None, [_vp, _vp, _vp]),
    "vpb_clean_div_e": (None, [_vp, _vp, _vp, _i]),
    "vpb_compute_div_b_err": (None, [_vp, _vp]),
    "vpb_compute_rms_div_b_err": (None, [_vp, _vp, _vp]),
    "vpb_clean_div_b": (None, [_vp, _vp]),
    "vpb_compute_rhob": (None, [_vp, _vp, _vp, _i]),
    "vpb_compute_curl_b": (None, [_vp, _vp, _vp, _i]),
    "vpb_synchronize_tang_e_norm_b": (None, [_vp, _vp, _vp]),
    # (C) time-step driver
    "vpb_sim_create": (_vp, [_vp, _i, _i, _i, _i, _i, _i, _i]),
    "vpb_sim_destroy": (None, [_vp]),
    "vpb_sim_set_materials": (None, [_vp, _vp, _i]),
    "vpb_sim_define_species": (_i, [_vp, C.c_char_p, _f, _l, _l, _i]),
    "vpb_sim_load_thermal": (None, [_vp, _i, _i, _f, _f, C.c_ulonglong, _l]),
    "vpb_sim_initialize": (None, [_vp, _vp]),
    "vpb_sim_load_pairs_mt": (_l, [_vp, _vp, _i, _i, _l, _vp, _vp, C.c_double, C.c_double, C.c_double, C.c_double, _i, _l, _l]),
    "vpb_sim_set_particles": (None, [_vp, _i, _vp, _l]),
    "vpb_sim_get_particles": (_l, [_vp, _i, _vp, _l]),
    "vpb_sim_set_fields": (None, [_vp, _vp]),
    "vpb_sim_get_fields": (None, [_vp, _vp]),
    "vpb_sim_set_intervals": (None, [_vp, _i, _i, _i]),
    "vpb_sim_set_sort_lookahead": (None, [_vp, _i]),
    "vpb_sim_set_sync_shared_interval": (None, [_vp, _i]),
    "vpb_sim_last_errors": (None, [_vp, _vp]),
    "vpb_sim_advance": (None, [_vp, _i]),
    "vpb_sim_energies": (None, [_vp, _vp]),
    "vpb_sim_hydro": (None, [_vp, _i, _vp]),
    "vpb_sim_step": (_l, [_vp]),
    "vpb_sim_graph_replays": (_l, [_vp]),
    "vpb_sim_num_species": (_i, [_vp]),
    "vpb_sim_np": (_l, [_vp, _i]),
    "vpb_sim_domain": (_vp, [_vp]),
    "vpb_sim_field_array": (_vp, [_vp]),
    "vpb_sim_interpolator_array": (_vp, [_vp]),
    "vpb_sim_accumulator_array": (_vp, [_vp]),
    "vpb_sim_particle_array": (_vp, [_vp, _i]),
}

DATA_SYMBOLS = ("_standard_field_advance", "_vacuum_field_advance", "_standard_v4_field_advance",
                "_vacuum_v4_field_advance")

_lib = None


def load():
    """dlopen the in-tree library and attach signatures.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise RuntimeError("%s is missing: run `python -m old_vpic_b200.build` (there is no CPU fallback)" % SO)
    L = C.CDLL(SO)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)     # AttributeError if the library lacks a declared entry point
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def field_methods(L, which=0):
    """The library's field-advance vtable as callables (same layout as the reference's)."""
    tab = abi.FieldAdvanceMethods.from_address(L.vpb_field_advance_table(which))
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

    class Methods:
        pass

    m = Methods()
    for n, (res, args) in sigs.items():
        setattr(m, n, C.CFUNCTYPE(res, *args)(getattr(tab, n)))
    return m
