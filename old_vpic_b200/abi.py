"""numpy / ctypes mirrors of include/vpic_b200_abi.h (the reference's struct layouts).

Every dtype cites the reference declaration it mirrors; tests/test_abi.py checks
sizes and offsets against the reference compiled from source.
"""
import ctypes as C

import numpy as np

# src/species_advance/species_advance.h:28-34
particle_dtype = np.dtype(
    [("dx", "f4"), ("dy", "f4"), ("dz", "f4"), ("i", "i4"), ("ux", "f4"), ("uy", "f4"), ("uz", "f4"), ("q", "f4"),
     ("tag", "i8"), ("tag2", "i8")], align=True)
# species_advance.h:39-42
mover_dtype = np.dtype([("dispx", "f4"), ("dispy", "f4"), ("dispz", "f4"), ("i", "i4")], align=True)
# species_advance.h:48-55
injector_dtype = np.dtype(
    [("dx", "f4"), ("dy", "f4"), ("dz", "f4"), ("i", "i4"), ("ux", "f4"), ("uy", "f4"), ("uz", "f4"), ("q", "f4"),
     ("dispx", "f4"), ("dispy", "f4"), ("dispz", "f4"), ("sp_id", "i4")], align=True)
# src/sf_interface/sf_interface.h:45-58
interpolator_dtype = np.dtype(
    [(n, "f4") for n in ("ex", "dexdy", "dexdz", "d2exdydz", "ey", "deydz", "deydx", "d2eydzdx", "ez", "dezdx", "dezdy",
                         "d2ezdxdy", "cbx", "dcbxdx", "cby", "dcbydy", "cbz", "dcbzdz")] + [("_pad", "f4", (2,))], align=True)
# sf_interface.h:28-38
hydro_dtype = np.dtype([(n, "f4") for n in ("jx", "jy", "jz", "rho", "px", "py", "pz", "ke", "txx", "tyy", "tzz", "tyz", "tzx", "txy")] +
                       [("_pad", "f4", (2,))], align=True)
assert hydro_dtype.itemsize == 64
# sf_interface.h:68-77
accumulator_dtype = np.dtype([("jx", "f4", (4,)), ("jy", "f4", (4,)), ("jz", "f4", (4,))], align=True)
# src/field_advance/field_advance.h:159-171
field_dtype = np.dtype(
    [(n, "f4") for n in ("ex", "ey", "ez", "div_e_err", "cbx", "cby", "cbz", "div_b_err", "tcax", "tcay", "tcaz", "rhob",
                         "jfx", "jfy", "jfz", "rhof")] +
    [(n, "u2") for n in ("ematx", "ematy", "ematz", "nmat", "fmatx", "fmaty", "fmatz", "cmat")], align=True)
# src/field_advance/standard/sfa_private.h:24-32
material_coefficient_dtype = np.dtype(
    [(n, "f4") for n in ("decayx", "drivex", "decayy", "drivey", "decayz", "drivez", "rmux", "rmuy", "rmuz",
                         "nonconductive", "epsx", "epsy", "epsz")] + [("pad", "f4", (3,))], align=True)

assert particle_dtype.itemsize == 48 and mover_dtype.itemsize == 16 and injector_dtype.itemsize == 48
assert interpolator_dtype.itemsize == 80 and accumulator_dtype.itemsize == 48 and field_dtype.itemsize == 80
assert material_coefficient_dtype.itemsize == 64

FIELD_FLOATS = ("ex", "ey", "ez", "div_e_err", "cbx", "cby", "cbz", "div_b_err", "tcax", "tcay", "tcaz", "rhob",
                "jfx", "jfy", "jfz", "rhof")

# grid.h:57-66
PEC_FIELDS, SYMMETRIC_FIELDS, PMC_FIELDS, ABSORB_FIELDS = -1, -2, -3, -4
REFLECT_PARTICLES, ABSORB_PARTICLES = -1, -2


def boundary(i, j, k):
    """grid.h:55 BOUNDARY(i,j,k): index into grid_t.bc[27]."""
    return (i + 1) + 3 * ((j + 1) + 3 * (k + 1))


class GridStruct(C.Structure):
    """src/grid/grid.h:112-167 (240 bytes)."""
    _fields_ = [("mp", C.c_void_p), ("dt", C.c_float), ("cvac", C.c_float), ("eps0", C.c_float), ("damp", C.c_float),
                ("x0", C.c_float), ("y0", C.c_float), ("z0", C.c_float), ("x1", C.c_float), ("y1", C.c_float),
                ("z1", C.c_float), ("dx", C.c_float), ("dy", C.c_float), ("dz", C.c_float), ("rdx", C.c_float),
                ("rdy", C.c_float), ("rdz", C.c_float), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("bc", C.c_int * 27), ("range", C.c_void_p), ("neighbor", C.c_void_p), ("rangel", C.c_int64),
                ("rangeh", C.c_int64), ("nb", C.c_int), ("boundary", C.c_void_p)]


class SpeciesStruct(C.Structure):
    """src/species_advance/species_advance.h:61-93 (name[] is resized on allocation; 8 bytes here)."""
    pass


SpeciesStruct._fields_ = [("id", C.c_int32), ("np", C.c_int), ("max_np", C.c_int), ("p", C.c_void_p), ("nm", C.c_int),
                          ("max_nm", C.c_int), ("pm", C.c_void_p), ("q_m", C.c_float), ("sort_interval", C.c_int),
                          ("sort_out_of_place", C.c_int), ("partition", C.c_void_p), ("next", C.POINTER(SpeciesStruct)),
                          ("name", C.c_char * 8)]

assert C.sizeof(GridStruct) == 240 and GridStruct.bc.offset == 84 and GridStruct.neighbor.offset == 200
assert SpeciesStruct.p.offset == 16 and SpeciesStruct.partition.offset == 56 and SpeciesStruct.name.offset == 72


class SpeciesState(C.Structure):
    """include/vpic_b200.h vpb_species_state_t: one species' device arrays for vpb_boundary_p."""
    _fields_ = [("p", C.c_void_p), ("pm", C.c_void_p), ("np", C.c_int), ("max_np", C.c_int), ("nm", C.c_int),
                ("max_nm", C.c_int), ("id", C.c_int), ("_pad", C.c_int)]


class FieldAdvanceMethods(C.Structure):
    """src/field_advance/field_advance.h:185-302: 20 function pointers."""
    NAMES = ("new_field", "delete_field", "new_material_coefficients", "delete_material_coefficients", "advance_b",
             "advance_e", "energy_f", "clear_jf", "synchronize_jf", "clear_rhof", "synchronize_rho", "compute_rhob",
             "compute_curl_b", "synchronize_tang_e_norm_b", "compute_div_e_err", "compute_rms_div_e_err", "clean_div_e",
             "compute_div_b_err", "compute_rms_div_b_err", "clean_div_b")
    _fields_ = [(n, C.c_void_p) for n in NAMES]


assert C.sizeof(FieldAdvanceMethods) == 160


def ptr(a):
    """void* of a numpy array (or None)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def aligned_empty(n, dtype, align=128):
    """Like the reference's MALLOC_ALIGNED(x, n, 128) (util_base.h:237-245)."""
    dtype = np.dtype(dtype)
    raw = np.empty(n * dtype.itemsize + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n * dtype.itemsize].view(dtype)


def aligned_zeros(n, dtype, align=128):
    a = aligned_empty(n, dtype, align)
    a.view(np.uint8)[:] = 0
    return a
