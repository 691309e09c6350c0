"""GPU parity: particle-side kernels through the reference-named C ABI entry points
(libvpic_b200.so) against the CPU oracle on identical seeded inputs.

Bar (north_star): every integer -- voxel indices, mover counts and indices -- and,
because the kernels follow the reference's scalar arithmetic without FMA, every
particle float is BIT-EXACT.  Accumulators are float sums whose ORDER differs
(atomics): tolerance 2e-5 of the largest accumulator entry, the same class of
difference the reference shows between -tpp settings.
"""
import ctypes as C

import os

import numpy as np
import pytest

from helpers import (abi, assert_bits_equal, host_grid, max_rel, random_fields, random_interpolator,
                     random_particles)
from old_vpic_b200.abi import ptr

pytestmark = pytest.mark.gpu

ACC_TOL = 2e-5
# tests of code that has not run on hardware yet


def acc_floats(a):
    return np.ascontiguousarray(a).view(np.float32).reshape(-1, 12)


@pytest.mark.parametrize("deposit,tma,store,cps", [(1, 2, 0, 5), (0, 2, 0, 5), (1, 2, 0, 4), (1, 2, 0, 6), (0, 2, 0, 6), (1, 2, 1, 4), (0, 2, 1, 4),
                                                   (1, 1, 0, 5), (0, 1, 0, 5), (1, 0, 0, 5), (0, 0, 0, 5)])
@pytest.mark.parametrize("kind", ["periodic", "metal", "absorbing"])
@pytest.mark.parametrize("n,np_,sort", [((6, 5, 4), 5000, True), ((8, 1, 6), 7001, False), ((1, 1, 16), 300, True),
                                        ((16, 16, 16), 16 * 16 * 16 * 40, True)])
def test_advance_p(vpb, orc, deposit, tma, store, cps, kind, n, np_, sort):
    g = host_grid(n, kind)
    rng = np.random.default_rng(21)
    p = random_particles(rng, g, np_, vth=0.6, sort=sort, edge_frac=0.02)
    fi = random_interpolator(rng, g, amp=0.3)
    q_m, max_nm = -1.0, np_
    vpb.vpb_set_tuning(b"advance_p.deposit", deposit)
    vpb.vpb_set_tuning(b"advance_p.tma", tma)
    vpb.vpb_set_tuning(b"advance_p.stream_store", store)
    vpb.vpb_set_tuning(b"advance_p.stream_cps", cps)
    p_o, p_g = p.copy(), p.copy()
    a_o = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    a_g = a_o.copy()
    pm_o = abi.aligned_zeros(max_nm, abi.mover_dtype)
    pm_g = pm_o.copy()
    nm_o = orc.orc_advance_p(ptr(p_o), np_, q_m, ptr(pm_o), max_nm, ptr(a_o), ptr(fi), g.ref())
    nm_g = vpb.advance_p(ptr(p_g), np_, q_m, ptr(pm_g), max_nm, ptr(a_g), ptr(fi), g.ref())
    vpb.vpb_set_tuning(b"advance_p.deposit", 1)
    vpb.vpb_set_tuning(b"advance_p.tma", 2)
    vpb.vpb_set_tuning(b"advance_p.stream_store", 0)
    vpb.vpb_set_tuning(b"advance_p.stream_cps", 5)
    assert nm_g == nm_o
    if kind == "absorbing":
        assert nm_o > 0
    assert_bits_equal(p_g, p_o, "particles")
    assert_bits_equal(pm_g[:nm_g], pm_o[:nm_o], "movers (ordered by particle index)")
    assert (p_o["i"] != p["i"]).sum() > 0
    assert max_rel(acc_floats(a_g), acc_floats(a_o)) < ACC_TOL
    # charge conservation property, independent of the oracle: every accumulator entry that the oracle
    # left untouched must be untouched on the GPU too
    untouched = ~np.any(acc_floats(a_o) != 0, axis=1)
    assert not np.any(acc_floats(a_g)[untouched] != 0)


def test_advance_p_empty_and_tail(vpb, orc):
    g = host_grid((4, 4, 4))
    rng = np.random.default_rng(1)
    fi = random_interpolator(rng, g)
    a = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    pm = abi.aligned_zeros(8, abi.mover_dtype)
    p = random_particles(rng, g, 1)
    assert vpb.advance_p(ptr(p), 0, -1.0, ptr(pm), 8, ptr(a), ptr(fi), g.ref()) == 0
    assert not np.any(acc_floats(a))
    for np_ in (1, 31, 255, 256, 257):
        p = random_particles(rng, g, np_, vth=0.3)
        p_o, p_g = p.copy(), p.copy()
        a_o, a_g = a.copy(), a.copy()
        orc.orc_advance_p(ptr(p_o), np_, 1.0, ptr(pm), 8, ptr(a_o), ptr(fi), g.ref())
        vpb.advance_p(ptr(p_g), np_, 1.0, ptr(pm), 8, ptr(a_g), ptr(fi), g.ref())
        assert_bits_equal(p_g, p_o, "np=%d" % np_)
        assert max_rel(acc_floats(a_g), acc_floats(a_o)) < ACC_TOL


def test_advance_p_managed_memory_in_place(vpb, orc):
    """Arrays from vpb_malloc_managed are used in place (the util_malloc_aligned integration mode)."""
    g = host_grid((6, 6, 6))
    rng = np.random.default_rng(2)
    np_ = 4000
    p = random_particles(rng, g, np_, vth=0.5)
    fi = random_interpolator(rng, g, amp=0.2)

    def managed(arr):
        addr = vpb.vpb_malloc_managed(arr.nbytes)
        view = np.ctypeslib.as_array(C.cast(addr, C.POINTER(C.c_uint8)), shape=(arr.nbytes,)).view(arr.dtype)
        view[:] = arr
        return view

    p_m, fi_m = managed(p), managed(fi)
    a_m = managed(abi.aligned_zeros(g.nv, abi.accumulator_dtype))
    pm_m = managed(abi.aligned_zeros(np_, abi.mover_dtype))
    sz = (C.c_size_t * 2)()
    vpb.vpb_staging_bytes(C.byref(sz, 0), C.byref(sz, 8))
    nm = vpb.advance_p(ptr(p_m), np_, -1.0, ptr(pm_m), np_, ptr(a_m), ptr(fi_m), g.ref())
    vpb.vpb_staging_bytes(C.byref(sz, 0), C.byref(sz, 8))
    assert sz[0] == 0 and sz[1] == 0, "managed arrays must not be staged"
    p_o = p.copy()
    a_o = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    pm_o = abi.aligned_zeros(np_, abi.mover_dtype)
    assert nm == orc.orc_advance_p(ptr(p_o), np_, -1.0, ptr(pm_o), np_, ptr(a_o), ptr(fi), g.ref())
    assert_bits_equal(p_m, p_o, "particles (managed)")
    assert max_rel(acc_floats(a_m), acc_floats(a_o)) < ACC_TOL


@pytest.mark.parametrize("hot_only", [0, 1])
@pytest.mark.parametrize("kind", ["periodic", "absorbing"])
def test_advance_p_streamed_in_pieces(vpb, orc, kind, hot_only):
    """Large host arrays go through the device in pieces (H2D / kernel / D2H overlapped); same results,
    movers still ordered by particle index across pieces.  hot_only: 2-D copies of the 32 hot bytes of every record
    (tuning dropin.hot_only); the tags on the host must come through untouched."""
    g = host_grid((10, 9, 8), kind)
    rng = np.random.default_rng(23)
    np_ = 20000 + 17
    p = random_particles(rng, g, np_, vth=0.6, edge_frac=0.02)
    fi = random_interpolator(rng, g, amp=0.3)
    vpb.vpb_set_tuning(b"dropin.piece", 2048)
    vpb.vpb_set_tuning(b"dropin.hot_only", hot_only)
    try:
        p_o, p_g = p.copy(), p.copy()
        a_o = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
        a_g = a_o.copy()
        pm_o = abi.aligned_zeros(np_, abi.mover_dtype)
        pm_g = pm_o.copy()
        nm_o = orc.orc_advance_p(ptr(p_o), np_, -1.0, ptr(pm_o), np_, ptr(a_o), ptr(fi), g.ref())
        nm_g = vpb.advance_p(ptr(p_g), np_, -1.0, ptr(pm_g), np_, ptr(a_g), ptr(fi), g.ref())
    finally:
        vpb.vpb_set_tuning(b"dropin.piece", 4 << 20)
        vpb.vpb_set_tuning(b"dropin.hot_only", 0)
    assert nm_g == nm_o and (kind != "absorbing" or nm_o > 0)
    assert_bits_equal(p_g, p_o, "particles")
    assert_bits_equal(pm_g[:nm_g], pm_o[:nm_o], "movers")
    assert max_rel(acc_floats(a_g), acc_floats(a_o)) < ACC_TOL


@pytest.mark.parametrize("which", ["center_p", "uncenter_p"])
def test_center_uncenter(vpb, orc, which):
    g = host_grid((6, 5, 4))
    rng = np.random.default_rng(3)
    p = random_particles(rng, g, 3333, vth=0.4)
    fi = random_interpolator(rng, g, amp=0.3)
    p_o, p_g = p.copy(), p.copy()
    getattr(orc, "orc_" + which)(ptr(p_o), len(p), 0.7, ptr(fi), g.ref())
    getattr(vpb, which)(ptr(p_g), len(p), 0.7, ptr(fi), g.ref())
    assert_bits_equal(p_g, p_o, which)


def test_center_then_uncenter_round_trip(vpb):
    """Size-independent property: uncenter_p inverts center_p to rounding."""
    g = host_grid((8, 8, 8))
    rng = np.random.default_rng(4)
    p = random_particles(rng, g, 20000, vth=0.4)
    fi = random_interpolator(rng, g, amp=0.2)
    q = p.copy()
    vpb.center_p(ptr(q), len(q), -1.0, ptr(fi), g.ref())
    vpb.uncenter_p(ptr(q), len(q), -1.0, ptr(fi), g.ref())
    for k in ("ux", "uy", "uz"):
        assert np.max(np.abs(q[k] - p[k])) < 5e-6


def test_energy_p(vpb, orc):
    g = host_grid((6, 5, 4))
    rng = np.random.default_rng(5)
    p = random_particles(rng, g, 10001, vth=0.4)
    fi = random_interpolator(rng, g, amp=0.3)
    e_o = orc.orc_energy_p(ptr(p), len(p), -1.0, ptr(fi), g.ref())
    e_g = vpb.energy_p(ptr(p), len(p), -1.0, ptr(fi), g.ref())
    assert e_g == pytest.approx(e_o, rel=1e-12)   # fp64 accumulation, different order


def test_accumulate_rho_p(vpb, orc):
    g = host_grid((6, 5, 4), "metal")
    rng = np.random.default_rng(6)
    p = random_particles(rng, g, 4000)
    f = random_fields(rng, g)
    f_o, f_g = f.copy(), f.copy()
    orc.orc_accumulate_rho_p(ptr(f_o), ptr(p), len(p), g.ref())
    vpb.accumulate_rho_p(ptr(f_g), ptr(p), len(p), g.ref())
    for k in abi.field_dtype.names:
        if k != "rhof":
            assert np.array_equal(np.ascontiguousarray(f_g[k]).view(np.uint8), np.ascontiguousarray(f_o[k]).view(np.uint8)), k
    assert max_rel(f_g["rhof"], f_o["rhof"]) < ACC_TOL


@pytest.mark.parametrize("n,np_", [((6, 5, 4), 3000), ((16, 16, 16), 150000), ((3, 3, 3), 5000), ((4, 4, 4), 0)])
def test_sort_p(vpb, orc, n, np_):
    """Stable counting sort: bit-exact against the reference's out-of-place sort (sort_p.c:61-77)."""
    g = host_grid(n)
    rng = np.random.default_rng(7)
    p = random_particles(rng, g, max(np_, 1), sort=False)
    sp = abi.SpeciesStruct()
    sp.np, sp.max_np, sp.p = np_, max(np_, 1), p.ctypes.data
    sp.sort_interval, sp.sort_out_of_place = 20, 1
    part_g = np.full(g.nv + 1, -1, np.int32)
    sp.partition = part_g.ctypes.data
    p_in = p.copy()
    vpb.sort_p(C.byref(sp), g.ref())
    if np_ == 0:
        return
    p_o = abi.aligned_zeros(np_, abi.particle_dtype)
    part_o = np.zeros(g.nv + 1, np.int32)
    orc.orc_sort_p(ptr(p_in), ptr(p_o), np_, ptr(part_o), g.ref())
    assert np.array_equal(part_g, part_o)
    assert_bits_equal(p[:np_], p_o, "sorted particles (tags included)")
    assert np.all(np.diff(p["i"][:np_]) >= 0)
    assert part_g[-1] == np_


@pytest.mark.parametrize("np_,dense", [(40000, (5000, 1025, 33)), (3000, (3000,)), (70000, (20000, 20000, 1500))])
def test_sort_p_dense_voxels(vpb, orc, np_, dense):
    """Voxels that hold thousands of particles (localised loads, sheets): the stable rank of a segment above 1024 goes
    through the block-wide bitonic path (vpb_particles.cu sort_rank_big_kernel) instead of the quadratic count; still
    the reference's out-of-place order, bit for bit."""
    g = host_grid((6, 5, 4))
    rng = np.random.default_rng(17)
    p = random_particles(rng, g, np_, sort=False)
    vox = rng.choice(np.unique(p["i"]), len(dense), replace=False)
    at = 0
    idx = rng.permutation(np_)
    for v, n in zip(vox, dense):
        p["i"][idx[at:at + n]] = v
        at += n
    p["tag"] = np.arange(np_)
    sp = abi.SpeciesStruct()
    sp.np, sp.max_np, sp.p = np_, np_, p.ctypes.data
    sp.sort_interval, sp.sort_out_of_place = 20, 1
    part_g = np.full(g.nv + 1, -1, np.int32)
    sp.partition = part_g.ctypes.data
    p_in = p.copy()
    vpb.sort_p(C.byref(sp), g.ref())
    p_o = abi.aligned_zeros(np_, abi.particle_dtype)
    part_o = np.zeros(g.nv + 1, np.int32)
    orc.orc_sort_p(ptr(p_in), ptr(p_o), np_, ptr(part_o), g.ref())
    assert np.array_equal(part_g, part_o)
    assert_bits_equal(p[:np_], p_o, "sorted particles (tags included)")


@pytest.mark.parametrize("store", [0, 1])
@pytest.mark.parametrize("kind", ["periodic", "metal"])
def test_layer_b_wide_interpolator(vpb, orc, kind, store):
    """The device-resident layouts (96-byte interpolator records, planar field array) give the same bits as the
    reference layouts: load_interpolator -> advance_p / center_p / energy_p on device arrays (layer B)."""
    from old_vpic_b200.sim import DevArray, FieldArray
    from helpers import random_fields
    n, np_ = (9, 7, 5), 20000
    g = host_grid(n, kind)
    rng = np.random.default_rng(5)
    p = random_particles(rng, g, np_, vth=0.5, sort=True, edge_frac=0.02)
    f = random_fields(rng, g)
    q_m = -1.0
    # oracle on the reference layouts
    fi_o = abi.aligned_zeros(g.nv, abi.interpolator_dtype)
    orc.orc_load_interpolator(ptr(fi_o), ptr(f), g.ref())
    p_o = p.copy()
    a_o = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    pm_o = abi.aligned_zeros(np_, abi.mover_dtype)
    en_o = orc.orc_energy_p(ptr(p_o), np_, q_m, ptr(fi_o), g.ref())
    nm_o = orc.orc_advance_p(ptr(p_o), np_, q_m, ptr(pm_o), np_, ptr(a_o), ptr(fi_o), g.ref())
    # device, layer B
    dom = vpb.vpb_domain_create(g.ref(), 0, 1)
    vpb.vpb_domain_set_field_layout(dom, 1)
    vpb.vpb_domain_set_interpolator_layout(dom, 1)
    assert vpb.vpb_interpolator_bytes(dom) == 96 * g.nv
    d_f = FieldArray(vpb, dom, g.nv)
    d_f.upload(f)
    assert_bits_equal(d_f.download(), f, "planar round trip")
    d_fi = DevArray(vpb, 96 * g.nv, np.uint8)
    vpb.vpb_load_interpolator(dom, d_fi.ptr, d_f.ptr)
    wide = d_fi.download().view(np.float32).reshape(g.nv, 24)
    assert_bits_equal(wide[:, :18].copy(), fi_o.view(np.float32).reshape(g.nv, 20)[:, :18].copy(), "wide interpolator")
    d_p, d_pm, d_a, d_nm = DevArray(vpb, np_, abi.particle_dtype), DevArray(vpb, np_, abi.mover_dtype), \
        DevArray(vpb, g.nv, abi.accumulator_dtype), DevArray(vpb, 4, np.int32)
    d_en = DevArray(vpb, 2, np.float64)
    d_p.upload(p)
    vpb.vpb_energy_p(dom, d_p.ptr, np_, q_m, d_fi.ptr, d_en.ptr)
    en_g = float(d_en.download(1)[0]) * g.struct.cvac ** 2 / q_m
    assert en_g == pytest.approx(en_o, rel=1e-12)
    vpb.vpb_set_tuning(b"advance_p.stream_store", store)
    vpb.vpb_advance_p(dom, d_p.ptr, np_, q_m, d_pm.ptr, np_, d_a.ptr, d_fi.ptr, d_nm.ptr)
    vpb.vpb_set_tuning(b"advance_p.stream_store", 0)
    nm_g = int(d_nm.download(1)[0])
    assert nm_g == nm_o
    assert_bits_equal(d_p.download(), p_o, "particles")
    assert_bits_equal(d_pm.download(nm_g), pm_o[:nm_o], "movers")
    assert max_rel(acc_floats(d_a.download()), acc_floats(a_o)) < ACC_TOL
    # center_p then uncenter_p against the oracle's
    p2_o = p.copy()
    orc.orc_center_p(ptr(p2_o), np_, q_m, ptr(fi_o), g.ref())
    d_p.upload(p)
    vpb.vpb_center_p(dom, d_p.ptr, np_, q_m, d_fi.ptr)
    assert_bits_equal(d_p.download(), p2_o, "center_p")
    orc.orc_uncenter_p(ptr(p2_o), np_, q_m, ptr(fi_o), g.ref())
    vpb.vpb_uncenter_p(dom, d_p.ptr, np_, q_m, d_fi.ptr)
    assert_bits_equal(d_p.download(), p2_o, "uncenter_p")
    for arr in (d_f, d_fi, d_p, d_pm, d_a, d_nm, d_en):
        arr.free()
    vpb.vpb_domain_destroy(dom)


# ---------------------------------------------------------------------------------------------------------
# Component-plane particle layout + two-particles-per-lane advance_p (vpb_advance_p_pair.cu, packed f32x2)
# ---------------------------------------------------------------------------------------------------------
class particle_planes:
    """Layer-A calls inside this context stage their particle array as component planes on the device."""

    def __init__(self, vpb, cps=4, pipe=1, merge=1, variant=-1):
        self.vpb, self.cps, self.pipe, self.merge, self.variant = vpb, cps, pipe, merge, variant

    def __enter__(self):
        self.vpb.vpb_set_tuning(b"dropin.particle_planes", 1)
        self.vpb.vpb_set_tuning(b"advance_p.pair_cps", self.cps)
        self.vpb.vpb_set_tuning(b"advance_p.pair_pipe", self.pipe)
        self.vpb.vpb_set_tuning(b"advance_p.pair_merge", self.merge)
        self.vpb.vpb_set_tuning(b"advance_p.pair_variant", self.variant)

    def __exit__(self, *a):
        self.vpb.vpb_set_tuning(b"advance_p.pair_variant", -1)      # back to the library's default
        self.vpb.vpb_set_tuning(b"dropin.particle_planes", 0)
        self.vpb.vpb_set_tuning(b"advance_p.pair_cps", 4)
        self.vpb.vpb_set_tuning(b"advance_p.pair_pipe", 1)
        self.vpb.vpb_set_tuning(b"advance_p.pair_merge", 1)


@pytest.mark.parametrize("cps,pipe,merge", [(2, 1, 1), (3, 1, 1), (4, 1, 1), (4, 0, 1), (5, 1, 1), (5, 0, 0), (4, 1, 0)])
@pytest.mark.parametrize("kind", ["periodic", "metal", "absorbing"])
@pytest.mark.parametrize("n,np_,sort,vth", [((6, 5, 4), 5000, True, 0.6), ((8, 1, 6), 7001, False, 0.6), ((1, 1, 16), 300, True, 0.6),
                                            ((16, 16, 16), 16 * 16 * 16 * 40, True, 0.6), ((12, 12, 12), 12 * 12 * 12 * 64, True, 0.1),
                                            ((7, 6, 5), 20001, True, 3.0)])
def test_advance_p_pair(vpb, orc, cps, pipe, merge, kind, n, np_, sort, vth):
    """Bit-exact particles and movers from the packed kernel: every multiply/add is the scalar one per half, the
    packed sqrt/div refinements are the scalar operators' own."""
    g = host_grid(n, kind)
    rng = np.random.default_rng(31)
    p = random_particles(rng, g, np_, vth=vth, sort=sort, edge_frac=0.02)
    fi = random_interpolator(rng, g, amp=0.3)
    q_m, max_nm = -1.0, np_
    p_o, p_g = p.copy(), p.copy()
    a_o = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    a_g = a_o.copy()
    pm_o = abi.aligned_zeros(max_nm, abi.mover_dtype)
    pm_g = pm_o.copy()
    nm_o = orc.orc_advance_p(ptr(p_o), np_, q_m, ptr(pm_o), max_nm, ptr(a_o), ptr(fi), g.ref())
    with particle_planes(vpb, cps, pipe, merge):
        nm_g = vpb.advance_p(ptr(p_g), np_, q_m, ptr(pm_g), max_nm, ptr(a_g), ptr(fi), g.ref())
    assert nm_g == nm_o
    if kind == "absorbing":
        assert nm_o > 0
    assert_bits_equal(p_g, p_o, "particles")
    assert_bits_equal(pm_g[:nm_g], pm_o[:nm_o], "movers (ordered by particle index)")
    assert max_rel(acc_floats(a_g), acc_floats(a_o)) < ACC_TOL
    untouched = ~np.any(acc_floats(a_o) != 0, axis=1)
    assert not np.any(acc_floats(a_g)[untouched] != 0)


@pytest.mark.parametrize("q_m", [-1.0, 1.0 / 1836.0, 3e-9, 0.0])
def test_advance_p_pair_extreme_operands(vpb, orc, q_m):
    """Operands outside the packed fast paths (huge, tiny and zero momenta and charge-to-mass ratios) take the scalar
    operators: still bit-exact."""
    g = host_grid((5, 4, 3), "metal")
    rng = np.random.default_rng(32)
    np_ = 4099
    p = random_particles(rng, g, np_, vth=0.5, sort=True)
    scale = 10.0 ** rng.uniform(-30, 17, size=np_)
    pick = rng.random(np_) < 0.3
    for k in ("ux", "uy", "uz"):
        p[k][pick] = (p[k][pick] * scale[pick]).astype(np.float32)
    p["ux"][::97] = 0
    p["uy"][::97] = 0
    p["uz"][::97] = 0
    fi = random_interpolator(rng, g, amp=0.3)
    fi_flat = fi.view(np.float32).reshape(g.nv, 20)
    fi_flat[::7, :] = 0          # field-free voxels: v3 = 0/..., tiny rotations
    p_o, p_g = p.copy(), p.copy()
    a_o = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    a_g = a_o.copy()
    pm_o = abi.aligned_zeros(np_, abi.mover_dtype)
    pm_g = pm_o.copy()
    nm_o = orc.orc_advance_p(ptr(p_o), np_, q_m, ptr(pm_o), np_, ptr(a_o), ptr(fi), g.ref())
    with particle_planes(vpb):
        nm_g = vpb.advance_p(ptr(p_g), np_, q_m, ptr(pm_g), np_, ptr(a_g), ptr(fi), g.ref())
    assert nm_g == nm_o
    assert_bits_equal(p_g, p_o, "particles")
    assert_bits_equal(pm_g[:nm_g], pm_o[:nm_o], "movers")


def test_advance_p_pair_tails(vpb, orc):
    g = host_grid((4, 4, 4))
    rng = np.random.default_rng(33)
    fi = random_interpolator(rng, g)
    a = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    pm = abi.aligned_zeros(8, abi.mover_dtype)
    with particle_planes(vpb):
        for np_ in (1, 2, 31, 63, 64, 65, 127, 255, 256, 257, 1023):
            p = random_particles(rng, g, np_, vth=0.3)
            p_o, p_g = p.copy(), p.copy()
            a_o, a_g = a.copy(), a.copy()
            orc.orc_advance_p(ptr(p_o), np_, 1.0, ptr(pm), 8, ptr(a_o), ptr(fi), g.ref())
            vpb.advance_p(ptr(p_g), np_, 1.0, ptr(pm), 8, ptr(a_g), ptr(fi), g.ref())
            assert_bits_equal(p_g, p_o, "np=%d" % np_)
            assert max_rel(acc_floats(a_g), acc_floats(a_o)) < ACC_TOL


# Kernel variants (vpb_advance_p_pair.cu: FULL fast path for whole chunks -- the default --, LEAN index-only mover ring):
# same bar as the kernel they replace.


@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("kind", ["periodic", "metal", "absorbing"])
@pytest.mark.parametrize("n,np_,sort,vth", [((6, 5, 4), 5000, True, 0.6), ((8, 1, 6), 7001, False, 0.6), ((1, 1, 16), 300, True, 0.6),
                                            ((16, 16, 16), 16 * 16 * 16 * 40, True, 0.6), ((12, 12, 12), 12 * 12 * 12 * 64, True, 0.1),
                                            ((7, 6, 5), 20001, True, 3.0), ((4, 4, 4), 64 * 50, True, 0.6), ((4, 4, 4), 63, True, 0.6)])
def test_advance_p_pair_variants(vpb, orc, variant, kind, n, np_, sort, vth):
    g = host_grid(n, kind)
    rng = np.random.default_rng(31)
    p = random_particles(rng, g, np_, vth=vth, sort=sort, edge_frac=0.02)
    fi = random_interpolator(rng, g, amp=0.3)
    q_m, max_nm = -1.0, np_
    p_o, p_g = p.copy(), p.copy()
    a_o = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    a_g = a_o.copy()
    pm_o = abi.aligned_zeros(max_nm, abi.mover_dtype)
    pm_g = pm_o.copy()
    nm_o = orc.orc_advance_p(ptr(p_o), np_, q_m, ptr(pm_o), max_nm, ptr(a_o), ptr(fi), g.ref())
    with particle_planes(vpb, variant=variant):
        nm_g = vpb.advance_p(ptr(p_g), np_, q_m, ptr(pm_g), max_nm, ptr(a_g), ptr(fi), g.ref())
    assert nm_g == nm_o
    assert_bits_equal(p_g, p_o, "particles")
    assert_bits_equal(pm_g[:nm_g], pm_o[:nm_o], "movers (ordered by particle index)")
    assert max_rel(acc_floats(a_g), acc_floats(a_o)) < ACC_TOL
    untouched = ~np.any(acc_floats(a_o) != 0, axis=1)
    assert not np.any(acc_floats(a_g)[untouched] != 0)


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_advance_p_pair_variants_extreme_and_tails(vpb, orc, variant):
    """Every kernel variant (0: validity tests in every chunk, the default until the end of round 1; 1: FULL, the
    default now; 2: LEAN; 3: both) on extreme operands and every tail length.  Variants 1-3 ran on a B200 at the end of
    round 1 (profiles/r1s_last_calls_pytest.txt); variant 0 is what all earlier runs measured."""
    g = host_grid((5, 4, 3), "metal")
    rng = np.random.default_rng(32)
    fi = random_interpolator(rng, g, amp=0.3)
    for np_ in (1, 2, 63, 64, 65, 127, 128, 129, 4099):
        p = random_particles(rng, g, np_, vth=0.5, sort=True)
        scale = 10.0 ** rng.uniform(-30, 17, size=np_)
        pick = rng.random(np_) < 0.3
        for k in ("ux", "uy", "uz"):
            p[k][pick] = (p[k][pick] * scale[pick]).astype(np.float32)
        p_o, p_g = p.copy(), p.copy()
        a_o = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
        a_g = a_o.copy()
        pm_o = abi.aligned_zeros(np_, abi.mover_dtype)
        pm_g = pm_o.copy()
        nm_o = orc.orc_advance_p(ptr(p_o), np_, -1.0, ptr(pm_o), np_, ptr(a_o), ptr(fi), g.ref())
        with particle_planes(vpb, variant=variant):
            nm_g = vpb.advance_p(ptr(p_g), np_, -1.0, ptr(pm_g), np_, ptr(a_g), ptr(fi), g.ref())
        assert nm_g == nm_o, np_
        assert_bits_equal(p_g, p_o, "particles np=%d" % np_)
        assert_bits_equal(pm_g[:nm_g], pm_o[:nm_o], "movers np=%d" % np_)


def test_particle_planes_other_kernels(vpb, orc):
    """center_p, uncenter_p, energy_p and accumulate_rho_p through the plane accessors."""
    g = host_grid((6, 5, 4), "metal")
    rng = np.random.default_rng(34)
    p = random_particles(rng, g, 3333, vth=0.4)
    fi = random_interpolator(rng, g, amp=0.3)
    f = random_fields(rng, g)
    with particle_planes(vpb):
        for which in ("center_p", "uncenter_p"):
            p_o, p_g = p.copy(), p.copy()
            getattr(orc, "orc_" + which)(ptr(p_o), len(p), 0.7, ptr(fi), g.ref())
            getattr(vpb, which)(ptr(p_g), len(p), 0.7, ptr(fi), g.ref())
            assert_bits_equal(p_g, p_o, which)
        e_o = orc.orc_energy_p(ptr(p), len(p), -1.0, ptr(fi), g.ref())
        assert vpb.energy_p(ptr(p), len(p), -1.0, ptr(fi), g.ref()) == pytest.approx(e_o, rel=1e-12)
        f_o, f_g = f.copy(), f.copy()
        orc.orc_accumulate_rho_p(ptr(f_o), ptr(p), len(p), g.ref())
        vpb.accumulate_rho_p(ptr(f_g), ptr(p), len(p), g.ref())
        assert max_rel(f_g["rhof"], f_o["rhof"]) < ACC_TOL


@pytest.mark.parametrize("n,np_", [((6, 5, 4), 3000), ((16, 16, 16), 150000), ((3, 3, 3), 5001)])
def test_sort_p_planes(vpb, orc, n, np_):
    """Layer B: the stable counting sort on component planes, bit-exact against the reference's out-of-place sort."""
    from old_vpic_b200.sim import DevArray, ParticleArray
    g = host_grid(n)
    rng = np.random.default_rng(35)
    p = random_particles(rng, g, np_, sort=False)
    p_o = abi.aligned_zeros(np_, abi.particle_dtype)
    part_o = np.zeros(g.nv + 1, np.int32)
    orc.orc_sort_p(ptr(p.copy()), ptr(p_o), np_, ptr(part_o), g.ref())
    dom = vpb.vpb_domain_create(g.ref(), 0, 1)
    vpb.vpb_domain_set_particle_layout(dom, (np_ + 63) // 64 * 64)
    d_in, d_out = ParticleArray(vpb, dom, np_), ParticleArray(vpb, dom, np_)
    d_part = DevArray(vpb, g.nv + 1, np.int32)
    d_in.upload(p)
    assert_bits_equal(d_in.download(np_), p, "plane round trip")
    vpb.vpb_sort_p(dom, d_in.ptr, d_out.ptr, np_, d_part.ptr)
    assert np.array_equal(d_part.download(), part_o)
    assert_bits_equal(d_out.download(np_), p_o, "sorted particles (tags included)")
    # the in-place variant the device-resident driver uses (records staged in the scratch array)
    vpb.vpb_memset(d_part.ptr, 0xff, 4 * (g.nv + 1))
    vpb.vpb_sort_p_planes(dom, d_in.ptr, d_out.ptr, np_, d_part.ptr)
    assert np.array_equal(d_part.download(), part_o)
    assert_bits_equal(d_in.download(np_), p_o, "sorted in place")
    for arr in (d_in, d_out, d_part):
        arr.free()
    vpb.vpb_domain_destroy(dom)


def test_sort_p_planes_lookahead(vpb):
    """Look-ahead sort key (vpb_sort_p_planes_ahead): a grouping of the SAME particles by the voxel they reach `L`
    steps ahead at their present velocity; partition[] delimits the groups; L = 0 is the plain (stable) sort."""
    from old_vpic_b200.sim import DevArray, ParticleArray
    n, np_, L = (10, 9, 8), 60001, 7
    g = host_grid(n)
    rng = np.random.default_rng(36)
    p = random_particles(rng, g, np_, vth=0.5, sort=False)
    dom = vpb.vpb_domain_create(g.ref(), 0, 1)
    vpb.vpb_domain_set_particle_layout(dom, (np_ + 63) // 64 * 64)
    d_p, d_tmp = ParticleArray(vpb, dom, np_), ParticleArray(vpb, dom, np_)
    d_part = DevArray(vpb, g.nv + 1, np.int32)
    d_p.upload(p)
    vpb.vpb_sort_p_planes_ahead(dom, d_p.ptr, d_tmp.ptr, np_, d_part.ptr, L)
    out, part = d_p.download(np_), d_part.download()
    # same particles
    order = np.argsort(out["tag"], kind="stable")
    assert_bits_equal(out[order], p[np.argsort(p["tag"], kind="stable")], "multiset")
    # expected keys, computed the same way in float32
    s = g.struct
    sx, sy = n[0] + 2, n[1] + 2

    def keys(q):
        v = q["i"].astype(np.int64)
        ix, iy, iz = v % sx, (v // sx) % sy, v // (sx * sy)
        u2 = (q["ux"] * q["ux"] + (q["uy"] * q["uy"] + q["uz"] * q["uz"])).astype(np.float32)
        rg = (np.float32(1) / np.sqrt(np.float32(1) + u2)).astype(np.float32)
        res = []
        for c, d, u, rd, nn in ((ix, q["dx"], q["ux"], s.rdx, n[0]), (iy, q["dy"], q["uy"], s.rdy, n[1]), (iz, q["dz"], q["uz"], s.rdz, n[2])):
            k = np.float32(2.0 * L * s.cvac * s.dt * rd)
            pos = (d + k * u * rg + np.float32(1)) * np.float32(0.5)
            res.append(np.clip(c + np.floor(pos).astype(np.int64), 1, nn))
        return res[0] + sx * (res[1] + sy * res[2])

    k_out = keys(out)
    # rsqrt on the device is approximate: a key may differ where the predicted position sits on a cell face; the
    # grouping the device produced must be sorted by ITS keys, which partition[] reveals
    dev_key = np.repeat(np.arange(g.nv), np.diff(part))
    assert len(dev_key) == np_ and part[-1] == np_
    assert np.mean(dev_key != k_out) < 1e-3
    assert np.all(np.diff(dev_key) >= 0)
    # (the order inside a group is the slot-claim order, not specified)
    # L = 0 reproduces the plain sort
    d_p.upload(p)
    vpb.vpb_sort_p_planes_ahead(dom, d_p.ptr, d_tmp.ptr, np_, d_part.ptr, 0)
    ref = d_p.download(np_)
    d_p.upload(p)
    vpb.vpb_sort_p_planes(dom, d_p.ptr, d_tmp.ptr, np_, d_part.ptr)
    assert_bits_equal(ref, d_p.download(np_), "lookahead 0")
    for arr in (d_p, d_tmp, d_part):
        arr.free()
    vpb.vpb_domain_destroy(dom)


def group_tables(vpb, n):
    """key(x,y,z) = fx[x] + fy[y] + fz[z] of the grouped sort (host code of the library, csrc/vpb_sort_group.cu)"""
    fx, fy, fz = (np.zeros(m + 2, np.int32) for m in n)
    nkeys = vpb.vpb_sort_group_order(n[0], n[1], n[2], fx.ctypes.data, fy.ctypes.data, fz.ctypes.data)
    return fx, fy, fz, int(nkeys)


@pytest.mark.parametrize("n,np_,L,vth", [((10, 9, 8), 60001, 0, 0.5), ((10, 9, 8), 60001, 7, 0.5), ((33, 1, 17), 150000, 0, 0.3),
                                          ((16, 16, 16), 300000, 4, 0.2), ((5, 4, 3), 63, 0, 0.1), ((40, 36, 20), 2048 * 37 + 5, 3, 0.4),
                                          ((5, 4, 3), 40000, 2, 0.3)])      # ~670 per group: ranks beyond the packed 8 bits
@pytest.mark.parametrize("variant", [2, 3, 1, 0])
def test_sort_p_planes_grouped(vpb, n, np_, L, vth, variant):
    """The device-resident driver's sort (vpb_sort_p_planes_grouped): the same particles, bit for bit, grouped by the
    voxel they occupy (L = 0) or reach L steps ahead, groups in the brick-Morton order of vpb_sort_group_order,
    partition[] = first particle of every group.  Several sizes: ragged last chunk, fewer particles than a warp, more
    chunks than resident CTAs' worth of keys, a 2-D grid.  All three forms of the move pass (sort.group_variant: 2 =
    inverse permutation + destination-ordered gather, the default; 1 = direct scatter; 0 = chunks staged in shared memory)."""
    from old_vpic_b200.sim import DevArray, ParticleArray
    g = host_grid(n)
    rng = np.random.default_rng(41)
    p = random_particles(rng, g, np_, vth=vth, sort=False)
    p["tag"] = np.arange(np_)
    fx, fy, fz, nkeys = group_tables(vpb, n)
    dom = vpb.vpb_domain_create(g.ref(), 0, 1)
    assert vpb.vpb_sort_group_keys(dom) == nkeys
    vpb.vpb_domain_set_particle_layout(dom, (np_ + 63) // 64 * 64)
    d_p, d_out = ParticleArray(vpb, dom, np_), ParticleArray(vpb, dom, np_)
    d_part = DevArray(vpb, nkeys + 1, np.int32)
    d_p.upload(p)
    vpb.vpb_set_tuning(b"sort.group_variant", min(variant, 2))   # 3: the gather form with separate key and rank arrays
    vpb.vpb_set_tuning(b"sort.pack_rank", int(variant == 2))
    try:
        vpb.vpb_sort_p_planes_grouped(dom, d_p.ptr, d_out.ptr, np_, d_part.ptr, L)
    finally:
        vpb.vpb_set_tuning(b"sort.group_variant", 2)
        vpb.vpb_set_tuning(b"sort.pack_rank", 1)
    out, part = d_out.download(np_), d_part.download()
    assert_bits_equal(d_p.download(np_), p, "the input is left alone")
    # same particles, every record intact
    assert_bits_equal(out[np.argsort(out["tag"], kind="stable")], p, "multiset")
    assert part[0] == 0 and part[-1] == np_ and np.all(np.diff(part) >= 0)
    sx, sy = n[0] + 2, n[1] + 2
    s = g.struct

    def keys(q):
        v = q["i"].astype(np.int64)
        c = [v % sx, (v // sx) % sy, v // (sx * sy)]
        if L:
            u2 = (q["ux"] * q["ux"] + (q["uy"] * q["uy"] + q["uz"] * q["uz"])).astype(np.float32)
            rg = (np.float32(1) / np.sqrt(np.float32(1) + u2)).astype(np.float32)
            for a, (d, u, rd) in enumerate(((q["dx"], q["ux"], s.rdx), (q["dy"], q["uy"], s.rdy), (q["dz"], q["uz"], s.rdz))):
                k = np.float32(2.0 * L * s.cvac * s.dt * rd)
                pos = (d + k * u * rg + np.float32(1)) * np.float32(0.5)
                c[a] = np.clip(c[a] + np.floor(pos).astype(np.int64), 1, n[a])
        return fx[c[0]].astype(np.int64) + fy[c[1]] + fz[c[2]]

    dev_key = np.repeat(np.arange(nkeys), np.diff(part))        # the key the device gave every output slot
    assert len(dev_key) == np_
    k_out = keys(out)
    if L == 0:
        assert np.array_equal(dev_key, k_out)                   # integer arithmetic: exact
    else:
        # rsqrt on the device is approximate: a key may differ where the predicted position sits on a cell face
        assert np.mean(dev_key != k_out) < 1e-3
    for arr in (d_p, d_out, d_part):
        arr.free()
    vpb.vpb_domain_destroy(dom)


@pytest.mark.parametrize("n", [(70, 3, 2), (32, 2, 2), (33, 1, 3), (5, 4, 3)])
@pytest.mark.parametrize("stage", [1, 0])
def test_wide_interpolator_store_paths(vpb, orc, n, stage):
    """load_interpolator into the 96-byte device records: through shared memory as whole lines (sf.stage_store = 1, the
    default) or straight from the registers (0) -- the 18 coefficients bit-identical to the oracle's, the six pad floats of
    interior records zero, ghost records untouched."""
    from old_vpic_b200.sim import DevArray, FieldArray
    from helpers import random_fields
    g = host_grid(n)
    f = random_fields(np.random.default_rng(77), g)
    fi_o = abi.aligned_zeros(g.nv, abi.interpolator_dtype)
    orc.orc_load_interpolator(ptr(fi_o), ptr(f), g.ref())
    dom = vpb.vpb_domain_create(g.ref(), 0, 1)
    vpb.vpb_domain_set_field_layout(dom, 1)
    vpb.vpb_domain_set_interpolator_layout(dom, 1)
    d_f = FieldArray(vpb, dom, g.nv)
    d_f.upload(f)
    d_fi = DevArray(vpb, 96 * g.nv, np.uint8)
    d_fi.upload(np.full(96 * g.nv, 0x5a, np.uint8))
    vpb.vpb_set_tuning(b"sf.stage_store", stage)
    try:
        vpb.vpb_load_interpolator(dom, d_fi.ptr, d_f.ptr)
    finally:
        vpb.vpb_set_tuning(b"sf.stage_store", 1)
    got = d_fi.download().view(np.float32).reshape(g.nv, 24)
    sz, sy, sx = g.shape
    iz, iy, ix = np.meshgrid(np.arange(sz), np.arange(sy), np.arange(sx), indexing="ij")
    interior = ((ix >= 1) & (ix <= n[0]) & (iy >= 1) & (iy <= n[1]) & (iz >= 1) & (iz <= n[2])).reshape(-1)
    want = fi_o.view(np.float32).reshape(g.nv, 20)[:, :18]
    assert_bits_equal(got[interior, :18].copy(), want[interior].copy(), "coefficients")
    assert not np.any(got[interior, 18:].view(np.uint32))
    assert np.all(got[~interior].view(np.uint8) == 0x5a)
    d_f.free(); d_fi.free()
    vpb.vpb_domain_destroy(dom)
