"""GPU parity: boundary_p (single rank: absorbing walls) and the one-mover entry points, against the oracle.

After a removal the reference's serial back-fill loop fixes an ORDER of the survivors that depends on its
visiting order (boundary_p.c:243-247); the device back-fill is order-free.  So the particle
arrays are compared as multisets keyed by the particle tag; counts are exact; rhob within the float-sum
tolerance."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import abi, assert_bits_equal, host_grid, max_rel, random_fields, random_interpolator, random_particles
from old_vpic_b200.abi import ptr

pytestmark = pytest.mark.gpu


def by_tag(p):
    return p[np.argsort(p["tag"], kind="stable")]


@pytest.mark.parametrize("n,np_", [((6, 5, 4), 4000), ((12, 12, 12), 60000), ((1, 1, 16), 500)])
def test_boundary_p_absorbing(vpb, orc, n, np_):
    g = host_grid(n, "absorbing")
    rng = np.random.default_rng(31)
    p = random_particles(rng, g, np_, vth=0.7, edge_frac=0.02)
    fi = random_interpolator(rng, g, amp=0.2)
    f = random_fields(rng, g)
    a = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    pm = abi.aligned_zeros(np_, abi.mover_dtype)
    # both sides start from the oracle's post-advance_p state (particles bit-identical by test_advance_p)
    nm = orc.orc_advance_p(ptr(p), np_, -1.0, ptr(pm), np_, ptr(a), ptr(fi), g.ref())
    assert nm > 0
    p_o, f_o = p.copy(), f.copy()
    out = [abi.aligned_zeros(nm + 1, abi.injector_dtype) for _ in range(6)]
    outp = (C.c_void_p * 6)(*[o.ctypes.data for o in out])
    n_out = (C.c_int * 6)()
    np_o = orc.orc_boundary_p_pack(ptr(p_o), np_, ptr(pm), nm, 0, ptr(f_o), g.ref(), 0, 1, outp, n_out)
    assert sum(n_out) == 0 and np_o == np_ - nm
    # device, through the reference-named entry point on a species_t
    p_g, f_g, pm_g = p.copy(), f.copy(), pm.copy()
    sp = abi.SpeciesStruct()
    sp.id, sp.np, sp.max_np, sp.p = 0, np_, np_, p_g.ctypes.data
    sp.nm, sp.max_nm, sp.pm = nm, np_, pm_g.ctypes.data
    vpb.boundary_p(C.byref(sp), ptr(f_g), ptr(a), g.ref(), None)
    assert sp.np == np_o and sp.nm == 0
    assert_bits_equal(by_tag(p_g[:sp.np]), by_tag(p_o[:np_o]), "surviving particles (multiset by tag)")
    assert max_rel(f_g["rhob"], f_o["rhob"]) < 2e-5
    for k in abi.field_dtype.names:
        if k != "rhob":
            assert np.array_equal(np.ascontiguousarray(f_g[k]), np.ascontiguousarray(f_o[k])), k


def test_boundary_p_nothing_to_do(vpb):
    g = host_grid((4, 4, 4))
    rng = np.random.default_rng(1)
    p = random_particles(rng, g, 100)
    pm = abi.aligned_zeros(16, abi.mover_dtype)
    sp = abi.SpeciesStruct()
    sp.id, sp.np, sp.max_np, sp.p, sp.nm, sp.max_nm, sp.pm = 0, 100, 100, p.ctypes.data, 0, 16, pm.ctypes.data
    before = p.copy()
    vpb.boundary_p(C.byref(sp), None, None, g.ref(), None)
    assert sp.np == 100 and sp.nm == 0
    assert_bits_equal(p, before, "untouched")


def test_move_p_and_accumulate_rhob_single(vpb, orc):
    g = host_grid((5, 4, 3), "metal")
    rng = np.random.default_rng(5)
    p = random_particles(rng, g, 32, vth=0.5)
    a_o = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    a_g = a_o.copy()
    p_o, p_g = p.copy(), p.copy()
    for k in range(0, 32, 5):
        m = np.zeros(1, abi.mover_dtype)
        m["dispx"], m["dispy"], m["dispz"] = rng.uniform(-1.5, 1.5, 3)
        m["i"] = k
        m_o, m_g = m.copy(), m.copy()
        assert orc.orc_move_p(ptr(p_o), ptr(m_o), ptr(a_o), g.ref()) == vpb.move_p(ptr(p_g), ptr(m_g), ptr(a_g), g.ref())
        assert_bits_equal(m_g, m_o, "mover")
    assert_bits_equal(p_g, p_o, "particles")
    assert max_rel(a_g.view(np.float32), a_o.view(np.float32)) < 2e-5
    f = random_fields(rng, g)
    f_o, f_g = f.copy(), f.copy()
    one = p[3:4].copy()
    orc.orc_accumulate_rhob(ptr(f_o), ptr(one), g.ref())
    vpb.accumulate_rhob(ptr(f_g), ptr(one), g.ref())
    assert max_rel(f_g["rhob"], f_o["rhob"]) < 1e-6


class _BoundaryT(C.Structure):
    """grid.h:50-53 boundary_t"""
    _fields_ = [("handler", C.c_void_p), ("params", C.c_char * 1024)]


def _handler_table(R, ut_perp, ut_para, ids):
    """grid_t.boundary for two of the reference's own handlers (src/boundary): 0 = maxwellian_reflux, 1 = absorb_tally"""
    t = (_BoundaryT * 2)()
    t[0].handler = C.cast(R.maxwellian_reflux, C.c_void_p).value
    mr = np.zeros(64, np.float32)                      # maxwellian_reflux_t: ut_perp[32], ut_para[32] by species id
    mr[:len(ut_perp)], mr[32:32 + len(ut_para)] = ut_perp, ut_para
    C.memmove(C.addressof(t[0]) + _BoundaryT.params.offset, mr.ctypes.data, mr.nbytes)
    t[1].handler = C.cast(R.absorb_tally, C.c_void_p).value
    at = np.zeros(65, np.int32)                        # absorb_tally_t: nspec, id[32], nabs[32]
    at[0], at[1:1 + len(ids)] = len(ids), ids
    C.memmove(C.addressof(t[1]) + _BoundaryT.params.offset, at.ctypes.data, at.nbytes)
    return t


def _tally(t):
    return np.frombuffer(C.string_at(C.addressof(t[1]) + _BoundaryT.params.offset + 4 * 33, 4 * 32), np.int32).copy()


def test_boundary_p_runs_the_decks_handlers_on_the_host(vpb, orc, ref_scalar):
    """boundary_p.c:271-277: a mover that ends on a cell face bound to one of the deck's custom handlers (neighbor code
    -3-k) is handed to that HOST callback and destroyed; the injectors the callbacks make are injected after the received
    buffers.  Here the whole -x wall refluxes (the reference's own maxwellian_reflux: draws from the host RNG, makes an
    injector) and the +x wall tallies (absorb_tally: counts, calls back into accumulate_rhob); the other walls absorb.
    The reference's boundary_p on the host against the library's reference-named boundary_p with the SAME callbacks:
    three rounds, two species -- counts, particle sets and pending movers exact (the handlers' random draws come in the
    reference's order: the generators end on the same word), tallies equal, rhob and accumulators to the float-sum
    tolerance."""
    from helpers import RefGrid
    R = ref_scalar
    R.new_mt_rng.restype, R.new_mt_rng.argtypes = C.c_void_p, [C.c_uint]
    R.mt_urand_fill.restype, R.mt_urand_fill.argtypes = None, [C.c_void_p, C.c_void_p, C.c_size_t]
    n = (6, 5, 4)
    g = RefGrid(R, n, "absorbing")
    sx, sy = n[0] + 2, n[1] + 2
    nbr = g.neighbor
    for z in range(1, n[2] + 1):
        for y in range(1, n[1] + 1):
            nbr[6 * (1 + sx * (y + sy * z)) + 0] = -3          # handler 0 behind the -x wall
            nbr[6 * (n[0] + sx * (y + sy * z)) + 3] = -4       # handler 1 behind the +x wall
    rng = np.random.default_rng(41)
    fi = random_interpolator(rng, g, amp=0.2)
    f0 = random_fields(rng, g)
    npk, cap = 4000, 6000
    start = []
    acc = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
    for sid, q in ((0, -1.0), (1, 0.5)):
        p = abi.aligned_zeros(cap, abi.particle_dtype)
        p[:npk] = random_particles(rng, g, npk, vth=0.8, q=q)
        pm = abi.aligned_zeros(cap, abi.mover_dtype)
        nm = orc.orc_advance_p(ptr(p), npk, q, ptr(pm), cap, ptr(acc), ptr(fi), g.ref())
        assert nm > 100
        start.append((sid, p, pm, nm))

    def species_list():
        arrays, sps = [], [abi.SpeciesStruct() for _ in start]
        for sp, (sid, p, pm, nm) in zip(sps, start):
            pc, pmc = p.copy(), pm.copy()
            arrays.append((pc, pmc))
            sp.id, sp.np, sp.max_np, sp.p = sid, npk, cap, pc.ctypes.data
            sp.nm, sp.max_nm, sp.pm = nm, cap, pmc.ctypes.data
            sp.q_m = 1.0
        sps[0].next = C.pointer(sps[1])
        return sps, arrays

    def hot_rows(p, k):
        a = np.ascontiguousarray(p[:k]).view(np.uint8).reshape(k, 48)[:, :32].copy().view(np.uint32).reshape(k, 8)
        return a[np.lexsort(a.T[::-1])]

    def movers(pm, p, k):
        if not k:
            return np.zeros((0, 11), np.uint32)
        rows = np.concatenate([np.ascontiguousarray(pm[:k]).view(np.uint32).reshape(k, 4)[:, :3],
                               np.ascontiguousarray(p[pm["i"][:k]]).view(np.uint8).reshape(k, 48)[:, :32].copy().view(np.uint32).reshape(k, 8)], axis=1)
        return rows[np.lexsort(rows.T[::-1])]

    sp_r, arr_r = species_list()
    sp_g, arr_g = species_list()
    f_r, f_g, a_r, a_g = f0.copy(), f0.copy(), acc.copy(), acc.copy()
    t_r, t_g = _handler_table(R, [0.3, 0.2], [0.25, 0.15], [0, 1]), _handler_table(R, [0.3, 0.2], [0.25, 0.15], [0, 1])
    rng_r, rng_g = R.new_mt_rng(5), R.new_mt_rng(5)
    g.struct.nb = 2
    refluxed = 0
    for rnd in range(3):
        g.struct.boundary = C.addressof(t_r)
        R.boundary_p(C.byref(sp_r[0]), ptr(f_r), ptr(a_r), g.ref(), C.c_void_p(rng_r))
        g.struct.boundary = C.addressof(t_g)
        vpb.boundary_p(C.byref(sp_g[0]), ptr(f_g), ptr(a_g), g.ref(), C.c_void_p(rng_g))
        for j in range(2):
            assert (sp_g[j].np, sp_g[j].nm) == (sp_r[j].np, sp_r[j].nm), ("counts", rnd, j, sp_g[j].np, sp_g[j].nm, sp_r[j].np, sp_r[j].nm)
            assert np.array_equal(hot_rows(arr_g[j][0], sp_g[j].np), hot_rows(arr_r[j][0], sp_r[j].np)), ("particle set", rnd, j)
            assert np.array_equal(movers(arr_g[j][1], arr_g[j][0], sp_g[j].nm), movers(arr_r[j][1], arr_r[j][0], sp_r[j].nm)), ("movers", rnd, j)
        assert np.array_equal(_tally(t_g), _tally(t_r))
        scale = max(float(np.abs(f_r["rhob"]).max()), 1e-30)
        assert float(np.abs(f_g["rhob"] - f_r["rhob"]).max()) <= 2e-5 * scale
        assert max_rel(a_g.view(np.float32).reshape(-1, 12), a_r.view(np.float32).reshape(-1, 12)) < 2e-5
        if rnd == 0:
            refluxed = sum(start[j][3] for j in range(2)) - sum(npk - sp_r[j].np for j in range(2))
    assert _tally(t_r)[:2].sum() > 50 and refluxed > 50          # both handlers were busy
    w_r, w_g = np.zeros(16, np.uint32), np.zeros(16, np.uint32)
    R.mt_urand_fill(C.c_void_p(rng_r), ptr(w_r), 16)
    R.mt_urand_fill(C.c_void_p(rng_g), ptr(w_g), 16)
    assert np.array_equal(w_r, w_g)                               # the same number of draws, in the same order


def test_custom_boundary_handlers_are_refused_without_their_host_program():
    """Only the reference-named boundary_p() can run the deck's handlers (host callbacks).  The device-resident driver
    has none: it refuses such a grid with the reference's ERROR convention (message, exit 1) instead of absorbing those
    particles silently.  Codes beyond grid->nb stay "unknown boundary interaction" (absorbed with a warning by boundary_p)."""
    import subprocess
    import sys
    code = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
from helpers import abi, host_grid
from old_vpic_b200 import lib
from old_vpic_b200.abi import ptr
from old_vpic_b200.sim import NativeSimulation
L = lib.load(); L.vpb_init(0)
g = host_grid((4, 4, 4), "metal")
v = g.voxel(1, 2, 2)
g.neighbor[6 * v + 0] = -3            # handler 0 on one -x face
g.struct.nb = int(sys.argv[1])
a = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
L.clear_accumulators(ptr(a), g.ref())  # the reference-named entry points take the grid
print("LAYER_A_OK")
sim = NativeSimulation(g, L=L)
print("ACCEPTED")
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = code % (root, os.path.join(root, "tests"))
    r = subprocess.run([sys.executable, "-c", src, "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "custom particle-boundary handlers" in r.stderr and "LAYER_A_OK" in r.stdout and \
        "ACCEPTED" not in r.stdout, (r.stdout, r.stderr[-2000:])
    r = subprocess.run([sys.executable, "-c", src, "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ACCEPTED" in r.stdout, (r.stdout, r.stderr[-2000:])
