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


def test_custom_boundary_handlers_are_refused():
    """boundary_p.c:271-277: a cell face bound to the deck's k-th custom handler (neighbor = -3-k, k < grid->nb) calls a
    host callback.  The library has no CPU fallback and must not absorb such particles silently: the grid is refused
    with the reference's ERROR convention (message, exit 1) the first time the hot path sees it.  Codes beyond grid->nb
    stay "unknown boundary interaction" (absorbed with a warning by boundary_p)."""
    import subprocess
    import sys
    code = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
from helpers import abi, host_grid
from old_vpic_b200 import lib
from old_vpic_b200.abi import ptr
L = lib.load(); L.vpb_init(0)
g = host_grid((4, 4, 4), "metal")
v = g.voxel(1, 2, 2)
g.neighbor[6 * v + 0] = -3            # handler 0 on one -x face
g.struct.nb = int(sys.argv[1])
a = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
L.clear_accumulators(ptr(a), g.ref())
print("ACCEPTED")
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = code % (root, os.path.join(root, "tests"))
    r = subprocess.run([sys.executable, "-c", src, "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "custom particle-boundary handlers" in r.stderr and "ACCEPTED" not in r.stdout, (r.stdout, r.stderr[-2000:])
    r = subprocess.run([sys.executable, "-c", src, "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ACCEPTED" in r.stdout, (r.stdout, r.stderr[-2000:])
