"""The traversal hint of advance_p (partition[] from the last sort) must not change results."""
import ctypes as C

import numpy as np
import pytest

from helpers import abi, assert_bits_equal, host_grid, max_rel, random_interpolator, random_particles
from old_vpic_b200.abi import ptr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,by", [((12, 10, 8), 16), ((12, 10, 8), 3), ((20, 1, 9), 4), ((1, 1, 16), 16)])
@pytest.mark.parametrize("kind", ["periodic", "absorbing"])
def test_sort_then_advance_uses_hint_and_matches_oracle(vpb, orc, n, by, kind):
    g = host_grid(n, kind)
    rng = np.random.default_rng(41)
    np_ = 30011
    p = random_particles(rng, g, np_, vth=0.6, sort=False)
    fi = random_interpolator(rng, g, amp=0.3)
    sp = abi.SpeciesStruct()
    sp.np, sp.max_np, sp.p = np_, np_, p.ctypes.data
    part = np.zeros(g.nv + 1, np.int32)
    sp.partition = part.ctypes.data
    vpb.vpb_set_tuning(b"advance_p.by", by)
    vpb.sort_p(C.byref(sp), g.ref())              # registers the layout for this array
    # drift: pretend a few steps passed by scrambling voxel indices of some particles to neighbours
    nx = g.n[0]
    if nx > 2:
        m = rng.random(np_) < 0.4
        x = p["i"] % (nx + 2)
        p["i"][m & (x > 1)] -= 1
    for ordered in (1, 0):
        vpb.vpb_set_tuning(b"advance_p.ordered", ordered)
        p_o, p_g = p.copy(), p.copy()
        a_o = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
        a_g = a_o.copy()
        pm_o = abi.aligned_zeros(np_, abi.mover_dtype)
        pm_g = pm_o.copy()
        nm_o = orc.orc_advance_p(ptr(p_o), np_, -1.0, ptr(pm_o), np_, ptr(a_o), ptr(fi), g.ref())
        # same host array pointer as the one sort_p saw -> the hint is found
        keep = p.copy()
        nm_g = vpb.advance_p(ptr(p), np_, -1.0, ptr(pm_g), np_, ptr(a_g), ptr(fi), g.ref())
        p_g[:] = p
        p[:] = keep
        assert nm_g == nm_o
        assert_bits_equal(p_g, p_o, "particles (ordered=%d)" % ordered)
        assert_bits_equal(pm_g[:nm_g], pm_o[:nm_o], "movers")
        assert max_rel(a_g.view(np.float32), a_o.view(np.float32)) < 2e-5
    vpb.vpb_set_tuning(b"advance_p.ordered", 1)
    vpb.vpb_set_tuning(b"advance_p.by", 16)
