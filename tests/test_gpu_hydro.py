"""GPU parity: hydro moments (accumulate_hydro_p, local_adjust_hydro, synchronize_hydro) through the
reference-named C ABI against the CPU oracle (pinned bit-exact to the reference by
tests/test_oracle_vs_ref.py::test_accumulate_hydro_p / test_synchronize_hydro).

Per-particle contributions are bit-identical (same scalar arithmetic, the reference's double-literal
expressions included); node sums differ only in the order of float additions (atomics): tolerance 2e-5 of the
largest entry of each moment.  The boundary operations (x2, lw*mine + rw*theirs) are bit-exact."""
import numpy as np
import pytest

from helpers import abi, assert_bits_equal, host_grid, random_interpolator, random_particles
from old_vpic_b200.abi import ptr

pytestmark = pytest.mark.gpu

TOL = 2e-5
NAMES = ("jx", "jy", "jz", "rho", "px", "py", "pz", "ke", "txx", "tyy", "tzz", "tyz", "tzx", "txy")


def check_moments(h_g, h_o):
    for n in NAMES:
        scale = float(np.max(np.abs(h_o[n])))
        assert scale > 0, n
        assert float(np.max(np.abs(h_g[n] - h_o[n]))) <= TOL * scale, n
    assert not np.any(h_g["_pad"])


@pytest.mark.parametrize("planes", [0, 1])
@pytest.mark.parametrize("kind", ["periodic", "metal", "absorbing"])
@pytest.mark.parametrize("n,np_", [((6, 5, 4), 5000), ((8, 1, 6), 7001), ((12, 12, 12), 12 * 12 * 12 * 16)])
def test_accumulate_hydro_p(vpb, orc, planes, kind, n, np_):
    g = host_grid(n, kind)
    rng = np.random.default_rng(51)
    p = random_particles(rng, g, np_, vth=0.8)
    p["q"] = rng.uniform(0.5, 1.5, np_).astype(np.float32)
    fi = random_interpolator(rng, g, amp=0.4)
    vpb.vpb_set_tuning(b"dropin.particle_planes", planes)
    try:
        for q_m in (-1.0, 0.25):
            h_o = abi.aligned_zeros(g.nv, abi.hydro_dtype)
            h_g = h_o.copy()
            orc.orc_accumulate_hydro_p(ptr(h_o), ptr(p), np_, q_m, ptr(fi), g.ref())
            vpb.accumulate_hydro_p(ptr(h_g), ptr(p), np_, q_m, ptr(fi), g.ref())
            check_moments(h_g, h_o)
            # nodes the oracle left untouched stay untouched
            flat_o = h_o.view(np.float32).reshape(-1, 16)
            flat_g = h_g.view(np.float32).reshape(-1, 16)
            assert not np.any(flat_g[~np.any(flat_o != 0, axis=1)])
    finally:
        vpb.vpb_set_tuning(b"dropin.particle_planes", 0)


def test_accumulate_hydro_p_single_particle_bits(vpb, orc):
    """One particle: no summation order involved, every moment of every node must match bit for bit."""
    g = host_grid((5, 4, 3))
    rng = np.random.default_rng(52)
    fi = random_interpolator(rng, g, amp=0.4)
    for trial in range(20):
        p = random_particles(rng, g, 1, vth=1.5)
        h_o = abi.aligned_zeros(g.nv, abi.hydro_dtype)
        h_g = h_o.copy()
        orc.orc_accumulate_hydro_p(ptr(h_o), ptr(p), 1, -0.7, ptr(fi), g.ref())
        vpb.accumulate_hydro_p(ptr(h_g), ptr(p), 1, -0.7, ptr(fi), g.ref())
        assert_bits_equal(h_g, h_o, "trial %d" % trial)


@pytest.mark.parametrize("kind", ["periodic", "metal", "absorbing"])
@pytest.mark.parametrize("n", [(6, 5, 4), (8, 1, 6), (1, 1, 16)])
def test_synchronize_hydro(vpb, orc, kind, n):
    g = host_grid(n, kind)
    rng = np.random.default_rng(53)
    h = abi.aligned_zeros(g.nv, abi.hydro_dtype)
    h.view(np.float32)[:] = rng.standard_normal(h.view(np.float32).shape).astype(np.float32)
    h_o, h_g = h.copy(), h.copy()
    orc.orc_local_adjust_hydro(ptr(h_o), g.ref(), 1)
    vpb.local_adjust_hydro(ptr(h_g), g.ref())
    assert_bits_equal(h_g, h_o, "local_adjust_hydro")
    h_o, h_g = h.copy(), h.copy()
    orc.orc_synchronize_hydro(ptr(h_o), g.ref(), 0, 1)
    vpb.synchronize_hydro(ptr(h_g), g.ref())
    assert_bits_equal(h_g, h_o, "synchronize_hydro")


def test_clear_and_structors(vpb):
    import ctypes as C
    g = host_grid((4, 3, 2))
    addr = vpb.new_hydro(g.ref())
    view = np.ctypeslib.as_array(C.cast(addr, C.POINTER(C.c_float)), shape=(g.nv * 16,))
    assert not np.any(view)
    view[:] = 1.0
    vpb.clear_hydro(addr, g.ref())
    assert not np.any(view)
    vpb.delete_hydro(addr)


def test_charge_is_conserved(vpb):
    """Size-independent property: after synchronize_hydro on a periodic box the rho moments of the distinct nodes sum
    to the total charge per cell volume."""
    n = (16, 16, 16)
    g = host_grid(n, "periodic")
    rng = np.random.default_rng(54)
    np_ = 16 ** 3 * 8
    p = random_particles(rng, g, np_, vth=0.3)
    fi = random_interpolator(rng, g, amp=0.1)
    h = abi.aligned_zeros(g.nv, abi.hydro_dtype)
    vpb.accumulate_hydro_p(ptr(h), ptr(p), np_, -1.0, ptr(fi), g.ref())
    vpb.synchronize_hydro(ptr(h), g.ref())
    rho = h["rho"].reshape(n[2] + 2, n[1] + 2, n[0] + 2)[1:n[2] + 1, 1:n[1] + 1, 1:n[0] + 1]
    s = g.struct
    total = float(rho.astype(np.float64).sum()) * s.dx * s.dy * s.dz
    assert total == pytest.approx(float(p["q"].astype(np.float64).sum()), rel=1e-5)
