"""Pins oracle/oracle_mt.c -- the restated Mersenne-Twister stream (util/mtrand/mtrand.c), mt_drand, the ziggurat
mt_drandn with its layer table rebuilt from make_zig.c's construction, and the particle load a deck makes from them
(vpic.hxx:491-505, misc.cxx:16-105) -- to the reference compiled from source: bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import abi, host_grid, loader
from old_vpic_b200.abi import ptr

_vp, _i, _d, _l = C.c_void_p, C.c_int, C.c_double, C.c_long


def oracle_mt():
    O = loader.oracle()
    for name, res, args in (("orc_mt_sizeof", _i, []), ("orc_mt_seed", None, [_vp, C.c_uint]), ("orc_mt_u32", C.c_uint32, [_vp]),
                            ("orc_mt_fill_u32", None, [_vp, _vp, _l]), ("orc_mt_fill_drand", None, [_vp, _vp, _l]),
                            ("orc_mt_fill_drandn", None, [_vp, _vp, _l]), ("orc_mt_draw", None, [_vp, C.c_char_p, _l, _vp]),
                            ("orc_mt_zig_table", None, [_vp, _vp, _vp]),
                            ("orc_load_thermal_pairs", _l, [_vp, _l, _vp, _vp, _d, _d, _d, _vp, _vp, _i, _vp, _vp, _i, _vp]),
                            ("orc_load_thermal_pairs_tagged", _l, [_vp, _l, _vp, _vp, _d, _d, _d, _vp, _vp, _i, _vp, _vp, _i, _vp, _l, _l])):
        fn = getattr(O, name)
        fn.restype, fn.argtypes = res, args
    return O


def new_rng(O, seed):
    rng = np.zeros(O.orc_mt_sizeof(), np.uint8)
    O.orc_mt_seed(ptr(rng), seed)
    return rng


needs_ref = pytest.mark.skipif(not loader.ref_available("scalar"), reason="oracle/_ref not built (needs /root/reference at build time)")


def ref_mt():
    R = loader.ref("scalar")
    for name, res, args in (("new_mt_rng", _vp, [C.c_uint]), ("delete_mt_rng", None, [_vp]), ("seed_mt_rng", None, [_vp, C.c_uint]),
                            ("mt_urand_fill", None, [_vp, _vp, C.c_size_t]), ("mt_drand_fill", None, [_vp, _vp, C.c_size_t]),
                            ("mt_drandn_fill", None, [_vp, _vp, C.c_size_t])):
        fn = getattr(R, name)
        fn.restype, fn.argtypes = res, args
    return R


@needs_ref
@pytest.mark.parametrize("seed", [0, 1, 7, 123456789, 0xffffffff])
def test_word_stream_and_uniform_deviates(seed):
    O, R = oracle_mt(), ref_mt()
    n = 5 * 624 + 17
    r = R.new_mt_rng(seed)
    want = np.zeros(n, np.uint32)
    R.mt_urand_fill(r, ptr(want), n)
    wd = np.zeros(1000, np.float64)
    R.mt_drand_fill(r, ptr(wd), len(wd))
    R.delete_mt_rng(r)
    rng = new_rng(O, seed)
    got = np.zeros(n, np.uint32)
    O.orc_mt_fill_u32(ptr(rng), ptr(got), n)
    gd = np.zeros(1000, np.float64)
    O.orc_mt_fill_drand(ptr(rng), ptr(gd), len(gd))
    assert np.array_equal(got, want)
    assert np.array_equal(gd.view(np.uint64), wd.view(np.uint64))


@needs_ref
def test_normal_deviates_including_rejections_and_tail():
    """5e6 draws: ~1.2 % take the rejection branch and ~1e-4 the tail layer (log/exp of the host libm on both sides)"""
    O, R = oracle_mt(), ref_mt()
    n = 5_000_000
    r = R.new_mt_rng(7)
    want = np.zeros(n, np.float64)
    R.mt_drandn_fill(r, ptr(want), n)
    tail_w = np.zeros(8, np.uint32)
    R.mt_urand_fill(r, ptr(tail_w), 8)            # the generators must also END at the same word
    R.delete_mt_rng(r)
    rng = new_rng(O, 7)
    got = np.zeros(n, np.float64)
    O.orc_mt_fill_drandn(ptr(rng), ptr(got), n)
    tail_g = np.zeros(8, np.uint32)
    O.orc_mt_fill_u32(ptr(rng), ptr(tail_g), 8)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    assert np.array_equal(tail_g, tail_w)
    assert (np.abs(got) > 3.6554204190269413).sum() > 100          # the tail branch was visited
    assert abs(got.std() - 1) < 2e-3 and abs(got.mean()) < 2e-3


def test_ziggurat_table_is_a_ziggurat():
    """the rebuilt layer table: equal-area layers under exp(-x^2/2) (make_zig.c:25-41), to the accuracy its double sqrt allows"""
    O = oracle_mt()
    x, y, r = np.zeros(257), np.zeros(257), np.zeros(1)
    O.orc_mt_zig_table(ptr(x), ptr(y), ptr(r))
    assert x[0] == 0 and y[0] == 1 and x[255] == r[0] and abs(r[0] - 3.6554204190269413) < 1e-15
    assert np.all(np.diff(x) > 0) and np.all(np.diff(y) < 0)
    assert np.allclose(y, np.exp(-0.5 * x * x), rtol=1e-13)
    v = r[0] * np.exp(-0.5 * r[0] ** 2) + np.exp(-0.5 * r[0] ** 2) / r[0]
    area = x[2:256] * (y[1:255] - y[2:256])
    assert np.allclose(area, v, rtol=1e-10)


@needs_ref
@pytest.mark.parametrize("cells,ppc", [(6, 5), (9, 3)])
def test_deck_load_matches_the_reference_initialize(cells, ppc, tmp_path):
    """The reference's own initialize() runs oracle/decks/thermal_c1.cxx (seed_rand(7), then per pair three uniform_rand
    and six maxwellian_rand through inject_particle) and dumps both particle arrays; the oracle's loader must leave
    the same bytes."""
    exe = os.path.join(loader.REF_DIR, "thermal_c1.op")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/thermal_c1.op not built")
    dump = tmp_path / "load.bin"
    env = dict(os.environ, VPB_DECK_CELLS=str(cells), VPB_DECK_PPC=str(ppc), VPB_DECK_STEPS="1", VPB_DECK_DUMP_LOAD=str(dump),
               VPB_DECK_ENERGIES="0")
    r = subprocess.run([exe, "-tpp=1"], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and dump.exists(), (r.stdout + r.stderr)[-2000:]
    raw = np.fromfile(dump, np.uint8)
    cnt = raw[:8].view(np.int32)
    n = cells ** 3 * ppc
    assert tuple(cnt) == (n, n)
    want_e = raw[8:8 + 48 * n].view(abi.particle_dtype)
    want_i = raw[8 + 48 * n:8 + 96 * n].view(abi.particle_dtype)
    O = oracle_mt()
    import old_vpic_b200.grid as G
    g = host_grid((cells,) * 3, "periodic", L=(float(cells),) * 3, dt=0.95 * G.courant_dt(1.0, 1.0, 1.0, frac=1.0))
    rng = new_rng(O, 7)
    pe, pi = abi.aligned_zeros(n + 8, abi.particle_dtype), abi.aligned_zeros(n + 8, abi.particle_dtype)
    npe, npi = C.c_int(0), C.c_int(0)
    lo, hi = np.zeros(3), np.full(3, float(cells))
    q = float(cells) ** 3 / n
    done = O.orc_load_thermal_pairs(ptr(rng), n, ptr(lo), ptr(hi), 0.1, 0.1, q, ptr(pe), C.byref(npe), n + 8, ptr(pi), C.byref(npi), n + 8,
                                    g.ref())
    assert done == n and npe.value == n and npi.value == n
    hot = ["dx", "dy", "dz", "i", "ux", "uy", "uz", "q", "tag"]
    for name in hot:
        assert np.array_equal(pe[name][:n].view(np.uint32 if pe[name].dtype.itemsize == 4 else np.uint64),
                              want_e[name].view(np.uint32 if pe[name].dtype.itemsize == 4 else np.uint64)), ("electron", name)
        assert np.array_equal(pi[name][:n].view(np.uint32 if pi[name].dtype.itemsize == 4 else np.uint64),
                              want_i[name].view(np.uint32 if pi[name].dtype.itemsize == 4 else np.uint64)), ("ion", name)


def test_library_builds_the_same_layer_table():
    """csrc/vpb_mt.cu rebuilds the ziggurat table on the host at start-up (no GPU involved): the same doubles as the
    oracle's, which test_normal_deviates_including_rejections_and_tail pins to the reference's"""
    from old_vpic_b200 import lib
    L, O = lib.load(), oracle_mt()
    a, b = [np.zeros(257), np.zeros(257), np.zeros(1)], [np.zeros(257), np.zeros(257), np.zeros(1)]
    L.vpb_mt_ziggurat_table(*[ptr(v) for v in a])
    O.orc_mt_zig_table(*[ptr(v) for v in b])
    for u, v in zip(a, b):
        assert np.array_equal(u.view(np.uint64), v.view(np.uint64))


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_against_the_committed_reference_fixtures():
    """tests/golden/ref_mt_stream.npz and ref_thermal_c1_load.npz were written by the reference itself
    (tests/golden/make_mt_golden.py); they travel to the GPU box, where /root/reference does not exist."""
    O = oracle_mt()
    G = np.load(os.path.join(GOLDEN, "ref_mt_stream.npz"))
    for seed in (0, 7, 0xfffffffe):
        rng = new_rng(O, seed)
        w, u, n = np.zeros(2000, np.uint32), np.zeros(500), np.zeros(4000)
        O.orc_mt_fill_u32(ptr(rng), ptr(w), len(w))
        O.orc_mt_fill_drand(ptr(rng), ptr(u), len(u))
        O.orc_mt_fill_drandn(ptr(rng), ptr(n), len(n))
        assert np.array_equal(w, G["words_%d" % seed])
        assert np.array_equal(u.view(np.uint64), G["drand_%d" % seed].view(np.uint64))
        assert np.array_equal(n.view(np.uint64), G["drandn_%d" % seed].view(np.uint64))
    rng = new_rng(O, 7)
    big = np.zeros(250000)
    O.orc_mt_fill_drandn(ptr(rng), ptr(big), len(big))
    assert np.array_equal(big[G["tail_where_7"]].view(np.uint64), G["tail_value_7"].view(np.uint64))
    D = np.load(os.path.join(GOLDEN, "ref_thermal_c1_load.npz"))
    cells, ppc = int(D["cells"]), int(D["ppc"])
    n = cells ** 3 * ppc
    import old_vpic_b200.grid as G2
    g = host_grid((cells,) * 3, "periodic", L=(float(cells),) * 3, dt=0.95 * G2.courant_dt(1.0, 1.0, 1.0, frac=1.0))
    rng = new_rng(O, int(D["seed"]))
    pe, pi = abi.aligned_zeros(n, abi.particle_dtype), abi.aligned_zeros(n, abi.particle_dtype)
    npe, npi = C.c_int(0), C.c_int(0)
    lo, hi = np.zeros(3), np.full(3, float(cells))
    q = float(cells) ** 3 / n
    assert O.orc_load_thermal_pairs(ptr(rng), n, ptr(lo), ptr(hi), float(D["vth"]), float(D["vth"]), q, ptr(pe), C.byref(npe), n, ptr(pi),
                                    C.byref(npi), n, g.ref()) == n
    for name in ("dx", "dy", "dz", "i", "ux", "uy", "uz", "q"):
        assert np.array_equal(pe[name].view(np.uint32), D["electron"][name].view(np.uint32)), ("electron", name)
        assert np.array_equal(pi[name].view(np.uint32), D["ion"][name].view(np.uint32)), ("ion", name)
