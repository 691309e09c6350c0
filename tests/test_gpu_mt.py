"""GPU: the reference's random-number stream and the deck's particle load on the device (csrc/vpb_mt.cu, SURVEY.md
8f-4) against the CPU oracle (oracle/oracle_mt.c, pinned bit for bit to the compiled reference and to the
reference's own initialize() by tests/test_oracle_mt.py): words, uniform and normal deviates, generator state
hand-over, and the loaded particle arrays -- all bit-exact."""
import ctypes as C

import numpy as np
import pytest

from helpers import abi, assert_bits_equal, host_grid
from old_vpic_b200.abi import ptr
from test_oracle_mt import new_rng, oracle_mt

pytestmark = pytest.mark.gpu


def dev(vpb, n, dtype):
    from old_vpic_b200.sim import DevArray
    return DevArray(vpb, n, dtype)


@pytest.mark.parametrize("seed", [0, 7, 0xfffffffe])
def test_word_stream(vpb, seed):
    O = oracle_mt()
    rng_o = new_rng(O, seed)
    rng = vpb.vpb_mt_create(seed)
    for n in (5, 624, 3 * 624 + 17, 1, 40000):          # across block boundaries, leftovers kept between calls
        want = np.zeros(n, np.uint32)
        O.orc_mt_fill_u32(ptr(rng_o), ptr(want), n)
        d = dev(vpb, n, np.uint32)
        vpb.vpb_mt_words(rng, d.ptr, n)
        assert np.array_equal(d.download(), want), n
        d.free()
    vpb.vpb_mt_destroy(rng)


@pytest.mark.parametrize("prog,n", [("U", 100000), ("N", 3000000), ("UUUNNNNNN", 400000), ("NUN", 7), ("NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN", 5000),
                                    ("N", 1)])
def test_deviates_match_the_host_calls(vpb, prog, n):
    """n records of a token program: the same doubles as the n*len host calls, rejection rounds and tail values
    included, and the generator left on the same word"""
    O = oracle_mt()
    rng_o = new_rng(O, 7)
    want = np.zeros(n * len(prog), np.float64)
    O.orc_mt_draw(ptr(rng_o), prog.encode(), n, ptr(want))
    rng = vpb.vpb_mt_create(7)
    d = dev(vpb, n * len(prog), np.float64)
    vpb.vpb_mt_draw(rng, prog.encode(), n, d.ptr)
    got = d.download()
    bad = np.flatnonzero(got.view(np.uint64) != want.view(np.uint64))
    assert len(bad) == 0, (len(bad), bad[:5], got[bad[:5]], want[bad[:5]])
    if prog == "N" and n > 1000000:
        assert (np.abs(got) > 3.6554204190269413).sum() > 100      # tail values were among them
    # the streams continue together
    w_o, w_g = np.zeros(100, np.uint32), dev(vpb, 100, np.uint32)
    O.orc_mt_fill_u32(ptr(rng_o), ptr(w_o), 100)
    vpb.vpb_mt_words(rng, w_g.ptr, 100)
    assert np.array_equal(w_g.download(), w_o)
    for a in (d, w_g):
        a.free()
    vpb.vpb_mt_destroy(rng)


def test_generator_state_hand_over(vpb):
    """set_mt_rng_state / get_mt_rng_state format (mtrand.c:74-124): a host generator in mid-block (odd word position)
    goes to the device, the device draws, the state comes back and the host stream continues where the device stopped"""
    O = oracle_mt()
    rng_o = new_rng(O, 99)
    skip = np.zeros(1001, np.uint32)
    O.orc_mt_fill_u32(ptr(rng_o), ptr(skip), len(skip))
    rng = vpb.vpb_mt_create(1)
    vpb.vpb_mt_set_state(rng, ptr(rng_o))               # {next, state[624]} little-endian = the oracle's struct
    n = 50000
    want = np.zeros(2 * n, np.float64)
    O.orc_mt_draw(ptr(rng_o), b"UN", n, ptr(want))
    d = dev(vpb, 2 * n, np.float64)
    vpb.vpb_mt_draw(rng, b"UN", n, d.ptr)
    assert np.array_equal(d.download().view(np.uint64), want.view(np.uint64))
    back = np.zeros(O.orc_mt_sizeof(), np.uint8)
    vpb.vpb_mt_get_state(rng, ptr(back))
    a, b = np.zeros(2000, np.uint32), np.zeros(2000, np.uint32)
    O.orc_mt_fill_u32(ptr(rng_o), ptr(a), len(a))
    O.orc_mt_fill_u32(ptr(back), ptr(b), len(b))
    assert np.array_equal(a, b)
    d.free()
    vpb.vpb_mt_destroy(rng)


@pytest.mark.parametrize("n_cells,ppc,topo,rank,planes", [((6, 6, 6), 5, (1, 1, 1), 0, False), ((12, 10, 8), 40, (1, 1, 1), 0, True),
                                                          ((8, 6, 4), 9, (2, 1, 1), 1, False), ((8, 6, 4), 9, (2, 2, 1), 2, True)])
def test_thermal_load_matches_the_serial_loop(vpb, n_cells, ppc, topo, rank, planes):
    """vpb_load_pairs_mt against the oracle's restatement of the deck's loop (inject_particle per particle, g++ argument
    order): both species, every byte that inject_particle writes.  On a rank of a decomposed box the particles that
    fall outside the local domain are skipped like the reference skips them."""
    from old_vpic_b200.sim import ParticleArray
    O = oracle_mt()
    g = host_grid(n_cells, "periodic", topo=topo, rank=rank)
    n = n_cells[0] * n_cells[1] * n_cells[2] * ppc
    cap = n + 64
    lo, hi = np.zeros(3), np.array(n_cells, np.float64)
    q = float(np.prod(n_cells)) / n
    rng_o = new_rng(O, 7)
    pe, pi = abi.aligned_zeros(cap, abi.particle_dtype), abi.aligned_zeros(cap, abi.particle_dtype)
    npe, npi = C.c_int(0), C.c_int(0)
    assert O.orc_load_thermal_pairs_tagged(ptr(rng_o), n, ptr(lo), ptr(hi), 0.1, 0.25, q, ptr(pe), C.byref(npe), cap, ptr(pi), C.byref(npi), cap,
                                           g.ref(), 100, 3) == n        # tag = 100 + 3k, k the loop counter (skipped iterations count)
    if topo == (1, 1, 1):
        assert npe.value == n
    else:
        assert 0 < npe.value < n
    dom = vpb.vpb_domain_create(g.ref(), rank, topo[0] * topo[1] * topo[2])
    if planes:
        vpb.vpb_domain_set_particle_layout(dom, (cap + 63) // 64 * 64)
    d_e, d_i = ParticleArray(vpb, dom, cap), ParticleArray(vpb, dom, cap)
    rng = vpb.vpb_mt_create(7)
    np2 = (C.c_int * 2)(0, 0)
    assert vpb.vpb_load_pairs_mt(dom, rng, n, ptr(lo), ptr(hi), 0.1, 0.25, -q, q, d_e.ptr, cap, d_i.ptr, cap, np2, 1, 100, 3) == n
    assert (np2[0], np2[1]) == (npe.value, npi.value)
    assert_bits_equal(d_e.download(np2[0]), pe[:np2[0]], "electrons")
    assert_bits_equal(d_i.download(np2[1]), pi[:np2[1]], "ions")
    # the generator is where the serial loop left it
    w_o, w_g = np.zeros(64, np.uint32), dev(vpb, 64, np.uint32)
    O.orc_mt_fill_u32(ptr(rng_o), ptr(w_o), 64)
    vpb.vpb_mt_words(rng, w_g.ptr, 64)
    assert np.array_equal(w_g.download(), w_o)
    for a in (d_e, d_i, w_g):
        a.free()
    vpb.vpb_mt_destroy(rng)
    vpb.vpb_domain_destroy(dom)


def test_device_run_from_the_seed_alone_reproduces_the_reference_deck(vpb):
    """oracle/decks/thermal_small.cxx on the reference alone (tests/golden/deck_thermal_small_energies.txt: seed_rand(7),
    the serial load loop, vpic_simulation::initialize(), 20 steps of advance() with both cleanings at step 10 and 20)
    against a run in which NOTHING comes from the host but the seed: the load from the reference's random-number stream on
    the device, vpb_sim_initialize (initialize.cxx:27-95: bound charge, uncenter_p) and the C++ step driver."""
    from old_vpic_b200.sim import NativeSimulation
    from test_gpu_deck import GOLD, read_energies
    want = read_energies(GOLD)
    n, ppc = 16, 8
    ne = n ** 3 * ppc
    g = host_grid((n, n, n), "periodic")
    sim = NativeSimulation(g, L=vpb)
    sim.set_intervals(10, 10)
    # the reference walks its species list from the species defined last (the energies file has "ion" before "electron")
    i = sim.define_species("ion", 1.0, int(1.5 * ne), sort_interval=5)
    e = sim.define_species("electron", -1.0, int(1.5 * ne), sort_interval=5)
    q = float(n) ** 3 / ne
    assert sim.load_pairs_mt(e, i, ne, [0, 0, 0], [n, n, n], 0.1, 0.1, -q, q, seed=7, tag_step=1) == ne     # the deck's tag = k
    assert (e.np, i.np) == (ne, ne)
    sim.set_fields(abi.aligned_zeros(g.nv, abi.field_dtype))
    errs = sim.initialize()
    assert errs[0] == 0 and errs[1] == 0          # no fields yet: nothing to synchronise, no div B
    got = [sim.energies()]
    for _ in range(20):
        sim.advance()
        got.append(sim.energies())
    got = np.array(got)
    assert want.shape == (21, 9) and got.shape == (21, 8)
    rel = np.abs(got - want[:, 1:]) / np.abs(want[:, 1:]).max(axis=0)
    assert rel.max() < 1e-4, rel.max(axis=0)
    # the kinetic energies of the freshly loaded plasma are sums over the very same particles: tighter
    assert np.all(rel[0, 6:] < 1e-6), rel[0]
    sim.free()


def test_device_against_the_committed_reference_fixtures(vpb):
    """The device stream and the device load against what the REFERENCE ITSELF produced (tests/golden/ref_mt_stream.npz,
    ref_thermal_c1_load.npz, written by tests/golden/make_mt_golden.py from oracle/_ref): words, mt_drand, mt_drandn, the
    tail-layer deviates among the first 250000 normals, and both particle arrays of the deck's load loop."""
    import os
    from old_vpic_b200.sim import ParticleArray
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    G = np.load(os.path.join(gold, "ref_mt_stream.npz"))
    for seed in (0, 7, 0xfffffffe):
        rng = vpb.vpb_mt_create(seed)
        w, u, n = dev(vpb, 2000, np.uint32), dev(vpb, 500, np.float64), dev(vpb, 4000, np.float64)
        vpb.vpb_mt_words(rng, w.ptr, 2000)
        vpb.vpb_mt_draw(rng, b"U", 500, u.ptr)
        vpb.vpb_mt_draw(rng, b"N", 4000, n.ptr)
        assert np.array_equal(w.download(), G["words_%d" % seed])
        assert np.array_equal(u.download().view(np.uint64), G["drand_%d" % seed].view(np.uint64))
        assert np.array_equal(n.download().view(np.uint64), G["drandn_%d" % seed].view(np.uint64))
        for a in (w, u, n):
            a.free()
        vpb.vpb_mt_destroy(rng)
    rng = vpb.vpb_mt_create(7)
    big = dev(vpb, 250000, np.float64)
    vpb.vpb_mt_draw(rng, b"N", 250000, big.ptr)
    got = big.download()
    where = np.flatnonzero(np.abs(got) > 3.6554204190269413)
    assert np.array_equal(where, G["tail_where_7"])
    assert np.array_equal(got[where].view(np.uint64), G["tail_value_7"].view(np.uint64))
    big.free()
    vpb.vpb_mt_destroy(rng)
    D = np.load(os.path.join(gold, "ref_thermal_c1_load.npz"))
    cells, ppc = int(D["cells"]), int(D["ppc"])
    n = cells ** 3 * ppc
    import old_vpic_b200.grid as G2
    g = host_grid((cells,) * 3, "periodic", L=(float(cells),) * 3, dt=0.95 * G2.courant_dt(1.0, 1.0, 1.0, frac=1.0))
    dom = vpb.vpb_domain_create(g.ref(), 0, 1)
    d_e, d_i = ParticleArray(vpb, dom, n), ParticleArray(vpb, dom, n)
    rng = vpb.vpb_mt_create(int(D["seed"]))
    np2 = (C.c_int * 2)(0, 0)
    lo, hi = np.zeros(3), np.full(3, float(cells))
    q = float(cells) ** 3 / n
    assert vpb.vpb_load_pairs_mt(dom, rng, n, ptr(lo), ptr(hi), float(D["vth"]), float(D["vth"]), -q, q, d_e.ptr, n, d_i.ptr, n, np2, 1, 0, 0) == n
    pe, pi = d_e.download(n), d_i.download(n)
    for name in ("dx", "dy", "dz", "i", "ux", "uy", "uz", "q"):
        assert np.array_equal(pe[name].view(np.uint32), D["electron"][name].view(np.uint32)), ("electron", name)
        assert np.array_equal(pi[name].view(np.uint32), D["ion"][name].view(np.uint32)), ("ion", name)
    for a in (d_e, d_i):
        a.free()
    vpb.vpb_mt_destroy(rng)
    vpb.vpb_domain_destroy(dom)


def test_draws_and_load_across_internal_batches(vpb):
    """vpb_mt_draw parses at most 2^24 records at a time and vpb_load_pairs_mt loads 2^23 iterations at a time: streams and
    loads longer than that continue seamlessly (16.8 M normal deviates; 8 392 704 load iterations on 16^3 cells)."""
    from old_vpic_b200.sim import ParticleArray
    O = oracle_mt()
    n = (1 << 24) + 12345
    rng_o = new_rng(O, 3)
    want = np.zeros(n, np.float64)
    O.orc_mt_fill_drandn(ptr(rng_o), ptr(want), n)
    rng = vpb.vpb_mt_create(3)
    d = dev(vpb, n, np.float64)
    vpb.vpb_mt_draw(rng, b"N", n, d.ptr)
    got = d.download()
    bad = np.flatnonzero(got.view(np.uint64) != want.view(np.uint64))
    assert len(bad) == 0, (len(bad), bad[:5])
    d.free()
    vpb.vpb_mt_destroy(rng)
    del want, got
    cells, ppc = 16, 2049
    g = host_grid((cells,) * 3, "periodic")
    pairs = cells ** 3 * ppc
    assert pairs > (1 << 23)
    lo, hi = np.zeros(3), np.full(3, float(cells))
    q = float(cells) ** 3 / pairs
    rng_o = new_rng(O, 11)
    pe, pi = abi.aligned_zeros(pairs, abi.particle_dtype), abi.aligned_zeros(pairs, abi.particle_dtype)
    npe, npi = C.c_int(0), C.c_int(0)
    assert O.orc_load_thermal_pairs_tagged(ptr(rng_o), pairs, ptr(lo), ptr(hi), 0.1, 0.1, q, ptr(pe), C.byref(npe), pairs, ptr(pi), C.byref(npi), pairs,
                                           g.ref(), 5, 1) == pairs
    dom = vpb.vpb_domain_create(g.ref(), 0, 1)
    vpb.vpb_domain_set_particle_layout(dom, (pairs + 63) // 64 * 64)
    d_e, d_i = ParticleArray(vpb, dom, pairs), ParticleArray(vpb, dom, pairs)
    rng = vpb.vpb_mt_create(11)
    np2 = (C.c_int * 2)(0, 0)
    assert vpb.vpb_load_pairs_mt(dom, rng, pairs, ptr(lo), ptr(hi), 0.1, 0.1, -q, q, d_e.ptr, pairs, d_i.ptr, pairs, np2, 1, 5, 1) == pairs
    assert (np2[0], np2[1]) == (pairs, pairs)
    assert_bits_equal(d_e.download(pairs), pe, "electrons")
    assert_bits_equal(d_i.download(pairs), pi, "ions")
    for a in (d_e, d_i):
        a.free()
    vpb.vpb_mt_destroy(rng)
    vpb.vpb_domain_destroy(dom)
