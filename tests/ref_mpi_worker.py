"""Worker for tests/test_ref_multirank.py: rank VPIC_SHIM_RANK of VPIC_SHIM_NPROC processes, each running the
REFERENCE ITSELF (oracle/_ref/libvpic_ref_scalar.so) over the shared-memory MPI shim (oracle/mpi_shim), i.e. the
reference's own remote.c / boundary_p.c / hydro.c exchange code between real processes.  Beside it every rank
replays all W ranks with the CPU oracle in-process (tests/orc_cluster.py) and requires its own rank's state to be
bit-identical to what the reference produced -- this pins the oracle's multi-rank pieces (face pack/unpack,
injector pack/inject, receive order) and the host mirror's decomposed grid_t against the real thing.

Never run on the GPU box (reads oracle/_ref only; no /root/reference at run time)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import (abi, assert_bits_equal, courant_dt, host_grid, interior_voxels, loader, random_fields,  # noqa: E402
                     random_interpolator, vacuum_coefficients)
from old_vpic_b200.abi import ptr  # noqa: E402
from orc_cluster import OracleCluster  # noqa: E402


def ref_grid(L, gn, kind, topo, dt, damp):
    p = L.new_grid()
    a = (p, 0.0, 0.0, 0.0, float(gn[0]), float(gn[1]), float(gn[2]), gn[0], gn[1], gn[2], topo[0], topo[1], topo[2])
    if kind in ("periodic", "sheet"):
        L.partition_periodic_box(*a)
    elif kind == "metal":
        L.partition_metal_box(*a)
    else:
        L.partition_absorbing_box(*a, abi.ABSORB_PARTICLES)
    if kind == "sheet":
        # the trecon-part box (turbulence.cxx:283-290): periodic in x and y, conducting walls that reflect particles on the
        # ranks that own the bottom and the top of the box -- BASELINE configs[4]'s 2x2x2 decomposition, scaled down
        rank = int(os.environ["VPIC_SHIM_RANK"])
        pz = rank // (topo[0] * topo[1])
        for sgn, edge in ((-1, 0), (1, topo[2] - 1)):
            if pz == edge:
                L.set_fbc(p, abi.boundary(0, 0, sgn), abi.PEC_FIELDS)
                L.set_pbc(p, abi.boundary(0, 0, sgn), abi.REFLECT_PARTICLES)
    s = abi.GridStruct.from_address(p)
    s.dt, s.cvac, s.eps0, s.damp = dt, 1.0, 1.0, damp
    return p, s


def mirror_grid(gn, kind, topo, rank, dt, damp):
    if kind != "sheet":
        return host_grid(gn, kind, topo=topo, rank=rank, dt=dt, damp=damp)
    g = host_grid(gn, "periodic", topo=topo, rank=rank, dt=dt, damp=damp)
    for sgn, edge in ((-1, 0), (1, topo[2] - 1)):
        if g.coords[2] == edge:
            g.set_fbc(abi.boundary(0, 0, sgn), abi.PEC_FIELDS)
            g.set_pbc(abi.boundary(0, 0, sgn), abi.REFLECT_PARTICLES)
    return g


def make_particles(seed, g, n, cap, q):
    rng = np.random.default_rng(seed)
    p = abi.aligned_zeros(cap, abi.particle_dtype)
    p["i"][:n] = np.sort(rng.choice(interior_voxels(g), n))
    for k in ("dx", "dy", "dz"):
        p[k][:n] = rng.uniform(-1, 1, n).astype(np.float32)
    for k in ("ux", "uy", "uz"):
        p[k][:n] = (0.9 * rng.standard_normal(n)).astype(np.float32)
    p["q"][:n] = q
    p["tag"] = np.arange(cap) + 1000 * seed       # the whole capacity: stale tags of injected slots are compared too
    return p


def main():
    rank, W = int(os.environ["VPIC_SHIM_RANK"]), int(os.environ["VPIC_SHIM_NPROC"])
    topo = tuple(int(x) for x in os.environ["REFW_TOPO"].split(","))
    kind = os.environ.get("REFW_KIND", "periodic")
    gn = tuple(int(x) for x in os.environ.get("REFW_GN", "8,6,4").split(","))
    assert topo[0] * topo[1] * topo[2] == W
    L = loader.ref("scalar")                       # mp_init -> MPI_Init of the shim: attaches to the shm file
    O = loader.oracle()
    M = loader.ref_methods(L, 0)
    dt = courant_dt(1.0, 1.0, 1.0)
    gp, gs = ref_grid(L, gn, kind, topo, dt, 0.01)
    gref = C.c_void_p(gp)
    grids = [mirror_grid(gn, kind, topo, k, dt, 0.01) for k in range(W)]
    g = grids[rank]

    # 0. the host mirror's grid_t for this rank is the reference's (partition.c, ops.c:join_grid)
    for name, _ in abi.GridStruct._fields_:
        if name in ("mp", "range", "neighbor", "boundary"):
            continue
        a, b = getattr(gs, name), getattr(g.struct, name)
        if name == "bc":
            a, b = list(a), list(b)
        assert a == b, ("grid_t." + name, rank, a, b)
    rr = np.ctypeslib.as_array(C.cast(gs.range, C.POINTER(C.c_int64)), shape=(W + 1,))
    assert np.array_equal(rr, g.range), ("range", rr, g.range)
    nb = np.ctypeslib.as_array(C.cast(gs.neighbor, C.POINTER(C.c_int64)), shape=(6 * g.nv,))
    assert np.array_equal(nb, g.neighbor), "neighbor[]"

    cl = OracleCluster(O, grids)
    m = vacuum_coefficients(3, np.random.default_rng(3))
    fs = [random_fields(np.random.default_rng(100 + k), grids[k], n_mat=3) for k in range(W)]
    f_r = fs[rank].copy()

    def check(what):
        assert_bits_equal(fs[rank], f_r, "%s (rank %d)" % (what, rank))

    # 1. field advance + divergence cleaning, every method that talks to a neighbour
    for frac in (0.5, 1.0):
        M.advance_b(ptr(f_r), gref, frac); cl.advance_b(fs, frac); check("advance_b")
    for _ in range(2):
        M.advance_e(ptr(f_r), ptr(m), gref); cl.advance_e(fs, m); check("advance_e")
    M.synchronize_jf(ptr(f_r), gref); cl.synchronize_jf(fs); check("synchronize_jf")
    M.synchronize_rho(ptr(f_r), gref); cl.synchronize_rho(fs); check("synchronize_rho")
    e_r = M.synchronize_tang_e_norm_b(ptr(f_r), gref)
    e_o = cl.synchronize_tang_e_norm_b(fs); check("synchronize_tang_e_norm_b")
    assert abs(e_o - e_r) <= 1e-12 * abs(e_r), ("sync err", e_o, e_r)
    M.compute_div_e_err(ptr(f_r), ptr(m), gref); cl.compute_div_e_err(fs, m); check("compute_div_e_err")
    M.clean_div_e(ptr(f_r), ptr(m), gref); cl.clean_div_e(fs, m); check("clean_div_e")
    M.compute_div_b_err(ptr(f_r), gref); cl.compute_div_b_err(fs); check("compute_div_b_err")
    M.clean_div_b(ptr(f_r), gref); cl.clean_div_b(fs); check("clean_div_b")
    M.compute_rhob(ptr(f_r), ptr(m), gref); cl.compute_div_e_err(fs, m, rhob=True); check("compute_rhob")
    M.compute_curl_b(ptr(f_r), ptr(m), gref); cl.compute_curl_b(fs, m); check("compute_curl_b")
    en_r = np.zeros(6)
    M.energy_f(ptr(en_r), ptr(f_r), ptr(m), gref)
    np.testing.assert_allclose(cl.energy_f(fs, m), en_r, rtol=1e-13)

    # 2. particle migration: two species, advance_p leaves movers on the remote faces, then num_comm_round=3
    #    boundary_p calls (vpic.cxx:17, advance.cxx:58-66).  Survivors, their ORDER, the movers still pending,
    #    rhob and the accumulator are compared after every call.
    npk, cap = 16 * 188, 6000      # whole bundles of 16: everything through pipeline 0 (advance_p.cxx:41)
    species, accs, fis = [], [], []
    for k in range(W):
        gk = grids[k]
        fi = random_interpolator(np.random.default_rng(200 + k), gk, amp=0.2)
        acc = abi.aligned_zeros(gk.nv, abi.accumulator_dtype)
        sl = []
        for sid, q in ((0, -1.0), (1, 0.5)):
            p = make_particles(10 * k + sid + 1, gk, npk, cap, q)
            pm = abi.aligned_zeros(cap, abi.mover_dtype)
            acc1 = abi.aligned_zeros(gk.nv, abi.accumulator_dtype)     # this species alone, for a bit-exact comparison
            if k == rank:      # the reference's advance_p on its decomposed grid: movers stop at the remote faces
                p2, pm2 = p.copy(), pm.copy()
                a2 = abi.aligned_zeros((1 + L.refh_n_pipeline()) * ((gk.nv + 1) // 2 * 2), abi.accumulator_dtype)
                nm2 = L.advance_p(ptr(p2), npk, q, ptr(pm2), cap, ptr(a2), ptr(fi), gref)
                L.reduce_accumulators(ptr(a2), gref)
            nm = O.orc_advance_p(ptr(p), npk, q, ptr(pm), cap, ptr(acc1), ptr(fi), gk.ref())
            if k == rank:
                assert nm2 == nm and nm > 0, (nm2, nm)
                assert_bits_equal(p2[:npk], p[:npk], "advance_p particles (rank %d)" % rank)
                assert_bits_equal(pm2[:nm], pm[:nm], "advance_p movers (rank %d)" % rank)
                assert_bits_equal(a2[:gk.nv], acc1, "advance_p accumulator (rank %d)" % rank)
            acc.view(np.float32)[:] += acc1.view(np.float32)
            sl.append({"id": sid, "p": p, "np": npk, "pm": pm, "nm": nm})
        species.append(sl); accs.append(acc); fis.append(fi)
    assert sum(s["nm"] for s in species[rank]) > 0
    # the reference's copies for this rank, as a species_t list in its own layout
    p_r = [s["p"].copy() for s in species[rank]]
    pm_r = [s["pm"].copy() for s in species[rank]]
    a_r = abi.aligned_zeros((1 + L.refh_n_pipeline()) * ((g.nv + 1) // 2 * 2), abi.accumulator_dtype)
    a_r[:g.nv] = accs[rank]
    sps = [abi.SpeciesStruct() for _ in species[rank]]
    for j, (sp, s) in enumerate(zip(sps, species[rank])):
        sp.id, sp.np, sp.max_np, sp.p = s["id"], s["np"], cap, p_r[j].ctypes.data
        sp.nm, sp.max_nm, sp.pm = s["nm"], cap, pm_r[j].ctypes.data
        sp.q_m = 1.0
        sp.name = b"sp%d" % j
    sps[0].next = C.pointer(sps[1])
    tot0 = sum(s["np"] for sl in species for s in sl)
    for rnd in range(3):
        L.boundary_p(C.byref(sps[0]), ptr(f_r), ptr(a_r), gref, None)
        cl.boundary_p(species, fs, accs)
        for j, (sp, s) in enumerate(zip(sps, species[rank])):
            assert (sp.np, sp.nm) == (s["np"], s["nm"]), ("counts", rank, rnd, j, sp.np, sp.nm, s["np"], s["nm"])
            assert_bits_equal(p_r[j][:sp.np], s["p"][:s["np"]], "particles sp%d round %d rank %d" % (j, rnd, rank))
            assert_bits_equal(pm_r[j][:sp.nm], s["pm"][:s["nm"]], "movers sp%d round %d rank %d" % (j, rnd, rank))
        check("rhob after boundary_p round %d" % rnd)
        assert_bits_equal(a_r[:g.nv], accs[rank], "accumulator after boundary_p round %d" % rnd)
        pending = sum(s["nm"] for sl in species for s in sl)
        if rnd == 0 and sum(t > 1 for t in topo) > 1:
            assert pending > 0          # some injected particles reach a second remote face: the later rounds do work
        if rnd == 2:
            assert pending == 0
    if kind == "periodic":
        assert sum(s["np"] for sl in species for s in sl) == tot0      # nothing lost in flight

    # 3. hydro moments: node planes shared with the neighbour are summed on both sides (hydro.c:62-140)
    hs = []
    for k in range(W):
        h = abi.aligned_zeros(grids[k].nv, abi.hydro_dtype)
        s = species[k][0]
        O.orc_accumulate_hydro_p(ptr(h), ptr(s["p"]), s["np"], -1.0, ptr(fis[k]), grids[k].ref())
        hs.append(h)
    h_r = hs[rank].copy()
    L.synchronize_hydro(ptr(h_r), gref)
    cl.synchronize_hydro(hs)
    assert_bits_equal(hs[rank], h_r, "synchronize_hydro (rank %d)" % rank)
    print("REF_MPI_OK rank=%d world=%d topo=%s kind=%s" % (rank, W, topo, kind), flush=True)


if __name__ == "__main__":
    main()
