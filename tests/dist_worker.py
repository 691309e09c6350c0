"""Worker for tests/test_dist_gloo.py: world_size ranks on CPU (gloo), one x-slab each of a periodic box.
Runs the CPU ORACLE per rank and moves the face messages / particle injectors between ranks with
torch.distributed, using the same conventions as the device path (vpb_faces.cu / vpb_boundary.cu /
vpb_comm.cuh): a message packed for face F is consumed through the neighbour's face (F+3)%6; sends are
posted by face 0..5, receives by face 3,4,5,0,1,2.  Every rank also runs the single-domain oracle on the
whole box and compares its slab with it."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import abi, host_grid, loader, random_fields  # noqa: E402
from old_vpic_b200.abi import ptr  # noqa: E402

RORDER = (3, 4, 5, 0, 1, 2)
FB = [abi.boundary(-1, 0, 0), abi.boundary(0, -1, 0), abi.boundary(0, 0, -1), abi.boundary(1, 0, 0), abi.boundary(0, 1, 0),
      abi.boundary(0, 0, 1)]


def exchange(bufs_out, peers, rank, sizes, dtype=np.float32):
    """bufs_out[face] -> what arrives through each face (dict face -> array)."""
    reqs, got = [], {}
    for face in range(6):
        if face in bufs_out and peers[face] != rank:
            reqs.append(dist.isend(torch.from_numpy(bufs_out[face]), peers[face], tag=face))
    for face in RORDER:
        if face in bufs_out:
            if peers[face] == rank:
                got[face] = bufs_out[(face + 3) % 6].copy()
            else:
                t = torch.empty(sizes[face], dtype=torch.from_numpy(np.zeros(1, dtype)).dtype)
                reqs.append(dist.irecv(t, peers[face], tag=(face + 3) % 6))
                got[face] = t
    for r in reqs:
        r.wait()
    return {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in got.items()}


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    O = loader.oracle()
    gn = (4 * world, 4, 3)
    gg = host_grid(gn, "periodic")                                  # the whole box on one rank
    g = host_grid(gn, "periodic", topo=(world, 1, 1), rank=rank, dt=gg.struct.dt)
    peers = [g.struct.bc[b] for b in FB]
    rng = np.random.default_rng(5)
    F = random_fields(rng, gg)
    nx = g.n[0]
    x0 = rank * nx

    def slab(fglob):
        """this rank's voxels (ghosts included) cut out of the global array; ghost planes are poisoned"""
        out = abi.aligned_zeros(g.nv, abi.field_dtype)
        o3, f3 = out.reshape(g.shape), fglob.reshape(gg.shape)
        o3[:, :, :] = f3[:, :, x0:x0 + nx + 2]
        for k in abi.FIELD_FLOATS:
            o3[k][:, :, 0] = 777.0
            o3[k][:, :, nx + 1] = -777.0
        return out

    def face_exchange(kind, f, faces):
        out = {}
        for face in faces:
            b = np.zeros(O.orc_face_message_floats(kind, face, g.ref()), np.float32)
            O.orc_face_pack(kind, face, ptr(f), g.ref(), ptr(b))
            out[face] = b
        sizes = {face: len(out[face]) for face in out}
        got = exchange(out, peers, rank, sizes)
        err = 0.0
        for face in RORDER:
            if face in got:
                err += O.orc_face_unpack(kind, face, ptr(f), g.ref(), ptr(np.ascontiguousarray(got[face])))
        return err

    def compare(f_loc, f_glob, what, xr, comps):
        a3, b3 = f_loc.reshape(g.shape), f_glob.reshape(gg.shape)
        for k in comps:
            x = a3[k][1:-1, 1:-1, xr[0]:xr[1]]
            y = b3[k][1:-1, 1:-1, x0 + xr[0]:x0 + xr[1]]
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), (what, k, rank)

    m = abi.aligned_zeros(1, abi.material_coefficient_dtype)
    for k in ("decayx", "drivex", "decayy", "drivey", "decayz", "drivez", "rmux", "rmuy", "rmuz", "nonconductive", "epsx",
              "epsy", "epsz"):
        m[k] = 1.0
    # 1. advance_e with tang-B ghosts from the neighbours
    f, fg = slab(F), F.copy()
    O.orc_advance_e(ptr(fg), ptr(m), gg.ref(), 0)
    face_exchange(loader.GHOST_TANG_B, f, range(6))
    O.orc_local_ghost_tang_b(ptr(f), g.ref(), world)
    O.orc_advance_e_update(ptr(f), ptr(m), g.ref(), 0)
    O.orc_local_adjust_tang_e(ptr(f), g.ref(), world)
    compare(f, fg, "advance_e", (1, nx + 1), ("ex", "ey", "ez", "tcax", "tcay", "tcaz"))
    # 3. particle migration: three waves of fresh particles.  Wave 0 through the reference's protocol (counts, then
    #    payloads of exactly that size); it also tells both sides of every face what passes through it per round, from
    #    which the capacities of the FUSED rounds follow (vpb_boundary.cu: one fixed-capacity message per face whose
    #    record 0 is a header {count, capacity, round, magic}; a face over its capacity sends the rest in a second,
    #    exactly sized message).  Wave 1 runs the fused rounds, wave 2 the same with capacities clamped to 4 injectors
    #    so that the second message is exercised.  Same checks for every wave.
    from old_vpic_b200.grid import interior_voxels
    MAGIC = 0x76706221
    fi_g = abi.aligned_zeros(gg.nv, abi.interpolator_dtype)
    fi_l = abi.aligned_zeros(g.nv, abi.interpolator_dtype)
    gsx = gn[0] + 2
    remote = [f_ for f_ in range(6) if peers[f_] != rank]

    def capacity_for(ns, nr, cap_max):
        m = max(ns, nr)
        cap = m + m // 2 + 64
        return min(cap, cap_max) if cap_max else cap

    def inject(got, counts, Pl, nl, pml, nm, al_):
        for face in RORDER:
            if face in got and counts[face]:
                inj = np.ascontiguousarray(got[face]).view(abi.injector_dtype)[:counts[face]]
                npc = C.c_int(nl)
                nm += O.orc_boundary_p_inject(ptr(Pl), C.byref(npc), ptr(pml), nm, ptr(np.ascontiguousarray(inj)), len(inj), 0, ptr(al_), g.ref())
                nl = npc.value
        return nl, nm

    def exact_round(Pl, nl, pml, nm, al_, fl):
        outs = [abi.aligned_zeros(max(nm, 1), abi.injector_dtype) for _ in range(6)]
        outp = (C.c_void_p * 6)(*[o.ctypes.data for o in outs])
        n_out = (C.c_int * 6)()
        nl = O.orc_boundary_p_pack(ptr(Pl), nl, ptr(pml), nm, 0, ptr(fl), g.ref(), rank, world, outp, n_out)
        counts = {f_: np.array([n_out[f_]], np.int32) for f_ in remote}
        got_n = exchange(counts, peers, rank, {f_: 1 for f_ in counts}, np.int32)
        pay = {f_: outs[f_][:n_out[f_]].view(np.uint8).reshape(-1).copy() if n_out[f_] else np.zeros(0, np.uint8) for f_ in counts}
        got = exchange(pay, peers, rank, {f_: int(got_n[f_][0]) * 48 for f_ in counts}, np.uint8)
        nr = {f_: int(got_n[f_][0]) for f_ in counts}
        nl, nm = inject(got, nr, Pl, nl, pml, 0, al_)
        return nl, nm, {f_: int(n_out[f_]) for f_ in remote}, nr

    def fused_round(Pl, nl, pml, nm, al_, fl, slot, cap, cap_max):
        outs = [abi.aligned_zeros(max(nm, 1), abi.injector_dtype) for _ in range(6)]
        outp = (C.c_void_p * 6)(*[o.ctypes.data for o in outs])
        n_out = (C.c_int * 6)()
        nl = O.orc_boundary_p_pack(ptr(Pl), nl, ptr(pml), nm, 0, ptr(fl), g.ref(), rank, world, outp, n_out)
        msg = {}
        for f_ in remote:
            m_ = np.full((cap[f_] + 1) * 48, 0xAB, np.uint8)           # the unused part of a message is never read
            m_[:16] = np.array([n_out[f_], cap[f_], slot, MAGIC], np.int32).view(np.uint8)
            k = min(int(n_out[f_]), cap[f_])
            m_[48:48 + 48 * k] = outs[f_][:k].view(np.uint8).reshape(-1)
            msg[f_] = m_
        got = exchange(msg, peers, rank, {f_: (cap[f_] + 1) * 48 for f_ in msg}, np.uint8)
        ns, nr, first = {}, {}, {}
        for f_ in remote:
            h = np.ascontiguousarray(got[f_][:16]).view(np.int32)
            assert (int(h[1]), int(h[2]), int(h[3])) == (cap[f_], slot, MAGIC), ("header", rank, f_, h, cap[f_], slot)
            ns[f_], nr[f_] = int(n_out[f_]), int(h[0])
            first[f_] = np.ascontiguousarray(got[f_][48:48 + 48 * min(nr[f_], cap[f_])])
        nl, nm = inject(first, {f_: min(nr[f_], cap[f_]) for f_ in remote}, Pl, nl, pml, 0, al_)
        seconds = 0
        if any(ns[f_] > cap[f_] or nr[f_] > cap[f_] for f_ in remote):
            more = {f_: (outs[f_][cap[f_]:ns[f_]].view(np.uint8).reshape(-1).copy() if ns[f_] > cap[f_] else np.zeros(0, np.uint8)) for f_ in remote}
            rest = exchange(more, peers, rank, {f_: max(nr[f_] - cap[f_], 0) * 48 for f_ in remote}, np.uint8)
            nl, nm = inject(rest, {f_: max(nr[f_] - cap[f_], 0) for f_ in remote}, Pl, nl, pml, nm, al_)
            seconds = 1
        for f_ in remote:
            m_ = max(ns[f_], nr[f_])
            if m_ + m_ // 4 > cap[f_]:
                cap[f_] = capacity_for(ns[f_], nr[f_], cap_max)
        return nl, nm, seconds

    caps = [None, None, None]
    for wave, cap_max in ((0, 0), (1, 0), (2, 4)):
        npg = 2000
        rngp = np.random.default_rng(9 + wave)
        P = abi.aligned_zeros(npg, abi.particle_dtype)
        P["i"] = np.sort(rngp.choice(interior_voxels(gg), npg))
        for k in ("dx", "dy", "dz"):
            P[k] = rngp.uniform(-1, 1, npg).astype(np.float32)
        for k in ("ux", "uy", "uz"):
            P[k] = (0.8 * rngp.standard_normal(npg)).astype(np.float32)
        P["q"] = 1.0
        P["tag"] = np.arange(npg)
        # global run
        Pg = P.copy()
        ag = abi.aligned_zeros(gg.nv, abi.accumulator_dtype)
        pmg = abi.aligned_zeros(npg, abi.mover_dtype)
        assert O.orc_advance_p(ptr(Pg), npg, 1.0, ptr(pmg), npg, ptr(ag), ptr(fi_g), gg.ref()) == 0
        # local run: my particles are those whose voxel x lies in my slab
        gx = P["i"] % gsx
        mine = (gx > x0) & (gx <= x0 + nx)
        cap_p = npg
        Pl = abi.aligned_zeros(cap_p, abi.particle_dtype)
        nl = int(mine.sum())
        Pl[:nl] = P[mine]
        gy = (P["i"][mine] // gsx) % (gn[1] + 2)
        gz = P["i"][mine] // (gsx * (gn[1] + 2))
        Pl["i"][:nl] = (gx[mine] - x0) + (nx + 2) * (gy + (gn[1] + 2) * gz)
        al_ = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
        pml = abi.aligned_zeros(cap_p, abi.mover_dtype)
        nm = O.orc_advance_p(ptr(Pl), nl, 1.0, ptr(pml), cap_p, ptr(al_), ptr(fi_l), g.ref())
        fl = abi.aligned_zeros(g.nv, abi.field_dtype)
        seconds = 0
        for rnd in range(3):                       # num_comm_round (vpic.cxx:17)
            if wave == 0:
                nl, nm, ns, nr = exact_round(Pl, nl, pml, nm, al_, fl)
                caps[rnd] = {f_: capacity_for(ns[f_], nr[f_], 0) for f_ in remote}
            else:
                if cap_max:
                    caps[rnd] = {f_: min(c_, cap_max) for f_, c_ in caps[rnd].items()}
                nl, nm, sec = fused_round(Pl, nl, pml, nm, al_, fl, rnd, caps[rnd], cap_max)
                seconds += sec
        assert nm == 0
        if wave == 2:
            flag = torch.tensor([seconds])
            dist.all_reduce(flag)
            assert int(flag) > 0, "the clamped capacities never forced a second message"
        # global particle count is conserved and every particle I hold equals the global run's, bit for bit
        tot = torch.tensor([nl])
        dist.all_reduce(tot)
        assert int(tot) == npg, (int(tot), npg)
        # tags are not carried by injectors (boundary_p.c:488-491), so match on the hot state instead
        def key(p, xoff, sx):
            lx = p["i"] % sx
            rest = p["i"] // sx
            gi = (lx + xoff) + gsx * rest
            return np.stack([gi.astype(np.int64)] + [p[c].view(np.uint32).astype(np.int64) for c in ("dx", "dy", "dz", "ux", "uy", "uz")], 1)
        kg = {tuple(r) for r in key(Pg, 0, gsx)}
        for r in key(Pl[:nl], x0, nx + 2):
            assert tuple(r) in kg, ("particle not in the single-domain result", rank, wave, r)
        # accumulators of my interior voxels agree with the global run (same per-particle contributions)
        a3, b3 = al_.view(np.float32).reshape(g.shape + (12,)), ag.view(np.float32).reshape(gg.shape + (12,))
        np.testing.assert_allclose(a3[1:-1, 1:-1, 1:nx + 1], b3[1:-1, 1:-1, x0 + 1:x0 + nx + 1], rtol=0, atol=2e-5 * np.abs(b3).max())
    # 3. currents: unload_accumulator + synchronize_jf (x, y, z passes).  Each rank only holds its own
    #    particles' share of the plane it shares with its neighbour; after the exchange both copies must be
    #    identical and equal to the single-domain result up to float summation order.
    fg, f = abi.aligned_zeros(gg.nv, abi.field_dtype), abi.aligned_zeros(g.nv, abi.field_dtype)
    O.orc_unload_accumulator(ptr(fg), ptr(ag), gg.ref())
    O.orc_synchronize_jf(ptr(fg), gg.ref())
    O.orc_unload_accumulator(ptr(f), ptr(al_), g.ref())
    O.orc_local_adjust_jf(ptr(f), g.ref(), world)
    for X in range(3):
        face_exchange(loader.SYNC_JF, f, (X, X + 3))
    a3, b3 = f.reshape(g.shape), fg.reshape(gg.shape)
    for k, hi in (("jfx", nx + 1), ("jfy", nx + 2), ("jfz", nx + 2)):   # jfx lives on x edges 1..nx; jfy,jfz also on the shared plane
        np.testing.assert_allclose(a3[k][1:-1, 1:-1, 1:hi], b3[k][1:-1, 1:-1, x0 + 1:x0 + hi], rtol=0,
                                   atol=2e-5 * max(np.abs(b3[k]).max(), 1e-30))
    mine_hi = torch.from_numpy(np.ascontiguousarray(np.stack([a3[k][:, :, nx + 1] for k in ("jfy", "jfz")])))
    mine_lo = torch.from_numpy(np.ascontiguousarray(np.stack([a3[k][:, :, 1] for k in ("jfy", "jfz")])))
    other_lo = torch.empty_like(mine_lo)
    reqs = [dist.isend(mine_lo, (rank - 1) % world, tag=77), dist.irecv(other_lo, (rank + 1) % world, tag=77)]
    for r_ in reqs:
        r_.wait()
    assert torch.equal(mine_hi, other_lo), "shared plane differs between the two ranks after synchronize_jf"
    # 4. hydro moments (hydro_p.c, sf_interface/hydro.c): every rank deposits its own particles, doubles the nodes on
    #    local faces, then exchanges the face planes axis by axis; the nodes it owns or shares must equal the
    #    single-domain result up to float summation order, and the two copies of a shared plane must be identical.
    rngf = np.random.default_rng(13)
    fi_glob = abi.aligned_zeros(gg.nv, abi.interpolator_dtype)
    for name in abi.interpolator_dtype.names:
        if name != "_pad":
            fi_glob[name] = (0.2 * rngf.standard_normal(gg.nv)).astype(np.float32)
    fi_loc = abi.aligned_zeros(g.nv, abi.interpolator_dtype)
    fi_loc.reshape(g.shape)[:, :, :] = fi_glob.reshape(gg.shape)[:, :, x0:x0 + nx + 2]
    hg, hl = abi.aligned_zeros(gg.nv, abi.hydro_dtype), abi.aligned_zeros(g.nv, abi.hydro_dtype)
    O.orc_accumulate_hydro_p(ptr(hg), ptr(Pg), npg, 1.0, ptr(fi_glob), gg.ref())
    O.orc_synchronize_hydro(ptr(hg), gg.ref(), 0, 1)
    O.orc_accumulate_hydro_p(ptr(hl), ptr(Pl), nl, 1.0, ptr(fi_loc), g.ref())
    O.orc_local_adjust_hydro(ptr(hl), g.ref(), world)
    for X in range(3):
        out = {}
        for face in (X, X + 3):
            b = np.zeros(O.orc_hydro_face_floats(face, g.ref()), np.float32)
            O.orc_hydro_face_pack(face, ptr(hl), g.ref(), ptr(b))
            out[face] = b
        got = exchange(out, peers, rank, {face: len(out[face]) for face in out})
        for face in (X + 3, X):                 # the reference unpacks the +X side first (hydro.c:111-112)
            O.orc_hydro_face_unpack(face, ptr(hl), g.ref(), ptr(np.ascontiguousarray(got[face])))
    a4 = hl.view(np.float32).reshape(g.shape + (16,))[1:, 1:, 1:nx + 2, :14]
    b4 = hg.view(np.float32).reshape(gg.shape + (16,))[1:, 1:, x0 + 1:x0 + nx + 2, :14]
    hs = np.abs(hg.view(np.float32).reshape(-1, 16)[:, :14]).max(axis=0)
    assert np.all(np.abs(a4 - b4) <= 2e-5 * hs), ("hydro", rank, float(np.max(np.abs(a4 - b4) / hs)))
    h_hi = torch.from_numpy(np.ascontiguousarray(a4[:, :, nx, :]))
    h_lo = torch.from_numpy(np.ascontiguousarray(a4[:, :, 0, :]))
    o_lo = torch.empty_like(h_lo)
    reqs = [dist.isend(h_lo, (rank - 1) % world, tag=78), dist.irecv(o_lo, (rank + 1) % world, tag=78)]
    for r_ in reqs:
        r_.wait()
    assert torch.equal(h_hi, o_lo), "shared node plane differs between the two ranks after synchronize_hydro"
    dist.barrier()
    if rank == 0:
        print("DIST_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
