"""Multi-GPU worker (torchrun, one rank per GPU, NCCL): PER-CALL parity of the decomposed path.  Every rank replays
all W ranks with the CPU oracle in-process (tests/orc_cluster.py, pinned rank for rank against the reference itself by
tests/test_ref_multirank.py) and requires what its GPU produced -- through the reference-named entry points, halos and
migration over NCCL -- to match the oracle's state of its own rank:
  * every field method that talks to a neighbour: bit-exact;
  * three boundary_p rounds with two species: counts exact, the SET of particles each rank holds bit-exact in its hot
    32 bytes (the order of survivors is the device's own; tags are not carried by injectors,
    boundary_p.c:488-491), pending movers exact as a set, rhob and the accumulator within the float-sum tolerance;
  * synchronize_hydro: bit-exact.
    torchrun --nproc-per-node 2 tests/dist_gpu_percall_worker.py     (also 4: 2x2x1)
Green on 2 B200s (profiles/r2b_summary_2gpu.txt, r2k_summary_2gpu_dist.txt).
VPB_BOUNDARY_FUSED=2 sends the reference-named boundary_p() through the driver's fused migration rounds
(vpb_boundary_p_round) so that the same oracle comparison covers them."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import (abi, assert_bits_equal, courant_dt, host_grid, interior_voxels, loader, max_rel, random_fields,  # noqa: E402
                     random_interpolator, vacuum_coefficients)
from old_vpic_b200 import lib  # noqa: E402
from old_vpic_b200.abi import ptr  # noqa: E402
from orc_cluster import OracleCluster  # noqa: E402

TOL = 2e-5


def hot_rows(p, n):
    """the 32 hot bytes of the first n particles as sortable rows"""
    a = np.ascontiguousarray(p[:n]).view(np.uint8).reshape(n, 48)[:, :32].copy().view(np.uint32).reshape(n, 8)
    return a[np.lexsort(a.T[::-1])]


def make_particles(seed, g, n, cap, q):
    rng = np.random.default_rng(seed)
    p = abi.aligned_zeros(cap, abi.particle_dtype)
    p["i"][:n] = np.sort(rng.choice(interior_voxels(g), n))
    for k in ("dx", "dy", "dz"):
        p[k][:n] = rng.uniform(-1, 1, n).astype(np.float32)
    for k in ("ux", "uy", "uz"):
        p[k][:n] = (0.9 * rng.standard_normal(n)).astype(np.float32)
    p["q"][:n] = q
    p["tag"] = np.arange(cap) + 1000 * seed
    return p


def main():
    rank, W, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    L = lib.load()
    L.vpb_init(local)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = (C.c_uint8 * 128)()
        L.vpb_comm_unique_id(buf)
        uid = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    L.vpb_comm_init(rank, W, (C.c_uint8 * 128)(*uid.cpu().tolist()))
    L.vpb_set_world(W)
    O = loader.oracle()
    M = lib.field_methods(L, 0)
    for kind, topo, gn in (("periodic", {2: (2, 1, 1), 4: (2, 2, 1)}[W], (8, 6, 4)), ("absorbing", {2: (1, 1, 2), 4: (2, 1, 2)}[W], (6, 4, 6))):
        dt = courant_dt(1.0, 1.0, 1.0)
        grids = [host_grid(gn, kind, topo=topo, rank=k, dt=dt, damp=0.01) for k in range(W)]
        g = grids[rank]
        cl = OracleCluster(O, grids)
        m = vacuum_coefficients(3, np.random.default_rng(3))
        L.vpb_register_material_coefficients(ptr(m), 3)
        fs = [random_fields(np.random.default_rng(100 + k), grids[k], n_mat=3) for k in range(W)]
        f_g = fs[rank].copy()

        def check(what):
            assert_bits_equal(f_g, fs[rank], "%s (%s, rank %d)" % (what, kind, rank))

        for frac in (0.5, 1.0):
            M.advance_b(ptr(f_g), g.ref(), frac); cl.advance_b(fs, frac); check("advance_b")
        for _ in range(2):
            M.advance_e(ptr(f_g), ptr(m), g.ref()); cl.advance_e(fs, m); check("advance_e")
        M.synchronize_jf(ptr(f_g), g.ref()); cl.synchronize_jf(fs); check("synchronize_jf")
        M.synchronize_rho(ptr(f_g), g.ref()); cl.synchronize_rho(fs); check("synchronize_rho")
        e_g = M.synchronize_tang_e_norm_b(ptr(f_g), g.ref())
        e_o = cl.synchronize_tang_e_norm_b(fs); check("synchronize_tang_e_norm_b")
        assert abs(e_g - e_o) <= 1e-12 * abs(e_o), (e_g, e_o)
        M.compute_div_e_err(ptr(f_g), ptr(m), g.ref()); cl.compute_div_e_err(fs, m); check("compute_div_e_err")
        M.clean_div_e(ptr(f_g), ptr(m), g.ref()); cl.clean_div_e(fs, m); check("clean_div_e")
        M.compute_div_b_err(ptr(f_g), g.ref()); cl.compute_div_b_err(fs); check("compute_div_b_err")
        M.clean_div_b(ptr(f_g), g.ref()); cl.clean_div_b(fs); check("clean_div_b")
        M.compute_rhob(ptr(f_g), ptr(m), g.ref()); cl.compute_div_e_err(fs, m, rhob=True); check("compute_rhob")
        M.compute_curl_b(ptr(f_g), ptr(m), g.ref()); cl.compute_curl_b(fs, m); check("compute_curl_b")
        en_g = np.zeros(6)
        M.energy_f(ptr(en_g), ptr(f_g), ptr(m), g.ref())
        np.testing.assert_allclose(en_g, cl.energy_f(fs, m), rtol=1e-12)

        # particle migration, two waves of fresh particles on the same grid: with VPB_BOUNDARY_FUSED=2 the first wave's three
        # rounds run the reference's exact protocol (and tell both sides of every face what passes through it), the second
        # wave's run the fused fixed-capacity messages (VPB_BOUNDARY_CAP_MAX=16 forces their second message as well)
        for wave in range(2):
            npk, cap = 16 * 188, 6000
            species, accs, fis = [], [], []
            for k in range(W):
                gk = grids[k]
                fi = random_interpolator(np.random.default_rng(200 + k + 50 * wave), gk, amp=0.2)
                acc = abi.aligned_zeros(gk.nv, abi.accumulator_dtype)
                sl = []
                for sid, q in ((0, -1.0), (1, 0.5)):
                    p = make_particles(1000 * wave + 10 * k + sid + 1, gk, npk, cap, q)
                    pm = abi.aligned_zeros(cap, abi.mover_dtype)
                    nm = O.orc_advance_p(ptr(p), npk, q, ptr(pm), cap, ptr(acc), ptr(fi), gk.ref())
                    sl.append({"id": sid, "p": p, "np": npk, "pm": pm, "nm": nm})
                species.append(sl); accs.append(acc); fis.append(fi)
            p_g = [s["p"].copy() for s in species[rank]]
            pm_g = [s["pm"].copy() for s in species[rank]]
            a_g = accs[rank].copy()
            sps = [abi.SpeciesStruct() for _ in species[rank]]
            for j, (sp, s) in enumerate(zip(sps, species[rank])):
                sp.id, sp.np, sp.max_np, sp.p = s["id"], s["np"], cap, p_g[j].ctypes.data
                sp.nm, sp.max_nm, sp.pm = s["nm"], cap, pm_g[j].ctypes.data
                sp.q_m = 1.0
            sps[0].next = C.pointer(sps[1])
            rhob_scale = None
            for rnd in range(3):
                L.boundary_p(C.byref(sps[0]), ptr(f_g), ptr(a_g), g.ref(), None)
                cl.boundary_p(species, fs, accs)
                for j, (sp, s) in enumerate(zip(sps, species[rank])):
                    assert (sp.np, sp.nm) == (s["np"], s["nm"]), ("counts", kind, rank, rnd, j, sp.np, sp.nm, s["np"], s["nm"])
                    assert np.array_equal(hot_rows(p_g[j], sp.np), hot_rows(s["p"], s["np"])), ("particle set", kind, rank, rnd, j)
                    # a pending mover names its particle by index, and the order of the array is the device's: compare what it
                    # points at together with its displacement
                    def movers(pm, p, n):
                        rows = np.concatenate([np.ascontiguousarray(pm[:n]).view(np.uint32).reshape(n, 4)[:, :3],
                                               np.ascontiguousarray(p[pm["i"][:n]]).view(np.uint8).reshape(n, 48)[:, :32].copy().view(np.uint32).reshape(n, 8)], axis=1) if n else np.zeros((0, 11), np.uint32)
                        return rows[np.lexsort(rows.T[::-1])] if n else rows
                    assert np.array_equal(movers(pm_g[j], p_g[j], sp.nm), movers(s["pm"], s["p"], s["nm"])), ("movers", kind, rank, rnd, j)
                others = [n for n in abi.FIELD_FLOATS if n != "rhob"]
                for n in others:
                    assert np.array_equal(f_g[n].view(np.uint32), fs[rank][n].view(np.uint32)), n
                rhob_scale = max(float(np.abs(fs[rank]["rhob"]).max()), 1e-30)
                assert float(np.abs(f_g["rhob"] - fs[rank]["rhob"]).max()) <= TOL * rhob_scale
                f_g["rhob"] = fs[rank]["rhob"]          # keep the later bit-exact field checks independent of the float-sum order
                assert max_rel(a_g.view(np.float32).reshape(-1, 12), accs[rank].view(np.float32).reshape(-1, 12)) < TOL
        # hydro
        hs = []
        for k in range(W):
            h = abi.aligned_zeros(grids[k].nv, abi.hydro_dtype)
            s = species[k][0]
            O.orc_accumulate_hydro_p(ptr(h), ptr(s["p"]), s["np"], -1.0, ptr(fis[k]), grids[k].ref())
            hs.append(h)
        h_g = hs[rank].copy()
        L.synchronize_hydro(ptr(h_g), g.ref())
        cl.synchronize_hydro(hs)
        assert_bits_equal(h_g, hs[rank], "synchronize_hydro (%s, rank %d)" % (kind, rank))
        dist.barrier()
    if rank == 0:
        print("PERCALL_OK world=%d" % W, flush=True)
    L.vpb_comm_finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
