"""TEST INFRASTRUCTURE: an in-process emulation of W ranks built from the CPU oracle's pieces.

Every rank's state lives in this process; a message packed for face F on rank k is handed to the rank
g_k.bc[F] and consumed there through its face (F+3)%6, in the order the reference completes its receives
(remote.c / boundary_p.c / hydro.c, cited per method).  tests/ref_mpi_worker.py runs the real reference
on W processes beside it and compares rank for rank, bit for bit."""
import ctypes as C

import numpy as np

from helpers import abi, loader
from old_vpic_b200.abi import ptr

FB = [abi.boundary(-1, 0, 0), abi.boundary(0, -1, 0), abi.boundary(0, 0, -1), abi.boundary(1, 0, 0), abi.boundary(0, 1, 0),
      abi.boundary(0, 0, 1)]
RORDER = (3, 4, 5, 0, 1, 2)       # receive ports complete in the order -x,-y,-z,+x,+y,+z of the SENDER's face


class OracleCluster:
    def __init__(self, O, grids):
        self.O, self.g, self.W = O, grids, len(grids)

    def peer(self, k, face):
        b = self.g[k].struct.bc[FB[face]]
        return b if 0 <= b < self.W else None

    # -- remote.c:33-140 (begin_/end_remote_ghost_*): all six faces in flight at once -------------------------
    def ghost(self, kind, fs, local_ghost):
        O, msgs = self.O, {}
        for k in range(self.W):
            for face in range(6):
                dst = self.peer(k, face)
                if dst is None:
                    continue
                b = np.zeros(O.orc_face_message_floats(kind, face, self.g[k].ref()), np.float32)
                O.orc_face_pack(kind, face, ptr(fs[k]), self.g[k].ref(), ptr(b))
                msgs[(dst, (face + 3) % 6)] = b
        for k in range(self.W):
            local_ghost(ptr(fs[k]), self.g[k].ref(), self.W)
        for k in range(self.W):
            for face in RORDER:
                if (k, face) in msgs:
                    O.orc_face_unpack(kind, face, ptr(fs[k]), self.g[k].ref(), ptr(msgs[(k, face)]))

    # -- remote.c:281-405 (synchronize_*): x pass, then y, then z -------------------------------------------------
    def sync(self, kind, fs):
        O, err = self.O, [0.0] * self.W
        for X in range(3):
            msgs = {}
            for k in range(self.W):
                for face in (X, X + 3):
                    dst = self.peer(k, face)
                    if dst is None:
                        continue
                    b = np.zeros(O.orc_face_message_floats(kind, face, self.g[k].ref()), np.float32)
                    O.orc_face_pack(kind, face, ptr(fs[k]), self.g[k].ref(), ptr(b))
                    msgs[(dst, (face + 3) % 6)] = b
            for k in range(self.W):
                for face in (X + 3, X):
                    if (k, face) in msgs:
                        err[k] += O.orc_face_unpack(kind, face, ptr(fs[k]), self.g[k].ref(), ptr(msgs[(k, face)]))
        return err

    def each(self, fn, fs, *a):
        for k in range(self.W):
            fn(ptr(fs[k]), *a, self.g[k].ref(), self.W)

    # -- the field-advance methods, composed as the reference composes them ---------------------------------------
    def advance_b(self, fs, frac):
        for k in range(self.W):
            self.O.orc_advance_b(ptr(fs[k]), self.g[k].ref(), frac, self.W)

    def advance_e(self, fs, m):
        O = self.O
        self.ghost(loader.GHOST_TANG_B, fs, O.orc_local_ghost_tang_b)
        for k in range(self.W):
            O.orc_advance_e_update(ptr(fs[k]), ptr(m), self.g[k].ref(), 0)
        self.each(O.orc_local_adjust_tang_e, fs)

    def compute_curl_b(self, fs, m):
        O = self.O
        self.ghost(loader.GHOST_TANG_B, fs, O.orc_local_ghost_tang_b)
        for k in range(self.W):
            O.orc_curl_b_update(ptr(fs[k]), ptr(m), self.g[k].ref())

    def compute_div_e_err(self, fs, m, rhob=False):
        O = self.O
        self.ghost(loader.GHOST_NORM_E, fs, O.orc_local_ghost_norm_e)
        for k in range(self.W):
            O.orc_div_e_err_update(ptr(fs[k]), ptr(m), self.g[k].ref(), 1 if rhob else 0)
        self.each(O.orc_local_adjust_rhob if rhob else O.orc_local_adjust_div_e, fs)

    def clean_div_e(self, fs, m):
        O = self.O
        for k in range(self.W):
            O.orc_clean_div_e_update(ptr(fs[k]), ptr(m), self.g[k].ref())
        self.each(O.orc_local_adjust_tang_e, fs)

    def compute_div_b_err(self, fs):
        for k in range(self.W):
            self.O.orc_compute_div_b_err(ptr(fs[k]), self.g[k].ref())

    def clean_div_b(self, fs):
        O = self.O
        self.ghost(loader.GHOST_DIV_B, fs, O.orc_local_ghost_div_b)
        for k in range(self.W):
            O.orc_clean_div_b_update(ptr(fs[k]), self.g[k].ref())
        self.each(O.orc_local_adjust_norm_b, fs)

    def synchronize_jf(self, fs):
        self.each(self.O.orc_local_adjust_jf, fs)
        self.sync(loader.SYNC_JF, fs)

    def synchronize_rho(self, fs):
        self.each(self.O.orc_local_adjust_rhof, fs)
        self.each(self.O.orc_local_adjust_rhob, fs)
        self.sync(loader.SYNC_RHO, fs)

    def synchronize_tang_e_norm_b(self, fs):
        self.each(self.O.orc_local_adjust_tang_e, fs)
        self.each(self.O.orc_local_adjust_norm_b, fs)
        return sum(self.sync(loader.SYNC_TEB, fs))

    def energy_f(self, fs, m):
        tot = np.zeros(6)
        for k in range(self.W):
            en = np.zeros(6)
            self.O.orc_energy_f(ptr(en), ptr(fs[k]), ptr(m), self.g[k].ref())
            tot += en
        return tot

    # -- sf_interface/hydro.c:62-140: x, y, z passes; the +X side is unpacked first -------------------------------
    def synchronize_hydro(self, hs):
        O = self.O
        for k in range(self.W):
            O.orc_local_adjust_hydro(ptr(hs[k]), self.g[k].ref(), self.W)
        for X in range(3):
            msgs = {}
            for k in range(self.W):
                for face in (X, X + 3):
                    dst = self.peer(k, face)
                    if dst is None:
                        continue
                    b = np.zeros(O.orc_hydro_face_floats(face, self.g[k].ref()), np.float32)
                    O.orc_hydro_face_pack(face, ptr(hs[k]), self.g[k].ref(), ptr(b))
                    msgs[(dst, (face + 3) % 6)] = b
            for k in range(self.W):
                for face in (X + 3, X):
                    if (k, face) in msgs:
                        O.orc_hydro_face_unpack(face, ptr(hs[k]), self.g[k].ref(), ptr(msgs[(k, face)]))

    # -- boundary_p.c:77-505, one call: species = per rank a list of dicts {id,p,np,pm,nm} ------------------------
    def boundary_p(self, species, fs, accs):
        O, msgs = self.O, {}
        for k in range(self.W):
            g = self.g[k]
            total_nm = sum(s["nm"] for s in species[k])
            outs = [abi.aligned_zeros(total_nm + 1, abi.injector_dtype) for _ in range(6)]
            used = [0] * 6
            for s in species[k]:                       # faces' buffers fill across the species list (:196-323)
                outp = (C.c_void_p * 6)(*[outs[f].ctypes.data + 48 * used[f] for f in range(6)])
                n_out = (C.c_int * 6)()
                s["np"] = O.orc_boundary_p_pack(ptr(s["p"]), s["np"], ptr(s["pm"]), s["nm"], s["id"], ptr(fs[k]), g.ref(), k,
                                                self.W, outp, n_out)
                s["nm"] = 0
                for f in range(6):
                    used[f] += n_out[f]
            for face in range(6):
                dst = self.peer(k, face)
                if dst is None or dst == k:            # SHARED_REMOTELY excludes the rank itself (:101-102)
                    assert used[face] == 0
                    continue
                msgs[(dst, (face + 3) % 6)] = outs[face][:used[face]].copy()
        for k in range(self.W):
            for face in RORDER:                        # rf2b order (:91-96), injectors walked in reverse (:470-497)
                inj = msgs.get((k, face))
                if inj is None or len(inj) == 0:
                    continue
                for n in range(len(inj) - 1, -1, -1):
                    s = [t for t in species[k] if t["id"] == inj["sp_id"][n]][0]
                    npc = C.c_int(s["np"])
                    one = inj[n:n + 1].copy()
                    s["nm"] += O.orc_boundary_p_inject(ptr(s["p"]), C.byref(npc), ptr(s["pm"]), s["nm"], ptr(one), 1, s["id"],
                                                       ptr(accs[k]), g_ref(self.g[k]))
                    s["np"] = npc.value


def g_ref(g):
    return g.ref()
