#!/usr/bin/env python
"""Generates tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (pdlfs/old-vpic compiled from
source into oracle/_ref, scalar flavour) on small seeded inputs.  The reference ships no golden
vectors, so these are the pinned ones: inputs and the reference's outputs, both stored.

    python tests/golden/make_golden.py        # needs oracle/_ref (i.e. /root/reference at build time)

The fixtures travel to machines where the reference is absent; tests/test_golden.py checks the CPU
oracle (always) and the CUDA path (-m gpu) against them.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402
from helpers import RefGrid, abi, loader, random_fields, random_interpolator, random_particles, vacuum_coefficients  # noqa: E402
from old_vpic_b200.abi import ptr  # noqa: E402

CASES = [("periodic", (5, 4, 3)), ("metal", (6, 1, 4)), ("absorbing", (4, 4, 4))]
FBC = {"periodic": None, "metal": abi.PEC_FIELDS, "absorbing": abi.ABSORB_FIELDS}


def grid_record(g):
    s = g.struct
    return dict(n=np.array(g.n), dt=np.float32(s.dt), damp=np.float32(s.damp), bc=np.array(list(s.bc)), neighbor=g.neighbor.copy())


def main():
    L = loader.ref("scalar", tpp=1)
    M = loader.ref_methods(L, 0)
    for kind, n in CASES:
        rng = np.random.default_rng(abs(hash((kind, n))) % 2 ** 31 if False else sum(n) * 7 + len(kind))
        g = RefGrid(L, n, kind, damp=0.01)
        out = {"grid_" + k: v for k, v in grid_record(g).items()}
        out["kind"] = np.array(kind)
        # --- particles -------------------------------------------------------
        np_ = 16 * 24
        p = random_particles(rng, g, np_, vth=0.6, edge_frac=0.03)
        fi = random_interpolator(rng, g, amp=0.3)
        stride = (g.nv + 1) // 2 * 2
        a = abi.aligned_zeros((1 + L.refh_n_pipeline()) * stride, abi.accumulator_dtype)
        pm = abi.aligned_zeros(np_, abi.mover_dtype)
        out["adv_p_in"], out["adv_fi"] = p.copy(), fi.copy()
        nm = L.advance_p(ptr(p), np_, -1.0, ptr(pm), np_, ptr(a), ptr(fi), g.ref())
        L.reduce_accumulators(ptr(a), g.ref())
        out["adv_p_out"], out["adv_pm_out"], out["adv_a_out"], out["adv_nm"] = p.copy(), pm[:nm].copy(), a[:g.nv].copy(), np.array(nm)
        q = out["adv_p_in"].copy()
        L.center_p(ptr(q), np_, 0.7, ptr(fi), g.ref())
        out["center_out"] = q.copy()
        L.uncenter_p(ptr(q), np_, 0.7, ptr(fi), g.ref())
        out["uncenter_out"] = q.copy()
        out["energy_p"] = np.array(L.energy_p(ptr(out["adv_p_in"]), np_, -1.0, ptr(fi), g.ref()))
        # --- fields ----------------------------------------------------------
        f = random_fields(rng, g, n_mat=3)
        m = vacuum_coefficients(3, rng)
        out["f_in"], out["m"] = f.copy(), m.copy()
        fi2 = abi.aligned_zeros(g.nv, abi.interpolator_dtype)
        L.load_interpolator(ptr(fi2), ptr(f), g.ref())
        out["load_interp_out"] = fi2.copy()
        f1 = f.copy()
        L.unload_accumulator(ptr(f1), ptr(a), g.ref())
        out["unload_out"] = f1.copy()
        L.accumulate_rho_p(ptr(f1), ptr(out["adv_p_out"]), np_, g.ref())
        out["rho_p_out"] = f1.copy()
        seq = []
        for name, call in [
            ("synchronize_jf", lambda: M.synchronize_jf(ptr(f1), g.ref())),
            ("advance_b", lambda: M.advance_b(ptr(f1), g.ref(), 0.5)),
            ("advance_e", lambda: M.advance_e(ptr(f1), ptr(m), g.ref())),
            ("synchronize_rho", lambda: M.synchronize_rho(ptr(f1), g.ref())),
            ("compute_div_e_err", lambda: M.compute_div_e_err(ptr(f1), ptr(m), g.ref())),
            ("clean_div_e", lambda: M.clean_div_e(ptr(f1), ptr(m), g.ref())),
            ("compute_div_b_err", lambda: M.compute_div_b_err(ptr(f1), g.ref())),
            ("clean_div_b", lambda: M.clean_div_b(ptr(f1), g.ref())),
            ("compute_curl_b", lambda: M.compute_curl_b(ptr(f1), ptr(m), g.ref())),
            ("compute_rhob", lambda: M.compute_rhob(ptr(f1), ptr(m), g.ref())),
        ]:
            call()
            out["seq_" + name] = f1.copy()
            seq.append(name)
        out["seq_order"] = np.array(seq)
        out["sync_teb_err"] = np.array(M.synchronize_tang_e_norm_b(ptr(f1), g.ref()))
        out["seq_synchronize_tang_e_norm_b"] = f1.copy()
        en = np.zeros(6)
        M.energy_f(ptr(en), ptr(f1), ptr(m), g.ref())
        out["energy_f"] = en
        out["rms_div_e"] = np.array(M.compute_rms_div_e_err(ptr(f1), g.ref()))
        out["rms_div_b"] = np.array(M.compute_rms_div_b_err(ptr(f1), g.ref()))
        path = os.path.join(HERE, "ref_%s.npz" % kind)
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


def make_hydro():
    """ref_hydro_<kind>.npz: hydro moments of the advance_p fixture's particles (hydro_p.c, sf_interface/hydro.c),
    written separately so that the older fixtures stay byte-identical."""
    L = loader.ref("scalar", tpp=1)
    for kind, n in CASES:
        z = np.load(os.path.join(HERE, "ref_%s.npz" % kind))
        g = RefGrid(L, n, kind, damp=0.01)
        p, fi = abi.aligned_empty(len(z["adv_p_in"]), abi.particle_dtype), abi.aligned_empty(g.nv, abi.interpolator_dtype)
        p[:], fi[:] = z["adv_p_in"], z["adv_fi"]
        out = {}
        for tag, q_m in (("e", -1.0), ("i", 0.25)):
            h = abi.aligned_zeros(g.nv, abi.hydro_dtype)
            L.accumulate_hydro_p(ptr(h), ptr(p), len(p), q_m, ptr(fi), g.ref())
            out["hydro_%s_accumulated" % tag] = h.copy()
            L.synchronize_hydro(ptr(h), g.ref())
            out["hydro_%s_synchronized" % tag] = h.copy()
        path = os.path.join(HERE, "ref_hydro_%s.npz" % kind)
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    if "--hydro-only" not in sys.argv:
        main()
    make_hydro()
