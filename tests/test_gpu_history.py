"""Per-run parity: the device time-step driver (csrc/vpb_step.cu, the call order of advance.cxx) against
the same loop run on the CPU with the oracle kernels -- and, where oracle/_ref is present, with the
reference's own compiled kernels -- from identical particles.  Energy histories (6 field + 2 kinetic, the
columns of dump_energies, dump.cxx:37-78) must agree within 1e-4 relative over the first 20 steps
(north_star: "single-precision relative tolerance"; the reference's own -tpp/V4 spread is <=2.5e-5 at 20
steps).  Particle counts and per-step mover counts are exact."""
import ctypes as C

import numpy as np
import pytest

from helpers import abi, host_grid, loader, random_particles
from old_vpic_b200.abi import ptr
from old_vpic_b200.sim import NativeSimulation

pytestmark = pytest.mark.gpu

STEPS, SORT = 20, 5


def cpu_history(K, g, species, steps, clean_e=0, clean_b=0, f_init=None, sync_shared=0, errors=None, state=None, sort=None):
    """K: dict of kernels with the oracle's calling convention.  errors: a list that receives, per cleaning / synchronising
    step, what advance.cxx:160,168,182,190,205 would print.  state: a dict that receives the final field and interpolator
    arrays ("f", "fi"); the species' arrays are updated in place.  sort: sort interval (default SORT)."""
    sort = SORT if sort is None else sort
    f = abi.aligned_zeros(g.nv, abi.field_dtype)
    if f_init is not None:
        f[:] = f_init
    fi = abi.aligned_zeros(g.nv, abi.interpolator_dtype)
    a = abi.aligned_zeros(K["n_acc"](g), abi.accumulator_dtype)
    m = abi.aligned_zeros(1, abi.material_coefficient_dtype)
    for k in ("decayx", "drivex", "decayy", "drivey", "decayz", "drivez", "rmux", "rmuy", "rmuz", "nonconductive", "epsx",
              "epsy", "epsz"):
        m[k] = 1.0
    K["load_interpolator"](fi, f, g)
    hist = []
    for step in range(steps):
        K["clear_accumulators"](a, g)
        for sp in species:
            if sort > 0 and step % sort == 0:
                sp["p"] = K["sort"](sp["p"], g)
        for sp in species:
            pm = abi.aligned_zeros(len(sp["p"]), abi.mover_dtype)
            nm = K["advance_p"](sp["p"], sp["q_m"], pm, a, fi, g)
            assert nm == 0
        K["reduce_accumulators"](a, g)
        K["clear_jf"](f, g)
        K["unload_accumulator"](f, a, g)
        K["synchronize_jf"](f, g)
        K["advance_b"](f, g, 0.5)
        K["advance_e"](f, m, g)
        K["advance_b"](f, g, 0.5)
        rec = {"step": step}
        if clean_e > 0 and step % clean_e == 0:                      # advance.cxx:151-173
            K["clear_rhof"](f, g)
            for sp in species:
                K["accumulate_rho_p"](f, sp["p"], g)
            K["synchronize_rho"](f, g)
            K["compute_div_e_err"](f, m, g)
            err = K["rms_div_e_err"](f, g)
            rec["div_e"] = [err]
            if err > 0:
                K["clean_div_e"](f, m, g)
                K["compute_div_e_err"](f, m, g)
                err = K["rms_div_e_err"](f, g)
                rec["div_e"].append(err)
                if err > 0:
                    K["clean_div_e"](f, m, g)
        if clean_b > 0 and step % clean_b == 0:                      # :177-195
            K["compute_div_b_err"](f, g)
            err = K["rms_div_b_err"](f, g)
            rec["div_b"] = [err]
            if err > 0:
                K["clean_div_b"](f, g)
                K["compute_div_b_err"](f, g)
                err = K["rms_div_b_err"](f, g)
                rec["div_b"].append(err)
                if err > 0:
                    K["clean_div_b"](f, g)
        if sync_shared > 0 and step % sync_shared == 0:              # :199-208
            rec["desync"] = K["synchronize_tang_e_norm_b"](f, g)
        if errors is not None and len(rec) > 1:
            errors.append(rec)
        K["load_interpolator"](fi, f, g)
        en = np.zeros(6)
        K["energy_f"](en, f, m, g)
        hist.append(list(en) + [K["energy_p"](sp["p"], sp["q_m"], fi, g) for sp in species])
    if state is not None:
        state["f"], state["fi"] = f, fi
    return np.array(hist)


def oracle_kernels(O):
    def rms(fn):
        def f_(f, g):
            loc = np.zeros(2)
            fn(ptr(loc), ptr(f), g.ref())            # one rank: {err^2 sum * dV, volume} (compute_rms_div_e_err.c:150-160)
            return float(g.struct.eps0 * np.sqrt(loc[0] / loc[1]))
        return f_

    def sort(p, g):
        out = abi.aligned_zeros(len(p), abi.particle_dtype)
        part = np.zeros(g.nv + 1, np.int32)
        O.orc_sort_p(ptr(p), ptr(out), len(p), ptr(part), g.ref())
        return out
    return {
        "n_acc": lambda g: g.nv,
        "load_interpolator": lambda fi, f, g: O.orc_load_interpolator(ptr(fi), ptr(f), g.ref()),
        "clear_accumulators": lambda a, g: O.orc_clear_accumulators(ptr(a), g.ref()),
        "reduce_accumulators": lambda a, g: None,
        "sort": sort,
        "advance_p": lambda p, q_m, pm, a, fi, g: O.orc_advance_p(ptr(p), len(p), q_m, ptr(pm), len(pm), ptr(a), ptr(fi), g.ref()),
        "clear_jf": lambda f, g: O.orc_clear_jf(ptr(f), g.ref()),
        "unload_accumulator": lambda f, a, g: O.orc_unload_accumulator(ptr(f), ptr(a), g.ref()),
        "synchronize_jf": lambda f, g: O.orc_synchronize_jf(ptr(f), g.ref()),
        "advance_b": lambda f, g, frac: O.orc_advance_b(ptr(f), g.ref(), frac, 1),
        "advance_e": lambda f, m, g: O.orc_advance_e(ptr(f), ptr(m), g.ref(), 0),
        "clear_rhof": lambda f, g: O.orc_clear_rhof(ptr(f), g.ref()),
        "accumulate_rho_p": lambda f, p, g: O.orc_accumulate_rho_p(ptr(f), ptr(p), len(p), g.ref()),
        "synchronize_rho": lambda f, g: O.orc_synchronize_rho(ptr(f), g.ref()),
        "compute_div_e_err": lambda f, m, g: O.orc_compute_div_e_err(ptr(f), ptr(m), g.ref()),
        "clean_div_e": lambda f, m, g: O.orc_clean_div_e(ptr(f), ptr(m), g.ref()),
        "compute_div_b_err": lambda f, g: O.orc_compute_div_b_err(ptr(f), g.ref()),
        "clean_div_b": lambda f, g: O.orc_clean_div_b(ptr(f), g.ref()),
        "rms_div_e_err": rms(O.orc_rms_div_e_err_local),
        "rms_div_b_err": rms(O.orc_rms_div_b_err_local),
        "synchronize_tang_e_norm_b": lambda f, g: float(O.orc_synchronize_tang_e_norm_b(ptr(f), g.ref())),
        "energy_f": lambda en, f, m, g: O.orc_energy_f(ptr(en), ptr(f), ptr(m), g.ref()),
        "energy_p": lambda p, q_m, fi, g: O.orc_energy_p(ptr(p), len(p), q_m, ptr(fi), g.ref()),
    }


def make_species(g, ppc, seed):
    rng = np.random.default_rng(seed)
    n = g.n[0] * g.n[1] * g.n[2] * ppc
    e = random_particles(rng, g, n, vth=0.1, sort=True, q=-1.0 / ppc)
    i = random_particles(rng, g, n, vth=0.1, sort=True, q=+1.0 / ppc)
    i["dx"], i["dy"], i["dz"], i["i"] = e["dx"], e["dy"], e["dz"], e["i"]   # ions co-located: neutral start
    return [{"p": e, "q_m": -1.0}, {"p": i, "q_m": 1.0}]


def gpu_history(vpb, g, species, steps, clean_e=0, clean_b=0, sync_shared=0, f_init=None, lookahead=0, sort=SORT, **layouts):
    """The same run through the library's time-step driver (csrc/vpb_step.cu)."""
    sim = NativeSimulation(g, L=vpb, **layouts)
    sim.set_intervals(clean_e, clean_b, sync_shared=sync_shared)
    sim.set_sort_lookahead(lookahead)
    for k, sp in enumerate(species):
        s = sim.define_species("s%d" % k, sp["q_m"], len(sp["p"]) + 64, sort_interval=sort)
        sim.set_particles(s, sp["p"])
    # also loads the interpolator (initialize.cxx:67)
    sim.set_fields(abi.aligned_zeros(g.nv, abi.field_dtype) if f_init is None else f_init)
    hist = []
    for _ in range(steps):
        sim.advance()
        hist.append(sim.energies())
    return np.array(hist), sim


@pytest.mark.parametrize("kind,n,clean", [("periodic", (16, 16, 16), 0), ("metal", (14, 1, 12), 0), ("periodic", (12, 10, 8), 5)])
def test_energy_history_matches_cpu(vpb, orc, kind, n, clean):
    g = host_grid(n, kind)
    ppc = 8
    h_cpu = cpu_history(oracle_kernels(orc), g, make_species(g, ppc, 3), STEPS, clean, clean)
    h_gpu, sim = gpu_history(vpb, g, make_species(g, ppc, 3), STEPS, clean, clean)
    assert h_cpu.shape == h_gpu.shape == (STEPS, 8)
    scale = np.abs(h_cpu).max(axis=0)
    rel = np.abs(h_gpu - h_cpu) / np.maximum(scale, 1e-300)
    assert rel.max() < 1e-4, rel.max(axis=0)
    # particle number is conserved and kinetic energy is what dominates
    assert [sp.np for sp in sim.species] == [g.n[0] * g.n[1] * g.n[2] * ppc] * 2
    assert h_cpu[-1, 6] > 0 and h_cpu[-1, 7] > 0
