import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from oracle import loader
from old_vpic_b200 import lib, abi
from old_vpic_b200.abi import ptr
from old_vpic_b200.sim import NativeSimulation
import test_gpu_harris as H
from test_gpu_history import SORT, oracle_kernels
vpb = lib.load(); vpb.vpb_init(0); orc = loader.oracle()
K = oracle_kernels(orc)
nx, nz, ce = 24, 20, 5
g = H.trecon_grid(nx, nz); f0 = H.sheet_fields(g)
species = H.sheet_species(g, 16, 9)
sim = NativeSimulation(g, L=vpb); sim.set_intervals(ce, 0)
for k, sp in enumerate(H.sheet_species(g, 16, 9)):
    s = sim.define_species("s%d" % k, sp["q_m"], len(sp["p"]) + 64, sort_interval=SORT); sim.set_particles(s, sp["p"])
sim.set_fields(f0)
f = abi.aligned_zeros(g.nv, abi.field_dtype); f[:] = f0
fi = abi.aligned_zeros(g.nv, abi.interpolator_dtype)
a = abi.aligned_zeros(g.nv, abi.accumulator_dtype)
m = abi.aligned_zeros(1, abi.material_coefficient_dtype)
for k in ("decayx", "drivex", "decayy", "drivey", "decayz", "drivez", "rmux", "rmuy", "rmuz", "nonconductive", "epsx", "epsy", "epsz"):
    m[k] = 1.0
K["load_interpolator"](fi, f, g)
def report(step, tag, fc):
    fg = sim.get_fields()
    out = []
    for c in ("ex", "ey", "ez", "cbx", "cby", "cbz", "jfx", "jfy", "jfz", "rhof", "rhob", "div_e_err", "tcax"):
        d = np.abs(fg[c] - fc[c]); sc = max(float(np.abs(fc[c]).max()), 1e-30)
        out.append("%s %.1e" % (c, d.max() / sc))
    print(step, tag, " ".join(out), flush=True)
for step in range(8):
    K["clear_accumulators"](a, g)
    for sp in species:
        if step % SORT == 0: sp["p"] = K["sort"](sp["p"], g)
    for sp in species:
        pm = abi.aligned_zeros(len(sp["p"]), abi.mover_dtype)
        assert K["advance_p"](sp["p"], sp["q_m"], pm, a, fi, g) == 0
    K["clear_jf"](f, g); K["unload_accumulator"](f, a, g); K["synchronize_jf"](f, g)
    K["advance_b"](f, g, 0.5); K["advance_e"](f, m, g); K["advance_b"](f, g, 0.5)
    if ce and step % ce == 0:
        K["clear_rhof"](f, g)
        for sp in species: K["accumulate_rho_p"](f, sp["p"], g)
        K["synchronize_rho"](f, g)
        for _ in range(2):
            K["compute_div_e_err"](f, m, g); K["clean_div_e"](f, m, g)
    K["load_interpolator"](fi, f, g)
    sim.advance()
    report(step, "after", f)
