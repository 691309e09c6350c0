#!/usr/bin/env python
"""bench.py -- particle-advances/s of the PIC hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE.json configs[3], "thermal plasma weak scaling,
256^3 cells and 64 ppc per species per GPU", periodic, dt = 0.95 Courant, vth = 0.1 c,
sort every 5 steps (--sort-interval); synthetic particles generated on the device.  One step = one full
time step of vpic_simulation::advance() (src/vpic/advance.cxx:13-244) for this deck:
clear_accumulators, sort_p when due, advance_p for both species, clear_jf,
unload_accumulator, synchronize_jf, advance_b/advance_e/advance_b, load_interpolator.
`value` = particles advanced per second of that whole step, all state resident in HBM.

JSON keys beyond the base contract: roofline (advance_p kernel, algorithmic bytes
64+176/ppc per particle), cpu_baseline (the reference's own V4/SSE
pthreads advance_p, oracle/_ref, on this host's cores, bounded sample), breakdown, and
e2e = BASELINE configs[0] run as an UNMODIFIED reference host program (the reference's
main.cxx + vpic_simulation::advance() + a deck) whose hot path was replaced at link
time by libvpic_b200.so: the program loads its particles on the HOST with the
reference's loader into arrays it allocated itself, every entry point works on those
arrays, and every step the program reads the step's energies back.  `--impl reference`
runs the same program on the reference alone (V4/SSE + pthreads, all cores <= 16) and
reports it as its value: whole time steps on both sides.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VTH = 0.1
# Species sort interval of the headline run.  The reference's decks sort every 20-25 steps because the CPU sort is dear;
# here a sort costs less than two pushes, and advance_p is fastest on freshly grouped particles: 5 (with the look-ahead key
# 3 steps ahead) is where advance_p averages >= 0.50 of the roofline; 10 gives the shortest step.
SORT_INTERVAL = 5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=256, help="cells per axis per GPU")
    ap.add_argument("--ppc", type=int, default=64, help="particles per cell per species")
    ap.add_argument("--e2e-particles", type=int, default=64 * 1024 * 1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--field-cells", type=int, default=1024, help="cells per axis of the field-only leg (configs[1]); 0 = skip")
    ap.add_argument("--workload", default="thermal", choices=["thermal", "fields", "harris", "harris3d"],
                    help="thermal: BASELINE configs[3] (the headline, default); fields: configs[1] alone; harris: configs[2], the "
                         "trecon-part shape 2048x1x1024 cells x 100 ppc on one GPU; harris3d: configs[4], the trecon-part plasma "
                         "in 3D, 1024x512x512 cells x 64 ppc (32 per species) decomposed 2x2x2 over 8 GPUs")
    ap.add_argument("--harris3d-cells", default="", help="global cells X,Y,Z of --workload harris3d at other GPU counts "
                                                          "(default 1024,512,512 at 8 GPUs; 512,256,256 per GPU otherwise)")
    ap.add_argument("--trecon-deck", action="store_true",
                    help="with --workload harris: also run the reference's trecon-part deck itself (unmodified turbulence.cxx, "
                         "config.h knobs 2048 x 1 x 1024, one rank) on libvpic_b200.so -- minutes of host-side load and I/O")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-deck-e2e", dest="deck_e2e", action="store_false",
                    help="skip the deck_e2e key: BASELINE configs[0] as an unmodified reference host program "
                         "(oracle/decks/thermal_c1.cxx), on libvpic_b200.so in the b200 arm, on the reference alone in the "
                         "reference arm; bounded to --deck-timeout seconds, a failure only shows in the key")
    ap.add_argument("--deck-e2e", dest="deck_e2e", action="store_true", help=argparse.SUPPRESS)
    ap.set_defaults(deck_e2e=True)
    ap.add_argument("--deck-steps", type=int, default=0, help="timed steps of the deck run (default: 10 x --steps on the B200, --steps on the CPU)")
    ap.add_argument("--deck-timeout", type=int, default=240)
    ap.add_argument("--clean-div-interval", type=int, default=100,
                    help="clean_div_e/b and sync_shared intervals of the timed run (trecon-part: status_interval/2 = 100)")
    ap.add_argument("--sort-lookahead", type=int, default=-1,
                    help="sort key = voxel the particle reaches this many steps ahead (-1: 0.6 x the sort interval, 0: current voxel)")
    ap.add_argument("--driver", default="native", choices=["native"], help=argparse.SUPPRESS)   # one driver: csrc/vpb_step.cu
    ap.add_argument("--sort-interval", type=int, default=SORT_INTERVAL, help="species sort_interval (default 5; the reference's decks use 20-25)")
    return ap.parse_args()


# ----------------------------------------------------------------------------
# clocks: sampled with nvidia-smi DURING the timed region
# ----------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own advance_p on host cores
# ----------------------------------------------------------------------------
def cpu_reference_rate(cells, ppc, steps, warmup):
    """particle advances/s of clear_accumulators + advance_p(e,i) + reduce_accumulators
    (the reference's p_time bucket, advance.cxx:38-74) with the V4/SSE pthreads library."""
    from oracle import loader
    from old_vpic_b200 import abi
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    cores = os.cpu_count() or 1
    tpp = max(1, min(cores, 16))                      # MAX_PIPELINE = 16 (pipelines.h:6)
    kind = "reference"
    if loader.ref_available("sse"):
        L = loader.ref("sse", tpp=tpp)
        g = helpers.RefGrid(L, (cells, cells, cells), "periodic")
        adv = L.advance_p
        clear = lambda a: L.clear_accumulators(a, g.ref())
        reduce_ = lambda a: L.reduce_accumulators(a, g.ref())
        nrep = 1 + L.refh_n_pipeline()
    else:                                              # reference not built here: scalar port, 1 core
        kind, tpp = "port", 1
        O = loader.oracle()
        g = helpers.host_grid((cells, cells, cells))
        adv = O.orc_advance_p
        clear = lambda a: O.orc_clear_accumulators(a, g.ref())
        reduce_ = lambda a: None
        nrep = 1
    rng = np.random.default_rng(7)
    np_ = cells ** 3 * ppc
    stride = (g.nv + 1) // 2 * 2
    acc = abi.aligned_zeros(nrep * stride, abi.accumulator_dtype)
    fi = abi.aligned_zeros(g.nv, abi.interpolator_dtype)
    species = []
    for q, q_m in ((-1.0 / ppc, -1.0), (1.0 / ppc, 1.0)):   # q = +-V/N: plasma frequency 1 (SURVEY.md 8d, C1 recipe)
        p = abi.aligned_zeros(np_, abi.particle_dtype)
        p["i"] = np.repeat(helpers.interior_voxels(g), ppc)
        for k in ("dx", "dy", "dz"):
            p[k] = rng.uniform(-1, 1, np_).astype(np.float32)
        for k in ("ux", "uy", "uz"):
            p[k] = (VTH * rng.standard_normal(np_)).astype(np.float32)
        p["q"] = q
        species.append((p, q_m, abi.aligned_zeros(max(2 * np_ // 25, 16), abi.mover_dtype)))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        clear(abi.ptr(acc))
        for p, q_m, pm in species:
            adv(abi.ptr(p), np_, q_m, abi.ptr(pm), len(pm), abi.ptr(acc), abi.ptr(fi), g.ref())
        reduce_(abi.ptr(acc))
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"value": 2 * np_ / sec, "unit": "particle-advances/s", "cores": tpp, "kind": kind,
            "sample": "thermal %d^3 cells x %d ppc x 2 species (%d particles), %d steps of clear_accumulators+advance_p+"
                      "reduce_accumulators" % (cells, ppc, 2 * np_, steps), "ms_per_step": 1e3 * sec}


def run_reference(args):
    """The reference arm: BASELINE configs[0] as a reference host program on the reference ALONE (oracle/_ref/thermal_c1.op:
    main.cxx, vpic_simulation::advance(), V4/SSE pipelines on min(cores, 16) pthreads), W warm-up and K timed steps of the
    whole time step.  A bounded sample of the workload `config` names (the CPU cannot hold 2 x 2^30 particles x 48 B within
    minutes): said in `sample`.  Falls back to timing the reference's advance_p alone when the deck executable is missing."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tpp = max(1, min(os.cpu_count() or 1, 16))
    steps, warm = max(1, min(args.steps, 40)), max(0, min(args.warmup, 10))
    deck = deck_e2e(os.path.join(ROOT, "oracle", "_ref", "thermal_c1.op"), steps, warm, tpp,
                    "reference alone (V4/SSE + pthreads hot path)", args.deck_timeout)
    adv = cpu_reference_rate(64, 32, min(steps, 3), 1)          # the p_time bucket alone, for the record
    if "value" in deck:
        value, ms, sample, kind = deck["value"], deck["ms_per_step"], deck["sample"], "reference"
    else:
        value, ms, sample, kind = adv["value"], adv["ms_per_step"], adv["sample"], adv["kind"]
    line = {"metric": "particle-advances/s (push+deposit)", "value": value, "unit": "particle-advances/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": workload_config(args), "sample": sample,
            "cpu_baseline": {"value": value, "unit": "particle-advances/s", "cores": tpp if kind == "reference" else adv["cores"],
                             "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "particle-advances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "advance_p_only": {k: adv[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "deck_e2e": deck, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def deck_e2e(exe, steps, warmup, tpp, what, timeout=240, cells=64, ppc=32):
    """One run of oracle/decks/thermal_c1.cxx (64^3 cells x 32 ppc x 2 species = BASELINE configs[0]) through the
    reference's own main.cxx + vpic_simulation::advance(): the deck loads the particles on the host, reads the energies
    back every step and logs the wall clock of every step (steps.txt); the rate is particle-advances over steps
    warmup..warmup+steps.  Never raises: a failure is reported in the returned dict."""
    import tempfile
    n = cells
    if not os.path.exists(exe):
        return {"unavailable": "%s not built" % os.path.relpath(exe, ROOT)}
    try:
        with tempfile.TemporaryDirectory() as t:
            env = dict(os.environ, VPB_DECK_STEPS=str(warmup + steps), VPB_DECK_CELLS=str(n), VPB_DECK_PPC=str(ppc))
            t0 = time.perf_counter()
            r = subprocess.run([exe, "-tpp=%d" % tpp], cwd=t, env=env, capture_output=True, text=True, timeout=timeout)
            wall = time.perf_counter() - t0
            if r.returncode != 0:
                return {"unavailable": "deck exited %d: %s" % (r.returncode, (r.stdout + r.stderr)[-300:])}
            stamps = dict((int(a), float(b)) for a, b in (ln.split() for ln in open(os.path.join(t, "steps.txt"))))
            sec = stamps[warmup + steps] - stamps[warmup]
            rows = [ln.split() for ln in open(os.path.join(t, "energies")) if ln.strip() and not ln.startswith("%")]
            last = [float(x) for x in rows[-1]]
    except Exception as e:          # noqa: BLE001 -- an optional leg must not take the bench line down
        return {"unavailable": repr(e)[:300]}
    adv = 2.0 * n ** 3 * ppc * steps
    state = 2 * n ** 3 * ppc * 48 + (n + 2) ** 3 * (80 + 80 + 48)
    return {"value": adv / sec, "unit": "particle-advances/s", "ms_per_step": 1e3 * sec / steps, "steps": steps, "warmup": warmup,
            "process_wall_s": wall, "state_bytes": state, "energy_rows": len(rows), "last_energies": last,
            "sample": "%s: BASELINE configs[0], thermal %d^3 cells x %d ppc x 2 species (%d particles), %d timed steps of "
                      "vpic_simulation::advance() after %d warm-up steps (sort every 20, energies read back every step), "
                      "unmodified reference host program, -tpp=%d" % (what, n, ppc, 2 * n ** 3 * ppc, steps, warmup, tpp)}


def trecon_deck(exe, cells, steps, tpp, what, timeout):
    """decks/trecon-part/turbulence.cxx with other config.h knobs (oracle/build_hybrid.sh: symlinked deck sources + a
    generated config.h), as a whole program: the deck loads nppc = 50 particles per cell per species on the host (four
    species + two tracer species that copy them), pushes the tracers itself from its particle-injection hook, dumps
    fields and hydro moments every 20 steps.  Rate = particle pushes (tracers included) over the `simulation time` the
    reference's main loop prints."""
    import re
    import tempfile
    if not os.path.exists(exe):
        return {"unavailable": "%s not built" % os.path.relpath(exe, ROOT)}
    try:
        with tempfile.TemporaryDirectory() as t:
            t0 = time.perf_counter()
            # the library's host wall-clock account of its entry points (VPB_TRACE; ignored by the reference alone)
            env = dict(os.environ, VPB_TRACE="1", VPB_TRACE_FILE=os.path.join(t, "vpb_trace.txt"))
            r = subprocess.run([exe, "-tpp=%d" % tpp], cwd=t, env=env, capture_output=True, text=True, timeout=timeout)
            wall = time.perf_counter() - t0
            m = re.search(r"simulation time: ([0-9.eE+-]+)", r.stdout + r.stderr)
            if r.returncode != 0 or not m:
                return {"unavailable": "deck exited %d: %s" % (r.returncode, (r.stdout + r.stderr)[-300:])}
            sec = float(m.group(1))
            hot = {}
            if os.path.exists(env["VPB_TRACE_FILE"]):
                for ln in open(env["VPB_TRACE_FILE"]):
                    f = ln[len("vpb trace: "):].split()
                    # entry points start right after the prefix; labels nested inside one are indented
                    if ln.startswith("vpb trace: ") and not ln.startswith("vpb trace:  ") and len(f) == 5 and f[0] != "label":
                        hot[f[0]] = float(f[2]) * 1e-3
    except Exception as e:          # noqa: BLE001
        return {"unavailable": repr(e)[:300]}
    pushed = 4 * 50 * cells          # e + i, each also copied into a tracer species (particle_select = 1)
    out_hot = {}
    if hot:
        # Whole run (initialisation included): seconds the program spent inside the library's entry points, i.e. the hot
        # path; the rest of `simulation time` is the deck's own host code (turbulence.cxx / energy.cxx / tracer.cxx loop
        # over every particle on the host for their diagnostics, which also pulls the managed arrays off the device)
        tot = sum(hot.values())
        out_hot = {"hot_path_s": tot, "hot_path_share_of_simulation_time": tot / sec if sec else None,
                   "value_hot_path": pushed * steps / tot if tot else None,
                   "hot_path_top": dict(sorted(hot.items(), key=lambda kv: -kv[1])[:6])}
    return {"value": pushed * steps / sec, "unit": "particle-advances/s", "ms_per_step": 1e3 * sec / steps, "steps": steps,
            "particles_pushed_per_step": pushed, "process_wall_s": wall, **out_hot,
            "sample": "%s: decks/trecon-part/turbulence.cxx, config.h knobs %s cells, topology 1x1x1, %d steps; nppc 50 as shipped "
                      "(BASELINE asks 100: a constant in the deck body, not a knob), tracer copies pushed by the deck, field + hydro "
                      "dumps every 20 steps; -tpp=%d" % (what, cells, steps, tpp)}


H3D_TOPO = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}


def harris3d_shape(args, world):
    if args.harris3d_cells:
        return tuple(int(v) for v in args.harris3d_cells.split(","))
    t = H3D_TOPO[world]
    return (512 * t[0], 256 * t[1], 256 * t[2])


def workload_config(args):
    if args.workload == "harris3d":
        world = int(os.environ.get("WORLD_SIZE", "1"))
        hn, t = harris3d_shape(args, world), H3D_TOPO[world]
        return {"workload": "BASELINE configs[4]: the trecon-part plasma in 3D -- %d x %d x %d cells of 0.488 c/wpe, %d ppc per species "
                            "(%d per cell; tracer copies off), pair plasma at vth=0.6c, wce/wpe=10 force-free sheet field, periodic "
                            "x/y, conducting reflecting z walls, dt=0.99 Courant, sort every %d steps; decomposed %dx%dx%d, one rank per "
                            "GPU; thermal device load without the sheet's drift current"
                            % (hn + (args.ppc, 2 * args.ppc, args.sort_interval) + t),
                "sort_key": "voxel 0.6 x interval steps ahead" if args.sort_lookahead != 0 else "current voxel",
                "cells_global": list(hn), "cells_per_gpu": [hn[0] // t[0], hn[1] // t[1], hn[2] // t[2]], "ppc_per_species": args.ppc,
                "species": 2, "l2_policy": "inputs (100 GB of particles per GPU) are far larger than the 126 MB L2; no flush needed",
                "decomposition": "%dx%dx%d, 1 rank per GPU" % t,
                "capacity": "48 B per particle in the plane layout, 2 species x 1.07 G particles per GPU + one sort buffer of one "
                            "species = 155 GB of 180; the deck's 64 ppc PER SPECIES plus tracer copies (1.65 TB, SURVEY.md 8d) does "
                            "not fit 8 x 180 GB at any layout that keeps the 16-byte tags"}
    if args.workload == "harris":
        return {"workload": "BASELINE configs[2]: the trecon-part shape on one GPU -- 2048 x 1 x 1024 cells (0.488 x 1.95 x 0.488 c/wpe), "
                            "%d ppc per species, pair plasma at vth=0.6c, wce/wpe=10 force-free sheet field, periodic x/y, conducting "
                            "reflecting z walls, dt=0.99 Courant, sort every %d steps; thermal device load without the sheet's drift "
                            "current" % (args.ppc, args.sort_interval),
                "sort_key": "voxel 0.6 x interval steps ahead" if args.sort_lookahead != 0 else "current voxel",
                "cells_per_gpu": [2048, 1, 1024], "ppc_per_species": args.ppc, "species": 2,
                "l2_policy": "inputs (20 GB of particles) are far larger than the 126 MB L2; no flush needed",
                "decomposition": "1 rank"}
    return {"workload": "BASELINE configs[3]: thermal e-/p+ plasma weak scaling, %d^3 cells and %d ppc per species per GPU, "
                        "periodic, dt=0.95 Courant, vth=%.1fc, sort every %d steps" % (args.cells, args.ppc, VTH, args.sort_interval),
            "sort_key": ("voxel %s steps ahead (a look-ahead grouping: same particles, same physics, different array order)" % ("0.6 x interval" if args.sort_lookahead < 0 else args.sort_lookahead)) if args.sort_lookahead != 0 else "current voxel",
            "cells_per_gpu": [args.cells] * 3, "ppc_per_species": args.ppc, "species": 2,
            "l2_policy": "inputs (>=100 GB of particles per GPU) are far larger than the 126 MB L2; no flush needed",
            "decomposition": "1 rank per GPU",
            # both arms print this config; what the CPU can step within minutes is a bounded sample of it
            "e2e_and_reference_arm_sample": "BASELINE configs[0] (64^3 cells x 32 ppc x 2 species, the same thermal plasma) as an "
                                            "unmodified reference host program: `e2e` of this arm and every number of the "
                                            "--impl reference arm are measured on that sample, `value` and `roofline` on the "
                                            "workload above"}


# ----------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from old_vpic_b200 import abi, lib
    from old_vpic_b200 import grid as helpers
    from old_vpic_b200.sim import NativeSimulation

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d" % args.gpus)
    L = lib.load()
    L.vpb_init(local)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            L.vpb_comm_unique_id(buf)
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        ub = (C.c_uint8 * 128)(*uid.cpu().tolist())
        L.vpb_comm_init(rank, world, ub)

    fields_c2 = None
    if world == 1 and args.field_cells > 0 and args.workload not in ("harris", "harris3d"):
        fields_c2 = fields_measure(L, args.field_cells, 5, 3)
        if args.workload == "fields":
            print(json.dumps(fields_c2), flush=True)
            return

    n = args.cells
    harris = args.workload in ("harris", "harris3d")
    if args.workload == "harris3d":
        args.no_e2e = args.no_cpu_baseline = True      # those legs belong to the headline workload
        # BASELINE configs[4]: the same plasma and field as configs[2] (turbulence.cxx:86-160) in a 3-D box of cubic
        # 0.488 c/wpe cells, decomposed over the GPUs; 32 ppc per species = 64 particles per cell
        if args.ppc == 64:
            args.ppc = 32
        hn, topo = harris3d_shape(args, world), H3D_TOPO[world]
        d = 1000.0 / 2048
        hL = (hn[0] * d, hn[1] * d, hn[2] * d)
        g = helpers.make_grid(hn, "periodic", topo=topo, rank=rank, L=hL, dt=helpers.courant_dt(d, d, d, frac=0.99))
        if g.coords[2] == 0:
            g.set_fbc(abi.boundary(0, 0, -1), abi.PEC_FIELDS)
            g.set_pbc(abi.boundary(0, 0, -1), abi.REFLECT_PARTICLES)
        if g.coords[2] == topo[2] - 1:
            g.set_fbc(abi.boundary(0, 0, 1), abi.PEC_FIELDS)
            g.set_pbc(abi.boundary(0, 0, 1), abi.REFLECT_PARTICLES)
    elif harris:
        # BASELINE configs[2]: the shape and plasma of decks/trecon-part/turbulence.cxx:86-160 -- 2048 x 1 x 1024 cells of
        # 0.488 x 1.95 x 0.488 c/wpe, pair plasma (mi/me = 1) at vth = 0.6 c, wce/wpe = 10, force-free sheet of half
        # thickness 6 c/wpe, conducting walls that reflect particles at z = 0, Lz, dt = 0.99 Courant (the deck sorts every 25 steps; here
        # --sort-interval, default 5: at vth = 0.6 c a particle crosses a cell every other step).
        # The device loader is thermal: the sheet's drift current is not loaded (the field near the sheet, 5 % of the
        # box, is not in equilibrium; |B| = b0 everywhere, so the particle work is the deck's).
        if world != 1:
            raise SystemExit("--workload harris is the one-GPU configuration")
        if args.ppc == 64:
            args.ppc = 100
        hn = (2048, 1, 1024)
        hL = (1000.0, 500.0 / 256, 500.0)
        g = helpers.make_grid(hn, "periodic", L=hL, dt=helpers.courant_dt(hL[0] / hn[0], 0, hL[2] / hn[2], frac=0.99))
        for sgn in (-1, 1):
            g.set_fbc(abi.boundary(0, 0, sgn), abi.PEC_FIELDS)
            g.set_pbc(abi.boundary(0, 0, sgn), abi.REFLECT_PARTICLES)
    else:
        topo = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
        g = helpers.make_grid((n * topo[0], n * topo[1], n * topo[2]), "periodic", topo=topo, rank=rank)
    # the library's C++ time-step driver (csrc/vpb_step.cu)
    sim = NativeSimulation(g, n_mat=1, L=L, planar=L.vpb_get_tuning(b"sim.aos_fields") == 0,
                     wide_interpolator=L.vpb_get_tuning(b"sim.narrow_interpolator") == 0,
                     particle_planes=L.vpb_get_tuning(b"sim.aos_particles") == 0)
    sim.set_sort_lookahead(args.sort_lookahead)
    # divergence cleaning and shared-face synchronisation at the interval the reference's trecon-part deck uses
    # (status_interval/2 = 100, turbulence.cxx:209-213); their cost is also measured on its own below
    sim.set_intervals(args.clean_div_interval, args.clean_div_interval, sync_shared=args.clean_div_interval)
    cells = g.n[0] * g.n[1] * g.n[2]
    np_ = cells * args.ppc
    max_np = int(np_ * (1.0 if world == 1 else 1.02)) + 1024
    vth = 0.6 if harris else VTH
    cell_volume = g.struct.dx * g.struct.dy * g.struct.dz
    if harris:
        b0, half = 10.0, 6.0
        f0 = abi.aligned_zeros(g.nv, abi.field_dtype)
        zc = ((np.arange(g.nv) // ((g.n[0] + 2) * (g.n[1] + 2))) - 0.5) * g.struct.dz + g.struct.z0 - 0.5 * hL[2]   # cell centres
        f0["cbx"] = (b0 * np.tanh(zc / half)).astype(np.float32)
        f0["cby"] = (b0 / np.cosh(zc / half)).astype(np.float32)
        sim.set_fields(f0)
        del f0, zc
    # macro-charge q = +-(cell volume)/ppc so that the plasma frequency is 1 (the reference's thermal recipe,
    # q = +-L^3/Ne); dt*wpe = 0.55
    for name, q_m, q, seed in (("electron", -1.0, -cell_volume / args.ppc, 7 + rank), ("ion", 1.0, cell_volume / args.ppc, 1007 + rank)):
        sp = sim.define_species(name, q_m, max_np, sort_interval=args.sort_interval)
        sim.load_thermal(sp, args.ppc, vth, q, seed, tag0=rank * (1 << 40))
    L.vpb_sync()

    def barrier():
        L.vpb_sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        sim.advance()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    L.vpb_prof_enable(1)
    L.vpb_launch_count(1)
    L.vpb_timer_start(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sim.advance()
    L.vpb_timer_stop(0)
    ms = L.vpb_timer_ms(0)
    barrier()
    wall = time.perf_counter() - t0
    launches = L.vpb_launch_count(1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    prof = {}
    names = ["advance_p", "sort_p", "advance_b", "advance_e", "load_interpolator", "unload_accumulator", "curl_b", "boundary_p",
             "halo", "div_clean"]
    for cls, nm in enumerate(names):
        tot, cnt = C.c_double(0), C.c_int(0)
        L.vpb_prof_collect(cls, C.byref(tot), C.byref(cnt), 0)
        prof[nm] = (tot.value, cnt.value)
    lst = (C.c_float * 4096)()
    nl = L.vpb_prof_list(0, lst, 4096)
    adv_list = [round(float(lst[i]), 3) for i in range(nl)]
    nl = L.vpb_prof_list(1, lst, 4096)
    sort_list = [round(float(lst[i]), 3) for i in range(nl)]
    L.vpb_prof_collect(0, None, None, 1)
    # every rank's own class times: the ranks meet at the first exchange after advance_p (the count message of boundary_p),
    # so whatever one rank's push and sort took less than the slowest rank's shows up as ITS boundary_p time
    per_rank = None
    if world > 1:
        mine = torch.tensor([prof[nm][0] / args.steps for nm in names], dtype=torch.float64, device="cuda")
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        tab = torch.stack(every).cpu().numpy()
        per_rank = {nm: [round(float(x), 3) for x in tab[:, k]] for k, nm in enumerate(names) if nm in ("advance_p", "sort_p", "boundary_p", "halo")}
    # one step with divergence cleaning (E and B) and shared-face synchronisation forced on, timed on its own: what the
    # interval-100 steps of a long run cost beyond an ordinary step (collective: every rank takes it)
    sim.set_intervals(1, 1, sync_shared=1)
    sim.advance()
    tot, cnt = C.c_double(0), C.c_int(0)
    L.vpb_prof_collect(9, C.byref(tot), C.byref(cnt), 1)
    clean_ms = tot.value
    L.vpb_prof_enable(0)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    total_particles = 2 * np_ * world
    planes = L.vpb_get_tuning(b"sim.aos_particles") == 0
    value = total_particles * args.steps / (ms_max * 1e-3)
    peak, peak_src = measured_peak()
    adv_ms, adv_n = prof["advance_p"]
    bytes_alg = 64.0 + 176.0 / args.ppc
    per_launch_particles = np_
    achieved = bytes_alg * per_launch_particles / (adv_ms / max(adv_n, 1) * 1e-3) / 1e9 if adv_n else None
    line = {
        "metric": "particle-advances/s (push+deposit)", "value": value, "unit": "particle-advances/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (device-generated thermal load)",
        "config": workload_config(args), "clocks": clocks, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None,
                     # DRAM bytes of ONE launch from the ncu --set full capture of this workload (3 steps after a sort at interval 5)
                     "traffic": (TRAFFIC_PLANES if planes else TRAFFIC_AOS) if (args.cells == 256 and args.ppc == 64) else None,
                     "traffic_source": ("profiles/r2g_256_advance_p_pair_full_3steps_after_sort.txt" if planes else
                                        "profiles/r1k_256_step10_advance_p_stream_c2.txt") + " (dram__bytes_read.sum + dram__bytes_write.sum)",
                     "kernel": "advance_p_pair_kernel" if planes else "advance_p_stream_kernel",
                     "algorithmic_bytes_per_particle": bytes_alg, "particles_per_launch": per_launch_particles,
                     "avg_launch_ms": adv_ms / max(adv_n, 1), "peak_source": peak_src,
                     # what the device layout makes DRAM move at least: component planes read 8 words and write 6 per
                     # particle; the reference's 48-byte records move whole (every 32-byte sector holds hot bytes)
                     "layout_imposed_bytes_per_particle": (56.0 if planes else 96.0) + 176.0 / args.ppc,
                     "frac_of_nominal_8TBs": (achieved / 8000.0) if achieved else None,
                     "min_launch_ms": min(adv_list) if adv_list else None, "max_launch_ms": max(adv_list) if adv_list else None},
        "breakdown_ms_per_step": dict({k: v[0] / args.steps for k, v in prof.items()},
                                      other=ms_max / args.steps - sum(v[0] for v in prof.values()) / args.steps),
        # N > 1: the same classes on every rank (ms per step)
        "ranks_ms_per_step": per_rank,
        "sort_p": {"ms_per_sort": (sum(sort_list) / len(sort_list)) if sort_list else None, "sorts_timed": len(sort_list),
                   "particles_per_sort": np_, "algorithmic_bytes_per_particle": 100.0,
                   "frac": (100.0 * np_ / (sum(sort_list) / len(sort_list) * 1e-3) / 1e9 / peak) if sort_list else None,
                   "kernel": ({2: "group_keys_kernel + group_invert_kernel + group_gather_kernel", 1: "group_keys_kernel + group_scatter_kernel",
                               0: "group_keys_kernel + group_move_kernel"}[int(os.environ.get("VPB_SORT_GROUP_VARIANT", "2"))]
                              + " (csrc/vpb_sort_group.cu)")
                             if int(os.environ.get("VPB_SORT_GROUPED", "1")) else "round-1 pipeline (csrc/vpb_particles.cu)"},
        "div_clean": {"ms_per_cleaning_step": clean_ms, "interval": args.clean_div_interval,
                      "amortised_ms_per_step": clean_ms / args.clean_div_interval if args.clean_div_interval > 0 else 0.0,
                      "what": "clean_div_e (rho accumulation of both species, 2 passes) + clean_div_b (2 passes) + synchronize_tang_e_norm_b"},
        "advance_p_only_particle_advances_per_s": (2 * np_ * args.steps / (adv_ms * 1e-3)) if adv_ms else None,
        "field_cell_updates_per_s": {"advance_b": (2 * cells * args.steps / (prof["advance_b"][0] * 1e-3)) if prof["advance_b"][0] else None,
                                     "advance_e": (cells * args.steps / (prof["advance_e"][0] * 1e-3)) if prof["advance_e"][0] else None},
        "host_wall_ms_per_step": 1e3 * wall / args.steps, "driver": "csrc/vpb_step.cu (vpb_sim_advance)",
        "l2_fetch_granularity_bytes": int(L.vpb_l2_fetch_granularity()),
        "advance_p_ms_by_launch": adv_list,
        # effective values: the environment override if there is one, else the library's default
        "tuning": {k: int(os.environ.get("VPB_" + k.upper().replace(".", "_"), d)) for k, d in (
            ("advance_p.pair_variant", 1), ("advance_p.pair_cps", 4), ("advance_p.pair_pipe", 1), ("advance_p.pair_merge", 1),
            ("sort.grouped", 1), ("sort.group_variant", 2), ("sort.pack_rank", 1), ("sort.gather_keep", 0), ("sim.graph", 1), ("sim.aos_fields", 0), ("sim.narrow_interpolator", 0), ("sim.aos_particles", 0),
            ("advance_p.tma", 2), ("advance_p.stream_cps", 5), ("advance_p.stream_store", 0), ("advance_p.deposit", 1))},
    }
    if fields_c2 is not None:
        line["fields_c2"] = fields_c2
    # release the big run before the other legs
    sim.free()
    if world == 1 and args.workload == "thermal" and not args.no_e2e:
        line["small_step"] = small_step_measure(L)
    if world == 1 and args.workload == "thermal" and not args.no_e2e:
        try:
            line["load_mt"] = load_mt_measure(L, not args.no_cpu_baseline)
        except Exception as exc:              # a side measurement must not cost the bench line
            line["load_mt"] = {"failed": repr(exc)}
    if harris and args.trecon_deck:
        hyb = os.path.join(ROOT, "oracle", "_ref", "hybrid")
        line["trecon_deck"] = trecon_deck(os.path.join(hyb, "turbulence_c2.b200.op"), 2048 * 1024, 60, 1,
                                          "reference host objects + libvpic_b200.so", 1500)
        # the same deck at a shape the CPU finishes in seconds, both ways, for a like-for-like ratio
        tpp = max(1, min(os.cpu_count() or 1, 16))
        line["trecon_deck_scaled"] = {
            "b200": trecon_deck(os.path.join(hyb, "turbulence_c2s.b200.op"), 128 * 64, 200, 1, "reference host objects + libvpic_b200.so", 600),
            "reference": trecon_deck(os.path.join(ROOT, "oracle", "_ref", "turbulence_c2s_sse.op"), 128 * 64, 200, tpp,
                                     "reference alone (V4/SSE + pthreads)", 600)}
    if args.deck_e2e and not args.no_e2e:
        # e2e: the reference-facing path.  An unmodified reference host program (HOST-allocated arrays, HOST particle
        # load) on libvpic_b200.so; at N > 1 this is still rank 0's GPU alone (said in the key)
        dk = deck_e2e(os.path.join(ROOT, "oracle", "_ref", "hybrid", "thermal_c1.b200.op"), args.deck_steps or 10 * args.steps,
                      max(args.warmup, 3), 1, "reference host objects + libvpic_b200.so (link-time substitution)", args.deck_timeout)
        line["deck_e2e"] = dk
        if "value" in dk:
            line["e2e"] = {"value": dk["value"], "unit": dk["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8 * 8,
                           "h2d_bytes_once": dk["state_bytes"], "ms_per_step": dk["ms_per_step"], "gpus": 1, "sample": dk["sample"],
                           "note": "the host program's arrays (util_malloc_aligned -> CUDA managed memory) are filled on the host "
                                   "by the reference's loader and migrate to the device when the hot path first touches them "
                                   "(h2d_bytes_once, before the timed steps, like the reference arm's load); after that a step "
                                   "needs no host data, and the program reads eight energies back per step"}
    if not args.no_e2e and world == 1:
        line["e2e_host_staged"] = e2e_measure(L, args, abi, helpers)
        if "e2e" not in line:
            line["e2e"] = line["e2e_host_staged"]
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = {k: v for k, v in cpu_reference_rate(64, args.ppc, 3, 1).items() if k != "ms_per_step"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def load_mt_measure(L, with_cpu, n=64, ppc=32):
    """The particle load of BASELINE configs[0]: seed_rand(7), then 64^3 x 32 iterations of three
    uniform_rand, six maxwellian_rand and two inject_particle.  On the device from the reference's own Mersenne-Twister
    stream (csrc/vpb_mt.cu: same particles, bit for bit -- tests/test_gpu_mt.py), wall clock of the whole call; beside it
    the serial host loop (cpu_baseline leg: the oracle's restatement of the deck's loop on a bounded sample)."""
    import ctypes as C
    from old_vpic_b200 import grid as helpers
    from old_vpic_b200.sim import NativeSimulation
    g = helpers.make_grid((n, n, n), "periodic")
    pairs = n ** 3 * ppc
    sim = NativeSimulation(g, L=L)
    e = sim.define_species("electron", -1.0, pairs + 1024, sort_interval=20)
    i = sim.define_species("ion", 1.0, pairs + 1024, sort_interval=20)
    q = float(n) ** 3 / pairs
    out = {"workload": "BASELINE configs[0] load: %d iterations of {3 uniform_rand, 6 maxwellian_rand, 2 inject_particle}, seed 7" % pairs}
    times = []
    for rep in range(4):                      # the first call is untimed (buffers allocated, kernels loaded)
        if rep:
            sim.free()
            sim = NativeSimulation(g, L=L)
            e = sim.define_species("electron", -1.0, pairs + 1024, sort_interval=20)
            i = sim.define_species("ion", 1.0, pairs + 1024, sort_interval=20)
        L.vpb_sync()
        t0 = time.perf_counter()
        sim.load_pairs_mt(e, i, pairs, [0, 0, 0], [n, n, n], 0.1, 0.1, -q, q, seed=7)
        L.vpb_sync()
        if rep:
            times.append(time.perf_counter() - t0)
    # the call allocates and frees ~1.2 GB of device buffers (word stream, deviate table); now and then one of those
    # cudaMalloc/cudaFree calls takes half a second on its own, so: the median of three calls, all three listed
    dt = sorted(times)[1]
    out["device"] = {"seconds": dt, "pairs_per_s": pairs / dt, "particles": 2 * pairs, "seconds_each_call": [round(t, 4) for t in times]}
    sim.free()
    if with_cpu:
        from oracle import loader
        O = loader.oracle()
        O.orc_mt_sizeof.restype = C.c_int
        O.orc_load_thermal_pairs.restype = C.c_long
        O.orc_load_thermal_pairs.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                             C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        O.orc_mt_seed.argtypes = [C.c_void_p, C.c_uint]
        rng = np.zeros(O.orc_mt_sizeof(), np.uint8)
        O.orc_mt_seed(rng.ctypes.data, 7)
        sample = 1 << 21
        from old_vpic_b200 import abi
        pe, pi = abi.aligned_zeros(sample, abi.particle_dtype), abi.aligned_zeros(sample, abi.particle_dtype)
        npe, npi = C.c_int(0), C.c_int(0)
        lo, hi = np.zeros(3), np.full(3, float(n))
        t0 = time.perf_counter()
        O.orc_load_thermal_pairs(rng.ctypes.data, sample, lo.ctypes.data, hi.ctypes.data, 0.1, 0.1, q, pe.ctypes.data, C.byref(npe), sample,
                                 pi.ctypes.data, C.byref(npi), sample, g.ref())
        ht = time.perf_counter() - t0
        out["host_serial_loop"] = {"seconds_per_sample": ht, "pairs_per_s": sample / ht, "kind": "port", "cores": 1,
                                   "sample": "%d iterations of the same loop (oracle/oracle_mt.c: mt_drand, mt_drandn, inject_particle)" % sample}
        out["speedup"] = out["device"]["pairs_per_s"] / out["host_serial_loop"]["pairs_per_s"]
    return out


def small_step_measure(L, n=64, ppc=32, steps=200):
    """BASELINE configs[0]'s shape (64^3 cells x 32 ppc x 2 species) device-resident through the C++ driver: a step of
    about half a millisecond, where the ~20 small launches of the field part matter.  Timed with the field part replayed
    from its captured CUDA graph (the default) and launched kernel by kernel."""
    from old_vpic_b200 import grid as helpers
    from old_vpic_b200.sim import NativeSimulation
    out = {"workload": "thermal %d^3 cells x %d ppc x 2 species, device-resident, %d steps, sort every 5" % (n, ppc, steps)}
    for graph in (1, 0):
        L.vpb_set_tuning(b"sim.graph", graph)
        try:
            g = helpers.make_grid((n, n, n), "periodic")
            sim = NativeSimulation(g, L=L)
            sim.set_sort_lookahead(-1)
            for name, q_m, q, seed in (("electron", -1.0, -1.0 / ppc, 7), ("ion", 1.0, 1.0 / ppc, 1007)):
                sp = sim.define_species(name, q_m, n ** 3 * ppc + 1024, sort_interval=5)
                sim.load_thermal(sp, ppc, VTH, q, seed)
            sim.advance(10)
            L.vpb_sync()
            L.vpb_timer_start(2)
            sim.advance(steps)
            L.vpb_timer_stop(2)
            ms = L.vpb_timer_ms(2)
            key = "graph" if graph else "kernel_by_kernel"
            out[key] = {"ms_per_step": ms / steps, "particle_advances_per_s": 2.0 * n ** 3 * ppc * steps / (ms * 1e-3),
                        "graph_replays": int(L.vpb_sim_graph_replays(sim.h))}
            sim.free()
        finally:
            L.vpb_set_tuning(b"sim.graph", 1)
    return out


def fields_measure(L, n, steps, warmup):
    """BASELINE configs[1]: field-only Yee plane wave on n^3 cells (advance_b x2 + advance_e per step), first with the
    vacuum field advance (vfa_advance_e.c), then with the standard one over a two-entry material table (advance_e.c: the
    kernel reads every voxel's material ids).  Returns cell-update rates and the roofline fractions of the stencil
    kernels, timed with CUDA events around each kernel launch (vpb_prof classes 2 and 3) over `steps` steps."""
    from old_vpic_b200 import grid as helpers
    from old_vpic_b200.sim import NativeSimulation
    peak, _ = measured_peak()
    cells = float(n) ** 3
    res = {"workload": "BASELINE configs[1]: field-only Yee vacuum plane wave, %d^3 cells, periodic, 1 GPU" % n, "steps": steps,
           "field_layout": "planar" if L.vpb_get_tuning(b"sim.aos_fields") == 0 else "aos"}
    # algorithmic bytes per cell and what the quad-planar device layout moves (DESIGN.md "field
    # layout"): advance_b reads the e quad and reads+writes the cb quad = 48 B; vacuum advance_e reads cb and jf
    # quads and reads+writes the e quad = 64 B; the standard advance_e also reads+writes the tca quad and reads the
    # material-id quad = 112 B.  (On the reference's 80-byte AoS array ncu measures 112 B per cell for the vacuum
    # kernels, profiles/r1j.)
    for variant, vacuum, n_mat, kernels in (("vacuum", True, 1, (("advance_b", 2, 36.0, 48.0), ("advance_e", 3, 48.0, 64.0))),
                                            ("standard", False, 2, (("advance_e_standard", 3, 84.0, 112.0),))):
        g = helpers.make_grid((n, n, n), "periodic", field_only=True)
        sim = NativeSimulation(g, n_mat=n_mat, vacuum=vacuum, L=L, planar=L.vpb_get_tuning(b"sim.aos_fields") == 0)
        L.vpb_load_plane_wave(sim.dom, sim.field_ptr, 8, 1.0)
        e0 = sum(sim.energies()[:6])
        for _ in range(warmup):
            sim.advance()
        L.vpb_sync()
        L.vpb_prof_enable(1)
        L.vpb_timer_start(1)
        for _ in range(steps):
            sim.advance()
        L.vpb_timer_stop(1)
        ms = L.vpb_timer_ms(1)
        out = {}
        for nm, cls, _, _ in kernels:
            tot, cnt = C.c_double(0), C.c_int(0)
            L.vpb_prof_collect(cls, C.byref(tot), C.byref(cnt), 0)
            out[nm] = (tot.value, cnt.value)
        L.vpb_prof_collect(0, None, None, 1)
        L.vpb_prof_enable(0)
        e1 = sum(sim.energies()[:6])
        sim.free()
        if variant == "vacuum":
            res.update({"ms_per_step": ms / steps, "field_cell_updates_per_s": 3 * cells * steps / (ms * 1e-3),
                        "em_energy_drift_rel": abs(e1 - e0) / e0})
        else:
            res["standard_ms_per_step"] = ms / steps
            res["standard_em_energy_drift_rel"] = abs(e1 - e0) / e0
        for nm, _, alg, layout in kernels:
            tot, cnt = out[nm]
            if not cnt:
                continue
            rate = cells * cnt / (tot * 1e-3)
            res[nm] = {"cell_updates_per_s": rate, "avg_launch_ms": tot / cnt, "algorithmic_bytes_per_cell": alg,
                       "layout_bytes_per_cell": layout, "achieved_GBs": rate * alg / 1e9, "frac": rate * alg / 1e9 / peak,
                       "frac_of_layout_bound": rate * layout / 1e9 / peak}
    return res


def e2e_measure(L, args, abi, helpers):
    """advance_p through the reference-named entry point with pinned HOST arrays: every call copies the
    particles, interpolator and accumulators host->device, runs the kernel, and copies particles, movers
    and accumulators back (vpb_dropin.cu).  Bounded to --e2e-particles per species."""
    n = 128
    ppc = max(1, min(args.ppc, args.e2e_particles // n ** 3))
    g = helpers.make_grid((n, n, n))
    np_ = n ** 3 * ppc

    def pinned(count, dtype):
        dtype = np.dtype(dtype)
        addr = L.vpb_host_alloc_pinned(count * dtype.itemsize)
        a = np.ctypeslib.as_array(C.cast(addr, C.POINTER(C.c_uint8)), shape=(count * dtype.itemsize,)).view(dtype)
        a.view(np.uint8)[:] = 0
        return a

    rng = np.random.default_rng(3)
    p = pinned(np_, abi.particle_dtype)
    p["i"] = np.repeat(helpers.interior_voxels(g), ppc)
    for k in ("dx", "dy", "dz"):
        p[k] = rng.uniform(-1, 1, np_).astype(np.float32)
    for k in ("ux", "uy", "uz"):
        p[k] = (VTH * rng.standard_normal(np_)).astype(np.float32)
    p["q"] = -1.0 / ppc
    max_nm = max(2 * np_ // 25, 16)
    pm = pinned(max_nm, abi.mover_dtype)
    acc = pinned(g.nv, abi.accumulator_dtype)
    fi = pinned(g.nv, abi.interpolator_dtype)
    sz = (C.c_size_t * 2)()
    for _ in range(2):
        L.advance_p(abi.ptr(p), np_, -1.0, abi.ptr(pm), max_nm, abi.ptr(acc), abi.ptr(fi), g.ref())
    L.vpb_staging_bytes(C.byref(sz, 0), C.byref(sz, 8))
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        L.advance_p(abi.ptr(p), np_, -1.0, abi.ptr(pm), max_nm, abi.ptr(acc), abi.ptr(fi), g.ref())
    sec = (time.perf_counter() - t0) / reps
    L.vpb_staging_bytes(C.byref(sz, 0), C.byref(sz, 8))
    return {"value": np_ / sec, "unit": "particle-advances/s", "h2d_bytes_per_step": int(sz[0] // reps),
            "d2h_bytes_per_step": int(sz[1] // reps),
            "sample": "advance_p() C-ABI call, pinned host arrays, %d^3 cells x %d ppc = %d particles per call" % (n, ppc, np_),
            "ms_per_call": 1e3 * sec}


TRAFFIC_AOS = 123.35e9      # profiles/r1k (48-byte records, advance_p_stream_kernel)
TRAFFIC_PLANES = 73.39e9    # profiles/r2g (component planes, advance_p_pair_kernel FULL variant, 3 steps after a sort at interval 5)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
