"""Device-resident time-step driver: the call order of vpic_simulation::advance()
(src/vpic/advance.cxx:13-244) issued against layer (B) of libvpic_b200.so, with
every array resident in HBM between steps.  Host Python only sequences the calls
(about 15 per step); nothing in the loop touches particle or field data on the
host.
"""
import ctypes as C

import numpy as np

from . import abi, lib


class DevArray:
    """A typed device allocation owned by the library (cudaMalloc)."""

    def __init__(self, L, n, dtype):
        self.L, self.n, self.dtype = L, int(n), np.dtype(dtype)
        self.nbytes = self.n * self.dtype.itemsize
        self.ptr = L.vpb_dev_alloc(max(self.nbytes, 16))

    def upload(self, host):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        assert host.nbytes <= self.nbytes
        self.L.vpb_h2d(self.ptr, host.ctypes.data, host.nbytes)
        self.L.vpb_sync()

    def download(self, n=None):
        n = self.n if n is None else int(n)
        out = abi.aligned_empty(n, self.dtype)
        if n:
            self.L.vpb_d2h(out.ctypes.data, self.ptr, n * self.dtype.itemsize)
        self.L.vpb_sync()
        return out

    def free(self):
        if self.ptr:
            self.L.vpb_dev_free(self.ptr)
            self.ptr = None


class FieldArray:
    """The device field array of one domain in the domain's layout (include/vpic_b200.h "Device field layout").
    upload()/download() speak the reference's AoS field_t[nv] and convert on the device."""

    def __init__(self, L, dom, nv):
        self.L, self.dom, self.n, self.dtype = L, dom, int(nv), abi.field_dtype
        self.nbytes = int(L.vpb_field_bytes(dom))
        self.ptr = L.vpb_dev_alloc(self.nbytes)     # zero-filled

    def _planar(self):
        return bool(self.L.vpb_domain_field_layout(self.dom))

    def upload(self, host):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        assert len(host) == self.n
        L = self.L
        if not self._planar():
            L.vpb_h2d(self.ptr, host.ctypes.data, host.nbytes)
        else:
            tmp = L.vpb_dev_alloc(host.nbytes)
            L.vpb_h2d(tmp, host.ctypes.data, host.nbytes)
            L.vpb_field_convert(self.dom, self.ptr, tmp, 1)
            L.vpb_sync()
            L.vpb_dev_free(tmp)
        L.vpb_sync()

    def download(self):
        L = self.L
        out = abi.aligned_empty(self.n, self.dtype)
        if not self._planar():
            L.vpb_d2h(out.ctypes.data, self.ptr, out.nbytes)
        else:
            tmp = L.vpb_dev_alloc(out.nbytes)
            L.vpb_field_convert(self.dom, tmp, self.ptr, 0)
            L.vpb_d2h(out.ctypes.data, tmp, out.nbytes)
            L.vpb_sync()
            L.vpb_dev_free(tmp)
        L.vpb_sync()
        return out

    def free(self):
        if self.ptr:
            self.L.vpb_dev_free(self.ptr)
            self.ptr = None


class ParticleArray(DevArray):
    """One species' device particle array in the domain's particle layout (include/vpic_b200.h "Device particle
    layout").  upload()/download() speak the reference's particle_t[] and convert on the device."""

    def __init__(self, L, dom, capacity):
        self.dom = dom
        plane = int(L.vpb_domain_particle_layout(dom))
        assert plane == 0 or capacity <= plane, "capacity exceeds the domain's particle plane stride"
        DevArray.__init__(self, L, plane if plane else capacity, abi.particle_dtype)

    def _planes(self):
        return int(self.L.vpb_domain_particle_layout(self.dom)) > 0

    def upload(self, host):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        if not self._planes() or len(host) == 0:
            return DevArray.upload(self, host)
        L = self.L
        tmp = L.vpb_dev_alloc(host.nbytes)
        L.vpb_h2d(tmp, host.ctypes.data, host.nbytes)
        L.vpb_particle_convert(self.dom, self.ptr, tmp, len(host), 1)
        L.vpb_sync()
        L.vpb_dev_free(tmp)

    def download(self, n=None):
        n = self.n if n is None else int(n)
        if not self._planes() or n == 0:
            return DevArray.download(self, n)
        L = self.L
        out = abi.aligned_empty(n, self.dtype)
        tmp = L.vpb_dev_alloc(out.nbytes)
        L.vpb_particle_convert(self.dom, tmp, self.ptr, n, 0)
        L.vpb_d2h(out.ctypes.data, tmp, out.nbytes)
        L.vpb_sync()
        L.vpb_dev_free(tmp)
        return out


class Species:
    def __init__(self, L, dom, name, q_m, max_np, max_nm, sort_interval, sp_id):
        self.name, self.q_m, self.id = name, float(q_m), sp_id
        self.max_np, self.max_nm, self.sort_interval = int(max_np), int(max_nm), int(sort_interval)
        self.np = 0
        self.p = ParticleArray(L, dom, max_np)
        self.pm = DevArray(L, max_nm, abi.mover_dtype)
        self.nm = DevArray(L, 4, np.int32)
        self.partition = None        # int[nv+1] from the last sort of this species (traversal hint)


class Simulation:
    """One rank's share of a PIC run on one GPU."""

    def __init__(self, grid, n_mat=1, vacuum=False, L=None, planar=True, wide_interpolator=True, particle_planes=True):
        self.L = L or lib.load()
        self.L.vpb_init(-1)
        self.grid = grid
        self.dom = self.L.vpb_domain_create(grid.ref(), grid.rank, grid.nproc)
        # the field array never leaves the device in this driver: keep it in the planar layout
        self.L.vpb_domain_set_field_layout(self.dom, 1 if planar else 0)
        self.nv = grid.nv
        self.vacuum = vacuum
        self.f = FieldArray(self.L, self.dom, self.nv)
        # a field-only grid (no neighbor table, grid.py) carries no particles: no interpolator / accumulators
        # ... and the interpolator in 96-byte records (include/vpic_b200.h "Device interpolator layout")
        self.L.vpb_domain_set_interpolator_layout(self.dom, 1 if wide_interpolator else 0)
        self.fi = None if grid.field_only else DevArray(self.L, int(self.L.vpb_interpolator_bytes(self.dom)), np.uint8)
        self.a = None if grid.field_only else DevArray(self.L, self.nv + 1, abi.accumulator_dtype)
        self.m_host = None
        self.m = None
        self.n_mat = n_mat
        if not vacuum:
            from_vac = abi.aligned_zeros(n_mat, abi.material_coefficient_dtype)
            for k in ("decayx", "drivex", "decayy", "drivey", "decayz", "drivez", "rmux", "rmuy", "rmuz", "nonconductive",
                      "epsx", "epsy", "epsz"):
                from_vac[k] = 1.0
            self.m = DevArray(self.L, n_mat, abi.material_coefficient_dtype)
            self.m.upload(from_vac)
        # species arrays as component planes (two-particles-per-lane advance_p); the plane stride is fixed by the
        # first define_species() and shared by every species of the domain
        self.particle_planes = bool(particle_planes)
        self.species = []
        self.sort_tmp = None
        self.step = 0
        self.clean_div_e_interval = 0
        self.clean_div_b_interval = 0
        self.scalars = DevArray(self.L, 16, np.float64)

    @property
    def m_ptr(self):
        return None if self.m is None else self.m.ptr

    def free(self):
        """Release every device array of this run."""
        for arr in [self.f, self.fi, self.a, self.m, self.sort_tmp, self.scalars, getattr(self, "_hydro", None)] + \
                [x for sp in self.species for x in (sp.p, sp.pm, sp.nm, sp.partition)]:
            if arr is not None:
                arr.free()
        self.L.vpb_domain_destroy(self.dom)
        self.dom = None

    def define_species(self, name, q_m, max_np, max_nm=None, sort_interval=20):
        assert not self.grid.field_only, "a field-only grid cannot carry particles"
        max_nm = max_nm if max_nm is not None else max(2 * max_np // 25, 16)   # vpic.hxx:416-420
        if self.particle_planes and not self.species:
            self.L.vpb_domain_set_particle_layout(self.dom, (int(max_np) + 63) // 64 * 64)
        sp = Species(self.L, self.dom, name, q_m, max_np, max_nm, sort_interval, len(self.species))
        self.species.append(sp)
        return sp

    def load_thermal(self, sp, ppc, vth, q, seed, tag0=0):
        nx, ny, nz = self.grid.n
        sp.np = ppc * nx * ny * nz
        assert sp.np <= sp.max_np
        self.L.vpb_load_thermal(self.dom, sp.p.ptr, ppc, vth, q, seed, tag0)

    # -- pieces of advance.cxx ------------------------------------------------
    def sort(self, sp):
        L = self.L
        if self.sort_tmp is None or self.sort_tmp.n < sp.max_np:
            if self.sort_tmp is not None:
                self.sort_tmp.free()
            self.sort_tmp = ParticleArray(L, self.dom, sp.max_np)
        if sp.partition is None:
            sp.partition = DevArray(L, self.nv + 1, np.int32)
        if int(L.vpb_domain_particle_layout(self.dom)) > 0:
            L.vpb_sort_p_planes(self.dom, sp.p.ptr, self.sort_tmp.ptr, sp.np, sp.partition.ptr)   # sorted planes return to sp.p
        else:
            L.vpb_sort_p(self.dom, sp.p.ptr, self.sort_tmp.ptr, sp.np, sp.partition.ptr)
            sp.p, self.sort_tmp = self.sort_tmp, sp.p     # out-of-place: swap (sort_p.c:76-77)

    def advance_fields(self):
        L, dom, f = self.L, self.dom, self.f.ptr
        L.vpb_advance_b(dom, f, 0.5)                                    # advance.cxx:129
        L.vpb_advance_e(dom, f, self.m_ptr, self.n_mat, int(self.vacuum))   # :133
        L.vpb_advance_b(dom, f, 0.5)                                    # :147

    def advance(self):
        L, dom = self.L, self.dom
        if self.species:
            L.vpb_clear_accumulators(dom, self.a.ptr)                   # advance.cxx:38
        for sp in self.species:                                         # :43-51
            if sp.sort_interval > 0 and self.step % sp.sort_interval == 0:
                self.sort(sp)
        for sp in self.species:                                         # :70-73
            L.vpb_advance_p_ordered(dom, sp.p.ptr, sp.np, sp.q_m, sp.pm.ptr, sp.max_nm, self.a.ptr, self.fi.ptr, sp.nm.ptr,
                                    None if sp.partition is None else sp.partition.ptr)
        # reduce_accumulators (:74) is a no-op with one replica; boundary_p (:94-96): see Simulation.migrate
        self.migrate()
        L.vpb_clear_jf(dom, self.f.ptr)                                 # :109
        if self.species:
            L.vpb_unload_accumulator(dom, self.f.ptr, self.a.ptr)       # :110
        L.vpb_synchronize_jf(dom, self.f.ptr)                           # :112
        self.advance_fields()
        if self.clean_div_e_interval and self.step % self.clean_div_e_interval == 0:
            self.clean_div_e()
        if self.clean_div_b_interval and self.step % self.clean_div_b_interval == 0:
            self.clean_div_b()
        if self.species:
            L.vpb_load_interpolator(dom, self.fi.ptr, self.f.ptr)       # :214
        self.step += 1

    NUM_COMM_ROUND = 3     # vpic.cxx:17

    def _needs_boundary_p(self):
        """advance_p can only leave movers where a face is shared with another rank or absorbs particles."""
        if getattr(self, "_nbp", None) is None:
            g = self.grid
            remote = any(0 <= g.struct.bc[b] < g.nproc and g.struct.bc[b] != g.rank
                         for b in (abi.boundary(-1, 0, 0), abi.boundary(1, 0, 0), abi.boundary(0, -1, 0), abi.boundary(0, 1, 0),
                                   abi.boundary(0, 0, -1), abi.boundary(0, 0, 1)))
            absorbing = bool(np.any((g.neighbor < 0) & (g.neighbor != abi.REFLECT_PARTICLES)))
            self._nbp = remote or absorbing
        return self._nbp

    def migrate(self):
        """boundary_p x num_comm_round (advance.cxx:94-103)."""
        if not self.species or not self._needs_boundary_p():
            return
        L = self.L
        st = (abi.SpeciesState * len(self.species))()
        nms = self.mover_counts()                     # one small read-back (synchronises)
        for k, sp in enumerate(self.species):
            st[k].p, st[k].pm, st[k].np, st[k].max_np = sp.p.ptr, sp.pm.ptr, sp.np, sp.max_np
            st[k].nm, st[k].max_nm, st[k].id = nms[k], sp.max_nm, sp.id
        for _ in range(self.NUM_COMM_ROUND):
            L.vpb_boundary_p(self.dom, st, len(self.species), self.f.ptr, self.a.ptr)
        for k, sp in enumerate(self.species):
            sp.np = st[k].np
            if st[k].nm:
                print("Warning: ignoring %d unprocessed %s movers (increase num_comm_round)" % (st[k].nm, sp.name))

    def clean_div_e(self):                                               # advance.cxx:151-173
        L, dom, f = self.L, self.dom, self.f.ptr
        L.vpb_clear_rhof(dom, f)
        for sp in self.species:
            L.vpb_accumulate_rho_p(dom, f, sp.p.ptr, sp.np)
        L.vpb_synchronize_rho(dom, f)
        for _ in range(2):
            L.vpb_compute_div_e_err(dom, f, self.m_ptr, self.n_mat)
            L.vpb_clean_div_e(dom, f, self.m_ptr, self.n_mat)

    def clean_div_b(self):                                               # advance.cxx:177-195
        L, dom, f = self.L, self.dom, self.f.ptr
        for _ in range(2):
            L.vpb_compute_div_b_err(dom, f)
            L.vpb_clean_div_b(dom, f)

    def energies(self):
        """dump_energies (src/vpic/dump.cxx:37-78): 6 field energies then one kinetic energy per species."""
        L, dom = self.L, self.dom
        s = self.scalars
        L.vpb_energy_f(dom, self.f.ptr, self.m_ptr, self.n_mat, s.ptr)
        L.vpb_comm_allsum_d(s.ptr, 6)
        g = self.grid.struct
        out = list(s.download(6) * (0.5 * g.eps0 * g.dx * g.dy * g.dz))
        for sp in self.species:
            L.vpb_energy_p(dom, sp.p.ptr, sp.np, sp.q_m, self.fi.ptr, s.ptr)
            L.vpb_comm_allsum_d(s.ptr, 1)
            out.append(float(s.download(1)[0]) * g.cvac * g.cvac / sp.q_m)
        return out

    def hydro(self, sp, synchronize=True):
        """The 14 hydro moments of one species on the mesh nodes, as the dump path computes them
        (dump.cxx: clear_hydro, accumulate_hydro_p, synchronize_hydro): hydro_t[nv] on the host."""
        L, dom = self.L, self.dom
        if getattr(self, "_hydro", None) is None:
            self._hydro = DevArray(L, self.nv, abi.hydro_dtype)
        h = self._hydro
        L.vpb_clear_hydro(dom, h.ptr)
        L.vpb_accumulate_hydro_p(dom, h.ptr, sp.p.ptr, sp.np, sp.q_m, self.fi.ptr)
        if synchronize:
            L.vpb_synchronize_hydro(dom, h.ptr)
        return h.download()

    def mover_counts(self):
        return [int(sp.nm.download(1)[0]) for sp in self.species]


class NativeSimulation:
    """The same run through the library's own C++ time-step driver (csrc/vpb_step.cu, vpb_sim_*): Python only holds
    the handle.  Same constructor and the same handful of methods as Simulation, so callers can switch."""

    class _Sp:
        def __init__(self, owner, idx, name, q_m, max_np, max_nm, sort_interval):
            self.owner, self.id, self.name, self.q_m = owner, idx, name, float(q_m)
            self.max_np, self.max_nm, self.sort_interval = int(max_np), max_nm, int(sort_interval)

        @property
        def np(self):
            return int(self.owner.L.vpb_sim_np(self.owner.h, self.id))

    def __init__(self, grid, n_mat=1, vacuum=False, L=None, planar=True, wide_interpolator=True, particle_planes=True):
        self.L = L or lib.load()
        self.L.vpb_init(-1)
        self.grid = grid
        self.nv = grid.nv
        self.h = self.L.vpb_sim_create(grid.ref(), grid.rank, grid.nproc, n_mat, int(vacuum), int(planar), int(wide_interpolator),
                                       int(particle_planes))
        self.species = []

    def free(self):
        if self.h:
            self.L.vpb_sim_destroy(self.h)
            self.h = None

    def define_species(self, name, q_m, max_np, max_nm=None, sort_interval=20):
        idx = self.L.vpb_sim_define_species(self.h, name.encode(), q_m, int(max_np), int(max_nm or 0), int(sort_interval))
        sp = NativeSimulation._Sp(self, idx, name, q_m, max_np, max_nm, sort_interval)
        self.species.append(sp)
        return sp

    def load_thermal(self, sp, ppc, vth, q, seed, tag0=0):
        self.L.vpb_sim_load_thermal(self.h, sp.id, ppc, vth, q, seed, tag0)

    def set_particles(self, sp, host):
        host = np.ascontiguousarray(host, dtype=abi.particle_dtype)
        self.L.vpb_sim_set_particles(self.h, sp.id, host.ctypes.data, len(host))

    def get_particles(self, sp):
        out = abi.aligned_empty(sp.np, abi.particle_dtype)
        n = self.L.vpb_sim_get_particles(self.h, sp.id, out.ctypes.data, len(out))
        return out[:n]

    def set_fields(self, host):
        host = np.ascontiguousarray(host, dtype=abi.field_dtype)
        assert len(host) == self.nv
        self.L.vpb_sim_set_fields(self.h, host.ctypes.data)

    def get_fields(self):
        out = abi.aligned_empty(self.nv, abi.field_dtype)
        self.L.vpb_sim_get_fields(self.h, out.ctypes.data)
        return out

    def set_intervals(self, clean_div_e=0, clean_div_b=0, num_comm_round=3):
        self.L.vpb_sim_set_intervals(self.h, clean_div_e, clean_div_b, num_comm_round)

    def set_sort_lookahead(self, steps):
        """< 0: 0.6 x each species' sort interval; 0: off (the reference's sort key)"""
        self.L.vpb_sim_set_sort_lookahead(self.h, steps)

    def set_callbacks(self, particle_collisions=None, particle_injection=None, current_injection=None, field_injection=None,
                      diagnostics=None):
        """The deck's hooks at the points advance.cxx calls them (vpb_sim_set_callbacks); each is called as fn(self)."""
        CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)

        class Callbacks(C.Structure):
            _fields_ = [(n, CB) for n in ("particle_collisions", "particle_injection", "current_injection", "field_injection",
                                          "diagnostics")] + [("user", C.c_void_p)]

        fns = (particle_collisions, particle_injection, current_injection, field_injection, diagnostics)
        self._cb_keep = [CB((lambda f: (lambda user, s: f(self)))(f)) if f else CB() for f in fns]
        self._cb_struct = Callbacks(*self._cb_keep, None)
        self.L.vpb_sim_set_callbacks(self.h, C.byref(self._cb_struct))

    def advance(self, nsteps=1):
        self.L.vpb_sim_advance(self.h, nsteps)

    @property
    def step(self):
        return int(self.L.vpb_sim_step(self.h))

    @property
    def dom(self):
        return self.L.vpb_sim_domain(self.h)

    @property
    def field_ptr(self):
        """device field array in the domain's layout (for vpb_* calls on it)"""
        return self.L.vpb_sim_field_array(self.h)

    def energies(self):
        out = np.zeros(6 + len(self.species))
        self.L.vpb_sim_energies(self.h, out.ctypes.data)
        return list(out)

    def hydro(self, sp):
        out = abi.aligned_zeros(self.nv, abi.hydro_dtype)
        self.L.vpb_sim_hydro(self.h, sp.id, out.ctypes.data)
        return out
