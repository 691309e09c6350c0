"""Python handle on the library's device-resident time-step driver (csrc/vpb_step.cu, vpb_sim_*: the call order of
vpic_simulation::advance(), src/vpic/advance.cxx:13-244, with every array resident in HBM between steps) and typed
device arrays for tests that call layer (B) directly.  There is one driver, and it is the C++ one.
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


class NativeSimulation:
    """One rank's share of a PIC run on one GPU, through the library's C++ time-step driver (csrc/vpb_step.cu,
    vpb_sim_*): Python only holds the handle."""

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

    def load_pairs_mt(self, sp_a, sp_b, n, lo, hi, vth_a, vth_b, q_a, q_b, seed=None, rng=None, args_right_to_left=1, tag0=0, tag_step=0):
        """The thermal deck's load loop (seed_rand(seed); n x {3 uniform_rand, 2 x 3 maxwellian_rand -> inject_particle})
        from the reference's own random-number stream, on the device (include/vpic_b200.h vpb_load_pairs_mt)."""
        own = rng is None
        if own:
            rng = self.L.vpb_mt_create(int(seed))
        lo, hi = np.ascontiguousarray(lo, np.float64), np.ascontiguousarray(hi, np.float64)
        done = self.L.vpb_sim_load_pairs_mt(self.h, rng, sp_a.id, sp_b.id, int(n), lo.ctypes.data, hi.ctypes.data, vth_a, vth_b, q_a, q_b,
                                            int(args_right_to_left), int(tag0), int(tag_step))
        if own:
            self.L.vpb_mt_destroy(rng)
        return done

    def initialize(self):
        """vpic_simulation::initialize() after the deck's own part (initialize.cxx:27-95); returns the synchronisation,
        div B and div E errors the reference prints there."""
        out = np.zeros(3)
        self.L.vpb_sim_initialize(self.h, out.ctypes.data)
        return out

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

    def set_intervals(self, clean_div_e=0, clean_div_b=0, num_comm_round=3, sync_shared=0):
        self.L.vpb_sim_set_intervals(self.h, clean_div_e, clean_div_b, num_comm_round)
        self.L.vpb_sim_set_sync_shared_interval(self.h, sync_shared)

    def last_errors(self):
        """rms div E error before cleaning pass 1 / 2, the same for div B, domain desynchronisation error
        (what advance.cxx:160,168,182,190,205 print), latest values"""
        out = np.zeros(5)
        self.L.vpb_sim_last_errors(self.h, out.ctypes.data)
        return out

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
