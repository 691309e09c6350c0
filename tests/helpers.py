"""Shared fixtures for the parity tests: seeded synthetic states in the reference's
own data layouts, grids built either by the reference (oracle/_ref) or by the
host mirror (old_vpic_b200.grid), and comparison utilities."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from old_vpic_b200 import abi  # noqa: E402
from old_vpic_b200 import grid as hostgrid  # noqa: E402
from oracle import loader  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


courant_dt = hostgrid.courant_dt
host_grid = hostgrid.make_grid
interior_voxels = hostgrid.interior_voxels


class RefGrid:
    """grid_t owned by the reference (new_grid + partition_*), viewed through GridStruct."""

    def __init__(self, L, n, kind="periodic", Lbox=None, dt=None, damp=0.0, pbc=abi.ABSORB_PARTICLES):
        nx, ny, nz = n
        Lbox = Lbox or (float(nx), float(ny), float(nz))
        self.ptr = L.new_grid()
        a = (self.ptr, 0.0, 0.0, 0.0, Lbox[0], Lbox[1], Lbox[2], nx, ny, nz, 1, 1, 1)
        if kind == "periodic":
            L.partition_periodic_box(*a)
        elif kind == "metal":
            L.partition_metal_box(*a)
        elif kind == "absorbing":
            L.partition_absorbing_box(*a, pbc)
        else:
            raise ValueError(kind)
        self.struct = abi.GridStruct.from_address(self.ptr)
        s = self.struct
        dims = [d for d, m in ((s.dx, nx), (s.dy, ny), (s.dz, nz)) if m > 1] or [s.dx]
        s.dt = dt if dt is not None else courant_dt(*(dims + [0, 0])[:3])
        s.cvac, s.eps0, s.damp = 1.0, 1.0, damp
        self.rank, self.nproc = 0, 1

    @property
    def n(self):
        return self.struct.nx, self.struct.ny, self.struct.nz

    @property
    def nv(self):
        nx, ny, nz = self.n
        return (nx + 2) * (ny + 2) * (nz + 2)

    @property
    def shape(self):
        return self.struct.nz + 2, self.struct.ny + 2, self.struct.nx + 2

    def ref(self):
        return C.c_void_p(self.ptr)

    @property
    def neighbor(self):
        return np.ctypeslib.as_array(C.cast(self.struct.neighbor, C.POINTER(C.c_int64)), shape=(6 * self.nv,))


def gptr(g):
    """void* of a grid_t for ctypes calls, whichever kind of grid object it is."""
    return g.ref()


def random_particles(rng, g, np_, vth=0.1, sort=True, q=-1.0, edge_frac=0.0):
    """np_ particles uniformly over the interior voxels, Maxwellian momenta."""
    p = abi.aligned_zeros(np_, abi.particle_dtype)
    vox = interior_voxels(g)
    i = rng.choice(vox, size=np_)
    if sort:
        i = np.sort(i)
    p["i"] = i
    for k in ("dx", "dy", "dz"):
        p[k] = rng.uniform(-1, 1, np_).astype(np.float32)
    if edge_frac > 0:  # park some particles exactly on a cell face (the reference's +-1 exactness cases)
        m = rng.random(np_) < edge_frac
        p["dx"][m] = np.where(rng.random(m.sum()) < 0.5, -1.0, 1.0)
    for k in ("ux", "uy", "uz"):
        p[k] = (vth * rng.standard_normal(np_)).astype(np.float32)
    p["q"] = np.float32(q)
    p["tag"] = np.arange(np_, dtype=np.int64)
    p["tag2"] = rng.integers(0, 2 ** 40, np_)
    return p


def random_interpolator(rng, g, amp=0.05):
    f = abi.aligned_zeros(g.nv, abi.interpolator_dtype)
    for n in abi.interpolator_dtype.names:
        if n == "_pad":
            continue
        scale = amp * (0.2 if n.startswith("d") else 1.0)
        f[n] = (scale * rng.standard_normal(g.nv)).astype(np.float32)
    return f


def random_fields(rng, g, amp=0.1, n_mat=1):
    f = abi.aligned_zeros(g.nv, abi.field_dtype)
    for n in abi.FIELD_FLOATS:
        f[n] = (amp * rng.standard_normal(g.nv)).astype(np.float32)
    for n in ("ematx", "ematy", "ematz", "nmat", "fmatx", "fmaty", "fmatz", "cmat"):
        f[n] = rng.integers(0, n_mat, g.nv).astype(np.uint16) if n_mat > 1 else 0
    return f


def vacuum_coefficients(n_mat=1, rng=None):
    """material_coefficient_t[]: entry 0 is vacuum (sfa.c:127-168 with eps=mu=1, sigma=0); further
    entries are random passive materials for exercising the per-voxel lookups."""
    m = abi.aligned_zeros(n_mat, abi.material_coefficient_dtype)
    for k in ("decayx", "drivex", "decayy", "drivey", "decayz", "drivez", "rmux", "rmuy", "rmuz", "nonconductive",
              "epsx", "epsy", "epsz"):
        m[k] = 1.0
    if n_mat > 1:
        rng = rng or np.random.default_rng(0)
        for k in ("decayx", "decayy", "decayz"):
            m[k][1:] = rng.uniform(0.5, 1.0, n_mat - 1)
        for k in ("drivex", "drivey", "drivez", "rmux", "rmuy", "rmuz"):
            m[k][1:] = rng.uniform(0.3, 1.0, n_mat - 1)
        for k in ("epsx", "epsy", "epsz"):
            m[k][1:] = rng.uniform(1.0, 3.0, n_mat - 1)
        m["nonconductive"][1:] = rng.integers(0, 2, n_mat - 1)
    return m


def assert_bits_equal(a, b, what=""):
    """Bit-exact comparison of two structured or plain arrays (NaN payloads included)."""
    a8, b8 = np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8)
    if a8.shape != b8.shape or not np.array_equal(a8, b8):
        bad = np.nonzero(a8.reshape(-1) != b8.reshape(-1))[0] if a8.shape == b8.shape else []
        first = int(bad[0]) // a.dtype.itemsize if len(bad) else -1
        raise AssertionError("%s: %d bytes differ, first differing element %d: %r vs %r" % (
            what, len(bad), first, a.reshape(-1)[first] if first >= 0 else None, b.reshape(-1)[first] if first >= 0 else None))


def hot(p):
    """The 32 hot bytes of a particle array as an (n,8) uint32 view (tags excluded)."""
    return np.ascontiguousarray(p).view(np.uint32).reshape(-1, 12)[:, :8]


def max_rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


class MaterialStruct(C.Structure):
    """material_t (src/material/material.h:42-50) with room for a short name."""
    pass


MaterialStruct._fields_ = [("id", C.c_uint16)] + [(n, C.c_float) for n in (
    "epsx", "epsy", "epsz", "mux", "muy", "muz", "sigmax", "sigmay", "sigmaz", "zetax", "zetay", "zetaz")] + \
    [("next", C.POINTER(MaterialStruct)), ("name", C.c_char * 24)]

# name, eps(3), mu(3), sigma(3), zeta(3): vacuum, a lossy anisotropic dielectric, a very good conductor
MATERIAL_TABLE = [("vacuum", (1, 1, 1), (1, 1, 1), (0, 0, 0), (0, 0, 0)),
                  ("glass", (2.5, 3.0, 2.0), (1.0, 1.5, 1.25), (0.3, 0.0, 0.05), (0, 0, 0)),
                  ("metal", (1, 1, 1), (1, 1, 1), (1e9, 1e9, 1e9), (0, 0, 0))]


def material_list():
    """A material list built the way new_material does it (newest first, ids counting up from 0).  Returns the
    head and the list of structs (keep it alive)."""
    keep, head = [], None
    for k, (name, eps, mu, sig, zeta) in enumerate(MATERIAL_TABLE):
        m = MaterialStruct()
        m.id = k
        (m.epsx, m.epsy, m.epsz), (m.mux, m.muy, m.muz) = eps, mu
        (m.sigmax, m.sigmay, m.sigmaz), (m.zetax, m.zetay, m.zetaz) = sig, zeta
        m.name = name.encode()
        if head is not None:
            m.next = C.pointer(head)
        keep.append(m)
        head = m
    return head, keep
