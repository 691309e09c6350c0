"""Host-side construction of a grid_t for one rank of a Cartesian decomposition.

Mirrors the reference's setup API for this path -- size_grid / join_grid /
set_fbc / set_pbc (src/grid/ops.c:25-231) and partition_{periodic,absorbing,
metal}_box (src/grid/partition.c:36-260) -- with the same names, argument
meaning and results, so a grid built here is byte-compatible with one built by
the reference (tests/test_grid.py compares bc[], range[] and neighbor[]).
The arrays are numpy-owned; `Grid.struct` is the 240-byte grid_t handed to the
C ABI.
"""
import ctypes as C

import numpy as np

from . import abi


def _rank_to_index(rank, gpx, gpy, gpz):
    iy, ix = divmod(rank, gpx)          # partition.c:13-23
    iz, iy = divmod(iy, gpy)
    return ix, iy, iz


def _index_to_rank(ix, iy, iz, gpx, gpy, gpz):
    return (ix % gpx) + gpx * ((iy % gpy) + gpy * (iz % gpz))  # partition.c:25-33 (periodic wrap)


class Grid:
    """One rank's grid_t.  rank/nproc stand in for the reference's mp handle."""

    def __init__(self, rank=0, nproc=1):
        self.rank, self.nproc = rank, nproc
        self.struct = abi.GridStruct()
        for i in range(27):
            self.struct.bc[i] = abi.PEC_FIELDS      # grid_structors.c:21
        self.struct.bc[abi.boundary(0, 0, 0)] = rank
        self.range = None
        self.neighbor = None
        self.field_only = False     # True: no neighbor[] (6 int64 per voxel); particles cannot move on such a grid

    # convenience -----------------------------------------------------------
    @property
    def n(self):
        return self.struct.nx, self.struct.ny, self.struct.nz

    @property
    def shape(self):
        """numpy shape (z,y,x) of a voxel array, ghosts included."""
        return self.struct.nz + 2, self.struct.ny + 2, self.struct.nx + 2

    @property
    def nv(self):
        nx, ny, nz = self.n
        return (nx + 2) * (ny + 2) * (nz + 2)

    def ref(self):
        return C.byref(self.struct)

    def voxel(self, x, y, z):
        nx, ny, _ = self.n
        return x + (nx + 2) * (y + (ny + 2) * z)

    def set_units(self, dt, cvac=1.0, eps0=1.0, damp=0.0):
        s = self.struct
        s.dt, s.cvac, s.eps0, s.damp = dt, cvac, eps0, damp

    # ops.c:25-101 ------------------------------------------------------------
    def size_grid(self, lnx, lny, lnz, all_local_sizes=None):
        if lnx < 1 or lny < 1 or lnz < 1:
            raise ValueError("Bad local grid size")
        s = self.struct
        s.nx, s.ny, s.nz = lnx, lny, lnz
        for i in range(27):
            s.bc[i] = abi.PEC_FIELDS
        s.bc[abi.boundary(0, 0, 0)] = self.rank
        lnc = (lnx + 2) * (lny + 2) * (lnz + 2)
        # every rank's voxel count (the reference allgathers it; uniform decompositions pass None)
        counts = [lnc] * self.nproc if all_local_sizes is None else list(all_local_sizes)
        self.range = np.zeros(self.nproc + 1, dtype=np.int64)
        self.range[1:] = np.cumsum(counts)
        s.range = self.range.ctypes.data
        s.rangel = int(self.range[self.rank])
        s.rangeh = int(self.range[self.rank + 1]) - 1
        if self.field_only:
            s.neighbor = None
            return
        sx, sy, sz = lnx + 2, lny + 2, lnz + 2
        z, y, x = np.meshgrid(np.arange(sz), np.arange(sy), np.arange(sx), indexing="ij")
        lid = (x + sx * (y + sy * z)).astype(np.int64)
        nb = np.empty((sz, sy, sx, 6), dtype=np.int64)
        offs = (-1, -sx, -sx * sy, 1, sx, sx * sy)
        for f, o in enumerate(offs):
            nb[..., f] = s.rangel + lid + o
        refl = abi.REFLECT_PARTICLES
        nb[:, :, 1, 0] = refl
        nb[:, 1, :, 1] = refl
        nb[1, :, :, 2] = refl
        nb[:, :, lnx, 3] = refl
        nb[:, lny, :, 4] = refl
        nb[lnz, :, :, 5] = refl
        ghost = (x == 0) | (x == lnx + 1) | (y == 0) | (y == lny + 1) | (z == 0) | (z == lnz + 1)
        nb[ghost] = refl
        self.neighbor = abi.aligned_empty(6 * lnc, np.int64)
        self.neighbor[:] = nb.reshape(-1)
        s.neighbor = self.neighbor.ctypes.data

    def _face_cells(self, face):
        """(z,y,x) index arrays of the interior cells touching `face` (0..5 = -x,-y,-z,+x,+y,+z)."""
        nx, ny, nz = self.n
        rng = [np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1)]
        ax = face % 3
        rng[ax] = np.array([1 if face < 3 else (nx, ny, nz)[ax]])
        z, y, x = np.meshgrid(rng[2], rng[1], rng[0], indexing="ij")
        return z, y, x

    # ops.c:135-182 -----------------------------------------------------------
    def join_grid(self, bound, rank):
        face = {abi.boundary(-1, 0, 0): 0, abi.boundary(0, -1, 0): 1, abi.boundary(0, 0, -1): 2,
                abi.boundary(1, 0, 0): 3, abi.boundary(0, 1, 0): 4, abi.boundary(0, 0, 1): 5}.get(bound)
        if face is None:
            raise ValueError("Bad boundary")
        if not 0 <= rank < self.nproc:
            raise ValueError("Bad rank")
        s = self.struct
        s.bc[bound] = rank
        if self.field_only:
            return
        ln = list(self.n)
        ax = face % 3
        Y, Z = (ax + 1) % 3, (ax + 2) % 3
        rnc = int(self.range[rank + 1] - self.range[rank])
        if rnc % ((ln[Y] + 2) * (ln[Z] + 2)) != 0:
            raise ValueError("Remote face is incompatible")
        rn = list(ln)
        rn[ax] = rnc // ((ln[Y] + 2) * (ln[Z] + 2)) - 2
        z, y, x = self._face_cells(face)
        l = [x, y, z]
        r = [x.copy(), y.copy(), z.copy()]
        r[ax] = np.full_like(x, rn[ax] if face < 3 else 1)
        rid = r[0] + (rn[0] + 2) * (r[1] + (rn[1] + 2) * r[2])
        lid = l[0] + (ln[0] + 2) * (l[1] + (ln[1] + 2) * l[2])
        self.neighbor[6 * lid.reshape(-1) + face] = int(self.range[rank]) + rid.reshape(-1)

    # ops.c:184-198 -----------------------------------------------------------
    def set_fbc(self, bound, fbc):
        if 0 <= fbc < self.nproc:
            raise ValueError("Use join_grid")
        if fbc not in (abi.PEC_FIELDS, abi.SYMMETRIC_FIELDS, abi.PMC_FIELDS, abi.ABSORB_FIELDS):
            raise ValueError("Bad field bc")
        self.struct.bc[bound] = fbc

    # ops.c:200-231 -----------------------------------------------------------
    def set_pbc(self, bound, pbc):
        face = {abi.boundary(-1, 0, 0): 0, abi.boundary(0, -1, 0): 1, abi.boundary(0, 0, -1): 2,
                abi.boundary(1, 0, 0): 3, abi.boundary(0, 1, 0): 4, abi.boundary(0, 0, 1): 5}.get(bound)
        if face is None:
            raise ValueError("Bad boundary")
        if pbc not in (abi.ABSORB_PARTICLES, abi.REFLECT_PARTICLES):
            raise ValueError("Bad particle bc")
        if self.field_only:
            return
        z, y, x = self._face_cells(face)
        nx, ny, _ = self.n
        lid = x + (nx + 2) * (y + (ny + 2) * z)
        self.neighbor[6 * lid.reshape(-1) + face] = pbc


def partition_periodic_box(g, gx0, gy0, gz0, gx1, gy1, gz1, gnx, gny, gnz, gpx, gpy, gpz):
    """partition.c:36-85."""
    if gpx < 1 or gpy < 1 or gpz < 1 or gpx * gpy * gpz != g.nproc:
        raise ValueError("Bad topology")
    if gnx < 1 or gny < 1 or gnz < 1:
        raise ValueError("Bad res")
    if gnx % gpx or gny % gpy or gnz % gpz:
        raise ValueError("Incompatible res")
    px, py, pz = _rank_to_index(g.rank, gpx, gpy, gpz)
    s = g.struct
    s.dx, s.dy, s.dz = (gx1 - gx0) / gnx, (gy1 - gy0) / gny, (gz1 - gz0) / gnz
    s.rdx, s.rdy, s.rdz = gnx / (gx1 - gx0), gny / (gy1 - gy0), gnz / (gz1 - gz0)

    def lerp(a, b, i, n):
        f = float(i) / float(n)
        return a * (1 - f) + b * f

    s.x0, s.y0, s.z0 = lerp(gx0, gx1, px, gpx), lerp(gy0, gy1, py, gpy), lerp(gz0, gz1, pz, gpz)
    s.x1, s.y1, s.z1 = lerp(gx0, gx1, px + 1, gpx), lerp(gy0, gy1, py + 1, gpy), lerp(gz0, gz1, pz + 1, gpz)
    g.size_grid(gnx // gpx, gny // gpy, gnz // gpz)
    for (i, j, k) in ((-1, 0, 0), (0, -1, 0), (0, 0, -1), (1, 0, 0), (0, 1, 0), (0, 0, 1)):
        g.join_grid(abi.boundary(i, j, k), _index_to_rank(px + i, py + j, pz + k, gpx, gpy, gpz))
    g.topology = (gpx, gpy, gpz)
    g.coords = (px, py, pz)


def _outer_faces(g, gn, gp):
    px, py, pz = g.coords
    p = (px, py, pz)
    for ax in range(3):
        if gn[ax] <= 1:
            continue
        ijk = [0, 0, 0]
        if p[ax] == 0:
            ijk[ax] = -1
            yield abi.boundary(*ijk)
        if p[ax] == gp[ax] - 1:
            ijk[ax] = 1
            yield abi.boundary(*ijk)


def partition_absorbing_box(g, gx0, gy0, gz0, gx1, gy1, gz1, gnx, gny, gnz, gpx, gpy, gpz, pbc):
    """partition.c:88-120."""
    partition_periodic_box(g, gx0, gy0, gz0, gx1, gy1, gz1, gnx, gny, gnz, gpx, gpy, gpz)
    for b in _outer_faces(g, (gnx, gny, gnz), (gpx, gpy, gpz)):
        g.set_fbc(b, abi.ABSORB_FIELDS)
        g.set_pbc(b, pbc)


def partition_metal_box(g, gx0, gy0, gz0, gx1, gy1, gz1, gnx, gny, gnz, gpx, gpy, gpz):
    """partition.c:190-231."""
    partition_periodic_box(g, gx0, gy0, gz0, gx1, gy1, gz1, gnx, gny, gnz, gpx, gpy, gpz)
    for b in _outer_faces(g, (gnx, gny, gnz), (gpx, gpy, gpz)):
        g.set_fbc(b, abi.PEC_FIELDS)
        g.set_pbc(b, abi.REFLECT_PARTICLES)


def courant_dt(dx, dy, dz, cvac=1.0, frac=0.95):
    """dt = frac * courant_length / c, the decks' usual choice (vpic.hxx courant_length helper)."""
    inv = sum(1.0 / (d * d) for d in (dx, dy, dz) if d > 0)
    return frac / (cvac * np.sqrt(inv))


def make_grid(n, kind="periodic", topo=(1, 1, 1), rank=0, L=None, dt=None, damp=0.0, pbc=abi.ABSORB_PARTICLES,
              field_only=False):
    """One rank's grid_t for a global box of n=(gnx,gny,gnz) cells (cell size 1 unless L is given),
    cvac=eps0=1 and dt=0.95 Courant unless given."""
    nx, ny, nz = n
    L = L or (float(nx), float(ny), float(nz))
    g = Grid(rank=rank, nproc=topo[0] * topo[1] * topo[2])
    g.field_only = field_only
    args = (g, 0.0, 0.0, 0.0, L[0], L[1], L[2], nx, ny, nz, topo[0], topo[1], topo[2])
    if kind == "periodic":
        partition_periodic_box(*args)
    elif kind == "metal":
        partition_metal_box(*args)
    elif kind == "absorbing":
        partition_absorbing_box(*args, pbc)
    else:
        raise ValueError(kind)
    s = g.struct
    dims = [d for d, m in ((s.dx, nx), (s.dy, ny), (s.dz, nz)) if m > 1] or [s.dx]
    g.set_units(dt if dt is not None else courant_dt(*(dims + [0, 0])[:3]), 1.0, 1.0, damp)
    return g


def interior_voxels(g):
    """Local voxel index of every interior cell, x fastest."""
    nx, ny, nz = g.n
    z, y, x = np.meshgrid(np.arange(1, nz + 1), np.arange(1, ny + 1), np.arange(1, nx + 1), indexing="ij")
    return (x + (nx + 2) * (y + (ny + 2) * z)).reshape(-1).astype(np.int32)
