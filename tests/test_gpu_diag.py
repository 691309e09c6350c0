"""GPU: the deck-side particle diagnostics on the device (csrc/vpb_diag.cu) against numpy restatements
of the reference's host loops -- decks/trecon-part/energy.cxx:90-176 (per-cell energy bands + global log spectrum,
including the sweep-order quirk of its ghost-cell copies) and tracer.cxx:125-160 (dump_tracers records, including the
macro's `field[p->i]`)."""
import numpy as np
import pytest

from helpers import abi, host_grid, random_fields, random_particles
from old_vpic_b200.abi import ptr

pytestmark = pytest.mark.gpu


def spectrum_reference(p, g, dke, nex, eminp, emaxp, nbin):
    """energy.cxx:60,96-166 in numpy (float32 / float64 where the deck uses float / double)"""
    nv = g.nv
    sz, sy, sx = g.shape
    ux, uy, uz = (p[k].astype(np.float32) for k in ("ux", "uy", "uz"))
    gam2 = np.float64(1.0) + (ux * ux).astype(np.float64) + (uy * uy).astype(np.float64) + (uz * uz).astype(np.float64)
    ke = np.sqrt(gam2) - 1.0
    k = np.minimum((ke / dke).astype(np.int64), nex - 1)
    dist = np.zeros((nex, nv), np.float32)
    np.add.at(dist, (k, p["i"].astype(np.int64)), np.float32(1))
    log_emin = np.float64(np.log10(np.float32(eminp)))                       # log10 of a float: the float function
    dloge = np.float32((np.log10(np.float32(emaxp)) - np.log10(np.float32(eminp))) / np.float32(nbin))
    with np.errstate(divide="ignore"):
        kk = (np.log10(ke) - log_emin) / np.float64(dloge) + 1.0
    ok = np.isfinite(kk)
    kb = np.trunc(kk[ok]).astype(np.int64)
    kb = kb[(kb >= 0) & (kb <= nbin - 1)]
    edist = np.bincount(kb, minlength=nbin).astype(np.float32)
    # the single sweep: normalise, then (ghosts) copy the neighbour as it is at that moment
    d = dist.astype(np.float64).reshape(nex, sz, sy, sx)
    raw = d.copy()
    tot = raw.sum(axis=0)
    norm = np.where(tot > 0, raw / np.where(tot > 0, tot, 1), raw).astype(np.float32)
    out = norm.copy()
    iz, iy, ix = np.meshgrid(np.arange(sz), np.arange(sy), np.arange(sx), indexing="ij")
    ghost = (ix == 0) | (ix == sx - 1) | (iy == 0) | (iy == sy - 1) | (iz == 0) | (iz == sz - 1)
    xn, yn, zn = np.clip(ix, 1, sx - 2), np.clip(iy, 1, sy - 2), np.clip(iz, 1, sz - 2)
    v = ix + sx * (iy + sy * iz)
    nid = xn + sx * (yn + sy * zn)
    later = ghost & (nid > v)
    earlier = ghost & (nid < v)
    out[:, later] = raw.astype(np.float32)[:, zn[later], yn[later], xn[later]]
    out[:, earlier] = norm[:, zn[earlier], yn[earlier], xn[earlier]]
    return out.reshape(nex, nv), edist


@pytest.mark.parametrize("n,np_,vth", [((6, 5, 4), 20000, 0.6), ((12, 1, 9), 50001, 0.3)])
def test_energy_spectrum(vpb, n, np_, vth):
    g = host_grid(n)
    rng = np.random.default_rng(5)
    p = random_particles(rng, g, np_, vth=vth, sort=False)
    nex, nbin, eminp, emaxp = 7, 800, 0.0001, 10000.0
    dke = 3.0 * (vth * vth / 2.0) / nex                       # emax = 3 (in units of vth^2/2): the last band collects the tail
    want_dist, want_e = spectrum_reference(p, g, dke, nex, eminp, emaxp, nbin)
    dist = np.full((nex, g.nv), -1, np.float32)
    edist = np.full(nbin, -1, np.float32)
    vpb.vpb_deck_energy_spectrum(ptr(p), np_, dke, nex, dist.ctypes.data, eminp, emaxp, nbin, edist.ctypes.data, g.ref())
    assert edist.sum() == want_e.sum() and np.abs(edist - want_e).sum() <= 4       # a particle on a bin edge may fall either way
    assert np.array_equal(dist.sum(axis=0) > 0, want_dist.sum(axis=0) > 0)
    bad = np.abs(dist - want_dist) > 1e-6
    assert bad.sum() <= 8, int(bad.sum())
    # the quirk is reproduced: a low-side ghost holds raw counts, a high-side ghost fractions
    lo, hi = g.voxel(0, 2, 2) if n[1] > 1 else g.voxel(0, 1, 2), g.voxel(n[0] + 1, 2, 2) if n[1] > 1 else g.voxel(n[0] + 1, 1, 2)
    assert dist[:, lo].sum() >= 1.0 and abs(dist[:, hi].sum() - 1.0) < 1e-5
    # either output may be omitted
    e2 = np.zeros(nbin, np.float32)
    vpb.vpb_deck_energy_spectrum(ptr(p), np_, dke, nex, None, eminp, emaxp, nbin, e2.ctypes.data, g.ref())
    assert np.array_equal(e2, edist)


def test_tracer_records(vpb):
    g = host_grid((7, 6, 5))
    rng = np.random.default_rng(9)
    np_ = 3001
    p = random_particles(rng, g, np_, vth=0.4, sort=False)
    p["q"] = rng.integers(1, 1 << 20, np_).astype(np.float32)            # the tracer's tag lives in q (tracer.cxx:60-66)
    f = random_fields(rng, g)
    s = g.struct
    s.x0, s.y0, s.z0 = 0.25, -1.5, 3.0
    sx, sy = g.n[0] + 2, g.n[1] + 2
    v = p["i"].astype(np.int64)
    for first in (1, 0):
        out = np.zeros((np_, 13), np.float32)
        vpb.vpb_deck_tracer_records(ptr(p), np_, ptr(f), out.ctypes.data, first, g.ref())
        want = np.zeros((np_, 13), np.float32)
        want[:, 0] = p["q"]
        for c, (idx, d, h, o) in enumerate(((v % sx, p["dx"], s.dx, s.x0), ((v // sx) % sy, p["dy"], s.dy, s.y0),
                                            (v // (sx * sy), p["dz"], s.dz, s.z0))):
            # ( i%(nx+2) + (dx-1)/2.0 ) * grid->dx + grid->x0, tracer.cxx:117-119
            want[:, 1 + c] = ((idx.astype(np.float64) + (d - np.float32(1)).astype(np.float64) / 2.0) * np.float64(np.float32(h))
                              + np.float64(np.float32(o))).astype(np.float32)
        want[:, 4], want[:, 5], want[:, 6] = p["ux"], p["uy"], p["uz"]
        fv = np.full(np_, v[0]) if first else v
        for c, k in enumerate(("ex", "ey", "ez", "cbx", "cby", "cbz")):
            want[:, 7 + c] = f[k][fv]
        assert np.array_equal(out.view(np.uint32), want.view(np.uint32)), first
