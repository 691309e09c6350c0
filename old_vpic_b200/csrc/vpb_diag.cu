// vpb_diag.cu -- the particle diagnostics the reference's trecon decks compute in HOST loops over sp->p, as device
// kernels: with the species arrays resident on the GPU a host loop drags every particle across PCIe at
// each dump (and, with managed arrays, back again at the next advance_p); here only the result crosses.
//
//  * vpb_energy_spectrum      decks/trecon-part/energy.cxx:90-176: per-cell kinetic-energy histogram dist[k*nv + voxel]
//                             (nex linear bands of width dke, the last one open-ended), normalised per cell, ghost cells
//                             copied from their interior neighbour; and the global log-spaced spectrum edist[nbin].
//  * vpb_tracer_records       decks/trecon-part/tracer.cxx:125-160 (dump_tracers): thirteen floats per particle --
//                             q (the tracer's tag), global x/y/z, momentum, and E and cB of a voxel.
// Same arithmetic as the host loops (double where they use double); counts are float increments as in the reference
// (exact below 2^24 per bin).  Layer-A wrappers (host / managed arrays) at the end.
#include <math.h>
#include "vpb_common.cuh"
#include "vpb_pview.cuh"

namespace vpb {

// energy.cxx:104-118
__global__ void __launch_bounds__(256) spectrum_count_kernel(const PView p, int np, double dke, int nex, float *__restrict__ dist, long nv,
                                                             double log_eminp, double dloge, int nbin, float *__restrict__ edist) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) {
    const float4 u = p.mom(i);
    const double gam2 = 1.0 + (double)(u.x * u.x) + (double)(u.y * u.y) + (double)(u.z * u.z);   // float products, double sum
    const double ke = sqrt(gam2) - 1.0;
    if (dist) {
      int k = (int)(ke / dke);
      if (k > nex - 1) k = nex - 1;
      atomicAdd(dist + (size_t)k * (size_t)nv + p.voxel(i), 1.0f);
    }
    if (edist) {
      // k = (log10(ke)-log10(eminp))/dloge + 1, converted to int (truncation toward zero); ke == 0 gives -inf -> out of range
      const double kk = (log10(ke) - log_eminp) / dloge + 1.0;
      if (kk > -2147483648.0 && kk < 2147483647.0) {
        const int k = (int)kk;
        if (k <= nbin - 1 && k >= 0) atomicAdd(edist + k, 1.0f);
      }
    }
  }
}

// energy.cxx:124-166.  ONE sweep over all voxels in index order normalises a cell (np = sum_k dist in double, dist /= np)
// and, when the cell is a ghost, overwrites it with the values of its interior neighbour AS THEY ARE AT THAT MOMENT: a
// neighbour with a higher index has not been normalised yet, so low-side ghosts (x = 0, y = 0, z = 0 faces) receive raw
// counts and high-side ghosts receive fractions.  Reproduced as three passes: raw copies, normalisation of the interior,
// normalised copies.  (Ghost cells hold no particles, so their own normalisation is a no-op.)
__device__ __forceinline__ bool ghost_neighbor(const DomainDev &g, long v, long &nid) {
  const int ix = (int)(v % g.sx), iy = (int)((v / g.sx) % g.sy), iz = (int)(v / g.sxy);
  const bool ghost = ix == 0 || ix == g.sx - 1 || iy == 0 || iy == g.sy - 1 || iz == 0 || iz == g.sz - 1;
  const int xn = ix == 0 ? 1 : (ix == g.sx - 1 ? ix - 1 : ix), yn = iy == 0 ? 1 : (iy == g.sy - 1 ? iy - 1 : iy),
            zn = iz == 0 ? 1 : (iz == g.sz - 1 ? iz - 1 : iz);
  nid = xn + (long)g.sx * (yn + (long)g.sy * zn);
  return ghost;
}
__global__ void __launch_bounds__(256) spectrum_normalize_kernel(float *__restrict__ dist, const DomainDev g, int nex) {
  const long nv = g.nv;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (long)gridDim.x * blockDim.x) {
    long nid;
    if (ghost_neighbor(g, v, nid)) continue;
    double np = 0;
    for (int k = 0; k < nex; k++) np += dist[(size_t)k * nv + v];
    if (np > 0)
      for (int k = 0; k < nex; k++) dist[(size_t)k * nv + v] = (float)(dist[(size_t)k * nv + v] / np);
  }
}
// later = 1: ghosts whose neighbour comes later in the sweep (copied before the normalisation); 0: the others (after)
__global__ void __launch_bounds__(256) spectrum_ghost_kernel(float *__restrict__ dist, const DomainDev g, int nex, int later) {
  const long nv = g.nv;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (long)gridDim.x * blockDim.x) {
    long nid;
    if (!ghost_neighbor(g, v, nid) || (nid > v) != (later != 0)) continue;
    for (int k = 0; k < nex; k++) dist[(size_t)k * nv + v] = dist[(size_t)k * nv + nid];
  }
}

// tracer.cxx:117-119,140-157.  The macro reads the fields of `field[p->i]` with p the HEAD of the species array, i.e.
// the voxel of particle 0 for every record (`p[j].i` was surely meant); field_of_first = 1 reproduces the reference's
// files, 0 uses each particle's own voxel.
__global__ void __launch_bounds__(256) tracer_records_kernel(const PView p, int np, const vpb_field_t *__restrict__ f, const DomainDev g,
                                                             float x0, float y0, float z0, int field_of_first, float *__restrict__ out) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < np; j += gridDim.x * blockDim.x) {
    const float4 r = p.pos(j), u = p.mom(j);
    const int v = __float_as_int(r.w);
    const int ix = v % g.sx, iy = (v / g.sx) % g.sy, iz = v / g.sxy;
    float *o = out + 13 * (size_t)j;
    o[0] = u.w;
    // ( i%(nx+2) + (dx-1)/2.0 ) * grid->dx + grid->x0 : int + double, times float, plus float -> double, then (float)
    o[1] = (float)(((double)ix + ((double)(r.x - 1.f)) / 2.0) * (double)g.dx + (double)x0);
    o[2] = (float)(((double)iy + ((double)(r.y - 1.f)) / 2.0) * (double)g.dy + (double)y0);
    o[3] = (float)(((double)iz + ((double)(r.z - 1.f)) / 2.0) * (double)g.dz + (double)z0);
    o[4] = u.x; o[5] = u.y; o[6] = u.z;
    const int fv = field_of_first ? p.voxel(0) : v;
    const float4 e = *CFQ(f, g, fv, 0), b = *CFQ(f, g, fv, 1);
    o[7] = e.x; o[8] = e.y; o[9] = e.z; o[10] = b.x; o[11] = b.y; o[12] = b.z;
  }
}

static int diag_grid(long n) {
  long b = (n + 255) / 256;
  const long cap = (long)ctx().sm_count * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace vpb

using namespace vpb;

extern "C" {

// d_dist: float[nex * nvoxel] or NULL; d_edist: float[nbin] or NULL (device pointers; both are overwritten)
void vpb_energy_spectrum(vpb_domain_t *dom, const vpb_particle_t *d_p, int np, double dke, int nex, float *d_dist, double eminp,
                         double emaxp, int nbin, float *d_edist) {
  if (!dom) VPB_ERROR("Bad grid");
  if (np < 0) VPB_ERROR("Bad number of particles");
  if (np && !d_p) VPB_ERROR("Bad particle array");
  if (d_dist && (nex < 1 || !(dke > 0))) VPB_ERROR("Bad energy bands");
  if (d_edist && (nbin < 1 || !(eminp > 0) || !(emaxp > eminp))) VPB_ERROR("Bad spectrum bins");
  Context &c = ctx();
  const DomainDev &g = dom->d;
  if (d_dist) VPB_CUDA(cudaMemsetAsync(d_dist, 0, (size_t)nex * g.nv * sizeof(float), c.stream));
  if (d_edist) VPB_CUDA(cudaMemsetAsync(d_edist, 0, (size_t)nbin * sizeof(float), c.stream));
  // energy.cxx:36-38,60: float eminp, emaxp, dloge; log10 of a float is the float function there
  const float log_eminp = log10f((float)eminp);
  const float dloge = d_edist ? (log10f((float)emaxp) - log_eminp) / nbin : 1.f;
  if (np > 0 && (d_dist || d_edist)) {
    spectrum_count_kernel<<<diag_grid(np), 256, 0, c.stream>>>(PView(d_p, g.p_plane), np, dke, nex, d_dist, g.nv, (double)log_eminp,
                                                               (double)dloge, nbin, d_edist);
    count_launch();
  }
  if (d_dist) {
    spectrum_ghost_kernel<<<diag_grid(g.nv), 256, 0, c.stream>>>(d_dist, g, nex, 1);
    spectrum_normalize_kernel<<<diag_grid(g.nv), 256, 0, c.stream>>>(d_dist, g, nex);
    spectrum_ghost_kernel<<<diag_grid(g.nv), 256, 0, c.stream>>>(d_dist, g, nex, 0);
    count_launch(3);
  }
  VPB_CUDA(cudaGetLastError());
}

// d_out: float[13 * np] (device).  x0,y0,z0: grid->x0,y0,z0 (the local domain's corner).
void vpb_tracer_records(vpb_domain_t *dom, const vpb_particle_t *d_p, int np, const vpb_field_t *d_f, float x0, float y0, float z0,
                        int field_of_first, float *d_out) {
  if (!dom) VPB_ERROR("Bad grid");
  if (np < 0) VPB_ERROR("Bad number of particles");
  if (np == 0) return;
  if (!d_p || !d_f || !d_out) VPB_ERROR("Bad args");
  tracer_records_kernel<<<diag_grid(np), 256, 0, ctx().stream>>>(PView(d_p, dom->d.p_plane), np, d_f, dom->d, x0, y0, z0, field_of_first, d_out);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

}  // extern "C"
