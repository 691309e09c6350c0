// vpb_loader.cu -- device-side synthetic particle load for benchmarks and
// large-size property tests: `ppc` particles in every interior voxel (so the
// array is born voxel-sorted), positions uniform in the cell, momenta Maxwellian.
// The reference loads particles one by one on the host with a Mersenne Twister
// (src/vpic/misc.cxx:16-105); at 10^9 particles per GPU that serial loop is the
// start-up bottleneck SURVEY.md 8f(4) flags, and no parity claim depends on it,
// so this uses a counter-based generator (reproducible for a given seed and
// independent of launch geometry).
#include "vpb_common.cuh"
#include "vpb_pview.cuh"

namespace vpb {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(uint64_t r) { return ((r >> 40) + 0.5f) * (1.0f / 16777216.0f); }   // (0,1)

__global__ void __launch_bounds__(256) load_thermal_kernel(const PView p, long np, int ppc, float vth,
                                                           float q, uint64_t seed, long tag0, const DomainDev g) {
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < np; k += (long)gridDim.x * blockDim.x) {
    const long cell = k / ppc;
    const int x = 1 + (int)(cell % g.nx);
    const long r = cell / g.nx;
    const int y = 1 + (int)(r % g.ny), z = 1 + (int)(r / g.ny);
    const int v = x + g.sx * (y + g.sy * z);
    const uint64_t b = mix64(seed ^ mix64((uint64_t)k));
    const uint64_t r0 = mix64(b), r1 = mix64(b + 1), r2 = mix64(b + 2), r3 = mix64(b + 3), r4 = mix64(b + 4), r5 = mix64(b + 5),
                   r6 = mix64(b + 6), r7 = mix64(b + 7);
    const float dx = 2.f * u01(r0) - 1.f, dy = 2.f * u01(r1) - 1.f, dz = 2.f * u01(r2) - 1.f;
    // Box-Muller, three normals from two pairs
    const float a0 = sqrtf(-2.f * logf(u01(r3))), a1 = sqrtf(-2.f * logf(u01(r5)));
    float s0, c0, s1, c1;
    sincospif(2.f * u01(r4), &s0, &c0);
    sincospif(2.f * u01(r6), &s1, &c1);
    (void)r7; (void)s1;
    p.set_pos(k, make_float4(dx, dy, dz, __int_as_float(v)));
    p.set_mom(k, make_float4(vth * a0 * c0, vth * a0 * s0, vth * a1 * c1, q));
    const longlong2 tags = make_longlong2(tag0 + k, 0);
    p.set_tag(k, *reinterpret_cast<const float4 *>(&tags));
  }
}

// copy only dx,dy,dz,i (the first quad) from one species to another: co-located ions
__global__ void __launch_bounds__(256) copy_positions_kernel(const PView dst, const PView src, long np) {
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < np; k += (long)gridDim.x * blockDim.x)
    dst.set_pos(k, src.pos(k));
}

// x-propagating vacuum plane wave on the Yee mesh (ey on x nodes, cbz half a cell further): a synthetic
// field state for the field-only benchmark and its energy-conservation property test
__global__ void __launch_bounds__(256) load_plane_wave_kernel(vpb_field_t *__restrict__ f, const DomainDev g, double kdx, float amp) {
  const size_t nv = (size_t)g.sxy * g.sz;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (size_t)gridDim.x * blockDim.x) {
    const int ix = (int)(v % g.sx);
    const float4 z = make_float4(0, 0, 0, 0);
    *FQ(f, g, v, 0) = make_float4(0, amp * (float)cos(kdx * (ix - 1)), 0, 0);
    *FQ(f, g, v, 1) = make_float4(0, 0, amp * (float)cos(kdx * (ix - 0.5)), 0);
    *FQ(f, g, v, 2) = z; *FQ(f, g, v, 3) = z; *FQ(f, g, v, 4) = z;
  }
}

}  // namespace vpb

using namespace vpb;

extern "C" {

// np = ppc * nx*ny*nz particles written to d_p; tags are tag0 + index.
void vpb_load_thermal(vpb_domain_t *dom, vpb_particle_t *d_p, int ppc, float vth, float q, unsigned long long seed, long tag0) {
  if (!dom) VPB_ERROR("Bad grid");
  if (!d_p) VPB_ERROR("Bad particle array");
  if (ppc < 1) VPB_ERROR("Bad ppc");
  const DomainDev &g = dom->d;
  const long np = (long)ppc * g.nx * g.ny * g.nz;
  load_thermal_kernel<<<ctx().sm_count * 16, 256, 0, ctx().stream>>>(PView(d_p, g.p_plane), np, ppc, vth, q, seed, tag0, g);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

// mode = number of wavelengths across the local nx cells (periodic in x)
void vpb_load_plane_wave(vpb_domain_t *dom, vpb_field_t *d_f, int mode, float amp) {
  if (!dom) VPB_ERROR("Bad grid");
  if (!d_f) VPB_ERROR("Bad field");
  const DomainDev &g = dom->d;
  const double kdx = 2.0 * 3.14159265358979323846 * mode / g.nx;
  load_plane_wave_kernel<<<ctx().sm_count * 16, 256, 0, ctx().stream>>>(d_f, g, kdx, amp);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

void vpb_copy_positions(vpb_particle_t *d_dst, const vpb_particle_t *d_src, long np) {
  copy_positions_kernel<<<ctx().sm_count * 16, 256, 0, ctx().stream>>>(PView(d_dst, 0), PView(d_src, 0), np);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

// same, for arrays in the domain's particle layout
void vpb_copy_positions_dom(vpb_domain_t *dom, vpb_particle_t *d_dst, const vpb_particle_t *d_src, long np) {
  if (!dom) VPB_ERROR("Bad grid");
  copy_positions_kernel<<<ctx().sm_count * 16, 256, 0, ctx().stream>>>(PView(d_dst, dom->d.p_plane), PView(d_src, dom->d.p_plane), np);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

}  // extern "C"
