// vpb_hydro.cu -- hydro moments of a species on the mesh nodes (diagnostics):
//   accumulate_hydro_p   src/species_advance/standard/hydro_p.c:24-161
//   clear_hydro          src/sf_interface/sf_structors.c
//   local_adjust_hydro   src/sf_interface/hydro.c:146-184
//   synchronize_hydro    src/sf_interface/hydro.c:30-141 (NCCL send/recv in place of the reference's ports)
// hydro_t is 64 bytes per node (14 moments + 2 pad words), the same array shape as the fields.  The deposit is
// the push's half step (half E kick, half Boris rotation) followed by a trilinear scatter of 14 moments to the 8
// nodes of the particle's voxel: four vector REDs per node.  Arithmetic follows the reference's scalar loop,
// including the two expressions it evaluates in double (hydro_p.c:86,93); sums differ only in order (atomics).
#include "vpb_comm.cuh"
#include "vpb_pview.cuh"

namespace vpb {

constexpr int kHydroComp = 14;

__device__ __forceinline__ void red_add_v2(float *addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

struct HydroConst {
  float qdt_2mc, qdt_4mc2, c, r8V, mc_q;
};

__global__ void __launch_bounds__(256) hydro_p_kernel(vpb_hydro_t *__restrict__ h0, const PView p, int np, const HydroConst K,
                                                      const vpb_interpolator_t *__restrict__ f0, int fi_bytes, int sx, int sxy) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < np; k += gridDim.x * blockDim.x) {
    const float4 r0 = p.pos(k), r1 = p.mom(k);
    const float x = r0.x, y = r0.y, z = r0.z, q = r1.w;
    const int ii = __float_as_int(r0.w);
    const char *fp = reinterpret_cast<const char *>(f0) + (size_t)ii * fi_bytes;
    const float4 fe_x = ldg4(fp), fe_y = ldg4(fp + 16), fe_z = ldg4(fp + 32), fb_0 = ldg4(fp + 48);
    const float2 fb_1 = ldg2(fp + 64);
    float ux = r1.x, uy = r1.y, uz = r1.z;
    ux += K.qdt_2mc * ((fe_x.x + y * fe_x.y) + z * (fe_x.z + y * fe_x.w));
    uy += K.qdt_2mc * ((fe_y.x + z * fe_y.y) + x * (fe_y.z + z * fe_y.w));
    uz += K.qdt_2mc * ((fe_z.x + x * fe_z.y) + y * (fe_z.z + x * fe_z.w));
    const float bx = fb_0.x + x * fb_0.y, by = fb_0.z + y * fb_0.w, bz = fb_1.x + z * fb_1.y;
    float ke_mc = ux * ux + uy * uy + uz * uz;
    float cg = (float)sqrt((double)(1.f + ke_mc));           // hydro_p.c:86: float sum, double sqrt
    ke_mc *= K.c / (cg + 1.f);
    cg = K.c / cg;
    const float t0 = K.qdt_4mc2 * cg;
    const float b2 = bx * bx + by * by + bz * bz;
    const float t2 = t0 * t0 * b2;
    // hydro_p.c:93: w0*(1+(1./3.)*w2*(1+0.4*w2)) with double literals -> evaluated in double
    const float t3 = (float)((double)t0 * (1.0 + ((1. / 3.) * (double)t2) * (1.0 + 0.4 * (double)t2)));
    float t4 = t3 / (1.f + b2 * t3 * t3);
    t4 += t4;
    const float px_ = ux + t3 * (uy * bz - uz * by), py_ = uy + t3 * (uz * bx - ux * bz), pz_ = uz + t3 * (ux * by - uy * bx);
    ux += t4 * (py_ * bz - pz_ * by);
    uy += t4 * (pz_ * bx - px_ * bz);
    uz += t4 * (px_ * by - py_ * bx);
    const float vx = ux * cg, vy = uy * cg, vz = cg * uz;
    float w[8], t;
    w[0] = K.r8V * q;
    t = x * w[0];
    w[1] = w[0] + t;
    w[0] -= t;
    w[3] = 1.f + y;
    w[2] = w[0] * w[3];
    w[3] *= w[1];
    t = 1.f - y;
    w[0] *= t;
    w[1] *= t;
    w[7] = 1.f + z;
    w[4] = w[0] * w[7];
    w[5] = w[1] * w[7];
    w[6] = w[2] * w[7];
    w[7] *= w[3];
    t = 1.f - z;
    w[0] *= t;
    w[1] *= t;
    w[2] *= t;
    w[3] *= t;
#pragma unroll
    for (int nd = 0; nd < 8; nd++) {
      float *h = reinterpret_cast<float *>(h0 + (size_t)ii + (nd & 1) + ((nd >> 1) & 1) * sx + (nd >> 2) * sxy);
      float wn = w[nd];
      const float jx = wn * vx, jy = wn * vy, jz = wn * vz, rho = wn;
      wn *= K.mc_q;
      const float mx = wn * ux, my = wn * uy, mz = wn * uz;
      red_add_v4(h, jx, jy, jz, rho);
      red_add_v4(h + 4, mx, my, mz, wn * ke_mc);
      red_add_v4(h + 8, mx * vx, my * vy, mz * vz, my * vz);
      red_add_v2(h + 12, mz * vx, mx * vy);
    }
  }
}

// local_adjust_hydro: every node on a face with a local boundary condition is doubled, once per such face it
// lies on (edges twice, corners three times); doubling is exact, so the order of the faces does not matter.
__global__ void __launch_bounds__(256) hydro_adjust_kernel(vpb_hydro_t *__restrict__ h, int nx, int ny, int nz, int sx, int sy,
                                                           int lo_x, int lo_y, int lo_z, int hi_x, int hi_y, int hi_z) {
  const long nn = (long)(nx + 1) * (ny + 1) * (nz + 1);
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < nn; t += (long)gridDim.x * blockDim.x) {
    const int x = 1 + (int)(t % (nx + 1));
    const long r = t / (nx + 1);
    const int y = 1 + (int)(r % (ny + 1)), z = 1 + (int)(r / (ny + 1));
    const int k = (lo_x && x == 1) + (hi_x && x == nx + 1) + (lo_y && y == 1) + (hi_y && y == ny + 1) + (lo_z && z == 1) +
                  (hi_z && z == nz + 1);
    if (!k) continue;
    float *hv = reinterpret_cast<float *>(h + (size_t)x + (size_t)sx * ((size_t)y + (size_t)sy * z));
    for (int j = 0; j < k; j++)
      for (int c = 0; c < kHydroComp; c++) hv[c] *= 2.f;
  }
}

struct HydroPlane {   // node plane X == p, the two other coordinates 1..n+1, x fastest (hydro.c:15-23)
  int X, p, n[3], sx, sy;
  __device__ __forceinline__ long nodes() const { return (long)(n[(X + 1) % 3] + 1) * (n[(X + 2) % 3] + 1); }
  __device__ __forceinline__ size_t voxel(long e) const {
    int c[3];
    c[X] = p;
    // message order: the reference's loops run x fastest, then y, then z over the two free coordinates
    const int a = X == 0 ? 1 : 0, b = X == 2 ? 1 : 2;   // a: faster free axis, b: slower
    c[a] = 1 + (int)(e % (n[a] + 1));
    c[b] = 1 + (int)(e / (n[a] + 1));
    return (size_t)c[0] + (size_t)sx * ((size_t)c[1] + (size_t)sy * c[2]);
  }
};

__global__ void __launch_bounds__(256) hydro_pack_kernel(const vpb_hydro_t *__restrict__ h, const HydroPlane P, float *__restrict__ buf,
                                                         float dX) {
  const long n = P.nodes() * kHydroComp;
  if (blockIdx.x == 0 && threadIdx.x == 0) buf[0] = dX;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x)
    buf[1 + e] = reinterpret_cast<const float *>(h + P.voxel(e / kHydroComp))[e % kHydroComp];
}

__global__ void __launch_bounds__(256) hydro_unpack_kernel(vpb_hydro_t *__restrict__ h, const HydroPlane P, const float *__restrict__ buf,
                                                           float dX) {
  float rw = buf[0], lw = rw + dX;    // hydro.c:76-81
  rw /= lw;
  lw = dX / lw;
  lw += lw;
  rw += rw;
  const long n = P.nodes() * kHydroComp;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    float *dst = &reinterpret_cast<float *>(h + P.voxel(e / kHydroComp))[e % kHydroComp];
    *dst = lw * (*dst) + rw * buf[1 + e];
  }
}

static int face_bc(const DomainDev &g, int X, int s) {
  int ijk[3] = {0, 0, 0};
  ijk[X] = s;
  return g.bc[VPB_BOUNDARY(ijk[0], ijk[1], ijk[2])];
}

static int grid_for(long n) {
  long b = (n + 255) / 256;
  const long cap = (long)ctx().sm_count * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace vpb

using namespace vpb;

extern "C" {

void vpb_clear_hydro(vpb_domain_t *dom, vpb_hydro_t *d_h) {
  if (!d_h) VPB_ERROR("Bad hydro");
  if (!dom) VPB_ERROR("Bad grid");
  VPB_CUDA(cudaMemsetAsync(d_h, 0, (size_t)dom->d.nv * sizeof(vpb_hydro_t), ctx().stream));
}

void vpb_accumulate_hydro_p(vpb_domain_t *dom, vpb_hydro_t *d_h, const vpb_particle_t *d_p, int np, float q_m,
                            const vpb_interpolator_t *d_f) {
  if (!d_h) VPB_ERROR("Bad hydro");
  if (!d_p) VPB_ERROR("Bad particle array");
  if (np < 0) VPB_ERROR("Bad number of particles");
  if (!d_f) VPB_ERROR("Bad field");
  if (!dom) VPB_ERROR("Bad grid");
  if (np == 0) return;
  const DomainDev &g = dom->d;
  HydroConst K;   // hydro_p.c:48-52, same expressions and types
  K.qdt_2mc = (float)(0.5 * q_m * g.dt / g.cvac);
  K.qdt_4mc2 = (float)(0.25 * q_m * g.dt / (g.cvac * g.cvac));
  K.c = g.cvac;
  K.r8V = (float)(0.125 * g.rdx * g.rdy * g.rdz);
  K.mc_q = g.cvac / q_m;
  hydro_p_kernel<<<grid_for(np), 256, 0, ctx().stream>>>(d_h, PView(d_p, g.p_plane), np, K, d_f, g.fi_bytes, g.sx, g.sxy);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

void vpb_local_adjust_hydro(vpb_domain_t *dom, vpb_hydro_t *d_h) {
  if (!d_h) VPB_ERROR("Bad hydro");
  if (!dom) VPB_ERROR("Bad grid");
  const DomainDev &g = dom->d;
  int loc[6];
  bool any = false;
  for (int f = 0; f < 6; f++) {
    const int bc = face_bc(g, f % 3, f < 3 ? -1 : 1);
    loc[f] = (bc < 0 || bc > g.nproc) ? 1 : 0;   // hydro.c:155
    any |= loc[f] != 0;
  }
  if (!any) return;
  hydro_adjust_kernel<<<grid_for((long)(g.nx + 1) * (g.ny + 1) * (g.nz + 1)), 256, 0, ctx().stream>>>(
      d_h, g.nx, g.ny, g.nz, g.sx, g.sy, loc[0], loc[1], loc[2], loc[3], loc[4], loc[5]);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

// local_adjust_hydro, then for x, y, z in turn: both face planes are packed, exchanged, and each plane becomes
// lw*mine + rw*neighbour's (hydro.c:107-136; the second pass carries the edges the first one completed).
void vpb_synchronize_hydro(vpb_domain_t *dom, vpb_hydro_t *d_h) {
  if (!d_h) VPB_ERROR("Bad hydro");
  if (!dom) VPB_ERROR("Bad grid");
  vpb_local_adjust_hydro(dom, d_h);
  const DomainDev &g = dom->d;
  cudaStream_t st = ctx().stream;
  const int n[3] = {g.nx, g.ny, g.nz};
  const float cell[3] = {g.dx, g.dy, g.dz};
  for (int X = 0; X < 3; X++) {
    const int Y = (X + 1) % 3, Z = (X + 2) % 3;
    const size_t floats = 1 + (size_t)kHydroComp * (n[Y] + 1) * (n[Z] + 1);
    const size_t bytes = (floats * sizeof(float) + 255) & ~(size_t)255;
    const int bc_lo = face_bc(g, X, -1), bc_hi = face_bc(g, X, 1);
    const bool r_lo = bc_lo >= 0 && bc_lo < g.nproc, r_hi = bc_hi >= 0 && bc_hi < g.nproc;
    if (!r_lo && !r_hi) continue;
    // scratch: send_lo | send_hi | recv_lo | recv_hi
    float *buf = (float *)scratch(4 * bytes);
    float *send_lo = buf, *send_hi = (float *)((char *)buf + bytes), *recv_lo = (float *)((char *)buf + 2 * bytes),
          *recv_hi = (float *)((char *)buf + 3 * bytes);
    HydroPlane P;
    P.X = X; P.n[0] = n[0]; P.n[1] = n[1]; P.n[2] = n[2]; P.sx = g.sx; P.sy = g.sy;
    const int gx = grid_for((long)(floats - 1));
    if (r_lo) { P.p = 1; hydro_pack_kernel<<<gx, 256, 0, st>>>(d_h, P, send_lo, cell[X]); count_launch(); }
    if (r_hi) { P.p = n[X] + 1; hydro_pack_kernel<<<gx, 256, 0, st>>>(d_h, P, send_hi, cell[X]); count_launch(); }
    Xfer x[4];
    int nx = 0;
    // sends by face (lo, hi), receives by face (hi, lo): see vpb_comm.cuh on matching order
    if (r_lo && bc_lo != g.rank) x[nx++] = {send_lo, floats * 4, bc_lo, nullptr, 0, -1};
    if (r_hi && bc_hi != g.rank) x[nx++] = {send_hi, floats * 4, bc_hi, nullptr, 0, -1};
    if (r_hi && bc_hi != g.rank) x[nx++] = {nullptr, 0, -1, recv_hi, floats * 4, bc_hi};
    if (r_lo && bc_lo != g.rank) x[nx++] = {nullptr, 0, -1, recv_lo, floats * 4, bc_lo};
    if (nx) comm_exchange(x, nx);
    // what arrives through my +X face is what the +X neighbour packed from ITS plane 1; it updates my plane n+1
    // (the reference unpacks that one first); a face joined to this rank reads the opposite send buffer
    if (r_hi) { P.p = n[X] + 1; hydro_unpack_kernel<<<gx, 256, 0, st>>>(d_h, P, bc_hi == g.rank ? send_lo : recv_hi, cell[X]); count_launch(); }
    if (r_lo) { P.p = 1; hydro_unpack_kernel<<<gx, 256, 0, st>>>(d_h, P, bc_lo == g.rank ? send_hi : recv_lo, cell[X]); count_launch(); }
    VPB_CUDA(cudaGetLastError());
  }
}

}  // extern "C"
