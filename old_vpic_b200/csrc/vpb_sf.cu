// vpb_sf.cu -- species<->field coupling kernels (K5, K6, K10):
//   load_interpolator   src/sf_interface/load_interpolator.cxx:12-144
//   clear_accumulators  src/sf_interface/clear_accumulators.c:27-49
//   unload_accumulator  src/sf_interface/unload_accumulator.cxx:12-67
// Each voxel is one thread; x is the fastest thread index so the 80-byte field_t
// / interpolator_t records of a warp are contiguous.  All three are pure
// streaming stencils: expression order follows the reference's scalar code and
// the file is built with -fmad=false, so results are bit-identical.
#include "vpb_common.cuh"

namespace vpb {

// The 18 coefficients of one voxel (load_interpolator.cxx:45-120) as the six quads of the 96-byte device record
// (bytes 72..95 belong to nobody and are written as zeros).
__device__ __forceinline__ void interpolator_of_voxel(const vpb_field_t *__restrict__ f, const DomainDev &g, size_t v, float4 (&o)[6]) {
  const size_t sX = 1, sY = (size_t)g.sx, sZ = (size_t)g.sxy;
  const float4 e0 = __ldg(CFQ(f, g, v, 0)), b0 = __ldg(CFQ(f, g, v, 1));
  const float4 ex_ = __ldg(CFQ(f, g, v + sX, 0)), bx_ = __ldg(CFQ(f, g, v + sX, 1));
  const float4 ey_ = __ldg(CFQ(f, g, v + sY, 0)), by_ = __ldg(CFQ(f, g, v + sY, 1));
  const float4 ez_ = __ldg(CFQ(f, g, v + sZ, 0)), bz_ = __ldg(CFQ(f, g, v + sZ, 1));
  const float4 eyz = __ldg(CFQ(f, g, v + sY + sZ, 0));
  const float4 ezx = __ldg(CFQ(f, g, v + sZ + sX, 0));
  const float4 exy = __ldg(CFQ(f, g, v + sX + sY, 0));
  const float fourth = 0.25f, half = 0.5f;
  float w0, w1, w2, w3;
  w0 = e0.x; w1 = ey_.x; w2 = ez_.x; w3 = eyz.x;
  o[0].x = fourth * ((w3 + w0) + (w1 + w2));
  o[0].y = fourth * ((w3 - w0) + (w1 - w2));
  o[0].z = fourth * ((w3 - w0) - (w1 - w2));
  o[0].w = fourth * ((w3 + w0) - (w1 + w2));
  w0 = e0.y; w1 = ez_.y; w2 = ex_.y; w3 = ezx.y;
  o[1].x = fourth * ((w3 + w0) + (w1 + w2));
  o[1].y = fourth * ((w3 - w0) + (w1 - w2));
  o[1].z = fourth * ((w3 - w0) - (w1 - w2));
  o[1].w = fourth * ((w3 + w0) - (w1 + w2));
  w0 = e0.z; w1 = ex_.z; w2 = ey_.z; w3 = exy.z;
  o[2].x = fourth * ((w3 + w0) + (w1 + w2));
  o[2].y = fourth * ((w3 - w0) + (w1 - w2));
  o[2].z = fourth * ((w3 - w0) - (w1 - w2));
  o[2].w = fourth * ((w3 + w0) - (w1 + w2));
  w0 = b0.x; w1 = bx_.x;
  o[3].x = half * (w1 + w0); o[3].y = half * (w1 - w0);
  w0 = b0.y; w1 = by_.y;
  o[3].z = half * (w1 + w0); o[3].w = half * (w1 - w0);
  w0 = b0.z; w1 = bz_.z;
  o[4] = make_float4(half * (w1 + w0), half * (w1 - w0), 0.f, 0.f);
  o[5] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// One thread per interior voxel (1..nx, 1..ny, 1..nz).  The first 32 bytes of a
// field_t are {ex,ey,ez,div_e_err | cbx,cby,cbz,div_b_err}: two 128-bit loads.
// STAGE (96-byte device records only): a lane's six quads sit 96 bytes from its neighbour's, so each of the six store
// instructions of a warp touches 24 lines; instead the warp's 3072 contiguous output bytes go through shared memory
// (rows of seven quads: conflict-free) and leave as six stores of 512 contiguous bytes.
template <int STAGE>
__global__ void __launch_bounds__(256) load_interpolator_kernel(vpb_interpolator_t *__restrict__ fi,
                                                                const vpb_field_t *__restrict__ f, const DomainDev g) {
  const int x = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int y = 1 + blockIdx.y;
  const int z = 1 + blockIdx.z;
  if (STAGE) {
    __shared__ float4 stage[8][32 * 7];
    const int lane = threadIdx.x & 31, x0 = x - lane;
    float4 *sw = stage[threadIdx.x >> 5];
    const size_t v0 = (size_t)x0 + (size_t)g.sx * ((size_t)y + (size_t)g.sy * z);
    if (x <= g.nx) {
      float4 o[6];
      interpolator_of_voxel(f, g, v0 + lane, o);
#pragma unroll
      for (int k = 0; k < 6; k++) sw[lane * 7 + k] = o[k];
    }
    __syncwarp();
    const int nq = 6 * min(32, g.nx - x0 + 1);            // quads this warp owes (<= 0: the whole warp is past the row)
    float4 *ob = reinterpret_cast<float4 *>(reinterpret_cast<char *>(fi) + v0 * 96);
#pragma unroll
    for (int j = 0; j < 6; j++) {
      const int q = lane + 32 * j;
      if (q < nq) ob[q] = sw[(q / 6) * 7 + q % 6];
    }
    return;
  }
  if (x > g.nx) return;
  const size_t v = (size_t)x + (size_t)g.sx * ((size_t)y + (size_t)g.sy * z);
  float4 q[6];
  interpolator_of_voxel(f, g, v, q);
  float4 *o = reinterpret_cast<float4 *>(reinterpret_cast<char *>(fi) + v * (size_t)g.fi_bytes);
  o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = q[3];
  if (g.fi_bytes == 96) {
    // the padded device record: its last sector is written whole (a lone 8-byte store would make DRAM read the other
    // 24 bytes first); bytes 72..95 belong to nobody
    o[4] = q[4];
    o[5] = q[5];
  } else {
    *reinterpret_cast<float2 *>(o + 4) = make_float2(q[4].x, q[4].y);   // _pad[2] is left untouched, like the reference
  }
}

// One thread per voxel of (1..nx+1, 1..ny+1, 1..nz+1); jf += c*(4 quadrants).
__global__ void __launch_bounds__(256) unload_accumulator_kernel(vpb_field_t *__restrict__ f, const float4 *__restrict__ a,
                                                                 const DomainDev g, float cx, float cy, float cz) {
  const int x = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int y = 1 + blockIdx.y;
  const int z = 1 + blockIdx.z;
  if (x > g.nx + 1) return;
  const size_t v = (size_t)x + (size_t)g.sx * ((size_t)y + (size_t)g.sy * z);
  const size_t sX = 1, sY = g.sx, sZ = g.sxy;
  // accumulator_t = 3 float4: jx[4], jy[4], jz[4]
  const float4 a0x = __ldg(a + 3 * v), a0y = __ldg(a + 3 * v + 1), a0z = __ldg(a + 3 * v + 2);
  const float4 axy_ = __ldg(a + 3 * (v - sX) + 1), axz_ = __ldg(a + 3 * (v - sX) + 2);  // ax: jy, jz
  const float4 ayx_ = __ldg(a + 3 * (v - sY)), ayz_ = __ldg(a + 3 * (v - sY) + 2);      // ay: jx, jz
  const float4 azx_ = __ldg(a + 3 * (v - sZ)), azy_ = __ldg(a + 3 * (v - sZ) + 1);      // az: jx, jy
  const float4 ayz_x = __ldg(a + 3 * (v - sY - sZ));         // ayz: jx
  const float4 azx_y = __ldg(a + 3 * (v - sZ - sX) + 1);     // azx: jy
  const float4 axy_z = __ldg(a + 3 * (v - sX - sY) + 2);     // axy: jz
  float4 *jf = FQ(f, g, v, 3);                               // jfx,jfy,jfz,rhof
  float4 j = *jf;
  j.x += cx * (a0x.x + ayx_.y + azx_.z + ayz_x.w);
  j.y += cy * (a0y.x + azy_.y + axy_.z + azx_y.w);
  j.z += cz * (a0z.x + axz_.y + ayz_.z + axy_z.w);
  *jf = j;
}

}  // namespace vpb

using namespace vpb;

extern "C" {

void vpb_domain_set_interpolator_layout(vpb_domain_t *dom, int wide) {
  if (!dom) VPB_ERROR("Bad grid");
  dom->d.fi_bytes = wide ? 96 : 80;
}

size_t vpb_interpolator_bytes(const vpb_domain_t *dom) {
  if (!dom) VPB_ERROR("Bad grid");
  return (size_t)dom->d.nv * (size_t)dom->d.fi_bytes;
}

void vpb_load_interpolator(vpb_domain_t *dom, vpb_interpolator_t *d_fi, const vpb_field_t *d_f) {
  if (!d_fi) VPB_ERROR("Bad interpolator");
  if (!d_f) VPB_ERROR("Bad field");
  if (!dom) VPB_ERROR("Bad grid");
  const DomainDev &g = dom->d;
  const int tb = g.nx >= 256 ? 256 : (g.nx >= 128 ? 128 : (g.nx >= 64 ? 64 : 32));
  dim3 grid((g.nx + tb - 1) / tb, g.ny, g.nz);
  ProfScope prof(4);
  // sf.stage_store (tuning): the 96-byte records leave through shared memory as whole lines (1, default) or straight
  // from the registers (0)
  if (g.fi_bytes == 96 && tuning("sf.stage_store", 1)) load_interpolator_kernel<1><<<grid, tb, 0, ctx().stream>>>(d_fi, d_f, g);
  else load_interpolator_kernel<0><<<grid, tb, 0, ctx().stream>>>(d_fi, d_f, g);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

void vpb_clear_accumulators(vpb_domain_t *dom, vpb_accumulator_t *d_a) {
  if (!d_a) VPB_ERROR("Bad accumulator");
  if (!dom) VPB_ERROR("Bad grid");
  // one replica on the device (no per-pipeline copies): clear_accumulators.c:27-49
  VPB_CUDA(cudaMemsetAsync(d_a, 0, (size_t)dom->d.nv * sizeof(vpb_accumulator_t), ctx().stream));
  count_launch();
}

void vpb_unload_accumulator(vpb_domain_t *dom, vpb_field_t *d_f, const vpb_accumulator_t *d_a) {
  if (!d_f) VPB_ERROR("Bad field");
  if (!d_a) VPB_ERROR("Bad accumulator");
  if (!dom) VPB_ERROR("Bad grid");
  const DomainDev &g = dom->d;
  // same expressions and types as unload_accumulator.cxx:29-31 (double until the store)
  const float cx = (float)(0.25 * g.rdy * g.rdz / g.dt);
  const float cy = (float)(0.25 * g.rdz * g.rdx / g.dt);
  const float cz = (float)(0.25 * g.rdx * g.rdy / g.dt);
  const int n = g.nx + 1;
  const int tb = n >= 256 ? 256 : (n >= 128 ? 128 : (n >= 64 ? 64 : 32));
  dim3 grid((n + tb - 1) / tb, g.ny + 1, g.nz + 1);
  ProfScope prof(5);
  unload_accumulator_kernel<<<grid, tb, 0, ctx().stream>>>(d_f, reinterpret_cast<const float4 *>(d_a), g, cx, cy, cz);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

}  // extern "C"
