// vpb_fields.cu -- the standard field advance on the device (K7-K9, K11-K14, K16):
//   advance_b            src/field_advance/standard/advance_b.c:12-14,75-161
//   advance_e            advance_e.c:8-25,87-330   (+ vacuum/vfa_advance_e.c:7-9)
//   compute_curl_b       compute_curl_b.c:8-18
//   energy_f             energy_f.c:13-91,93-179
//   div-E / div-B clean  compute_div_e_err.c, compute_rms_div_e_err.c, clean_div_e.c,
//                        compute_div_b_err.c, compute_rms_div_b_err.c, clean_div_b.c, compute_rhob.c
//   clear_jf, clear_rhof sfa.c:188-235
// The reference walks the interior in pipelines and patches the surface planes
// on the host; here one launch covers the whole index range of each component
// (the updates are Jacobi-style: they only read the OTHER field), ghosts having
// been filled first (vpb_faces.cu).  One thread per voxel, x fastest; a field_t
// is five 16-byte quads {e,div_e | cb,div_b | tca,rhob | jf,rhof | material ids}
// and every access is a 128-bit load/store of one quad.  Expression order is the
// reference's; built with -fmad=false so results are bit-identical.
#include "vpb_common.cuh"
#include "vpb_faces.cuh"

namespace vpb {

__device__ __forceinline__ float comp(const float4 &q, int c) { return c == 0 ? q.x : (c == 1 ? q.y : q.z); }
__device__ __forceinline__ void set_comp(float4 &q, int c, float v) {
  if (c == 0) q.x = v; else if (c == 1) q.y = v; else q.z = v;
}
// 8 uint16 material ids packed in quad 4: [ematx ematy ematz nmat | fmatx fmaty fmatz cmat]
__device__ __forceinline__ int mat_of(const uint4 &q, int k) {
  const unsigned w = k < 2 ? q.x : (k < 4 ? q.y : (k < 6 ? q.z : q.w));
  return (k & 1) ? (w >> 16) : (w & 0xffffu);
}

struct VoxelIdx {
  int x, y, z;
  size_t v;
  bool ok;
};
// threads cover x in 1..nx+1, y in 1..ny+1, z in 1..nz+1
__device__ __forceinline__ VoxelIdx voxel_of_thread(const DomainDev &g) {
  VoxelIdx i;
  i.x = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  i.y = 1 + blockIdx.y;
  i.z = 1 + blockIdx.z;
  i.ok = i.x <= g.nx + 1;
  i.v = (size_t)i.x + (size_t)g.sx * ((size_t)i.y + (size_t)g.sy * i.z);
  return i;
}

#define QUAD(f, v, q) FQ(f, g, v, q)
#define CQUAD(f, v, q) CFQ(f, g, v, q)

// cbX -= pY*(eZ(+Y) - eZ) - pZ*(eY(+Z) - eY)  on X faces (X in 1..nX+1, others 1..n)
__global__ void __launch_bounds__(256) advance_b_kernel(vpb_field_t *__restrict__ f, const DomainDev g, float px, float py, float pz) {
  const VoxelIdx i = voxel_of_thread(g);
  if (!i.ok) return;
  const bool inx = i.x <= g.nx, iny = i.y <= g.ny, inz = i.z <= g.nz;
  const size_t sX = 1, sY = g.sx, sZ = g.sxy;
  const float4 e0 = *CQUAD(f, i.v, 0);
  float4 ex = e0, ey = e0, ez = e0;
  if (inx) ex = *CQUAD(f, i.v + sX, 0);
  if (iny) ey = *CQUAD(f, i.v + sY, 0);
  if (inz) ez = *CQUAD(f, i.v + sZ, 0);
  float4 b = *QUAD(f, i.v, 1);
  if (iny && inz) b.x -= (py * (ey.z - e0.z) - pz * (ez.y - e0.y));
  if (inz && inx) b.y -= (pz * (ez.x - e0.x) - px * (ex.z - e0.z));
  if (inx && iny) b.z -= (px * (ex.y - e0.y) - py * (ey.x - e0.x));
  *QUAD(f, i.v, 1) = b;
}

// MODE 0: standard advance_e; 1: vacuum advance_e; 2: compute_curl_b (tca only).
// UNIFORM: one material (ids not read).
template <int MODE, bool UNIFORM>
__global__ void __launch_bounds__(256) advance_e_kernel(vpb_field_t *__restrict__ f, const vpb_material_coefficient_t *__restrict__ m,
                                                        const DomainDev g, float px, float py, float pz, float damp, float cj) {
  const VoxelIdx i = voxel_of_thread(g);
  if (!i.ok) return;
  const bool inX[3] = {i.x <= g.nx, i.y <= g.ny, i.z <= g.nz};
  const size_t st[3] = {1, (size_t)g.sx, (size_t)g.sxy};
  const float p[3] = {px, py, pz};
  const float4 b0 = *CQUAD(f, i.v, 1);
  float4 bm[3];
  uint4 m0 = make_uint4(0, 0, 0, 0), mm[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    bm[a] = *CQUAD(f, i.v - st[a], 1);
    if (!UNIFORM) mm[a] = *reinterpret_cast<const uint4 *>(CQUAD(f, i.v - st[a], 4));
  }
  if (!UNIFORM) m0 = *reinterpret_cast<const uint4 *>(CQUAD(f, i.v, 4));
  float4 e = *QUAD(f, i.v, 0), tca = *QUAD(f, i.v, 2);
  float4 jf = make_float4(0, 0, 0, 0);
  if (MODE != 2) jf = *CQUAD(f, i.v, 3);
#pragma unroll
  for (int X = 0; X < 3; X++) {
    const int Y = (X + 1) % 3, Z = (X + 2) % 3;
    if (!inX[X]) continue;
    if (MODE == 1) {
      set_comp(e, X, comp(e, X) + ((p[Y] * (comp(b0, Z) - comp(bm[Y], Z)) - p[Z] * (comp(b0, Y) - comp(bm[Z], Y))) - cj * comp(jf, X)));
    } else {
      const vpb_material_coefficient_t *c0 = m, *cZ0 = m, *cZy = m, *cY0 = m, *cYz = m;
      if (!UNIFORM) {
        c0 = m + mat_of(m0, X);
        cZ0 = m + mat_of(m0, 4 + Z); cZy = m + mat_of(mm[Y], 4 + Z);
        cY0 = m + mat_of(m0, 4 + Y); cYz = m + mat_of(mm[Z], 4 + Y);
      }
      const float curl = p[Y] * (comp(b0, Z) * (&cZ0->rmux)[Z] - comp(bm[Y], Z) * (&cZy->rmux)[Z]) -
                         p[Z] * (comp(b0, Y) * (&cY0->rmux)[Y] - comp(bm[Z], Y) * (&cYz->rmux)[Y]);
      if (MODE == 2) {
        set_comp(tca, X, curl);
      } else {
        const float t = curl - damp * comp(tca, X);
        set_comp(tca, X, t);
        set_comp(e, X, (&c0->decayx)[2 * X] * comp(e, X) + (&c0->decayx)[2 * X + 1] * (t - cj * comp(jf, X)));
      }
    }
  }
  if (MODE != 2) *QUAD(f, i.v, 0) = e;
  if (MODE != 1) *QUAD(f, i.v, 2) = tca;
}

// The same update with every thread marching ZC voxels up the z axis: the cb quad (and the material ids) of voxel z are
// the z-1 neighbour of the next iteration and stay in registers, so the z-1 plane is never fetched again.  On a 1024^3
// box that plane is 34 MB of cb + material quads behind the current one; fetching it from L2 costs a quarter of the
// kernel's load wavefronts and, once the planes of all streams no longer fit the L2, DRAM traffic (profiles/r2w).
// Same expressions in the same order: bit-identical results.
template <int MODE, bool UNIFORM, int ZC>
__global__ void __launch_bounds__(256) advance_e_march_kernel(vpb_field_t *__restrict__ f, const vpb_material_coefficient_t *__restrict__ m,
                                                              const DomainDev g, float px, float py, float pz, float damp, float cj) {
  const int x = 1 + blockIdx.x * blockDim.x + threadIdx.x, y = 1 + blockIdx.y, z0 = 1 + blockIdx.z * ZC;
  if (x > g.nx + 1) return;
  const bool inx = x <= g.nx, iny = y <= g.ny;
  const size_t st[3] = {1, (size_t)g.sx, (size_t)g.sxy};
  const float p[3] = {px, py, pz};
  size_t v = (size_t)x + (size_t)g.sx * ((size_t)y + (size_t)g.sy * z0);
  float4 bz = *CQUAD(f, v - st[2], 1);
  uint4 mz = make_uint4(0, 0, 0, 0);
  if (!UNIFORM) mz = *reinterpret_cast<const uint4 *>(CQUAD(f, v - st[2], 4));
#pragma unroll 1
  for (int k = 0; k < ZC; k++, v += st[2]) {
    const int z = z0 + k;
    if (z > g.nz + 1) break;
    const bool inX[3] = {inx, iny, z <= g.nz};
    const float4 b0 = *CQUAD(f, v, 1);
    float4 bm[3];
    uint4 m0 = make_uint4(0, 0, 0, 0), mm[3];
    bm[0] = *CQUAD(f, v - st[0], 1);
    bm[1] = *CQUAD(f, v - st[1], 1);
    bm[2] = bz;
    if (!UNIFORM) {
      mm[0] = *reinterpret_cast<const uint4 *>(CQUAD(f, v - st[0], 4));
      mm[1] = *reinterpret_cast<const uint4 *>(CQUAD(f, v - st[1], 4));
      mm[2] = mz;
      m0 = *reinterpret_cast<const uint4 *>(CQUAD(f, v, 4));
    }
    float4 e = *QUAD(f, v, 0), tca = *QUAD(f, v, 2);
    float4 jf = make_float4(0, 0, 0, 0);
    if (MODE != 2) jf = *CQUAD(f, v, 3);
#pragma unroll
    for (int X = 0; X < 3; X++) {
      const int Y = (X + 1) % 3, Z = (X + 2) % 3;
      if (!inX[X]) continue;
      if (MODE == 1) {
        set_comp(e, X, comp(e, X) + ((p[Y] * (comp(b0, Z) - comp(bm[Y], Z)) - p[Z] * (comp(b0, Y) - comp(bm[Z], Y))) - cj * comp(jf, X)));
      } else {
        const vpb_material_coefficient_t *c0 = m, *cZ0 = m, *cZy = m, *cY0 = m, *cYz = m;
        if (!UNIFORM) {
          c0 = m + mat_of(m0, X);
          cZ0 = m + mat_of(m0, 4 + Z); cZy = m + mat_of(mm[Y], 4 + Z);
          cY0 = m + mat_of(m0, 4 + Y); cYz = m + mat_of(mm[Z], 4 + Y);
        }
        const float curl = p[Y] * (comp(b0, Z) * (&cZ0->rmux)[Z] - comp(bm[Y], Z) * (&cZy->rmux)[Z]) -
                           p[Z] * (comp(b0, Y) * (&cY0->rmux)[Y] - comp(bm[Z], Y) * (&cYz->rmux)[Y]);
        if (MODE == 2) {
          set_comp(tca, X, curl);
        } else {
          const float t = curl - damp * comp(tca, X);
          set_comp(tca, X, t);
          set_comp(e, X, (&c0->decayx)[2 * X] * comp(e, X) + (&c0->decayx)[2 * X + 1] * (t - cj * comp(jf, X)));
        }
      }
    }
    if (MODE != 2) *QUAD(f, v, 0) = e;
    if (MODE != 1) *QUAD(f, v, 2) = tca;
    bz = b0;
    mz = m0;
  }
}

// MODE 0: div_e_err (compute_div_e_err.c:7-11); 1: rhob (compute_rhob.c:8-12). Nodes 1..n+1.
template <int MODE, bool UNIFORM>
__global__ void __launch_bounds__(256) div_e_kernel(vpb_field_t *__restrict__ f, const vpb_material_coefficient_t *__restrict__ m,
                                                    const DomainDev g, float px, float py, float pz, float cj) {
  const VoxelIdx i = voxel_of_thread(g);
  if (!i.ok) return;
  const size_t st[3] = {1, (size_t)g.sx, (size_t)g.sxy};
  const float p[3] = {px, py, pz};
  float4 e0 = *QUAD(f, i.v, 0);
  float4 q2 = *QUAD(f, i.v, 2);
  const float4 q3 = *CQUAD(f, i.v, 3);
  uint4 m0 = make_uint4(0, 0, 0, 0);
  if (!UNIFORM) m0 = *reinterpret_cast<const uint4 *>(CQUAD(f, i.v, 4));
  float t[3];
#pragma unroll
  for (int X = 0; X < 3; X++) {
    const float4 em = *CQUAD(f, i.v - st[X], 0);
    float eps0 = m[0].epsx, eps1 = eps0;
    if (!UNIFORM) {
      const uint4 m1 = *reinterpret_cast<const uint4 *>(CQUAD(f, i.v - st[X], 4));
      eps0 = (&m[mat_of(m0, X)].epsx)[X];
      eps1 = (&m[mat_of(m1, X)].epsx)[X];
    } else {
      eps0 = eps1 = (&m[0].epsx)[X];
    }
    t[X] = p[X] * (eps0 * comp(e0, X) - eps1 * comp(em, X));
  }
  const float nc = UNIFORM ? m[0].nonconductive : m[mat_of(m0, 3)].nonconductive;
  if (MODE == 0) {
    e0.w = nc * (t[0] + t[1] + t[2] - cj * (q3.w + q2.w));
    *QUAD(f, i.v, 0) = e0;
  } else {
    q2.w = nc * (t[0] + t[1] + t[2] - q3.w);
    *QUAD(f, i.v, 2) = q2;
  }
}

// eX += driveX*pX*(div_e_err(+X) - div_e_err) on X edges (clean_div_e.c:6-13)
template <bool UNIFORM>
__global__ void __launch_bounds__(256) clean_div_e_kernel(vpb_field_t *__restrict__ f, const vpb_material_coefficient_t *__restrict__ m,
                                                          const DomainDev g, float px, float py, float pz) {
  const VoxelIdx i = voxel_of_thread(g);
  if (!i.ok) return;
  const bool inX[3] = {i.x <= g.nx, i.y <= g.ny, i.z <= g.nz};
  const size_t st[3] = {1, (size_t)g.sx, (size_t)g.sxy};
  const float p[3] = {px, py, pz};
  float4 e = *QUAD(f, i.v, 0);
  uint4 m0 = make_uint4(0, 0, 0, 0);
  if (!UNIFORM) m0 = *reinterpret_cast<const uint4 *>(CQUAD(f, i.v, 4));
#pragma unroll
  for (int X = 0; X < 3; X++) {
    if (!inX[X]) continue;
    const float dp = CQUAD(f, i.v + st[X], 0)->w;
    const float drive = (&m[UNIFORM ? 0 : mat_of(m0, X)].decayx)[2 * X + 1];
    set_comp(e, X, comp(e, X) + drive * p[X] * (dp - e.w));
  }
  *QUAD(f, i.v, 0) = e;
}

// cells 1..n (compute_div_b_err.c:44-46)
__global__ void __launch_bounds__(256) div_b_kernel(vpb_field_t *__restrict__ f, const DomainDev g, float px, float py, float pz) {
  const VoxelIdx i = voxel_of_thread(g);
  if (!i.ok || i.x > g.nx || i.y > g.ny || i.z > g.nz) return;
  float4 b0 = *QUAD(f, i.v, 1);
  const float4 bx = *CQUAD(f, i.v + 1, 1), by = *CQUAD(f, i.v + g.sx, 1), bz = *CQUAD(f, i.v + g.sxy, 1);
  b0.w = px * (bx.x - b0.x) + py * (by.y - b0.y) + pz * (bz.z - b0.z);
  *QUAD(f, i.v, 1) = b0;
}

// cbX += pX*(div_b_err - div_b_err(-X)) on X faces (clean_div_b.c:6-8)
__global__ void __launch_bounds__(256) clean_div_b_kernel(vpb_field_t *__restrict__ f, const DomainDev g, float px, float py, float pz) {
  const VoxelIdx i = voxel_of_thread(g);
  if (!i.ok) return;
  const bool inx = i.x <= g.nx, iny = i.y <= g.ny, inz = i.z <= g.nz;
  float4 b = *QUAD(f, i.v, 1);
  const float d0 = b.w;
  if (iny && inz) b.x += px * (d0 - CQUAD(f, i.v - 1, 1)->w);
  if (inz && inx) b.y += py * (d0 - CQUAD(f, i.v - g.sx, 1)->w);
  if (inx && iny) b.z += pz * (d0 - CQUAD(f, i.v - g.sxy, 1)->w);
  *QUAD(f, i.v, 1) = b;
}

// which: 0 clear jfx,jfy,jfz   1 clear rhof   (all voxels, ghosts included)
__global__ void __launch_bounds__(256) clear_quad3_kernel(vpb_field_t *__restrict__ f, const DomainDev g, int which) {
  const size_t nv = (size_t)g.nv;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (size_t)gridDim.x * blockDim.x) {
    float4 q = *QUAD(f, v, 3);
    if (which == 0) { q.x = 0; q.y = 0; q.z = 0; } else { q.w = 0; }
    *QUAD(f, v, 3) = q;
  }
}

__device__ __forceinline__ void block_sum_to(double *vals, int n, double *__restrict__ out) {
  __shared__ double ws[8][8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k = 0; k < n; k++) {
    const double s = warp_sum(vals[k]);
    if (lane == 0) ws[w][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < n) {
    double t = 0;
    for (int j = 0; j < (int)(blockDim.x >> 5); j++) t += ws[j][threadIdx.x];
    atomicAdd(out + threadIdx.x, t);
  }
}

// energy_f.c:51-69; cells 1..n; out[6] accumulates the unscaled sums
template <bool UNIFORM>
__global__ void __launch_bounds__(256) energy_f_kernel(const vpb_field_t *__restrict__ f, const vpb_material_coefficient_t *__restrict__ m,
                                                       const DomainDev g, double *__restrict__ out) {
  const VoxelIdx i = voxel_of_thread(g);
  double en[6] = {0, 0, 0, 0, 0, 0};
  if (i.ok && i.x <= g.nx && i.y <= g.ny && i.z <= g.nz) {
    const size_t st[3] = {1, (size_t)g.sx, (size_t)g.sxy};
#pragma unroll
    for (int X = 0; X < 3; X++) {
      const int Y = (X + 1) % 3, Z = (X + 2) % 3;
      const size_t ve[4] = {i.v, i.v + st[Y], i.v + st[Z], i.v + st[Y] + st[Z]};
      float se[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float e = comp(*CQUAD(f, ve[k], 0), X);
        float eps = (&m[0].epsx)[X];
        if (!UNIFORM) eps = (&m[mat_of(*reinterpret_cast<const uint4 *>(CQUAD(f, ve[k], 4)), X)].epsx)[X];
        se[k] = eps * e * e;
      }
      en[X] = 0.25 * (se[0] + se[1] + se[2] + se[3]);
      const size_t vb[2] = {i.v, i.v + st[X]};
      float sb[2];
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const float b = comp(*CQUAD(f, vb[k], 1), X);
        float rmu = (&m[0].rmux)[X];
        if (!UNIFORM) rmu = (&m[mat_of(*reinterpret_cast<const uint4 *>(CQUAD(f, vb[k], 4)), 4 + X)].rmux)[X];
        sb[k] = rmu * b * b;
      }
      en[3 + X] = 0.5 * (sb[0] + sb[1]);
    }
  }
  block_sum_to(en, 6, out);
}

// which 0: div_e_err over nodes 1..n+1 with surface weights (compute_rms_div_e_err.c);
// which 1: div_b_err over cells 1..n (compute_rms_div_b_err.c).  out[0] += sum.
__global__ void __launch_bounds__(256) rms_kernel(const vpb_field_t *__restrict__ f, const DomainDev g, int which, double *__restrict__ out) {
  const VoxelIdx i = voxel_of_thread(g);
  double s[1] = {0};
  if (i.ok) {
    if (which == 0) {
      const float e = CQUAD(f, i.v, 0)->w;
      const int nsurf = (i.x == 1 || i.x == g.nx + 1) + (i.y == 1 || i.y == g.ny + 1) + (i.z == 1 || i.z == g.nz + 1);
      if (nsurf == 0) s[0] = (double)(e * e);   // interior: float product (compute_rms_div_e_err.c:31)
      else s[0] = (nsurf == 1 ? 0.5 : nsurf == 2 ? 0.25 : 0.125) * (double)e * (double)e;
    } else if (i.x <= g.nx && i.y <= g.ny && i.z <= g.nz) {
      const float e = CQUAD(f, i.v, 1)->w;
      s[0] = (double)(e * e);
    }
  }
  block_sum_to(s, 1, out);
}

// AoS <-> planar copy, one thread per quad
__global__ void __launch_bounds__(256) field_convert_kernel(float4 *__restrict__ dst, const float4 *__restrict__ src, size_t nv,
                                                            size_t nvp, int to_planar) {
  const size_t n5 = 5 * nv;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n5; t += (size_t)gridDim.x * blockDim.x) {
    const size_t v = t / 5, q = t - 5 * v;
    if (to_planar) dst[q * nvp + v] = src[t];
    else dst[t] = src[q * nvp + v];
  }
}

static inline int tb_for(int n) { return n >= 256 ? 256 : (n >= 128 ? 128 : (n >= 64 ? 64 : 32)); }
static inline dim3 node_grid(const DomainDev &g, int tb) { return dim3((g.nx + 1 + tb - 1) / tb, g.ny + 1, g.nz + 1); }

}  // namespace vpb

using namespace vpb;

static const vpb_material_coefficient_t *uniform_vacuum() {
  // device copy of a single vacuum material (eps=mu=1, sigma=0): what vfa_* assume (vfa.c:60-75)
  static vpb_material_coefficient_t *d = nullptr;
  if (!d) {
    vpb_material_coefficient_t h = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, {0, 0, 0}};
    VPB_CUDA(cudaMalloc(&d, sizeof(h)));
    VPB_CUDA(cudaMemcpy(d, &h, sizeof(h), cudaMemcpyHostToDevice));
  }
  return d;
}

#define CHECK_FG()                      \
  if (!d_f) VPB_ERROR("Bad field");     \
  if (!dom) VPB_ERROR("Bad grid");      \
  const DomainDev &g = dom->d;          \
  cudaStream_t st = ctx().stream;       \
  const int tb = tb_for(g.nx + 1);      \
  const dim3 grid = node_grid(g, tb);   \
  (void)st; (void)tb; (void)grid

extern "C" {

// ---- field layout (DESIGN.md "field layout") ----
static size_t planar_quads(const DomainDev &g) { return ((size_t)g.nv + 7) & ~(size_t)7; }

void vpb_domain_set_field_layout(vpb_domain_t *dom, int planar) {
  if (!dom) VPB_ERROR("Bad grid");
  DomainDev &g = dom->d;
  if (planar) { g.fqv = 1; g.fqq = (long)planar_quads(g); }
  else { g.fqv = 5; g.fqq = 1; }
}

int vpb_domain_field_layout(const vpb_domain_t *dom) {
  if (!dom) VPB_ERROR("Bad grid");
  return dom->d.fqv == 1;
}

size_t vpb_field_bytes(const vpb_domain_t *dom) {
  if (!dom) VPB_ERROR("Bad grid");
  return dom->d.fqv == 1 ? 5 * planar_quads(dom->d) * sizeof(float4) : (size_t)dom->d.nv * sizeof(vpb_field_t);
}

void vpb_field_convert(vpb_domain_t *dom, vpb_field_t *d_dst, const vpb_field_t *d_src, int to_planar) {
  if (!dom) VPB_ERROR("Bad grid");
  if (!d_dst || !d_src) VPB_ERROR("Bad field");
  const DomainDev &g = dom->d;
  field_convert_kernel<<<ctx().sm_count * 16, 256, 0, ctx().stream>>>((float4 *)d_dst, (const float4 *)d_src, (size_t)g.nv,
                                                                       planar_quads(g), to_planar);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

void vpb_advance_b(vpb_domain_t *dom, vpb_field_t *d_f, float frac) {
  CHECK_FG();
  const float px = (g.nx > 1) ? frac * g.cvac * g.dt * g.rdx : 0;   // advance_b.c:112-114
  const float py = (g.ny > 1) ? frac * g.cvac * g.dt * g.rdy : 0;
  const float pz = (g.nz > 1) ? frac * g.cvac * g.dt * g.rdz : 0;
  {
    ProfScope prof(2);
    advance_b_kernel<<<grid, tb, 0, st>>>(d_f, g, px, py, pz);
  }
  count_launch();
  faces_local_adjust(dom, d_f, ADJ_NORM_B);
  VPB_CUDA(cudaGetLastError());
}

static void launch_e(int mode, vpb_domain_t *dom, vpb_field_t *d_f, const vpb_material_coefficient_t *d_m, int n_mat) {
  const DomainDev &g = dom->d;
  cudaStream_t st = ctx().stream;
  const int tb = tb_for(g.nx + 1);
  const dim3 grid = node_grid(g, tb);
  const float damp = (mode == 0) ? g.damp : 0.f;
  const float px = (g.nx > 1) ? (1 + damp) * g.cvac * g.dt * g.rdx : 0;   // advance_e.c:104-107
  const float py = (g.ny > 1) ? (1 + damp) * g.cvac * g.dt * g.rdy : 0;
  const float pz = (g.nz > 1) ? (1 + damp) * g.cvac * g.dt * g.rdz : 0;
  const float cj = g.dt / g.eps0;
  const bool uni = n_mat <= 1;
  ProfScope prof(mode == 2 ? 6 : 3);
  // fields.march_z (tuning): threads march 16 voxels up z with the z-1 neighbour in registers. 1: the standard advance_e
  // with a material table; 2: the vacuum and one-material variants as well; 0: one voxel per thread everywhere
  const int march = tuning("fields.march_z", 1);
  constexpr int ZC = 16;
  const dim3 mgrid(grid.x, grid.y, (g.nz + 1 + ZC - 1) / ZC);
  if (mode == 0 && !uni && march >= 1) { advance_e_march_kernel<0, false, ZC><<<mgrid, tb, 0, st>>>(d_f, d_m, g, px, py, pz, damp, cj); count_launch(); return; }
  if (mode == 0 && uni && march >= 2) { advance_e_march_kernel<0, true, ZC><<<mgrid, tb, 0, st>>>(d_f, d_m, g, px, py, pz, damp, cj); count_launch(); return; }
  if (mode == 1 && march >= 2) { advance_e_march_kernel<1, true, ZC><<<mgrid, tb, 0, st>>>(d_f, d_m, g, px, py, pz, damp, cj); count_launch(); return; }
  if (mode == 1) advance_e_kernel<1, true><<<grid, tb, 0, st>>>(d_f, d_m, g, px, py, pz, damp, cj);
  else if (mode == 0 && uni) advance_e_kernel<0, true><<<grid, tb, 0, st>>>(d_f, d_m, g, px, py, pz, damp, cj);
  else if (mode == 0) advance_e_kernel<0, false><<<grid, tb, 0, st>>>(d_f, d_m, g, px, py, pz, damp, cj);
  else if (uni) advance_e_kernel<2, true><<<grid, tb, 0, st>>>(d_f, d_m, g, px, py, pz, damp, cj);
  else advance_e_kernel<2, false><<<grid, tb, 0, st>>>(d_f, d_m, g, px, py, pz, damp, cj);
  count_launch();
}

void vpb_advance_e(vpb_domain_t *dom, vpb_field_t *d_f, const vpb_material_coefficient_t *d_m, int n_mat, int vacuum) {
  CHECK_FG();
  if (vacuum) {
    if (g.damp != 0) VPB_ERROR("Vacuum field advance does not support TCA radiation damping");   // vfa.c:66-67
    d_m = uniform_vacuum(); n_mat = 1;
  } else if (!d_m) VPB_ERROR("Bad material coefficients");
  { ProfScope prof(8); faces_ghost_exchange(dom, d_f, MSG_GHOST_TANG_B); }   // begin/local/end_remote_ghost_tang_b (advance_e.c:114-115,197)
  launch_e(vacuum ? 1 : 0, dom, d_f, d_m, n_mat);
  { ProfScope prof(8); faces_local_adjust(dom, d_f, ADJ_TANG_E); }
  VPB_CUDA(cudaGetLastError());
}

void vpb_compute_curl_b(vpb_domain_t *dom, vpb_field_t *d_f, const vpb_material_coefficient_t *d_m, int n_mat) {
  CHECK_FG();
  if (!d_m) { d_m = uniform_vacuum(); n_mat = 1; }
  faces_ghost_exchange(dom, d_f, MSG_GHOST_TANG_B);
  launch_e(2, dom, d_f, d_m, n_mat);
  VPB_CUDA(cudaGetLastError());
}

void vpb_clear_jf(vpb_domain_t *dom, vpb_field_t *d_f) {
  CHECK_FG();
  clear_quad3_kernel<<<ctx().sm_count * 8, 256, 0, st>>>(d_f, g, 0);
  count_launch();
}

void vpb_clear_rhof(vpb_domain_t *dom, vpb_field_t *d_f) {
  CHECK_FG();
  clear_quad3_kernel<<<ctx().sm_count * 8, 256, 0, st>>>(d_f, g, 1);
  count_launch();
}

void vpb_synchronize_jf(vpb_domain_t *dom, vpb_field_t *d_f) {
  CHECK_FG();
  faces_local_adjust(dom, d_f, ADJ_JF);
  faces_sync_passes(dom, d_f, MSG_SYNC_JF, nullptr);
}

void vpb_synchronize_rho(vpb_domain_t *dom, vpb_field_t *d_f) {
  CHECK_FG();
  faces_local_adjust(dom, d_f, ADJ_RHOF);
  faces_local_adjust(dom, d_f, ADJ_RHOB);
  faces_sync_passes(dom, d_f, MSG_SYNC_RHO, nullptr);
}

void vpb_synchronize_tang_e_norm_b(vpb_domain_t *dom, vpb_field_t *d_f, double *d_err) {
  CHECK_FG();
  if (d_err) VPB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(double), st));
  faces_local_adjust(dom, d_f, ADJ_TANG_E);
  faces_local_adjust(dom, d_f, ADJ_NORM_B);
  faces_sync_passes(dom, d_f, MSG_SYNC_TEB, d_err);
}

void vpb_energy_f(vpb_domain_t *dom, const vpb_field_t *d_f, const vpb_material_coefficient_t *d_m, int n_mat, double *d_en6) {
  CHECK_FG();
  if (!d_m) { d_m = uniform_vacuum(); n_mat = 1; }
  VPB_CUDA(cudaMemsetAsync(d_en6, 0, 6 * sizeof(double), st));
  if (n_mat <= 1) energy_f_kernel<true><<<grid, tb, 0, st>>>(d_f, d_m, g, d_en6);
  else energy_f_kernel<false><<<grid, tb, 0, st>>>(d_f, d_m, g, d_en6);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

void vpb_compute_div_e_err(vpb_domain_t *dom, vpb_field_t *d_f, const vpb_material_coefficient_t *d_m, int n_mat) {
  CHECK_FG();
  if (!d_m) { d_m = uniform_vacuum(); n_mat = 1; }
  faces_ghost_exchange(dom, d_f, MSG_GHOST_NORM_E);
  const float px = (g.nx > 1) ? g.rdx : 0, py = (g.ny > 1) ? g.rdy : 0, pz = (g.nz > 1) ? g.rdz : 0;
  const float cj = (float)(1. / g.eps0);
  if (n_mat <= 1) div_e_kernel<0, true><<<grid, tb, 0, st>>>(d_f, d_m, g, px, py, pz, cj);
  else div_e_kernel<0, false><<<grid, tb, 0, st>>>(d_f, d_m, g, px, py, pz, cj);
  count_launch();
  faces_local_adjust(dom, d_f, ADJ_DIV_E);
  VPB_CUDA(cudaGetLastError());
}

void vpb_compute_rhob(vpb_domain_t *dom, vpb_field_t *d_f, const vpb_material_coefficient_t *d_m, int n_mat) {
  CHECK_FG();
  if (!d_m) { d_m = uniform_vacuum(); n_mat = 1; }
  faces_ghost_exchange(dom, d_f, MSG_GHOST_NORM_E);
  const float px = (g.nx > 1) ? g.eps0 * g.rdx : 0, py = (g.ny > 1) ? g.eps0 * g.rdy : 0, pz = (g.nz > 1) ? g.eps0 * g.rdz : 0;
  if (n_mat <= 1) div_e_kernel<1, true><<<grid, tb, 0, st>>>(d_f, d_m, g, px, py, pz, 0.f);
  else div_e_kernel<1, false><<<grid, tb, 0, st>>>(d_f, d_m, g, px, py, pz, 0.f);
  count_launch();
  faces_local_adjust(dom, d_f, ADJ_RHOB);
  VPB_CUDA(cudaGetLastError());
}

static void marder(const DomainDev &g, float p[3]) {   // clean_div_e.c:39-45
  p[0] = (g.nx > 1) ? g.rdx : 0; p[1] = (g.ny > 1) ? g.rdy : 0; p[2] = (g.nz > 1) ? g.rdz : 0;
  const float alphadt = (float)(0.3888889 / (p[0] * p[0] + p[1] * p[1] + p[2] * p[2]));
  p[0] *= alphadt; p[1] *= alphadt; p[2] *= alphadt;
}

void vpb_clean_div_e(vpb_domain_t *dom, vpb_field_t *d_f, const vpb_material_coefficient_t *d_m, int n_mat) {
  CHECK_FG();
  if (!d_m) { d_m = uniform_vacuum(); n_mat = 1; }
  float p[3];
  marder(g, p);
  if (n_mat <= 1) clean_div_e_kernel<true><<<grid, tb, 0, st>>>(d_f, d_m, g, p[0], p[1], p[2]);
  else clean_div_e_kernel<false><<<grid, tb, 0, st>>>(d_f, d_m, g, p[0], p[1], p[2]);
  count_launch();
  faces_local_adjust(dom, d_f, ADJ_TANG_E);
  VPB_CUDA(cudaGetLastError());
}

void vpb_compute_div_b_err(vpb_domain_t *dom, vpb_field_t *d_f) {
  CHECK_FG();
  const float px = (g.nx > 1) ? g.rdx : 0, py = (g.ny > 1) ? g.rdy : 0, pz = (g.nz > 1) ? g.rdz : 0;
  div_b_kernel<<<grid, tb, 0, st>>>(d_f, g, px, py, pz);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

void vpb_clean_div_b(vpb_domain_t *dom, vpb_field_t *d_f) {
  CHECK_FG();
  float p[3];
  marder(g, p);
  faces_ghost_exchange(dom, d_f, MSG_GHOST_DIV_B);
  clean_div_b_kernel<<<grid, tb, 0, st>>>(d_f, g, p[0], p[1], p[2]);
  count_launch();
  faces_local_adjust(dom, d_f, ADJ_NORM_B);
  VPB_CUDA(cudaGetLastError());
}

// d_out[0] = this rank's sum(err^2 * weights) (unscaled); the caller finishes
// eps0*sqrt(sum*dV / (nx*ny*nz*dV)) after the allreduce (compute_rms_div_e_err.c:150-160)
void vpb_compute_rms_div_e_err(vpb_domain_t *dom, const vpb_field_t *d_f, double *d_out) {
  CHECK_FG();
  VPB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double), st));
  rms_kernel<<<grid, tb, 0, st>>>(d_f, g, 0, d_out);
  count_launch();
}

void vpb_compute_rms_div_b_err(vpb_domain_t *dom, const vpb_field_t *d_f, double *d_out) {
  CHECK_FG();
  VPB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double), st));
  rms_kernel<<<grid, tb, 0, st>>>(d_f, g, 1, d_out);
  count_launch();
}

}  // extern "C"
