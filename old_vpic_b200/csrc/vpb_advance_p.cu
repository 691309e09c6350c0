// vpb_advance_p.cu -- K1/K2: Boris push + 18-coefficient gather + charge-conserving
// deposit + in-kernel cell crossing (move_p) + ordered mover emission.
//
// Replaces src/species_advance/standard/advance_p.cxx:9-183,399-472 and
// move_p.c:20-136.  Arithmetic follows the reference's SCALAR pipeline operation
// by operation (this file is compiled with -fmad=false, IEEE sqrt and divide), so
// every particle's dx,dy,dz,i,ux,uy,uz and every mover come out bit-identical to
// the reference's scalar flavour; only the summation ORDER into an accumulator
// cell differs (atomics), which the reference itself does not fix across -tpp.
//
// Layout (DESIGN.md "advance_p"): one CTA walks 256-particle tiles of the 48-byte
// AoS array.  Phase 1: one particle per thread, two 128-bit loads of the 32 hot
// bytes, interpolator through the read-only path (voxel-sorted particles make
// that a warp-wide broadcast out of L1), push, in-cell particles are written
// back with two 128-bit stores and their 12 current contributions are combined
// across the warp (segmented shuffle reduction over runs of equal voxel) before
// three REDG.128 per run reach the accumulator in L2.  Out-of-cell particles are
// parked in a shared-memory queue.  Phase 2: the queue is drained densely, one
// mover per thread (move_p), so the data-dependent streak loop does not run at
// 1/8 lane efficiency inside phase 1.  Movers that hit something move_p cannot
// resolve are emitted in increasing particle index (boundary_p.c:168-176 needs
// that): in-tile rank by block scan, tile order by a post-pass over per-tile
// counts.
#include "vpb_common.cuh"
#include "vpb_scan.cuh"

namespace vpb {

constexpr int kTile = 256;

struct AdvanceArgs {
  vpb_particle_t *p;
  int np;
  int ntiles;
  float qdt_2mc, cdt_dx, cdt_dy, cdt_dz;
  float *a;                       // accumulator_t[nv] viewed as float[12*nv]
  const vpb_interpolator_t *f;
  const int32_t *nbr;
  vpb_particle_mover_t *tmp_pm;   // unordered staging, capacity max_nm
  int max_nm;
  int *counters;                  // [0] staging cursor  [1] movers ignored (overflow)
  int2 *tile_info;                // per tile: (start in tmp_pm, count)
};

// One quadrant set of the charge-conserving deposit (advance_p.cxx:136-155 /
// move_p.c:73-92): X is the current direction, Y,Z the transverse ones.
__device__ __forceinline__ void accumulate_j(float q, float uX, float dY, float dZ, float v5, float &o0, float &o1,
                                             float &o2, float &o3) {
  float v0, v1, v2, v3, v4;
  v4 = q * uX;
  v1 = v4 * dY;
  v0 = v4 - v1;
  v1 += v4;
  v4 = 1.0f + dZ;
  v2 = v0 * v4;
  v3 = v1 * v4;
  v4 = 1.0f - dZ;
  v0 *= v4;
  v1 *= v4;
  v0 += v5;
  v1 -= v5;
  v2 -= v5;
  v3 += v5;
  o0 = v0; o1 = v1; o2 = v2; o3 = v3;
}

struct Mover {
  float dx, dy, dz;
  int i;
  float ux, uy, uz, q;
  float dispx, dispy, dispz;
};

// move_p.c:20-136 on registers.  Returns 1 if the mover is still in use.
__device__ __noinline__ int move_p_dev(Mover &s, float *__restrict__ a0, const int32_t *__restrict__ nbr) {
  for (;;) {
    float s_midx = s.dx, s_midy = s.dy, s_midz = s.dz;
    float s_dispx = s.dispx, s_dispy = s.dispy, s_dispz = s.dispz;
    const float dirx = (s_dispx > 0) ? 1.0f : -1.0f;
    const float diry = (s_dispy > 0) ? 1.0f : -1.0f;
    const float dirz = (s_dispz > 0) ? 1.0f : -1.0f;
    const float big = (float)3.4e38;
    float v0 = (s_dispx == 0) ? big : (dirx - s_midx) / s_dispx;
    float v1 = (s_dispy == 0) ? big : (diry - s_midy) / s_dispy;
    float v2 = (s_dispz == 0) ? big : (dirz - s_midz) / s_dispz;
    float v3 = 2.0f;
    int type = 3;
    if (v0 < v3) { v3 = v0; type = 0; }
    if (v1 < v3) { v3 = v1; type = 1; }
    if (v2 < v3) { v3 = v2; type = 2; }
    v3 *= 0.5f;
    s_dispx *= v3; s_dispy *= v3; s_dispz *= v3;
    s_midx += s_dispx; s_midy += s_dispy; s_midz += s_dispz;
    // the reference multiplies by the DOUBLE constant (1./3.) here (move_p.c:71)
    const float v5 = (float)((double)(((s.q * s_dispx) * s_dispy) * s_dispz) * (1. / 3.));
    float *a = a0 + 12 * (size_t)s.i;
    float o0, o1, o2, o3;
    accumulate_j(s.q, s_dispx, s_midy, s_midz, v5, o0, o1, o2, o3);
    red_add_v4(a, o0, o1, o2, o3);
    accumulate_j(s.q, s_dispy, s_midz, s_midx, v5, o0, o1, o2, o3);
    red_add_v4(a + 4, o0, o1, o2, o3);
    accumulate_j(s.q, s_dispz, s_midx, s_midy, v5, o0, o1, o2, o3);
    red_add_v4(a + 8, o0, o1, o2, o3);
    s.dispx -= s_dispx; s.dispy -= s_dispy; s.dispz -= s_dispz;
    s.dx += s_dispx + s_dispx; s.dy += s_dispy + s_dispy; s.dz += s_dispz + s_dispz;
    if (type == 3) return 0;
    const float dir = (type == 0) ? dirx : (type == 1) ? diry : dirz;
    const int n = __ldg(nbr + 6 * (size_t)s.i + ((dir > 0) ? 3 : 0) + type);
    if (n < 0) {  // hit a boundary: put the particle exactly on it
      if (type == 0) s.dx = dir; else if (type == 1) s.dy = dir; else s.dz = dir;
      if (n != -1) return 1;  // only reflection is resolved locally
      if (type == 0) { s.ux = -s.ux; s.dispx = -s.dispx; }
      else if (type == 1) { s.uy = -s.uy; s.dispy = -s.dispy; }
      else { s.uz = -s.uz; s.dispz = -s.dispz; }
    } else {
      s.i = n;
      if (type == 0) s.dx = -dir; else if (type == 1) s.dy = -dir; else s.dz = -dir;
    }
  }
}

// Segmented warp reduction of the 12 deposit values over runs of consecutive
// lanes that share a voxel; the head lane of each run issues three REDG.128.
__device__ __forceinline__ void deposit_runs(float (&v)[12], int key, bool active, float *__restrict__ a0) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int k = active ? key : -1;
  const int kprev = __shfl_up_sync(full, k, 1);
  const unsigned heads = __ballot_sync(full, lane == 0 || k != kprev);
  if (heads != full) {
    const int nxt = __ffs((heads >> lane) >> 1);
    const int run_end = nxt ? lane + nxt - 1 : 31;
    const int maxlen = __reduce_max_sync(full, run_end - lane + 1);
    for (int d = 1; d < maxlen; d <<= 1) {
      const bool take = (lane + d <= run_end);
#pragma unroll
      for (int c = 0; c < 12; c++) {
        const float t = __shfl_down_sync(full, v[c], d);
        if (take) v[c] += t;
      }
    }
  }
  if (active && ((heads >> lane) & 1u)) {
    float *a = a0 + 12 * (size_t)key;
    red_add_v4(a, v[0], v[1], v[2], v[3]);
    red_add_v4(a + 4, v[4], v[5], v[6], v[7]);
    red_add_v4(a + 8, v[8], v[9], v[10], v[11]);
  }
}

template <int DEPOSIT>  // 0: one REDG.128 triple per particle, 1: warp run reduction first
__global__ void __launch_bounds__(kTile, 3) advance_p_kernel(const AdvanceArgs A) {
  __shared__ float4 q_pos[kTile];    // dx,dy,dz,i
  __shared__ float4 q_mom[kTile];    // ux,uy,uz,q
  __shared__ float4 q_disp[kTile];   // dispx,dispy,dispz, in-tile index
  __shared__ int mv_slot[kTile];     // in-tile particle index -> queue slot of its unresolved mover
  __shared__ int q_n;
  __shared__ int warp_cnt[kTile / 32];
  __shared__ int tile_start;

  const int tid = threadIdx.x;
  const float one = 1.f;
  const float one_third = (float)(1. / 3.);
  const float two_fifteenths = (float)(2. / 15.);
  const float qdt_2mc = A.qdt_2mc, cdt_dx = A.cdt_dx, cdt_dy = A.cdt_dy, cdt_dz = A.cdt_dz;

  for (int tile = blockIdx.x; tile < A.ntiles; tile += gridDim.x) {
    if (tid == 0) q_n = 0;
    mv_slot[tid] = -1;
    __syncthreads();

    const int k = tile * kTile + tid;
    const bool valid = k < A.np;
    bool inbnds = false;
    int ii = 0;
    float dep[12];
#pragma unroll
    for (int c = 0; c < 12; c++) dep[c] = 0.f;

    if (valid) {
      float4 *pp = reinterpret_cast<float4 *>(A.p + k);
      const float4 r0 = pp[0];
      const float4 r1 = pp[1];
      float dx = r0.x, dy = r0.y, dz = r0.z;
      ii = __float_as_int(r0.w);
      const char *fp = reinterpret_cast<const char *>(A.f + ii);
      const float4 fe_x = ldg4(fp);        // ex dexdy dexdz d2exdydz
      const float4 fe_y = ldg4(fp + 16);   // ey deydz deydx d2eydzdx
      const float4 fe_z = ldg4(fp + 32);   // ez dezdx dezdy d2ezdxdy
      const float4 fb_0 = ldg4(fp + 48);   // cbx dcbxdx cby dcbydy
      const float2 fb_1 = ldg2(fp + 64);   // cbz dcbzdz
      const float hax = qdt_2mc * ((fe_x.x + dy * fe_x.y) + dz * (fe_x.z + dy * fe_x.w));
      const float hay = qdt_2mc * ((fe_y.x + dz * fe_y.y) + dx * (fe_y.z + dz * fe_y.w));
      const float haz = qdt_2mc * ((fe_z.x + dx * fe_z.y) + dy * (fe_z.z + dx * fe_z.w));
      const float cbx = fb_0.x + dx * fb_0.y;
      const float cby = fb_0.z + dy * fb_0.w;
      const float cbz = fb_1.x + dz * fb_1.y;
      float ux = r1.x, uy = r1.y, uz = r1.z;
      const float q = r1.w;
      ux += hax; uy += hay; uz += haz;
      float v0 = qdt_2mc / sqrtf(one + (ux * ux + (uy * uy + uz * uz)));
      float v1 = cbx * cbx + (cby * cby + cbz * cbz);
      float v2 = (v0 * v0) * v1;
      float v3 = v0 * (one + v2 * (one_third + v2 * two_fifteenths));
      float v4 = v3 / (one + v1 * (v3 * v3));
      v4 += v4;
      v0 = ux + v3 * (uy * cbz - uz * cby);
      v1 = uy + v3 * (uz * cbx - ux * cbz);
      v2 = uz + v3 * (ux * cby - uy * cbx);
      ux += v4 * (v1 * cbz - v2 * cby);
      uy += v4 * (v2 * cbx - v0 * cbz);
      uz += v4 * (v0 * cby - v1 * cbx);
      ux += hax; uy += hay; uz += haz;
      const float4 mom = make_float4(ux, uy, uz, q);   // stored momentum
      v0 = one / sqrtf(one + (ux * ux + (uy * uy + uz * uz)));
      ux *= cdt_dx; uy *= cdt_dy; uz *= cdt_dz;
      ux *= v0; uy *= v0; uz *= v0;
      v0 = dx + ux; v1 = dy + uy; v2 = dz + uz;      // streak midpoint
      v3 = v0 + ux; v4 = v1 + uy;                     // new position
      float v5 = v2 + uz;
      inbnds = v3 <= one && v4 <= one && v5 <= one && -v3 <= one && -v4 <= one && -v5 <= one;
      if (inbnds) {
        pp[0] = make_float4(v3, v4, v5, r0.w);
        pp[1] = mom;
        dx = v0; dy = v1; dz = v2;
        v5 = q * ux * uy * uz * one_third;
        accumulate_j(q, ux, dy, dz, v5, dep[0], dep[1], dep[2], dep[3]);
        accumulate_j(q, uy, dz, dx, v5, dep[4], dep[5], dep[6], dep[7]);
        accumulate_j(q, uz, dx, dy, v5, dep[8], dep[9], dep[10], dep[11]);
      } else {
        const int slot = atomicAdd(&q_n, 1);
        q_pos[slot] = r0;
        q_mom[slot] = mom;
        q_disp[slot] = make_float4(ux, uy, uz, __int_as_float(tid));
      }
    }

    if (DEPOSIT == 0) {
      if (inbnds) {
        float *a = A.a + 12 * (size_t)ii;
        red_add_v4(a, dep[0], dep[1], dep[2], dep[3]);
        red_add_v4(a + 4, dep[4], dep[5], dep[6], dep[7]);
        red_add_v4(a + 8, dep[8], dep[9], dep[10], dep[11]);
      }
    } else {
      deposit_runs(dep, ii, inbnds, A.a);
    }

    __syncthreads();
    const int nq = q_n;
    int unresolved = 0;
    if (tid < nq) {   // phase 2: one mover per thread (nq <= kTile by construction)
      Mover s;
      const float4 a = q_pos[tid], b = q_mom[tid], c = q_disp[tid];
      s.dx = a.x; s.dy = a.y; s.dz = a.z; s.i = __float_as_int(a.w);
      s.ux = b.x; s.uy = b.y; s.uz = b.z; s.q = b.w;
      s.dispx = c.x; s.dispy = c.y; s.dispz = c.z;
      const int t = __float_as_int(c.w);
      unresolved = move_p_dev(s, A.a, A.nbr);
      float4 *pp = reinterpret_cast<float4 *>(A.p + (size_t)tile * kTile + t);
      pp[0] = make_float4(s.dx, s.dy, s.dz, __int_as_float(s.i));
      pp[1] = make_float4(s.ux, s.uy, s.uz, s.q);
      if (unresolved) {
        q_disp[tid] = make_float4(s.dispx, s.dispy, s.dispz, c.w);
        mv_slot[t] = tid;
      }
    }
    // ordered mover emission; the common case (nothing unresolved) costs one barrier
    const int any = __syncthreads_or(unresolved);
    if (any) {
      const int lane = tid & 31, w = tid >> 5;
      const int slot = mv_slot[tid];
      const unsigned b = __ballot_sync(0xffffffffu, slot >= 0);
      if (lane == 0) warp_cnt[w] = __popc(b);
      __syncthreads();
      int before = 0, total = 0;
#pragma unroll
      for (int j = 0; j < kTile / 32; j++) {
        const int cj = warp_cnt[j];
        if (j < w) before += cj;
        total += cj;
      }
      if (tid == 0) {
        tile_start = atomicAdd(&A.counters[0], total);
        A.tile_info[tile] = make_int2(tile_start, total);
      }
      __syncthreads();
      if (slot >= 0) {
        const int dst = tile_start + before + __popc(b & ((1u << lane) - 1u));
        if (dst < A.max_nm) {
          const float4 c = q_disp[slot];
          reinterpret_cast<float4 *>(A.tmp_pm)[dst] = make_float4(c.x, c.y, c.z, __int_as_float(tile * kTile + tid));
        } else {
          atomicAdd(&A.counters[1], 1);
        }
      }
      __syncthreads();
    } else if (tid == 0) {
      A.tile_info[tile] = make_int2(0, 0);
    }
  }
}

// Post-pass: tmp_pm holds each tile's movers contiguously (ordered inside the
// tile) but tiles landed in completion order.  tile_off = exclusive scan of the
// per-tile counts gives the final, particle-index-ordered position.
__global__ void gather_movers_kernel(const int2 *__restrict__ tile_info, const int *__restrict__ tile_off, int ntiles,
                                     const float4 *__restrict__ tmp_pm, float4 *__restrict__ pm, int max_nm,
                                     const int *__restrict__ counters, int *__restrict__ d_nm) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0 && d_nm) {
    const int total = counters[0];
    *d_nm = total < max_nm ? total : max_nm;
  }
  if (t >= ntiles) return;
  const int2 info = tile_info[t];
  const int off = tile_off[t];
  for (int j = 0; j < info.y; j++) {
    const int src = info.x + j, dst = off + j;
    if (src < max_nm && dst < max_nm) pm[dst] = tmp_pm[src];
  }
}

__global__ void extract_tile_counts_kernel(const int2 *__restrict__ tile_info, int *__restrict__ cnt, int ntiles) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < ntiles) cnt[t] = tile_info[t].y;
}

}  // namespace vpb

using namespace vpb;

extern "C" void vpb_advance_p(vpb_domain_t *dom, vpb_particle_t *d_p, int np, float q_m, vpb_particle_mover_t *d_pm,
                              int max_nm, vpb_accumulator_t *d_a, const vpb_interpolator_t *d_f, int *d_nm) {
  if (!dom) VPB_ERROR("Bad grid");
  if (!d_p) VPB_ERROR("Bad particle array");
  if (np < 0) VPB_ERROR("Bad number of particles");
  if (!d_pm) VPB_ERROR("Bad particle mover");
  if (max_nm < 0) VPB_ERROR("Bad number of movers");
  if (!d_a) VPB_ERROR("Bad accumulator");
  if (!d_f) VPB_ERROR("Bad interpolator");
  Context &c = ctx();
  const DomainDev &g = dom->d;
  if (np == 0) {
    if (d_nm) VPB_CUDA(cudaMemsetAsync(d_nm, 0, sizeof(int), c.stream));
    return;
  }
  AdvanceArgs A;
  A.p = d_p;
  A.np = np;
  A.ntiles = (np + kTile - 1) / kTile;
  // same expressions, same types as advance_p.cxx:425-428
  A.qdt_2mc = (float)(0.5 * q_m * g.dt / g.cvac);
  A.cdt_dx = g.cvac * g.dt * g.rdx;
  A.cdt_dy = g.cvac * g.dt * g.rdy;
  A.cdt_dz = g.cvac * g.dt * g.rdz;
  A.a = reinterpret_cast<float *>(d_a);
  A.f = d_f;
  A.nbr = g.nbr;
  A.max_nm = max_nm;
  // scratch: counters | tile_info | tile_cnt | tile_off | tmp_pm | scan scratch
  const size_t nt = (size_t)A.ntiles;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t off_info = 256, off_cnt = off_info + al(nt * sizeof(int2)), off_off = off_cnt + al(nt * sizeof(int)),
               off_tmp = off_off + al(nt * sizeof(int)), off_scan = off_tmp + al((size_t)max_nm * 16 + 16);
  char *s = (char *)scratch(off_scan + scan_scratch_bytes(A.ntiles));
  A.counters = (int *)s;
  A.tile_info = (int2 *)(s + off_info);
  int *tile_cnt = (int *)(s + off_cnt);
  int *tile_off = (int *)(s + off_off);
  A.tmp_pm = (vpb_particle_mover_t *)(s + off_tmp);
  void *scan_tmp = s + off_scan;
  VPB_CUDA(cudaMemsetAsync(A.counters, 0, 256, c.stream));

  const int per_sm = tuning("advance_p.ctas_per_sm", 6);
  int grid = c.sm_count * per_sm;
  if (grid > A.ntiles) grid = A.ntiles;
  {
    ProfScope prof(0);
    if (tuning("advance_p.deposit", 1) == 0)
      advance_p_kernel<0><<<grid, kTile, 0, c.stream>>>(A);
    else
      advance_p_kernel<1><<<grid, kTile, 0, c.stream>>>(A);
  }
  const int tb = 256, tg = (A.ntiles + tb - 1) / tb;
  extract_tile_counts_kernel<<<tg, tb, 0, c.stream>>>(A.tile_info, tile_cnt, A.ntiles);
  exclusive_scan_i32(tile_cnt, tile_off, A.ntiles, scan_tmp, c.stream);
  gather_movers_kernel<<<tg, tb, 0, c.stream>>>(A.tile_info, tile_off, A.ntiles, (const float4 *)A.tmp_pm, (float4 *)d_pm,
                                                max_nm, A.counters, d_nm);
  count_launch(3 + scan_launches(A.ntiles));
  VPB_CUDA(cudaGetLastError());
}
