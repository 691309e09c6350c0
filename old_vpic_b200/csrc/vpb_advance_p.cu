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
// Structure (DESIGN.md "advance_p"; profiles/ has the ncu history that led here):
//  * every WARP is independent -- no CTA barrier anywhere.  A warp walks chunks of
//    32 consecutive particles of the 48-byte AoS array; consecutive warps take
//    consecutive chunks, so a CTA streams a contiguous 12 KB window.
//  * phase 1, one particle per lane: two 128-bit loads of the 32 hot bytes, the
//    interpolator through the read-only path (voxel-sorted particles make that a
//    warp-wide broadcast out of L1), the push, two 128-bit stores, and the 12
//    current contributions combined across the warp by a segmented shuffle
//    reduction over runs of equal voxel before three REDG.128 per run reach L2.
//  * particles that leave their cell are appended to a per-warp shared-memory
//    ring; whenever it holds 32 of them the warp runs move_p on a FULL warp of
//    movers (the data-dependent streak loop would otherwise run at ~1/8 lane
//    efficiency inside phase 1).
//  * movers that hit something move_p cannot resolve are staged unordered and a
//    bit is set for their particle index; a post-pass ranks them by prefix
//    popcount so pm[] comes out in increasing particle index, which
//    boundary_p.c:168-176 relies on.  The post-pass kernels return at once when
//    nothing was staged (the usual case on a periodic or reflecting rank).
#include "vpb_advance_p.cuh"
#include "vpb_move_p.cuh"
#include "vpb_scan.cuh"

namespace vpb {

constexpr int kWarps = 8;            // warps per CTA
constexpr int kQueue = 64;           // ring capacity per warp (>= 31 + 32)

// Which chunks does work item t own?  A chunk belongs to the row that contains its first particle.
// Rows are walked in blocks of `by` consecutive y, all z for a block before the next block, so that a
// voxel and its z+-1 neighbours are touched within by rows (a few MB of particle stream) of each other
// and their interpolator/accumulator lines are still in L2 -- the array position of a particle only
// changes at a sort, its voxel drifts by about a cell every ten steps.
__device__ __forceinline__ void work_range(const AdvanceArgs &A, int t, int &c0, int &c1) {
  if (!A.partition) {
    c0 = A.chunk_lo + t * 64;
    c1 = c0 + 64 < A.chunk_hi ? c0 + 64 : A.chunk_hi;
    return;
  }
  int lo, hi;
  if (t == A.nwork - 1) {                 // particles appended since the sort
    lo = A.partition[A.nv];
    hi = A.np;
  } else {
    const int per_block = A.by * A.sz;
    const int yb = t / per_block, rem = t - yb * per_block;
    const int z = rem / A.by, y = yb * A.by + (rem - z * A.by);
    if (y >= A.sy) { c0 = c1 = 0; return; }
    const int v0 = A.sx * (y + A.sy * z);
    lo = A.partition[v0];
    hi = A.partition[v0 + A.sx];
  }
  if (lo > A.np) lo = A.np;
  if (hi > A.np) hi = A.np;
  c0 = (lo + 31) >> 5;
  c1 = (hi + 31) >> 5;
}

__device__ __forceinline__ void red3(float *a, const float (&v)[12]) {
  red_add_v4(a, v[0], v[1], v[2], v[3]);
  red_add_v4(a + 4, v[4], v[5], v[6], v[7]);
  red_add_v4(a + 8, v[8], v[9], v[10], v[11]);
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
  if (active && ((heads >> lane) & 1u)) red3(a0 + 12 * (size_t)key, v);
}

// Run move_p on up to 32 queued movers (one per lane), write the particles back and
// stage the unresolved ones.
__device__ __noinline__ void drain_movers(vpb_particle_t *__restrict__ p, float *__restrict__ acc, const int32_t *__restrict__ nbr,
                                          vpb_particle_mover_t *__restrict__ tmp_pm, int max_nm, int *__restrict__ counters,
                                          unsigned *__restrict__ bitmap, const float4 *q_pos, const float4 *q_mom,
                                          const float4 *q_disp, int head, int count) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int unresolved = 0, k = 0;
  Mover s;
  s.dispx = s.dispy = s.dispz = 0.f;
  if (lane < count) {
    const int e = (head + lane) & (kQueue - 1);
    const float4 a = q_pos[e], b = q_mom[e], c = q_disp[e];
    s.dx = a.x; s.dy = a.y; s.dz = a.z; s.i = __float_as_int(a.w);
    s.ux = b.x; s.uy = b.y; s.uz = b.z; s.q = b.w;
    s.dispx = c.x; s.dispy = c.y; s.dispz = c.z;
    k = __float_as_int(c.w);
    unresolved = move_p_dev(s, acc, nbr);
    float4 *pp = reinterpret_cast<float4 *>(p + k);
    pp[0] = make_float4(s.dx, s.dy, s.dz, __int_as_float(s.i));
    pp[1] = make_float4(s.ux, s.uy, s.uz, s.q);
  }
  const unsigned um = __ballot_sync(full, unresolved);
  if (um) {   // rare: stage {remaining displacement, particle index}, mark the particle
    int base = 0;
    if (lane == 0) base = atomicAdd(&counters[0], __popc(um));
    base = __shfl_sync(full, base, 0);
    if (unresolved) {
      const int dst = base + __popc(um & ((1u << lane) - 1u));
      if (dst < max_nm) {
        reinterpret_cast<float4 *>(tmp_pm)[dst] = make_float4(s.dispx, s.dispy, s.dispz, __int_as_float(k));
        atomicOr(&bitmap[k >> 5], 1u << (k & 31));
      } else {
        atomicAdd(&counters[1], 1);
      }
    }
  }
}

// DEPOSIT 0: one REDG.128 triple per particle, 1: warp run reduction first.
template <int DEPOSIT>
__global__ void __launch_bounds__(kWarps * 32, 4) advance_p_kernel(const AdvanceArgs A) {
  __shared__ float4 q_pos[kWarps][kQueue];    // dx,dy,dz,i
  __shared__ float4 q_mom[kWarps][kQueue];    // ux,uy,uz,q
  __shared__ float4 q_disp[kWarps][kQueue];   // dispx,dispy,dispz, particle index

  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float one = 1.f;
  const float one_third = (float)(1. / 3.);
  const float qdt_2mc = A.qdt_2mc, cdt_dx = A.cdt_dx, cdt_dy = A.cdt_dy, cdt_dz = A.cdt_dz;
  int q_head = 0, q_n = 0;   // warp-uniform ring state

  for (int t = blockIdx.x; t < A.nwork; t += gridDim.x) {
    int c_begin, c_end;
    work_range(A, t, c_begin, c_end);
    for (int chunk = c_begin + w; chunk < c_end; chunk += kWarps) {
      const int k = chunk * 32 + lane;
      const bool valid = k < A.np;
      bool inbnds = false, outbnds = false;
      int ii = 0;
      float dep[12];
#pragma unroll
      for (int c = 0; c < 12; c++) dep[c] = 0.f;
      float4 r0 = make_float4(0, 0, 0, 0), mom = r0;
      float hx = 0, hy = 0, hz = 0;

      if (valid) {
        float4 *pp = reinterpret_cast<float4 *>(A.p + k);
        r0 = pp[0];
        const float4 r1 = pp[1];
        float dx = r0.x, dy = r0.y, dz = r0.z;
        ii = __float_as_int(r0.w);
        const char *fp = reinterpret_cast<const char *>(A.f) + (size_t)ii * A.fi_bytes;
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
        const float two_fifteenths = (float)(2. / 15.);
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
        mom = make_float4(ux, uy, uz, q);   // stored momentum
        v0 = one / sqrtf(one + (ux * ux + (uy * uy + uz * uz)));
        ux *= cdt_dx; uy *= cdt_dy; uz *= cdt_dz;
        ux *= v0; uy *= v0; uz *= v0;
        v0 = dx + ux; v1 = dy + uy; v2 = dz + uz;      // streak midpoint
        v3 = v0 + ux; v4 = v1 + uy;                     // new position
        float v5 = v2 + uz;
        inbnds = v3 <= one && v4 <= one && v5 <= one && -v3 <= one && -v4 <= one && -v5 <= one;
        outbnds = !inbnds;
        if (inbnds) {
          pp[0] = make_float4(v3, v4, v5, r0.w);
          pp[1] = mom;
          dx = v0; dy = v1; dz = v2;
          v5 = q * ux * uy * uz * one_third;
          accumulate_j(q, ux, dy, dz, v5, dep[0], dep[1], dep[2], dep[3]);
          accumulate_j(q, uy, dz, dx, v5, dep[4], dep[5], dep[6], dep[7]);
          accumulate_j(q, uz, dx, dy, v5, dep[8], dep[9], dep[10], dep[11]);
        } else {
          hx = ux; hy = uy; hz = uz;
        }
      }

      if (DEPOSIT == 0) {
        if (inbnds) red3(A.a + 12 * (size_t)ii, dep);
      } else {
        deposit_runs(dep, ii, inbnds, A.a);
      }

      // park the out-of-cell particles (in particle order) in this warp's ring
      const unsigned om = __ballot_sync(full, outbnds);
      if (om) {
        if (outbnds) {
          const int e = (q_head + q_n + __popc(om & ((1u << lane) - 1u))) & (kQueue - 1);
          q_pos[w][e] = r0;
          q_mom[w][e] = mom;
          q_disp[w][e] = make_float4(hx, hy, hz, __int_as_float(k));
        }
        q_n += __popc(om);
        __syncwarp();
        if (q_n >= 32) {
          drain_movers(A.p, A.a, A.nbr, A.tmp_pm, A.max_nm, A.counters, A.bitmap, q_pos[w], q_mom[w], q_disp[w], q_head, 32);
          q_head = (q_head + 32) & (kQueue - 1);
          q_n -= 32;
          __syncwarp();
        }
      }
    }
  }
  if (q_n) drain_movers(A.p, A.a, A.nbr, A.tmp_pm, A.max_nm, A.counters, A.bitmap, q_pos[w], q_mom[w], q_disp[w], q_head, q_n);
}

// ---------------------------------------------------------------------------
// TMA variant.  The particle stream goes global -> shared -> global with 1-D bulk
// async copies (cp.async.bulk, SASS UBLKCP) carrying an L2 evict-first policy:
// each warp owns a ring of kStages 1536-byte tiles (32 particles x 48 B) with one
// mbarrier per tile, keeps kStages-1 tiles in flight ahead of the one it computes
// on, reads/writes its particle with conflict-free LDS.128/STS.128 (48-byte lane
// stride), and writes the whole tile back with one bulk store.  Full 128-byte
// lines travel to and from DRAM exactly once, and because the stream no longer
// competes for L2, the interpolator/accumulator lines the particles gather from
// and scatter to stay resident there even after the particles have drifted off
// their sorted voxels (profiles/: DRAM bytes per particle 195 -> ~100 when
// drifted).
// ---------------------------------------------------------------------------
constexpr int kStages = 3;
constexpr int kTileBytes = 32 * 48;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_load_tile(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_store_tile(void *dst, uint32_t src, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src), "r"(bytes), "l"(pol) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

struct TmaSmem {
  float4 tile[kWarps][kStages][96];
  float4 q_pos[kWarps][kQueue], q_mom[kWarps][kQueue], q_disp[kWarps][kQueue];
  uint64_t full[kWarps][kStages];
};

template <int DEPOSIT>
__global__ void __launch_bounds__(kWarps * 32, 3) advance_p_tma_kernel(const AdvanceArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TmaSmem &S = *reinterpret_cast<TmaSmem *>(smem_raw);

  const unsigned fullmask = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float one = 1.f;
  const float one_third = (float)(1. / 3.);
  const float qdt_2mc = A.qdt_2mc, cdt_dx = A.cdt_dx, cdt_dy = A.cdt_dy, cdt_dz = A.cdt_dz;
  int q_head = 0, q_n = 0;
  const uint64_t pol = l2_policy_evict_first();
  char *const gbase = reinterpret_cast<char *>(A.p);

  if (lane == 0) {
    for (int s = 0; s < kStages; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.full[w][s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  auto tile_bytes = [&](int chunk) -> uint32_t {
    const int n = A.np - chunk * 32;
    return (uint32_t)((n < 32 ? n : 32) * 48);
  };
  // Dynamic scheduling: chunks are handed out kGrab at a time from a global ticket counter, so the chunks in
  // flight on the chip are always the most recently issued ones however unevenly the warps progress.  With
  // a static stride the warps drift apart by whole z-planes within a launch and a voxel's interpolator /
  // accumulator lines are then touched at times too far apart to still be in L2 (profiles/r1f: 55-59 % L2
  // miss on gathers and REDs once particles have left their sorted voxels).
  constexpr int kGrab = 4;
  int g_cur = 0, g_end = 0;   // warp-uniform: current group of chunks [g_cur, g_end)
  auto next_chunk = [&]() -> int {
    if (g_cur >= g_end) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&A.counters[2], kGrab);
      base = __shfl_sync(fullmask, base, 0) + A.chunk_lo;
      g_cur = base;
      g_end = base + kGrab < A.chunk_hi ? base + kGrab : A.chunk_hi;
      if (g_cur >= g_end) { g_cur = g_end = A.chunk_hi; return -1; }
    }
    return g_cur++;
  };
  // chunks this warp will compute on next, oldest first; their tiles are in flight
  int pend[kStages - 1];
#pragma unroll
  for (int j = 0; j < kStages - 1; j++) {
    pend[j] = next_chunk();
    if (lane == 0 && pend[j] >= 0)
      tma_load_tile(smem_u32(&S.tile[w][j][0]), gbase + (size_t)pend[j] * kTileBytes, tile_bytes(pend[j]), smem_u32(&S.full[w][j]), pol);
  }

  for (int it = 0; pend[0] >= 0; ++it) {
    const int chunk = pend[0];
    const int stage = it % kStages;
    mbar_wait(smem_u32(&S.full[w][stage]), (uint32_t)((it / kStages) & 1));
    const int k = chunk * 32 + lane;
    const bool valid = k < A.np;
    bool inbnds = false, outbnds = false;
    int ii = 0;
    float dep[12];
#pragma unroll
    for (int c = 0; c < 12; c++) dep[c] = 0.f;
    float4 r0 = make_float4(0, 0, 0, 0), mom = r0;
    float hx = 0, hy = 0, hz = 0;
    float4 *tp = &S.tile[w][stage][lane * 3];

    if (valid) {
      r0 = tp[0];
      const float4 r1 = tp[1];
      float dx = r0.x, dy = r0.y, dz = r0.z;
      ii = __float_as_int(r0.w);
      const char *fp = reinterpret_cast<const char *>(A.f) + (size_t)ii * A.fi_bytes;
      const float4 fe_x = ldg4(fp);
      const float4 fe_y = ldg4(fp + 16);
      const float4 fe_z = ldg4(fp + 32);
      const float4 fb_0 = ldg4(fp + 48);
      const float2 fb_1 = ldg2(fp + 64);
      const float hax = qdt_2mc * ((fe_x.x + dy * fe_x.y) + dz * (fe_x.z + dy * fe_x.w));
      const float hay = qdt_2mc * ((fe_y.x + dz * fe_y.y) + dx * (fe_y.z + dz * fe_y.w));
      const float haz = qdt_2mc * ((fe_z.x + dx * fe_z.y) + dy * (fe_z.z + dx * fe_z.w));
      const float cbx = fb_0.x + dx * fb_0.y;
      const float cby = fb_0.z + dy * fb_0.w;
      const float cbz = fb_1.x + dz * fb_1.y;
      float ux = r1.x, uy = r1.y, uz = r1.z;
      const float q = r1.w;
      ux += hax; uy += hay; uz += haz;
      const float two_fifteenths = (float)(2. / 15.);
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
      mom = make_float4(ux, uy, uz, q);
      v0 = one / sqrtf(one + (ux * ux + (uy * uy + uz * uz)));
      ux *= cdt_dx; uy *= cdt_dy; uz *= cdt_dz;
      ux *= v0; uy *= v0; uz *= v0;
      v0 = dx + ux; v1 = dy + uy; v2 = dz + uz;
      v3 = v0 + ux; v4 = v1 + uy;
      float v5 = v2 + uz;
      inbnds = v3 <= one && v4 <= one && v5 <= one && -v3 <= one && -v4 <= one && -v5 <= one;
      outbnds = !inbnds;
      if (inbnds) {
        tp[0] = make_float4(v3, v4, v5, r0.w);
        tp[1] = mom;
        dx = v0; dy = v1; dz = v2;
        v5 = q * ux * uy * uz * one_third;
        accumulate_j(q, ux, dy, dz, v5, dep[0], dep[1], dep[2], dep[3]);
        accumulate_j(q, uy, dz, dx, v5, dep[4], dep[5], dep[6], dep[7]);
        accumulate_j(q, uz, dx, dy, v5, dep[8], dep[9], dep[10], dep[11]);
      } else {
        hx = ux; hy = uy; hz = uz;
      }
    }
    // the tile goes back as one bulk store (out-of-cell particles unchanged; drain_movers rewrites them later)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      tma_store_tile(gbase + (size_t)chunk * kTileBytes, smem_u32(&S.tile[w][stage][0]), tile_bytes(chunk), pol);
    }
    // take the next chunk and refill the stage used one iteration ago (its store must have finished
    // reading shared memory first)
    const int cn = next_chunk();
#pragma unroll
    for (int j = 0; j < kStages - 2; j++) pend[j] = pend[j + 1];
    pend[kStages - 2] = cn;
    if (lane == 0 && cn >= 0) {
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      const int sn = (it + kStages - 1) % kStages;
      tma_load_tile(smem_u32(&S.tile[w][sn][0]), gbase + (size_t)cn * kTileBytes, tile_bytes(cn), smem_u32(&S.full[w][sn]), pol);
    }

    if (DEPOSIT == 0) {
      if (inbnds) red3(A.a + 12 * (size_t)ii, dep);
    } else {
      deposit_runs(dep, ii, inbnds, A.a);
    }

    const unsigned om = __ballot_sync(fullmask, outbnds);
    if (om) {
      if (outbnds) {
        const int e = (q_head + q_n + __popc(om & ((1u << lane) - 1u))) & (kQueue - 1);
        S.q_pos[w][e] = r0;
        S.q_mom[w][e] = mom;
        S.q_disp[w][e] = make_float4(hx, hy, hz, __int_as_float(k));
      }
      q_n += __popc(om);
      __syncwarp();
      if (q_n >= 32) {
        // the movers' tiles must have landed in global memory before their final state is written over them
        if (lane == 0) {
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        __syncwarp();
        drain_movers(A.p, A.a, A.nbr, A.tmp_pm, A.max_nm, A.counters, A.bitmap, S.q_pos[w], S.q_mom[w], S.q_disp[w], q_head, 32);
        q_head = (q_head + 32) & (kQueue - 1);
        q_n -= 32;
        __syncwarp();
      }
    }
  }
  if (lane == 0) {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __syncwarp();
  if (q_n) drain_movers(A.p, A.a, A.nbr, A.tmp_pm, A.max_nm, A.counters, A.bitmap, S.q_pos[w], S.q_mom[w], S.q_disp[w], q_head, q_n);
}

// ---------------------------------------------------------------------------
// Streaming variant (default).  What the ncu captures of the kernel above showed (profiles/README.md):
// a fifth of the stall samples sat on the first use of the interpolator, a sixth on the scheduler's
// ticket atomic, an eighth on the proxy fence in front of the bulk store, and a third of the issued
// instructions were the segmented shuffle reduction.  Here
//  * particles still ARRIVE by bulk async copy (kStagesS tiles per warp in flight), but results leave
//    by two STG.128 per lane straight from registers (streaming policy): no proxy fence, no wait
//    before movers are drained, and the tile can be refilled as soon as it has been read;
//  * the interpolator of chunk n+1 is requested (voxel index read from its staged tile) BEFORE chunk n
//    is computed and held in registers across it;
//  * the next ticket is requested one grab ahead of its use;
//  * deposit: all in-cell lanes in one voxel (the usual case after a sort) -> halving butterfly
//    (54 instructions instead of ~180) ending in 12 scalar REDs from four lanes; many short runs
//    (drifted) -> no reduction at all, each lane issues its own three REDG.128; otherwise the
//    segmented reduction, with out-of-cell lanes made transparent so they do not split a run.
// ---------------------------------------------------------------------------
constexpr int kGrabS = 16;

struct Interp {
  float4 ex, ey, ez, b0;
  float2 b1;
};

// WIDE: 96-byte records (32-byte aligned): the 72 useful bytes arrive with two LDG.256 and one LDG.64, each
// of the record's three sectors requested once; otherwise the reference's 80-byte record, five loads.
template <int WIDE>
__device__ __forceinline__ void load_interp(Interp &I, const vpb_interpolator_t *f, int ii) {
  if (WIDE) {
    const char *fp = reinterpret_cast<const char *>(f) + (size_t)ii * 96;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(I.ex.x), "=f"(I.ex.y), "=f"(I.ex.z), "=f"(I.ex.w), "=f"(I.ey.x), "=f"(I.ey.y), "=f"(I.ey.z), "=f"(I.ey.w)
                 : "l"(fp));
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(I.ez.x), "=f"(I.ez.y), "=f"(I.ez.z), "=f"(I.ez.w), "=f"(I.b0.x), "=f"(I.b0.y), "=f"(I.b0.z), "=f"(I.b0.w)
                 : "l"(fp + 32));
    I.b1 = ldg2(fp + 64);
  } else {
    const char *fp = reinterpret_cast<const char *>(f + ii);
    I.ex = ldg4(fp);        // ex dexdy dexdz d2exdydz
    I.ey = ldg4(fp + 16);   // ey deydz deydx d2eydzdx
    I.ez = ldg4(fp + 32);   // ez dezdx dezdy d2ezdxdy
    I.b0 = ldg4(fp + 48);   // cbx dcbxdx cby dcbydy
    I.b1 = ldg2(fp + 64);   // cbz dcbzdz
  }
}

__device__ __forceinline__ void st_stream4(void *p, float4 v) { __stcs(reinterpret_cast<float4 *>(p), v); }

// Deposit for one chunk.  A few steps after a sort a warp's 32 particles are mostly still in one or two voxels
// with strays mixed in at random lanes, so runs of equal voxel are short even though one voxel dominates.
// The dominant voxel (the more populous of the first and the last in-cell lane's) is summed over the warp by
// a halving butterfly -- 18 shuffles instead of the 60 of a full 12-value reduction -- and leaves as 12 scalar
// REDs from four lanes; every other in-cell lane issues its own three REDG.128.
__device__ __forceinline__ void deposit_dominant(float (&v)[12], int key, bool active, float *__restrict__ a0) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned am = __ballot_sync(full, active);
  if (am == 0) return;
  const int ka = __shfl_sync(full, key, __ffs(am) - 1), kb = __shfl_sync(full, key, 31 - __clz(am));
  const unsigned ma = __ballot_sync(full, active && key == ka), mb = __ballot_sync(full, active && key == kb);
  const bool pick_a = __popc(ma) >= __popc(mb);
  const int k0 = pick_a ? ka : kb;
  const unsigned dm = pick_a ? ma : mb;
  const bool mine = (dm >> lane) & 1u;
  if (active && (!mine || __popc(dm) < 3)) red3(a0 + 12 * (size_t)key, v);
  if (__popc(dm) < 3) return;
  // component c = 4g+j of the dominant voxel ends up summed in the lanes with j = 2*bit4 + bit3
  const bool h16 = lane & 16, h8 = lane & 8;
  float w[6];
#pragma unroll
  for (int g = 0; g < 3; g++) {
#pragma unroll
    for (int jj = 0; jj < 2; jj++) {
      const float lo = mine ? v[4 * g + jj] : 0.f, hi = mine ? v[4 * g + 2 + jj] : 0.f;
      w[2 * g + jj] = (h16 ? hi : lo) + __shfl_xor_sync(full, h16 ? lo : hi, 16);
    }
  }
  float u[3];
#pragma unroll
  for (int g = 0; g < 3; g++) u[g] = (h8 ? w[2 * g + 1] : w[2 * g]) + __shfl_xor_sync(full, h8 ? w[2 * g] : w[2 * g + 1], 8);
#pragma unroll
  for (int d = 4; d >= 1; d >>= 1) {
#pragma unroll
    for (int g = 0; g < 3; g++) u[g] += __shfl_xor_sync(full, u[g], d);
  }
  if ((lane & 7) == 0) {
    float *a = a0 + 12 * (size_t)k0 + (lane >> 3);
    red_add(a, u[0]);
    red_add(a + 4, u[1]);
    red_add(a + 8, u[2]);
  }
}

// drain_movers for the streaming kernel: the updated momentum is read back from the particle array
// (stored by the main loop, as advance_p.cxx:131-133 does before it calls move_p)
__device__ __noinline__ void drain_movers_slim(vpb_particle_t *__restrict__ p, float *__restrict__ acc, const int32_t *__restrict__ nbr,
                                               vpb_particle_mover_t *__restrict__ tmp_pm, int max_nm, int *__restrict__ counters,
                                               unsigned *__restrict__ bitmap, const float4 *q_pos, const float4 *q_disp, int head,
                                               int count) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int unresolved = 0, k = 0;
  Mover s;
  s.dispx = s.dispy = s.dispz = 0.f;
  if (lane < count) {
    const int e = (head + lane) & (kQueue - 1);
    const float4 a = q_pos[e], c = q_disp[e];
    k = __float_as_int(c.w);
    float4 *pp = reinterpret_cast<float4 *>(p + k);
    const float4 b = __ldcg(pp + 1);
    s.dx = a.x; s.dy = a.y; s.dz = a.z; s.i = __float_as_int(a.w);
    s.ux = b.x; s.uy = b.y; s.uz = b.z; s.q = b.w;
    s.dispx = c.x; s.dispy = c.y; s.dispz = c.z;
    unresolved = move_p_dev(s, acc, nbr);
    pp[0] = make_float4(s.dx, s.dy, s.dz, __int_as_float(s.i));
    pp[1] = make_float4(s.ux, s.uy, s.uz, s.q);                   // a reflection flips a momentum component
  }
  const unsigned um = __ballot_sync(full, unresolved);
  if (um) {
    int base = 0;
    if (lane == 0) base = atomicAdd(&counters[0], __popc(um));
    base = __shfl_sync(full, base, 0);
    if (unresolved) {
      const int dst = base + __popc(um & ((1u << lane) - 1u));
      if (dst < max_nm) {
        reinterpret_cast<float4 *>(tmp_pm)[dst] = make_float4(s.dispx, s.dispy, s.dispz, __int_as_float(k));
        atomicOr(&bitmap[k >> 5], 1u << (k & 31));
      } else {
        atomicAdd(&counters[1], 1);
      }
    }
  }
}

// STORE 0: results leave by three STG.128 per lane (all 48 bytes of the record, so that no sector is written
//          partially -- a partial write makes L2 fetch the rest from DRAM first: +25 GB per launch in profiles/r1h);
// STORE 1: results are put back into the staged tile and the tile leaves as one bulk async store, which takes
//          the 96 store sectors per chunk off the LSU (its wavefronts were 70 % busy in profiles/r1i) at the
//          price of a proxy fence per chunk and one more stage (a tile is refilled one iteration after its store).
template <int STORE> struct StreamCfg { static constexpr int stages = STORE ? 5 : 4; };

constexpr int kWarpsS = 4;           // warps per CTA of the streaming kernel
template <int STORE>
struct StreamSmem {
  float4 tile[kWarpsS][StreamCfg<STORE>::stages][96];
  float4 q_pos[kWarpsS][kQueue], q_disp[kWarpsS][kQueue];   // the momentum waits in global memory (already stored)
  uint64_t full[kWarpsS][StreamCfg<STORE>::stages];
};

template <int DEPOSIT, int WIDE, int STORE, int CPS>
__global__ void __launch_bounds__(kWarpsS * 32, CPS) advance_p_stream_kernel(const AdvanceArgs A) {
  constexpr int NS = StreamCfg<STORE>::stages;
  constexpr int AHEAD = STORE ? NS - 1 : NS;      // tiles landed or landing, the current one included
  extern __shared__ __align__(128) unsigned char smem_raw[];
  StreamSmem<STORE> &S = *reinterpret_cast<StreamSmem<STORE> *>(smem_raw);

  const unsigned fullmask = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float one = 1.f;
  const float one_third = (float)(1. / 3.);
  const float two_fifteenths = (float)(2. / 15.);
  const float qdt_2mc = A.qdt_2mc, cdt_dx = A.cdt_dx, cdt_dy = A.cdt_dy, cdt_dz = A.cdt_dz;
  int q_head = 0, q_n = 0;
  const uint64_t pol = l2_policy_evict_first();
  char *const gbase = reinterpret_cast<char *>(A.p);

  if (lane == 0) {
    for (int s = 0; s < NS; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.full[w][s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  auto tile_bytes = [&](int chunk) -> uint32_t {
    const int n = A.np - chunk * 32;
    return (uint32_t)((n < 32 ? n : 32) * 48);
  };
  // dynamic scheduling (see advance_p_tma_kernel); the ticket for the NEXT group is already in flight
  // (inline PTX: nvcc turns a lane-0 atomicAdd into a warp-aggregated one whose result is shuffled out at
  // once, which would wait for the atomic right here)
  int g_cur = 0, g_end = 0;
  int ticket = 0;                                  // lane 0: result of the outstanding atomic
  auto take_ticket = [&]() {
    if (lane == 0) asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(ticket) : "l"(A.counters + 2), "r"(kGrabS) : "memory");
  };
  take_ticket();
  auto next_chunk = [&]() -> int {
    if (g_cur >= g_end) {
      const int base = __shfl_sync(fullmask, ticket, 0) + A.chunk_lo;
      if (base >= A.chunk_hi) { g_cur = g_end = A.chunk_hi; return -1; }
      take_ticket();
      g_cur = base;
      g_end = base + kGrabS < A.chunk_hi ? base + kGrabS : A.chunk_hi;
    }
    return g_cur++;
  };
  auto issue_load = [&](int chunk, int stage) {
    if (lane == 0 && chunk >= 0)
      tma_load_tile(smem_u32(&S.tile[w][stage][0]), gbase + (size_t)chunk * kTileBytes, tile_bytes(chunk), smem_u32(&S.full[w][stage]),
                    pol);
  };
  auto drain = [&](int count) {
    if (STORE) {   // the movers' tiles must have landed in global memory before their final state is written over them
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
      }
    }
    __syncwarp();
    drain_movers_slim(A.p, A.a, A.nbr, A.tmp_pm, A.max_nm, A.counters, A.bitmap, S.q_pos[w], S.q_disp[w], q_head, count);
  };

  // pend[j]: chunk whose tile sits (or is landing) in stage (it+j) % NS
  int pend[AHEAD];
#pragma unroll
  for (int j = 0; j < AHEAD; j++) {
    pend[j] = next_chunk();
    issue_load(pend[j], j);
  }
  Interp fa, fb;    // interpolators of the chunk being computed / of the next one; roles alternate
  fa.ex = fa.ey = fa.ez = fa.b0 = make_float4(0, 0, 0, 0);
  fa.b1 = make_float2(0, 0);
  fb = fa;
  if (pend[0] >= 0) {
    mbar_wait(smem_u32(&S.full[w][0]), 0);
    const int k = pend[0] * 32 + lane;
    const int ii = k < A.np ? __float_as_int(S.tile[w][0][lane * 3].w) : 0;
    load_interp<WIDE>(fa, A.f, ii);
  }

  int it = 0;
  // one chunk: compute with `cur`, request `nxt` for the following chunk.  Every lane runs the whole push; lanes
  // past the end of the array carry zeros and are masked out of every store.
  auto step = [&](const Interp &cur, Interp &nxt) {
    const int chunk = pend[0];
    const int stage = it % NS;
    const int k = chunk * 32 + lane;
    const bool valid = k < A.np;
    float4 *tp = &S.tile[w][stage][lane * 3];
    float4 r0 = make_float4(0, 0, 0, 0), r1 = r0;
    if (valid) { r0 = tp[0]; r1 = tp[1]; }

    // request the interpolator of the next chunk before computing on this one
    if (pend[1] >= 0) {
      const int sn = (it + 1) % NS;
      mbar_wait(smem_u32(&S.full[w][sn]), (uint32_t)(((it + 1) / NS) & 1));
      const int kn = pend[1] * 32 + lane;
      const int iin = kn < A.np ? __float_as_int(S.tile[w][sn][lane * 3].w) : 0;
      load_interp<WIDE>(nxt, A.f, iin);
    }

    const int ii = __float_as_int(r0.w);
    float dep[12];
    float dx = r0.x, dy = r0.y, dz = r0.z;
    const float hax = qdt_2mc * ((cur.ex.x + dy * cur.ex.y) + dz * (cur.ex.z + dy * cur.ex.w));
    const float hay = qdt_2mc * ((cur.ey.x + dz * cur.ey.y) + dx * (cur.ey.z + dz * cur.ey.w));
    const float haz = qdt_2mc * ((cur.ez.x + dx * cur.ez.y) + dy * (cur.ez.z + dx * cur.ez.w));
    const float cbx = cur.b0.x + dx * cur.b0.y;
    const float cby = cur.b0.z + dy * cur.b0.w;
    const float cbz = cur.b1.x + dz * cur.b1.y;
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
    const float4 mom = make_float4(ux, uy, uz, q);
    v0 = one / sqrtf(one + (ux * ux + (uy * uy + uz * uz)));
    ux *= cdt_dx; uy *= cdt_dy; uz *= cdt_dz;
    ux *= v0; uy *= v0; uz *= v0;
    v0 = dx + ux; v1 = dy + uy; v2 = dz + uz;
    v3 = v0 + ux; v4 = v1 + uy;
    float v5 = v2 + uz;
    const bool in_cell = v3 <= one && v4 <= one && v5 <= one && -v3 <= one && -v4 <= one && -v5 <= one;
    const bool inbnds = valid && in_cell, outbnds = valid && !in_cell;
    const float4 pos = in_cell ? make_float4(v3, v4, v5, r0.w) : r0;   // out-of-cell: old position until move_p has run
    if (STORE) {
      if (valid) { tp[0] = pos; tp[1] = mom; }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) tma_store_tile(gbase + (size_t)chunk * kTileBytes, smem_u32(&S.tile[w][stage][0]), tile_bytes(chunk), pol);
    } else if (valid) {
      float4 *pp = reinterpret_cast<float4 *>(A.p + k);
      st_stream4(pp, pos);
      st_stream4(pp + 1, mom);
      st_stream4(pp + 2, tp[2]);
    }
    // the 12 contributions (garbage in out-of-cell lanes, which are masked out of the deposit)
    v5 = q * ux * uy * uz * one_third;
    accumulate_j(q, ux, v1, v2, v5, dep[0], dep[1], dep[2], dep[3]);
    accumulate_j(q, uy, v2, v0, v5, dep[4], dep[5], dep[6], dep[7]);
    accumulate_j(q, uz, v0, v1, v5, dep[8], dep[9], dep[10], dep[11]);

    // refill: STORE 0 this stage (read by every lane: the push consumed r0/r1); STORE 1 the stage whose bulk
    // store was issued one iteration ago (it must have finished reading shared memory)
    __syncwarp();
    const int cn = next_chunk();
    if (STORE) {
      if (lane == 0 && cn >= 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      issue_load(cn, (it + NS - 1) % NS);
    } else {
      issue_load(cn, stage);
    }
#pragma unroll
    for (int j = 0; j < AHEAD - 1; j++) pend[j] = pend[j + 1];
    pend[AHEAD - 1] = cn;

    if (DEPOSIT == 0) {
      if (inbnds) red3(A.a + 12 * (size_t)ii, dep);
    } else {
      deposit_dominant(dep, ii, inbnds, A.a);
    }

    const unsigned om = __ballot_sync(fullmask, outbnds);
    if (om) {
      if (outbnds) {
        const int e = (q_head + q_n + __popc(om & ((1u << lane) - 1u))) & (kQueue - 1);
        S.q_pos[w][e] = r0;
        S.q_disp[w][e] = make_float4(ux, uy, uz, __int_as_float(k));
      }
      q_n += __popc(om);
      if (q_n >= 32) {
        drain(32);
        q_head = (q_head + 32) & (kQueue - 1);
        q_n -= 32;
        __syncwarp();
      }
    }
    ++it;
  };

  while (pend[0] >= 0) {
    step(fa, fb);
    if (pend[0] < 0) break;
    step(fb, fa);
  }
  if (q_n) drain(q_n);
  if (STORE) {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// ---- ordered mover emission (post-pass; every kernel leaves at once if nothing was staged) ----

__global__ void __launch_bounds__(256) mover_popc_kernel(const unsigned *__restrict__ bitmap, int nwords, int *__restrict__ cnt,
                                                         const int *__restrict__ counters) {
  if (counters[0] == 0) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += gridDim.x * blockDim.x) cnt[i] = __popc(bitmap[i]);
}

__global__ void __launch_bounds__(256) mover_place_kernel(const unsigned *__restrict__ bitmap, const int *__restrict__ word_off,
                                                          const float4 *__restrict__ tmp_pm, float4 *__restrict__ pm, int max_nm,
                                                          const int *__restrict__ counters, int *__restrict__ d_nm) {
  int total = counters[0];
  if (total > max_nm) total = max_nm;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0 && d_nm) *d_nm = total;
  for (int j = t; j < total; j += gridDim.x * blockDim.x) {
    const float4 m = tmp_pm[j];
    const int k = __float_as_int(m.w);
    const int dst = word_off[k >> 5] + __popc(bitmap[k >> 5] & ((1u << (k & 31)) - 1u));
    pm[dst] = m;
  }
}

}  // namespace vpb

namespace vpb {

// An advance_p call in three parts so that a caller can feed the particle array in pieces (the staged
// host-buffer path of vpb_dropin.cu overlaps H2D / kernel / D2H per piece).
void advance_p_begin(vpb_domain_t *dom, int np, float q_m, int max_nm, vpb_accumulator_t *d_a, const vpb_interpolator_t *d_f,
                     AdvanceJob &J, cudaStream_t st) {
  const DomainDev &g = dom->d;
  AdvanceArgs &A = J.A;
  A.p = nullptr;
  A.np = np;
  A.nchunks = (np + 31) / 32;
  // same expressions, same types as advance_p.cxx:425-428
  A.qdt_2mc = (float)(0.5 * q_m * g.dt / g.cvac);
  A.cdt_dx = g.cvac * g.dt * g.rdx;
  A.cdt_dy = g.cvac * g.dt * g.rdy;
  A.cdt_dz = g.cvac * g.dt * g.rdz;
  A.a = reinterpret_cast<float *>(d_a);
  A.f = d_f;
  A.fi_bytes = g.fi_bytes;
  A.nbr = g.nbr;
  A.max_nm = max_nm;
  A.sx = g.sx; A.sy = g.sy; A.sz = g.sz; A.nv = g.nv;
  A.by = tuning("advance_p.by", 16);
  if (A.by < 1) A.by = 1;
  // scratch: counters | bitmap[nwords] | word_cnt/off[nwords] | tmp_pm[max_nm] | scan scratch
  J.nwords = A.nchunks > 0 ? A.nchunks : 1;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t off_bits = 256, off_cnt = off_bits + al((size_t)J.nwords * 4), off_tmp = off_cnt + al((size_t)J.nwords * 4),
               off_scan = off_tmp + al((size_t)max_nm * 16 + 16);
  char *s = (char *)scratch(off_scan + scan_scratch_bytes(J.nwords));
  A.counters = (int *)s;
  A.bitmap = (unsigned *)(s + off_bits);
  J.word_off = (int *)(s + off_cnt);
  A.tmp_pm = (vpb_particle_mover_t *)(s + off_tmp);
  J.scan_tmp = s + off_scan;
  VPB_CUDA(cudaMemsetAsync(s, 0, off_bits + (size_t)J.nwords * 4, st));   // counters + bitmap
}

// particles k0..k1-1 (k0 a multiple of 32); d_base[k] must address particle k
void advance_p_range(AdvanceJob &J, vpb_particle_t *d_base, int k0, int k1, const int *d_partition, cudaStream_t st) {
  if (k1 <= k0) return;
  Context &c = ctx();
  AdvanceArgs A = J.A;
  A.p = d_base;
  A.np = k1;
  A.chunk_lo = k0 / 32;
  A.chunk_hi = (k1 + 31) / 32;
  A.partition = (k0 == 0 && k1 == J.A.np && tuning("advance_p.ordered", 1)) ? d_partition : nullptr;
  if (A.partition) A.nwork = ((A.sy + A.by - 1) / A.by) * A.by * A.sz + 1;
  else A.nwork = (A.chunk_hi - A.chunk_lo + 63) / 64;
  const int tma_mode = (reinterpret_cast<uintptr_t>(d_base) & 15) == 0 ? tuning("advance_p.tma", 2) : 0;
  if (tma_mode == 2) {
    const int nch = A.chunk_hi - A.chunk_lo;
    VPB_CUDA(cudaMemsetAsync(&A.counters[2], 0, sizeof(int), st));   // ticket counter of the dynamic scheduler
    const bool dep = tuning("advance_p.deposit", 1) != 0;
    const int wide = A.fi_bytes == 96;
    const int store = tuning("advance_p.stream_store", 0) ? 1 : 0;
    // CTAs of 4 warps; 4, 5 or 6 per SM = 16, 20 or 24 warps at <=128, <=96 or <=80 registers per thread
    int cps = store ? 4 : tuning("advance_p.stream_cps", 5);
    cps = cps < 5 ? 4 : (cps > 5 ? 6 : 5);
    typedef void (*kern_t)(AdvanceArgs);
    static const kern_t table[4][2][2] = {   // [cps-4, or 3 = bulk store][wide interpolator][deposit]
        {{advance_p_stream_kernel<0, 0, 0, 4>, advance_p_stream_kernel<1, 0, 0, 4>}, {advance_p_stream_kernel<0, 1, 0, 4>, advance_p_stream_kernel<1, 1, 0, 4>}},
        {{advance_p_stream_kernel<0, 0, 0, 5>, advance_p_stream_kernel<1, 0, 0, 5>}, {advance_p_stream_kernel<0, 1, 0, 5>, advance_p_stream_kernel<1, 1, 0, 5>}},
        {{advance_p_stream_kernel<0, 0, 0, 6>, advance_p_stream_kernel<1, 0, 0, 6>}, {advance_p_stream_kernel<0, 1, 0, 6>, advance_p_stream_kernel<1, 1, 0, 6>}},
        {{advance_p_stream_kernel<0, 0, 1, 4>, advance_p_stream_kernel<1, 0, 1, 4>}, {advance_p_stream_kernel<0, 1, 1, 4>, advance_p_stream_kernel<1, 1, 1, 4>}}};
    const int row = store ? 3 : cps - 4;
    const int smem = store ? (int)sizeof(StreamSmem<1>) : (int)sizeof(StreamSmem<0>);
    static bool attr_set = false;
    if (!attr_set) {
      for (int a = 0; a < 4; a++)
        for (int b = 0; b < 2; b++)
          for (int d = 0; d < 2; d++)
            VPB_CUDA(cudaFuncSetAttribute(table[a][b][d], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          a == 3 ? (int)sizeof(StreamSmem<1>) : (int)sizeof(StreamSmem<0>)));
      attr_set = true;
    }
    int grid = c.sm_count * cps;
    if (grid > (nch + kWarpsS - 1) / kWarpsS) grid = (nch + kWarpsS - 1) / kWarpsS;
    table[row][wide][dep]<<<grid, kWarpsS * 32, smem, st>>>(A);
  } else if (tma_mode == 1) {
    static bool attr_set = false;
    if (!attr_set) {
      VPB_CUDA(cudaFuncSetAttribute(advance_p_tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TmaSmem)));
      VPB_CUDA(cudaFuncSetAttribute(advance_p_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TmaSmem)));
      attr_set = true;
    }
    const int nch = A.chunk_hi - A.chunk_lo;
    VPB_CUDA(cudaMemsetAsync(&A.counters[2], 0, sizeof(int), st));   // ticket counter of the dynamic scheduler
    int grid = c.sm_count * tuning("advance_p.tma_ctas_per_sm", 3);
    if (grid > (nch + kWarps - 1) / kWarps) grid = (nch + kWarps - 1) / kWarps;
    if (tuning("advance_p.deposit", 1) == 0) advance_p_tma_kernel<0><<<grid, kWarps * 32, sizeof(TmaSmem), st>>>(A);
    else advance_p_tma_kernel<1><<<grid, kWarps * 32, sizeof(TmaSmem), st>>>(A);
  } else {
    int grid = c.sm_count * tuning("advance_p.ctas_per_sm", 4);
    if (grid > A.nwork) grid = A.nwork;
    if (tuning("advance_p.deposit", 1) == 0) advance_p_kernel<0><<<grid, kWarps * 32, 0, st>>>(A);
    else advance_p_kernel<1><<<grid, kWarps * 32, 0, st>>>(A);
  }
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

void advance_p_end(AdvanceJob &J, vpb_particle_mover_t *d_pm, int *d_nm, cudaStream_t st) {
  Context &c = ctx();
  const AdvanceArgs &A = J.A;
  const int tg = c.sm_count * 4;
  mover_popc_kernel<<<tg, 256, 0, st>>>(A.bitmap, J.nwords, J.word_off, A.counters);
  exclusive_scan_i32(J.word_off, J.word_off, J.nwords, J.scan_tmp, st, A.counters);
  mover_place_kernel<<<tg, 256, 0, st>>>(A.bitmap, J.word_off, (const float4 *)A.tmp_pm, (float4 *)d_pm, A.max_nm, A.counters, d_nm);
  count_launch(2 + scan_launches(J.nwords));
  VPB_CUDA(cudaGetLastError());
}

}  // namespace vpb

using namespace vpb;

extern "C" void vpb_advance_p_ordered(vpb_domain_t *dom, vpb_particle_t *d_p, int np, float q_m, vpb_particle_mover_t *d_pm,
                                      int max_nm, vpb_accumulator_t *d_a, const vpb_interpolator_t *d_f, int *d_nm,
                                      const int *d_partition) {
  if (!dom) VPB_ERROR("Bad grid");
  if (!d_p) VPB_ERROR("Bad particle array");
  if (np < 0) VPB_ERROR("Bad number of particles");
  if (!d_pm) VPB_ERROR("Bad particle mover");
  if (max_nm < 0) VPB_ERROR("Bad number of movers");
  if (!d_a) VPB_ERROR("Bad accumulator");
  if (!d_f) VPB_ERROR("Bad interpolator");
  Context &c = ctx();
  if (np == 0) {
    if (d_nm) VPB_CUDA(cudaMemsetAsync(d_nm, 0, sizeof(int), c.stream));
    return;
  }
  AdvanceJob J;
  advance_p_begin(dom, np, q_m, max_nm, d_a, d_f, J, c.stream);
  {
    ProfScope prof(0);
    if (dom->d.p_plane > 0) advance_p_pair_launch(J, reinterpret_cast<float *>(d_p), dom->d.p_plane, c.stream);
    else advance_p_range(J, d_p, 0, np, d_partition, c.stream);
  }
  advance_p_end(J, d_pm, d_nm, c.stream);
}

extern "C" void vpb_advance_p(vpb_domain_t *dom, vpb_particle_t *d_p, int np, float q_m, vpb_particle_mover_t *d_pm, int max_nm,
                              vpb_accumulator_t *d_a, const vpb_interpolator_t *d_f, int *d_nm) {
  vpb_advance_p_ordered(dom, d_p, np, q_m, d_pm, max_nm, d_a, d_f, d_nm, nullptr);
}

// number of movers the last vpb_advance_p had to drop because pm[] was full (advance_p.cxx:463-465)
extern "C" int vpb_advance_p_ignored(void) {
  Context &c = ctx();
  if (!c.scratch) return 0;
  VPB_CUDA(cudaMemcpyAsync(c.h_pinned_i, (char *)c.scratch + 4, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  VPB_CUDA(cudaStreamSynchronize(c.stream));
  return c.h_pinned_i[0];
}
