// vpb_sort_group.cu -- the device-resident driver's particle sort: component planes in, component planes out, in
// THREE streaming passes (keys, scan, move), particles grouped by voxel in a brick-Morton order of the voxels.
//
// What sort_p is for (sort_p.c:16-77): particles that share a voxel sit next to each other, so that advance_p's
// interpolator reads and accumulator updates of neighbouring particles hit the same lines.  The ORDER of the groups
// in the array has no meaning to the physics or to advance_p (every particle carries its voxel index); only the
// reference-named sort_p / vpb_sort_p (vpb_particles.cu) promise the reference's order and its partition[].
//
// Why not the voxel index as the key (round 1, vpb_sort_p_planes_ahead: 76 ms per 2^30 particles, 0.21 of the HBM
// roofline).  The array is nearly sorted: since the last sort a particle has moved a cell or two.  With x-fastest
// voxel indices a step of one cell in z is 258^2 voxels = 200 MB of particles away, far outside the 126 MB L2, so a
// permutation pass -- gather or scatter -- touches every DRAM line several times (ncu: 170 GB read for 51 GB of
// payload), and to keep the requests wide the planes had to be transposed to 48-byte records and back.
//
// Here the key of a voxel is  Morton(brick) * brick_volume + index inside the brick,  bricks of 8 x 4 x 4 voxels:
// spatial neighbours are neighbours in the array (for all but a geometrically small share of brick faces), so
//   * a chunk of 1024 consecutive source particles (an eighth of a brick) sends its particles to a few dozen
//     destination groups, most of them in the same or an adjacent brick;
//   * the chunk is bucketed in shared memory (hash table of its destination keys, one global atomic per distinct key
//     claims the slots) while its plane words arrive in shared memory by bulk async copies (cp.async.bulk), and they
//     leave in destination order: runs of a group's particles, i.e. mostly whole 32-byte sectors;
//   * the lines a chunk only partly fills are completed by chunks that run at almost the same time, while they are
//     still in L2.
// Traffic per particle: keys pass 28 B read (4 B without look-ahead) + 4 B written; move pass 4 + 48 B read, 48 B
// written: 132 B, against ~320 B before.  The key is separable, key = fx[x] + fy[y] + fz[z], three small tables built
// on the host (vpb_sort_group_order; also what the tests use to check the grouping).
#include <algorithm>
#include <vector>
#include "vpb_common.cuh"
#include "vpb_pview.cuh"
#include "vpb_scan.cuh"

namespace vpb {

// ---- host: the order tables ---------------------------------------------------------------------------------------
struct GroupOrder {
  std::vector<int> fx, fy, fz;   // per coordinate 0..n+1 (ghosts share the key of the nearest interior coordinate)
  long nkeys = 0;
  int b[3] = {1, 1, 1};
};

static int ceil_log2(long v) { int b = 0; while ((1L << b) < v) b++; return b; }

static GroupOrder make_group_order(int nx, int ny, int nz) {
  const int n[3] = {nx, ny, nz};
  GroupOrder o;
  // brick: 8 x 4 x 4 voxels, clipped to the grid; a clipped axis lets the others grow until the brick holds ~128 voxels
  int b[3] = {std::min(8, nx), std::min(4, ny), std::min(4, nz)};
  for (int pass = 0; pass < 8 && b[0] * b[1] * b[2] < 128; pass++)
    for (int a = 2; a >= 0; a--)
      if (b[0] * b[1] * b[2] < 128 && 2 * b[a] <= n[a]) b[a] *= 2;
  long nb[3];
  int w[3];
  for (int a = 0; a < 3; a++) { nb[a] = (n[a] + b[a] - 1) / b[a]; w[a] = ceil_log2(nb[a]); }
  const long bv = (long)b[0] * b[1] * b[2];
  const long nv_real = (long)nx * ny * nz;
  // Morton interleave of the brick coordinates (axes with fewer bits simply run out); when padding the brick counts to
  // powers of two would inflate the key space beyond 2x (+ a little), bricks are ordered row-major instead
  const bool morton = ((1L << (w[0] + w[1] + w[2])) * bv) <= 2 * nv_real + 4096;
  std::vector<long> spread[3];
  if (morton) {
    int pos[3][32];
    int out = 0;
    for (int bit = 0; bit < 31; bit++)
      for (int a = 0; a < 3; a++)
        if (bit < w[a]) pos[a][bit] = out++;
    for (int a = 0; a < 3; a++) {
      spread[a].resize(nb[a]);
      for (long v = 0; v < nb[a]; v++) {
        long s = 0;
        for (int bit = 0; bit < w[a]; bit++)
          if ((v >> bit) & 1) s |= 1L << pos[a][bit];
        spread[a][v] = s;
      }
    }
    o.nkeys = (1L << (w[0] + w[1] + w[2])) * bv;
  } else {
    const long stride[3] = {1, nb[0], nb[0] * nb[1]};
    for (int a = 0; a < 3; a++) {
      spread[a].resize(nb[a]);
      for (long v = 0; v < nb[a]; v++) spread[a][v] = v * stride[a];
    }
    o.nkeys = nb[0] * nb[1] * nb[2] * bv;
  }
  if (o.nkeys > 0x7fff0000L) VPB_ERROR("sort key space of %ld groups exceeds 2^31", o.nkeys);
  const long in_stride[3] = {1, b[0], (long)b[0] * b[1]};
  std::vector<int> *tab[3] = {&o.fx, &o.fy, &o.fz};
  for (int a = 0; a < 3; a++) {
    tab[a]->resize(n[a] + 2);
    for (int c = 0; c <= n[a] + 1; c++) {
      const int ci = std::min(std::max(c, 1), n[a]) - 1;       // 0-based interior coordinate
      (*tab[a])[c] = (int)(spread[a][ci / b[a]] * bv + (ci % b[a]) * in_stride[a]);
    }
    o.b[a] = b[a];
  }
  return o;
}

// ---- device -------------------------------------------------------------------------------------------------------
struct GroupKeyArgs {
  int L, sx, sy, nx, ny, nz;
  unsigned long long msx, msy;  // ceil(2^64 / sx), ceil(2^64 / sy): v / sx == umul64hi(v, msx) for every 32-bit v
  float kx, ky, kz;             // 2 * L * c dt / d{x,y,z} (vpb_particles.cu: SortAhead)
  const int *fx, *fy, *fz;
};

// Key of particle k: the group of its voxel (AHEAD = 0) or of the voxel it reaches in L steps at its present velocity,
// clamped to the interior (look-ahead grouping).  In two halves so that a thread can have the loads of
// several particles in flight before it uses any of them.
struct GroupRaw { int v; float dx, dy, dz, ux, uy, uz; };

template <int AHEAD>
__device__ __forceinline__ GroupRaw group_load(const PView &p, long k) {
  GroupRaw r;
  const size_t pl = (size_t)p.plane;
  const float *b = p.b + k;
  r.v = __float_as_int(__ldcs(b + 3 * pl));
  if (AHEAD) {
    r.dx = __ldcs(b); r.dy = __ldcs(b + pl); r.dz = __ldcs(b + 2 * pl);
    r.ux = __ldcs(b + 4 * pl); r.uy = __ldcs(b + 5 * pl); r.uz = __ldcs(b + 6 * pl);
  } else {
    r.dx = r.dy = r.dz = r.ux = r.uy = r.uz = 0.f;
  }
  return r;
}

template <int AHEAD>
__device__ __forceinline__ int group_key(const GroupRaw &r, const GroupKeyArgs &A) {
  const int v = r.v;
  const int t = (int)__umul64hi((unsigned long long)(unsigned)v, A.msx), ix = v - t * A.sx;
  const int iz = (int)__umul64hi((unsigned long long)(unsigned)t, A.msy), iy = t - iz * A.sy;
  int cx = ix, cy = iy, cz = iz;
  if (AHEAD) {
    const float rg = rsqrtf(1.f + (r.ux * r.ux + (r.uy * r.uy + r.uz * r.uz)));
    cx = ix + (int)floorf((r.dx + A.kx * r.ux * rg + 1.f) * 0.5f);
    cy = iy + (int)floorf((r.dy + A.ky * r.uy * rg + 1.f) * 0.5f);
    cz = iz + (int)floorf((r.dz + A.kz * r.uz * rg + 1.f) * 0.5f);
    cx = min(max(cx, 1), A.nx); cy = min(max(cy, 1), A.ny); cz = min(max(cz, 1), A.nz);
  }
  return __ldg(A.fx + cx) + __ldg(A.fy + cy) + __ldg(A.fz + cz);
}

// Both passes below are chains of dependent memory round trips (stream in -> table look-up -> atomic -> store): with one
// row per thread in flight they ran at the latency of that chain (2.2 us per warp and row, a third of the DRAM
// bandwidth; profiles/r2h), so every thread carries kKeyRows / kInvRows independent rows.
constexpr int kKeyRows = 4, kInvRows = 8;
template <int RANKS, int AHEAD>
__global__ void __launch_bounds__(256) group_keys_kernel(const PView p, int np, const GroupKeyArgs A, int *__restrict__ keys,
                                                         int *__restrict__ ranks, int *__restrict__ count) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const long tile = (long)blockDim.x * kKeyRows;
  const long ntiles = ((long)np + tile - 1) / tile;
  for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long k0 = t * tile + threadIdx.x;
    GroupRaw raw[kKeyRows];
#pragma unroll
    for (int j = 0; j < kKeyRows; j++) {          // every load of the tile first (rows past the end re-read the last particle)
      const long k = k0 + (long)j * blockDim.x;
      raw[j] = group_load<AHEAD>(p, k < np ? k : (long)np - 1);
    }
    int key[kKeyRows];
#pragma unroll
    for (int j = 0; j < kKeyRows; j++) {
      const long k = k0 + (long)j * blockDim.x;
      key[j] = k < np ? group_key<AHEAD>(raw[j], A) : -1 - lane;
    }
    unsigned peers[kKeyRows];
    int base[kKeyRows];
#pragma unroll
    for (int j = 0; j < kKeyRows; j++) {
      peers[j] = __match_any_sync(full, key[j]);
      base[j] = 0;
      if (key[j] >= 0 && lane == __ffs(peers[j]) - 1) {
        if (RANKS) base[j] = atomicAdd(count + key[j], __popc(peers[j]));
        else atomicAdd(count + key[j], __popc(peers[j]));   // (ptxas turns an unused result into a RED)
      }
    }
#pragma unroll
    for (int j = 0; j < kKeyRows; j++) {
      const long k = k0 + (long)j * blockDim.x;
      if (RANKS) base[j] = __shfl_sync(full, base[j], __ffs(peers[j]) - 1);
      if (k < np) {
        const int rank = base[j] + __popc(peers[j] & lt);
        if (RANKS == 2) {   // one word: key in the low 24 bits, rank in the high 8; the rare rank >= 255 goes to ranks[] as well
          keys[k] = key[j] | ((rank < 255 ? rank : 255) << 24);
          if (rank >= 255) ranks[k] = rank;
        } else {
          keys[k] = key[j];
          if (RANKS) ranks[k] = rank;
        }
      }
    }
  }
}

constexpr int kGsThreads = 256, kGsPer = 4, kGsChunk = kGsThreads * kGsPer, kGsHash = 2 * kGsChunk;

struct GroupSmem {
  union {
    float w[8][kGsChunk];            // the eight word planes of the chunk ...
    float4 pad_[2 * kGsChunk];
  } data;
  float4 tag[kGsChunk];              // ... and its tag plane, filled by bulk async copies (cp.async.bulk, SASS UBLKCP)
  int hkey[kGsHash];                 // key of the slot (-1 free); after the claim: first destination of the chunk's group
  int hcnt[kGsHash];                 // particles of the chunk with that key; after the claim: their first local position
  int ldst[kGsChunk];                // destination of the particle at local position l
  unsigned short lsrc[kGsChunk];     // its index inside the chunk
  int wsum[kGsThreads / 32];
  unsigned long long full;           // mbarrier: the chunk's 48 bytes per particle have landed
};

__device__ __forceinline__ uint32_t gs_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gs_bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void gs_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// pass 3: move.  cursor[] starts as a copy of partition[]; one atomicAdd per (chunk, distinct key) claims the slots.
// A CTA works on chunks of 1024 consecutive source particles.  Per chunk: one thread starts the bulk copies of the
// chunk's nine planes into shared memory; meanwhile all threads bucket the chunk by key (its keys were requested one
// chunk earlier), claim the destination ranges and lay out the local order; then the words leave shared memory in
// destination order.  Three CTAs per SM are in different phases at any time.
template <int EVICT_LAST>
__global__ void __launch_bounds__(kGsThreads, 3) group_move_kernel(const PView in, const PView out, int np, const int *__restrict__ keys,
                                                                   int *__restrict__ cursor) {
  extern __shared__ __align__(128) unsigned char gs_raw[];
  GroupSmem &S = *reinterpret_cast<GroupSmem *>(gs_raw);
  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const uint32_t bar = gs_smem_u32(&S.full);
  for (int s = tid; s < kGsHash; s += kGsThreads) { S.hkey[s] = -1; S.hcnt[s] = 0; }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint64_t pol_in = l2_policy_evict_first(), pol_out = l2_policy_evict_last();
  const int nchunks = (np + kGsChunk - 1) / kGsChunk;
  const size_t pli = (size_t)in.plane, plo = (size_t)out.plane;
  // keys of a chunk, requested a whole chunk ahead
  int key_next[kGsPer];
  auto load_keys = [&](int chunk) {
    const int c0 = chunk * kGsChunk;
#pragma unroll
    for (int j = 0; j < kGsPer; j++) {
      const int i = c0 + j * kGsThreads + tid;
      key_next[j] = (chunk < nchunks && i < np) ? __ldcs(keys + i) : -1;
    }
  };
  load_keys(blockIdx.x);
  uint32_t parity = 0;
  for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, parity ^= 1u) {
    const int c0 = chunk * kGsChunk;
    const int nvalid = np - c0 < kGsChunk ? np - c0 : kGsChunk;
    if (tid == 0) {
      // word planes: a multiple of 16 bytes (the plane stride is a multiple of 64 particles, so rounding the count up to
      // 4 stays inside the allocation); tags: 16 bytes each
      const uint32_t wbytes = (uint32_t)((nvalid + 3) & ~3) * 4u, tbytes = (uint32_t)nvalid * 16u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(8u * wbytes + tbytes) : "memory");
#pragma unroll
      for (int c = 0; c < 8; c++) gs_bulk_load(gs_smem_u32(&S.data.w[c][0]), in.b + (size_t)c * pli + c0, wbytes, bar, pol_in);
      gs_bulk_load(gs_smem_u32(&S.tag[0]), reinterpret_cast<const float4 *>(in.b + 8 * pli) + c0, tbytes, bar, pol_in);
    }
    // ---- bucket the chunk by key: (slot, rank inside the slot) per particle, one table insert per distinct key and warp
    int slot[kGsPer], rank[kGsPer];
#pragma unroll
    for (int j = 0; j < kGsPer; j++) {
      const bool valid = j * kGsThreads + tid < nvalid;
      const int key = valid ? key_next[j] : -1 - lane;
      const unsigned peers = __match_any_sync(full, key);
      const int leader = __ffs(peers) - 1;
      int s = 0, r0 = 0;
      if (valid && lane == leader) {
        s = (int)(((unsigned)key * 2654435761u) >> 19) & (kGsHash - 1);
        for (;;) {   // linear probing; at most half of the table is ever in use
          const int old = atomicCAS(&S.hkey[s], -1, key);
          if (old == -1 || old == key) break;
          s = (s + 1) & (kGsHash - 1);
        }
        r0 = atomicAdd(&S.hcnt[s], __popc(peers));
      }
      s = __shfl_sync(full, s, leader);
      r0 = __shfl_sync(full, r0, leader);
      slot[j] = valid ? s : -1;
      rank[j] = r0 + __popc(peers & lt);
    }
    load_keys(chunk + gridDim.x);                      // in flight until the next iteration
    __syncthreads();
    // ---- the particle of rank 0 speaks for its group: local offset by a block-wide exclusive scan of the group
    //      sizes, destination by ONE global atomic
    int gcnt[kGsPer], gkey[kGsPer], mine = 0;
#pragma unroll
    for (int j = 0; j < kGsPer; j++) {
      const bool lead = slot[j] >= 0 && rank[j] == 0;
      gcnt[j] = lead ? S.hcnt[slot[j]] : 0;
      gkey[j] = lead ? S.hkey[slot[j]] : -1;
      mine += gcnt[j];
    }
    int base[kGsPer];
#pragma unroll
    for (int j = 0; j < kGsPer; j++) base[j] = gkey[j] >= 0 ? atomicAdd(cursor + gkey[j], gcnt[j]) : 0;   // issued together
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(full, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) S.wsum[w] = incl;
    __syncthreads();                                   // every group size has been read; wsum is complete
    int run = incl - mine;
    for (int q = 0; q < w; q++) run += S.wsum[q];
#pragma unroll
    for (int j = 0; j < kGsPer; j++)
      if (gkey[j] >= 0) {
        S.hkey[slot[j]] = base[j];
        S.hcnt[slot[j]] = run;
        run += gcnt[j];
      }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kGsPer; j++)
      if (slot[j] >= 0) {
        const int l = S.hcnt[slot[j]] + rank[j];
        S.lsrc[l] = (unsigned short)(j * kGsThreads + tid);
        S.ldst[l] = S.hkey[slot[j]] + rank[j];
      }
    __syncthreads();
    // leave the table empty for the next chunk (nobody reads it any more)
#pragma unroll
    for (int j = 0; j < kGsPer; j++)
      if (slot[j] >= 0) { S.hkey[slot[j]] = -1; S.hcnt[slot[j]] = 0; }
    // ---- the chunk's words have landed: out they go, in destination order
    gs_mbar_wait(bar, parity);
    float4 *tdst = reinterpret_cast<float4 *>(out.b + 8 * plo);
#pragma unroll
    for (int j = 0; j < kGsPer; j++) {
      const int l = j * kGsThreads + tid;
      if (l < nvalid) {
        const int sidx = S.lsrc[l];
        const int d = S.ldst[l];
        float *dst = out.b + d;
        if (EVICT_LAST) {
#pragma unroll
          for (int c = 0; c < 8; c++)
            asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(dst + (size_t)c * plo), "f"(S.data.w[c][sidx]), "l"(pol_out) : "memory");
          st_hint4(tdst + d, S.tag[sidx], pol_out);
        } else {
#pragma unroll
          for (int c = 0; c < 8; c++) dst[(size_t)c * plo] = S.data.w[c][sidx];
          tdst[d] = S.tag[sidx];
        }
      }
    }
    __syncthreads();                                   // the data buffer and the table are free again
  }
}


// ---- pass 3, destination-driven (the default) -----------------------------------------------------------------------
// The chunk kernel above scatters 4-byte words: a destination sector (eight particles of one plane) is assembled from
// the partial writes of several CTAs, and whatever part of it has left L2 before the last writer arrives costs a
// partial DRAM write now and a read-modify-write later (ncu, 2^30 particles ten steps after a sort: 2.0 G write misses
// for 1.6 G distinct sectors, 87 GB read + 73 GB written for 56 + 52 GB of payload).  Reads have no such penalty, so
// the move is turned around: the key pass also returns every particle's rank inside its group, group_invert_kernel
// records src[partition[key] + rank] = particle; group_gather_kernel walks the DESTINATION in order -- src[] and every output plane are coalesced,
// full-line streaming stores -- and gathers the nine source words of a slot.  Blocks are dispatched in destination
// order, i.e. (the array being nearly sorted) in source order too, so the source sectors a warp touches are the ones
// its neighbours in time touch: they are fetched from DRAM once and served from L2/L1 afterwards.
template <int PACKED>
__global__ void __launch_bounds__(256) group_invert_kernel(int np, const int *__restrict__ keys, const int *__restrict__ ranks,
                                                           const int *__restrict__ partition, int *__restrict__ src) {
  const long tile = (long)blockDim.x * kInvRows;
  const long ntiles = ((long)np + tile - 1) / tile;
  for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long k0 = t * tile + threadIdx.x;
    int key[kInvRows], d[kInvRows];
#pragma unroll
    for (int j = 0; j < kInvRows; j++) {
      const long k = k0 + (long)j * blockDim.x;
      if (PACKED) {
        const unsigned w = k < np ? (unsigned)__ldcs(keys + k) : 0xffffffffu;
        key[j] = k < np ? (int)(w & 0xffffffu) : -1;
        d[j] = (int)(w >> 24);
        if (k < np && d[j] == 255) d[j] = __ldcs(ranks + k);
      } else {
        key[j] = k < np ? __ldcs(keys + k) : -1;
        d[j] = k < np ? __ldcs(ranks + k) : 0;
      }
    }
#pragma unroll
    for (int j = 0; j < kInvRows; j++)
      if (key[j] >= 0) d[j] += __ldg(partition + key[j]);
#pragma unroll
    for (int j = 0; j < kInvRows; j++)
      if (key[j] >= 0) src[d[j]] = (int)(k0 + (long)j * blockDim.x);
  }
}

template <int KEEP>
__global__ void __launch_bounds__(256) group_gather_kernel(const PView in, const PView out, int np, const int *__restrict__ src) {
  const int d = blockIdx.x * 256 + threadIdx.x;
  if (d >= np) return;
  const int s = __ldcs(src + d);
  const size_t pli = (size_t)in.plane, plo = (size_t)out.plane;
  const float *b = in.b + s;
  float w[8];
  float4 t;
  if (KEEP) {   // sort.gather_keep: ask L2 to hold on to the source sectors (their other words are wanted soon)
    const uint64_t pol = l2_policy_evict_last();
#pragma unroll
    for (int c = 0; c < 8; c++)
      asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(w[c]) : "l"(b + (size_t)c * pli), "l"(pol));
    t = ldg_hint4(reinterpret_cast<const float4 *>(in.b + 8 * pli) + s, pol);
  } else {
#pragma unroll
    for (int c = 0; c < 8; c++) w[c] = __ldg(b + (size_t)c * pli);
    t = __ldg(reinterpret_cast<const float4 *>(in.b + 8 * pli) + s);
  }
  float *o = out.b + d;
#pragma unroll
  for (int c = 0; c < 8; c++) __stcs(o + (size_t)c * plo, w[c]);
  __stcs(reinterpret_cast<float4 *>(out.b + 8 * plo) + d, t);
}

// ---- pass 3, source-driven without staging (sort.group_variant = 1, for A/B runs) -----------------------------------
// Thread per source particle: claim the slot, stream the nine words in, scatter them.  Per warp the stores touch the
// same sectors the chunk kernel's do (a warp's 32 consecutive particles go to about ten groups either way).
__global__ void __launch_bounds__(256) group_scatter_kernel(const PView in, const PView out, int np, const int *__restrict__ keys,
                                                            int *__restrict__ cursor) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const long k = (long)blockIdx.x * 256 + threadIdx.x;      // whole warps stay together: np is rounded up by the grid
  const bool valid = k < np;
  const int key = valid ? __ldcs(keys + k) : -1 - lane;
  const unsigned peers = __match_any_sync(full, key);
  const int leader = __ffs(peers) - 1;
  int base = 0;
  if (valid && lane == leader) base = atomicAdd(cursor + key, __popc(peers));
  base = __shfl_sync(full, base, leader);
  if (!valid) return;
  const int d = base + __popc(peers & lt);
  const size_t pli = (size_t)in.plane, plo = (size_t)out.plane;
  const float *b = in.b + k;
  float w[8];
#pragma unroll
  for (int c = 0; c < 8; c++) w[c] = __ldcs(b + (size_t)c * pli);
  const float4 t = __ldcs(reinterpret_cast<const float4 *>(in.b + 8 * pli) + k);
  float *o = out.b + d;
#pragma unroll
  for (int c = 0; c < 8; c++) o[(size_t)c * plo] = w[c];
  reinterpret_cast<float4 *>(out.b + 8 * plo)[d] = t;
}

struct GroupTables { int *dev = nullptr; long nkeys = 0; int nx = 0, ny = 0, nz = 0; };
static std::vector<std::pair<const vpb_domain_t *, GroupTables>> g_group_tables;

static const GroupTables &tables_of(vpb_domain_t *dom) {
  const DomainDev &g = dom->d;
  for (auto &e : g_group_tables)
    if (e.first == dom && e.second.nx == g.nx && e.second.ny == g.ny && e.second.nz == g.nz) return e.second;
  const GroupOrder o = make_group_order(g.nx, g.ny, g.nz);
  GroupTables t;
  t.nkeys = o.nkeys; t.nx = g.nx; t.ny = g.ny; t.nz = g.nz;
  std::vector<int> all;
  all.insert(all.end(), o.fx.begin(), o.fx.end());
  all.insert(all.end(), o.fy.begin(), o.fy.end());
  all.insert(all.end(), o.fz.begin(), o.fz.end());
  VPB_CUDA(cudaMalloc(&t.dev, all.size() * sizeof(int)));
  VPB_CUDA(cudaMemcpyAsync(t.dev, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice, ctx().stream));
  VPB_CUDA(cudaStreamSynchronize(ctx().stream));
  for (auto &e : g_group_tables)
    if (e.first == dom) { cudaFree(e.second.dev); e.second = t; return e.second; }
  g_group_tables.push_back({dom, t});
  return g_group_tables.back().second;
}

void sort_group_forget(const vpb_domain_t *dom) {
  for (size_t i = 0; i < g_group_tables.size(); i++)
    if (g_group_tables[i].first == dom) {
      cudaFree(g_group_tables[i].second.dev);
      g_group_tables.erase(g_group_tables.begin() + i);
      return;
    }
}

}  // namespace vpb

using namespace vpb;

extern "C" {

// The order of the groups: key(voxel x,y,z) = fx[x] + fy[y] + fz[z] for interior coordinates 1..n (tables of n+2
// entries; NULL = only the size is wanted).  Returns the number of keys; the partition of a grouped sort has one
// entry more.  Pure host code (no device needed).
long vpb_sort_group_order(int nx, int ny, int nz, int *fx, int *fy, int *fz) {
  if (nx < 1 || ny < 1 || nz < 1) VPB_ERROR("Bad grid size");
  const GroupOrder o = make_group_order(nx, ny, nz);
  if (fx) std::copy(o.fx.begin(), o.fx.end(), fx);
  if (fy) std::copy(o.fy.begin(), o.fy.end(), fy);
  if (fz) std::copy(o.fz.begin(), o.fz.end(), fz);
  return o.nkeys;
}

long vpb_sort_group_keys(vpb_domain_t *dom) {
  if (!dom) VPB_ERROR("Bad grid");
  return make_group_order(dom->d.nx, dom->d.ny, dom->d.nz).nkeys;
}

// Grouping sort of a component-plane array: d_out receives the particles of d_in grouped by the voxel they occupy
// (lookahead == 0) or will occupy `lookahead` steps from now, groups in the order of vpb_sort_group_order;
// d_partition (int[keys+1]) the first particle of every group.  The order inside a group is arbitrary.
void vpb_sort_p_planes_grouped(vpb_domain_t *dom, const vpb_particle_t *d_in, vpb_particle_t *d_out, int np, int *d_partition,
                               int lookahead) {
  if (!dom) VPB_ERROR("Bad grid");
  if (dom->d.p_plane <= 0) VPB_ERROR("the domain keeps its particles in the reference layout: use vpb_sort_p");
  if (!d_partition) VPB_ERROR("Bad partition");
  if (np < 0) VPB_ERROR("Bad number of particles");
  if (lookahead < 0) VPB_ERROR("Bad look-ahead");
  if (np > 0 && (!d_in || !d_out || d_in == d_out)) VPB_ERROR("Bad particle array");
  Context &c = ctx();
  ProfScope prof(1);
  const DomainDev &gd = dom->d;
  const GroupTables &T = tables_of(dom);
  const int nk1 = (int)T.nkeys + 1;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  // 0: chunks of 1024 staged in shared memory, scattered; 1: thread per source particle, scattered; 2: inverse
  // permutation, then a destination-ordered gather (default)
  const int variant = tuning("sort.group_variant", 2);
  const PView in(d_in, gd.p_plane), out(d_out, gd.p_plane);
  // scratch: cursor[keys+1] | (variant 2) src[np] | (variants 0, 1) keys[np] | scan.  Variant 2 keeps keys[] and ranks[]
  // in the first two planes of d_out: both are dead when the gather starts to write there.
  const size_t off_arr = al((size_t)nk1 * 4), off_scan = off_arr + al((size_t)np * 4 + 4);
  char *s = (char *)scratch(off_scan + scan_scratch_bytes(nk1));
  int *cursor = (int *)s, *arr = (int *)(s + off_arr);
  int *keys = variant == 2 ? reinterpret_cast<int *>(out.b) : arr;
  int *ranks = variant == 2 ? reinterpret_cast<int *>(out.b) + gd.p_plane : nullptr;
  int *src = arr;
  const bool packed = variant == 2 && T.nkeys <= (1L << 24) && tuning("sort.pack_rank", 1) != 0;
  GroupKeyArgs A;
  A.L = lookahead; A.sx = gd.sx; A.sy = gd.sy; A.nx = gd.nx; A.ny = gd.ny; A.nz = gd.nz;
  A.kx = 2.f * lookahead * gd.cvac * gd.dt * gd.rdx;
  A.ky = 2.f * lookahead * gd.cvac * gd.dt * gd.rdy;
  A.kz = 2.f * lookahead * gd.cvac * gd.dt * gd.rdz;
  A.fx = T.dev; A.fy = T.dev + gd.sx; A.fz = T.dev + gd.sx + gd.sy;
  A.msx = ~0ULL / (unsigned long long)gd.sx + 1ULL;
  A.msy = ~0ULL / (unsigned long long)gd.sy + 1ULL;
  VPB_CUDA(cudaMemsetAsync(cursor, 0, (size_t)nk1 * 4, c.stream));
  const long blocks = ((long)np + 255) / 256, cap = (long)c.sm_count * 16;
  auto grid_of = [&](int rows) { const long b = (blocks + rows - 1) / rows; return (int)(b < cap ? b : cap); };
  if (np > 0) {
    auto kern = variant != 2 ? (lookahead ? group_keys_kernel<0, 1> : group_keys_kernel<0, 0>)
                : packed     ? (lookahead ? group_keys_kernel<2, 1> : group_keys_kernel<2, 0>)
                             : (lookahead ? group_keys_kernel<1, 1> : group_keys_kernel<1, 0>);
    kern<<<grid_of(kKeyRows), 256, 0, c.stream>>>(in, np, A, keys, ranks, cursor);
  }
  exclusive_scan_i32(cursor, d_partition, nk1, s + off_scan, c.stream);   // partition[keys] = np
  count_launch(1 + scan_launches(nk1));
  if (np == 0) return;
  if (variant == 2) {
    if (packed) group_invert_kernel<1><<<grid_of(kInvRows), 256, 0, c.stream>>>(np, keys, ranks, d_partition, src);
    else group_invert_kernel<0><<<grid_of(kInvRows), 256, 0, c.stream>>>(np, keys, ranks, d_partition, src);
    if (tuning("sort.gather_keep", 0)) group_gather_kernel<1><<<(int)blocks, 256, 0, c.stream>>>(in, out, np, src);
    else group_gather_kernel<0><<<(int)blocks, 256, 0, c.stream>>>(in, out, np, src);
    count_launch(2);
  } else if (variant == 1) {
    VPB_CUDA(cudaMemcpyAsync(cursor, d_partition, (size_t)nk1 * 4, cudaMemcpyDeviceToDevice, c.stream));
    group_scatter_kernel<<<(int)blocks, 256, 0, c.stream>>>(in, out, np, keys, cursor);
    count_launch(2);
  } else {
    VPB_CUDA(cudaMemcpyAsync(cursor, d_partition, (size_t)nk1 * 4, cudaMemcpyDeviceToDevice, c.stream));
    static int ctas_per_sm[2] = {0, 0};
    const int ev = tuning("sort.evict_last", 0) ? 1 : 0;
    auto kern = ev ? group_move_kernel<1> : group_move_kernel<0>;
    if (!ctas_per_sm[ev]) {
      VPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GroupSmem)));
      VPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm[ev], kern, kGsThreads, sizeof(GroupSmem)));
      if (ctas_per_sm[ev] < 1) VPB_ERROR("group_move_kernel does not fit on this device");
    }
    const int nchunks = (np + kGsChunk - 1) / kGsChunk;
    const int want = ctas_per_sm[ev] * c.sm_count;
    kern<<<nchunks < want ? nchunks : want, kGsThreads, sizeof(GroupSmem), c.stream>>>(in, out, np, keys, cursor);
    count_launch(2);
  }
  VPB_CUDA(cudaGetLastError());
}

}  // extern "C"
