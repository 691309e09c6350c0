// vpb_particles.cu -- particle-side kernels besides the push:
//   center_p / uncenter_p   src/species_advance/standard/center_p.cxx:5-70, uncenter_p.cxx:5-72
//   energy_p                energy_p.cxx:5-48,124-157
//   accumulate_rho_p        rho_p.c:23-79
//   sort_p                  sort_p.c:16-77 (stable counting sort + partition[])
// Arithmetic follows the reference's scalar flavour (built with -fmad=false).
#include "vpb_common.cuh"
#include "vpb_pview.cuh"
#include "vpb_scan.cuh"

namespace vpb {

struct Gather {
  float hax, hay, haz, cbx, cby, cbz;
};

// advance_p.cxx:73-83: half E kick and B at the particle from the 18 coefficients
__device__ __forceinline__ Gather gather_fields(const vpb_interpolator_t *__restrict__ f0, int fi_bytes, int ii, float qdt_2mc,
                                                float dx, float dy, float dz) {
  const char *fp = reinterpret_cast<const char *>(f0) + (size_t)ii * fi_bytes;
  const float4 fe_x = ldg4(fp), fe_y = ldg4(fp + 16), fe_z = ldg4(fp + 32), fb_0 = ldg4(fp + 48);
  const float2 fb_1 = ldg2(fp + 64);
  Gather g;
  g.hax = qdt_2mc * ((fe_x.x + dy * fe_x.y) + dz * (fe_x.z + dy * fe_x.w));
  g.hay = qdt_2mc * ((fe_y.x + dz * fe_y.y) + dx * (fe_y.z + dz * fe_y.w));
  g.haz = qdt_2mc * ((fe_z.x + dx * fe_z.y) + dy * (fe_z.z + dx * fe_z.w));
  g.cbx = fb_0.x + dx * fb_0.y;
  g.cby = fb_0.z + dy * fb_0.w;
  g.cbz = fb_1.x + dz * fb_1.y;
  return g;
}

// advance_p.cxx:90-102 with rotation constant k (qdt_2mc, or +-qdt_4mc for the half rotations)
__device__ __forceinline__ void boris_rotate(float &ux, float &uy, float &uz, const Gather &g, float k) {
  const float one = 1.f, one_third = (float)(1. / 3.), two_fifteenths = (float)(2. / 15.);
  float v0 = k / sqrtf(one + (ux * ux + (uy * uy + uz * uz)));
  float v1 = g.cbx * g.cbx + (g.cby * g.cby + g.cbz * g.cbz);
  float v2 = (v0 * v0) * v1;
  const float v3 = v0 * (one + v2 * (one_third + v2 * two_fifteenths));
  float v4 = v3 / (one + v1 * (v3 * v3));
  v4 += v4;
  v0 = ux + v3 * (uy * g.cbz - uz * g.cby);
  v1 = uy + v3 * (uz * g.cbx - ux * g.cbz);
  v2 = uz + v3 * (ux * g.cby - uy * g.cbx);
  ux += v4 * (v1 * g.cbz - v2 * g.cby);
  uy += v4 * (v2 * g.cbx - v0 * g.cbz);
  uz += v4 * (v0 * g.cby - v1 * g.cbx);
}

// MODE 0: center_p (half kick then half rotate); MODE 1: uncenter_p (constants
// negated, half rotate then half kick).
template <int MODE>
__global__ void __launch_bounds__(256) center_kernel(const PView p, int np, float qdt_2mc, float qdt_4mc,
                                                     const vpb_interpolator_t *__restrict__ f0, int fi_bytes) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < np; k += gridDim.x * blockDim.x) {
    const float4 r0 = p.pos(k);
    float4 r1 = p.mom(k);
    const Gather g = gather_fields(f0, fi_bytes, __float_as_int(r0.w), qdt_2mc, r0.x, r0.y, r0.z);
    if (MODE == 0) { r1.x += g.hax; r1.y += g.hay; r1.z += g.haz; }
    boris_rotate(r1.x, r1.y, r1.z, g, qdt_4mc);
    if (MODE == 1) { r1.x += g.hax; r1.y += g.hay; r1.z += g.haz; }
    p.set_mom(k, r1);
  }
}

__global__ void __launch_bounds__(256) energy_p_kernel(const PView p, int np, float qdt_2mc,
                                                       const vpb_interpolator_t *__restrict__ f0, int fi_bytes,
                                                       double *__restrict__ out) {
  __shared__ double ws[8];
  double en = 0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < np; k += gridDim.x * blockDim.x) {
    const float4 r0 = p.pos(k), r1 = p.mom(k);
    const Gather g = gather_fields(f0, fi_bytes, __float_as_int(r0.w), qdt_2mc, r0.x, r0.y, r0.z);
    float v0 = r1.x + g.hax, v1 = r1.y + g.hay, v2 = r1.z + g.haz;   // energy_p.cxx:37-43
    v0 = v0 * v0 + v1 * v1 + v2 * v2;
    v0 /= sqrtf(1.f + v0) + 1.f;
    en += (double)v0 * (double)r1.w;
  }
  en = warp_sum(en);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = en;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int j = 0; j < 8; j++) t += ws[j];
    atomicAdd(out, t);
  }
}

// rho_p.c:43-78: trilinear deposit of q/8V onto the 8 nodes of the particle's voxel.
// A scalar RED costs ~1.3 LSU cycles per LANE wherever it goes (46 ms per 2^30 particles with eight per particle), so
// lanes of a warp that share a voxel first add their eight weights together: two rounds of pointer jumping along the
// list of lanes with the same voxel (MATCH.ANY) leave in every fourth lane of a list the sum of itself and the next
// three, and only those lanes issue REDs.  Changes the order of the float additions into a node, nothing else.
__global__ void __launch_bounds__(256) rho_p_kernel(vpb_field_t *__restrict__ f, const PView p, int np,
                                                    float r8V, const DomainDev g) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const long n_round = ((long)np + 31) & ~31L;                 // whole warps enter the loop together
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < n_round; k += (long)gridDim.x * blockDim.x) {
    const bool valid = k < np;
    const float4 r0 = valid ? p.pos(k) : make_float4(0.f, 0.f, 0.f, 0.f), r1 = valid ? p.mom(k) : make_float4(0.f, 0.f, 0.f, 0.f);
    float t, w[8];
    t = r0.x; w[0] = r8V * r1.w; t *= w[0]; w[1] = w[0] + t; w[0] -= t;
    t = r0.y; w[3] = 1 + t; w[2] = w[0] * w[3]; w[3] *= w[1]; t = 1 - t; w[0] *= t; w[1] *= t;
    t = r0.z; w[7] = 1 + t; w[4] = w[0] * w[7]; w[5] = w[1] * w[7]; w[6] = w[2] * w[7]; w[7] *= w[3];
    t = 1 - t; w[0] *= t; w[1] *= t; w[2] *= t; w[3] *= t;
    const int vox = valid ? __float_as_int(r0.w) : -1 - lane;
    const unsigned peers = __match_any_sync(full, vox);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    const unsigned above = lane < 31 ? (peers & ~((2u << lane) - 1u)) : 0u;
    int nxt = above ? __ffs(above) - 1 : -1;
#pragma unroll
    for (int round = 0; round < 2; round++) {
      const int from = nxt < 0 ? lane : nxt;
#pragma unroll
      for (int c = 0; c < 8; c++) {
        const float o = __shfl_sync(full, w[c], from);
        if (nxt >= 0) w[c] += o;
      }
      const int nn = __shfl_sync(full, nxt, from);
      nxt = nxt < 0 ? -1 : nn;
    }
    if (valid && (rank & 3) == 0) {
      float *rho = &FCOMP(f, g, vox, 15);                                     // rhof = component 15
      const size_t X = 4 * (size_t)g.fqv, Y = X * (size_t)g.sx, Z = X * (size_t)g.sxy;   // floats per voxel step
      red_add(rho, w[0]); red_add(rho + X, w[1]); red_add(rho + Y, w[2]); red_add(rho + X + Y, w[3]);
      red_add(rho + Z, w[4]); red_add(rho + Z + X, w[5]); red_add(rho + Z + Y, w[6]); red_add(rho + Z + Y + X, w[7]);
    }
  }
}

// ---------------------------------------------------------------------------
// sort_p: stable counting sort.  (1) histogram of voxel keys, (2) exclusive scan
// = partition[], (3) claim a slot per particle with an atomic cursor (order
// within a voxel arbitrary) recording only the SOURCE INDEX, (4) one warp per
// voxel re-ranks its segment by source index (which makes the permutation equal
// to the reference's stable scatter, sort_p.c:74), (5) gather the 48-byte
// records through the permutation with 128-bit accesses.
// ---------------------------------------------------------------------------
// Sort key.  L == 0: the particle's voxel (sort_p.c:46-59).  L > 0 ("look-ahead", device-resident driver only): the
// voxel the particle will be in L steps from now if it keeps its velocity, clamped to the local interior.  Sorting
// is a locality heuristic -- any order of the array is physically the same -- and between two sorts `interval`
// steps apart the particles of a chunk are closest to sharing a voxel when they were grouped by where they will be
// half-way: the age of the drift that advance_p sees runs 10,9,..,0,..,9 steps instead of 0..19.
struct SortAhead {
  int L, sx, sy, nx, ny, nz;
  float kx, ky, kz;        // 2 * L * c dt / d{x,y,z}: displacement in cell-local units (a cell spans [-1,1]) per unit u/gamma
};

__device__ __forceinline__ int sort_key_of(const float4 r, const float4 u, const SortAhead &A) {
  const int v = __float_as_int(r.w);
  if (A.L == 0) return v;
  const int ix = v % A.sx, t = v / A.sx, iy = t % A.sy, iz = t / A.sy;
  const float rg = rsqrtf(1.f + (u.x * u.x + (u.y * u.y + u.z * u.z)));
  int cx = ix + (int)floorf((r.x + A.kx * u.x * rg + 1.f) * 0.5f);
  int cy = iy + (int)floorf((r.y + A.ky * u.y * rg + 1.f) * 0.5f);
  int cz = iz + (int)floorf((r.z + A.kz * u.z * rg + 1.f) * 0.5f);
  cx = min(max(cx, 1), A.nx); cy = min(max(cy, 1), A.ny); cz = min(max(cz, 1), A.nz);
  return cx + A.sx * (cy + A.sy * cz);
}

__device__ __forceinline__ int sort_key(const PView &p, int k, const SortAhead &A) {
  if (A.L == 0) return p.voxel(k);
  return sort_key_of(p.pos(k), p.mom(k), A);
}

// histogram of the keys; with look-ahead the keys are kept for the claim pass
__global__ void __launch_bounds__(256) sort_hist_kernel(const PView p, int np, int *__restrict__ count, const SortAhead A,
                                                        int *__restrict__ keys) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < np; k += gridDim.x * blockDim.x) {
    const int key = sort_key(p, k, A);
    if (keys) keys[k] = key;
    atomicAdd(count + key, 1);
  }
}

// Slot claim, aggregated per CTA tile.  A tile of 1024 consecutive particles holds few distinct keys (the array is
// sorted up to the drift since the last sort), so the tile first counts its keys in a shared-memory hash table
// (native integer atomics), then sends ONE global atomic per distinct key to reserve that key's slots, and every
// particle takes base + its rank inside the tile's group.  The kernel was bound by the latency of one returning
// global atomic per group of equal keys in a WARP (17.9 ms per 2^30 drifted particles).
constexpr int kClaimTile = 1024, kClaimHash = 2048, kClaimPer = kClaimTile / 256;

__global__ void __launch_bounds__(256) sort_claim_kernel(const PView p, int np, int *__restrict__ cursor,
                                                         int *__restrict__ perm, const int *__restrict__ keys) {
  __shared__ int hkey[kClaimHash], hcnt[kClaimHash];
  const int ntiles = (np + kClaimTile - 1) / kClaimTile;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    for (int s = threadIdx.x; s < kClaimHash; s += 256) { hkey[s] = -1; hcnt[s] = 0; }
    __syncthreads();
    int slot[kClaimPer], rank[kClaimPer];
#pragma unroll
    for (int j = 0; j < kClaimPer; j++) {
      const int k = tile * kClaimTile + j * 256 + threadIdx.x;
      slot[j] = -1;
      if (k < np) {
        const int v = keys ? keys[k] : p.voxel(k);
        int s = (int)(((unsigned)v * 2654435761u) >> 21) & (kClaimHash - 1);
        for (;;) {   // linear probing; the table is at most half full
          const int old = atomicCAS(&hkey[s], -1, v);
          if (old == -1 || old == v) break;
          s = (s + 1) & (kClaimHash - 1);
        }
        slot[j] = s;
        rank[j] = atomicAdd(&hcnt[s], 1);
      }
    }
    __syncthreads();
    for (int s = threadIdx.x; s < kClaimHash; s += 256)
      if (hkey[s] >= 0) hcnt[s] = atomicAdd(cursor + hkey[s], hcnt[s]);     // count -> base of the tile's group
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kClaimPer; j++)
      if (slot[j] >= 0) perm[hcnt[slot[j]] + rank[j]] = tile * kClaimTile + j * 256 + threadIdx.x;
    __syncthreads();
  }
}

// INV = 0: perm_sorted[destination] = source (for a gather); INV = 1: perm_sorted[source] = destination (for a scatter)
// Segments of more than kRankBig particles (a voxel holding thousands: localised loads, targets, sheets) are not ranked
// here -- counting the smaller sources of each element is quadratic -- but listed in big[] (big[0] = how many,
// big[1..cap] = the voxels) for sort_rank_big_kernel; when the list is full the quadratic loop does run.
constexpr int kRankBig = 1024, kRankBigCap = 8191;
template <int INV>
__global__ void __launch_bounds__(256) sort_rank_kernel(const int *__restrict__ partition, int nv, const int *__restrict__ perm,
                                                        int *__restrict__ perm_sorted, int *__restrict__ big) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int v = blockIdx.x * wpb + (threadIdx.x >> 5); v < nv; v += gridDim.x * wpb) {
    const int b = partition[v], n = partition[v + 1] - b;
    if (n <= 0) continue;
    if (n <= 32) {
      const int mine = lane < n ? perm[b + lane] : 0x7fffffff;
      int r = 0;
      for (int j = 0; j < n; j++) r += (__shfl_sync(0xffffffffu, mine, j) < mine);
      if (lane < n) perm_sorted[INV ? mine : b + r] = INV ? b + r : mine;
    } else {
      if (n > kRankBig) {
        int slot = 0;
        if (lane == 0) slot = atomicAdd(big, 1);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot < kRankBigCap) {
          if (lane == 0) big[1 + slot] = v;
          continue;
        }
      }
      for (int i = lane; i < n; i += 32) {
        const int mine = perm[b + i];
        int r = 0;
        for (int j = 0; j < n; j++) r += (perm[b + j] < mine);
        perm_sorted[INV ? mine : b + r] = INV ? b + r : mine;
      }
    }
  }
}

// The listed segments, one block at a time: bitonic sort of perm[b .. b+n) in place (it is scratch), n log^2 n / 256
// steps per thread.  All compare-exchanges ascend (the first step of every merge mirrors the upper half), so the
// virtual +infinity padding beyond n never moves and is never touched.
template <int INV>
__global__ void __launch_bounds__(256) sort_rank_big_kernel(const int *__restrict__ partition, int *__restrict__ perm,
                                                            int *__restrict__ perm_sorted, const int *__restrict__ big) {
  const int nbig = big[0] < kRankBigCap ? big[0] : kRankBigCap;
  for (int s = blockIdx.x; s < nbig; s += gridDim.x) {
    const int v = big[1 + s];
    const int b = partition[v], n = partition[v + 1] - b;
    int *a = perm + b;
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int k = 2; k <= n2; k <<= 1) {
      for (int j = k >> 1; j >= 1; j >>= 1) {
        const bool flip = j == (k >> 1);
        for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
          const int lo = (t / j) * 2 * j + (t % j);
          const int hi = flip ? (t / j) * 2 * j + (2 * j - 1 - (t % j)) : lo + j;
          if (hi < n) {
            const int x = a[lo], y = a[hi];
            if (y < x) { a[lo] = y; a[hi] = x; }
          }
        }
        __syncthreads();
      }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) perm_sorted[INV ? a[i] : b + i] = INV ? b + i : a[i];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) sort_gather_kernel(const vpb_particle_t *__restrict__ in, vpb_particle_t *__restrict__ out,
                                                          int np, const int *__restrict__ perm) {
  // three 16-byte pieces per particle; consecutive threads take consecutive pieces of the output
  const long n3 = 3L * np;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n3; t += (long)gridDim.x * blockDim.x) {
    const int k = (int)(t / 3), piece = (int)(t - 3L * k);
    reinterpret_cast<float4 *>(out)[t] = __ldg(reinterpret_cast<const float4 *>(in + perm[k]) + piece);
  }
}

// Component planes, out of place: blockIdx.y = plane (0..7 the 4-byte components, 8 the 16-byte tag pairs); coalesced
// writes, reads gathered through the permutation.  Moving single 4-byte words through a permutation is slow once the
// particles have drifted (every word pulls its own 32-byte sector: 142 ms per 2^30 particles 20 steps after a sort;
// a 4-byte scatter is worse, 300 ms) -- the device-resident driver uses vpb_sort_p_planes below instead.
__global__ void __launch_bounds__(256) sort_gather_planes_kernel(const float *__restrict__ in, float *__restrict__ out, long plane_in,
                                                                 long plane_out, int np, const int *__restrict__ perm) {
  const int c = blockIdx.y;
  if (c < 8) {
    const float *src = in + (size_t)c * plane_in;
    float *dst = out + (size_t)c * plane_out;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < np; k += gridDim.x * blockDim.x) dst[k] = __ldg(src + perm[k]);
  } else {
    const float4 *src = reinterpret_cast<const float4 *>(in + 8 * (size_t)plane_in);
    float4 *dst = reinterpret_cast<float4 *>(out + 8 * (size_t)plane_out);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < np; k += gridDim.x * blockDim.x) dst[k] = __ldg(src + perm[k]);
  }
}

// planes -> 48-byte records, both sides coalesced: a warp takes 32 consecutive particles, every lane reads its
// particle's nine plane words, the 96 quads are transposed through shared memory (48-byte lane stride: conflict
// free) and leave as three fully coalesced 512-byte stores.
// HIST: the pass also computes every particle's sort key (it holds the words the key needs), counts it and keeps it
// for the claim pass, which saves the separate histogram pass over the planes.
template <int HIST>
__global__ void __launch_bounds__(256) planes_to_records_kernel(const PView in, float4 *__restrict__ rec, int np, int *__restrict__ count,
                                                                const SortAhead A, int *__restrict__ keys) {
  __shared__ float4 tile[8][96];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nblk = (np + 31) >> 5;
  for (int blk = blockIdx.x * 8 + w; blk < nblk; blk += gridDim.x * 8) {
    const int k = blk * 32 + lane;
    if (k < np) {
      const float4 r = in.pos(k), u = in.mom(k);
      tile[w][3 * lane] = r;
      tile[w][3 * lane + 1] = u;
      tile[w][3 * lane + 2] = in.tag(k);
      if (HIST) {
        const int key = sort_key_of(r, u, A);
        keys[k] = key;
        atomicAdd(count + key, 1);
      }
    }
    __syncwarp();
    const int nq = 3 * (np - blk * 32 < 32 ? np - blk * 32 : 32);
    float4 *o = rec + 96 * (size_t)blk;
#pragma unroll
    for (int j = 0; j < 3; j++)
      if (lane + 32 * j < nq) __stcs(o + lane + 32 * j, tile[w][lane + 32 * j]);
    __syncwarp();
  }
}

// The permutation moves whole 48-byte records (three 16-byte requests per particle, as for the reference layout):
// the planes are first copied to records in the scratch array (both sides coalesced), then records are gathered
// through the permutation and written straight back as planes.
// Records gathered through the permutation and written back as planes.  Array order: a particle's source lies within
// a few x-rows (y) or planes (z) of its destination, and the planes z+-1 are 400 MB of records away, so each source
// line comes from DRAM about three times (170 GB read per 2^30 particles 20 steps after a sort).  Visiting the
// output in y-blocked / z-inner order was measured: DRAM reads fall to 130 GB but the kernel gets slower (46-62 ms
// against 38) because the rows in flight no longer cover the chip (profiles/README.md).
__global__ void __launch_bounds__(256) sort_gather_records_to_planes_kernel(const float4 *__restrict__ rec, const PView out, int np,
                                                                            const int *__restrict__ perm) {
  const size_t pl = (size_t)out.plane;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < np; k += gridDim.x * blockDim.x) {
    const float4 *r = rec + 3 * (size_t)__ldcs(perm + k);
    const float4 a = __ldg(r), b = __ldg(r + 1), c = __ldg(r + 2);
    // streaming stores: the output should not push source lines out of L2
    float *o = out.b + k;
    __stcs(o, a.x); __stcs(o + pl, a.y); __stcs(o + 2 * pl, a.z); __stcs(o + 3 * pl, a.w);
    __stcs(o + 4 * pl, b.x); __stcs(o + 5 * pl, b.y); __stcs(o + 6 * pl, b.z); __stcs(o + 7 * pl, b.w);
    __stcs(reinterpret_cast<float4 *>(out.b + 8 * pl) + k, c);
  }
}

// AoS <-> component planes (uploads, downloads, tests)
__global__ void __launch_bounds__(256) particle_convert_kernel(const PView dst, const PView src, long np) {
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < np; k += (long)gridDim.x * blockDim.x) {
    dst.set_pos(k, src.pos(k));
    dst.set_mom(k, src.mom(k));
    dst.set_tag(k, src.tag(k));
  }
}

}  // namespace vpb

using namespace vpb;


static int grid_for(long n, int tb) {
  long b = (n + tb - 1) / tb;
  const long cap = (long)ctx().sm_count * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" {

void vpb_center_p(vpb_domain_t *dom, vpb_particle_t *d_p, int np, float q_m, const vpb_interpolator_t *d_f) {
  if (np < 0) VPB_ERROR("Bad number of particles");
  if (!d_f) VPB_ERROR("Bad interpolator");
  if (!dom) VPB_ERROR("Bad grid");
  if (np == 0) return;
  const float qdt_2mc = (float)(0.5 * q_m * dom->d.dt / dom->d.cvac);   // center_p.cxx:171
  const float qdt_4mc = (float)(0.5 * qdt_2mc);                         // center_p.cxx:15
  center_kernel<0><<<grid_for(np, 256), 256, 0, ctx().stream>>>(PView(d_p, dom->d.p_plane), np, qdt_2mc, qdt_4mc, d_f, dom->d.fi_bytes);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

void vpb_uncenter_p(vpb_domain_t *dom, vpb_particle_t *d_p, int np, float q_m, const vpb_interpolator_t *d_f) {
  if (np < 0) VPB_ERROR("Bad number of particles");
  if (!d_f) VPB_ERROR("Bad interpolator");
  if (!dom) VPB_ERROR("Bad grid");
  if (np == 0) return;
  const float fwd = (float)(0.5 * q_m * dom->d.dt / dom->d.cvac);       // uncenter_p.cxx:171
  const float qdt_2mc = -fwd;                                           // uncenter_p.cxx:14
  const float qdt_4mc = (float)(-0.5 * fwd);                            // uncenter_p.cxx:15
  center_kernel<1><<<grid_for(np, 256), 256, 0, ctx().stream>>>(PView(d_p, dom->d.p_plane), np, qdt_2mc, qdt_4mc, d_f, dom->d.fi_bytes);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

void vpb_energy_p(vpb_domain_t *dom, const vpb_particle_t *d_p, int np, float q_m, const vpb_interpolator_t *d_f, double *d_en) {
  if (np < 0) VPB_ERROR("Bad number of particles");
  if (!d_f) VPB_ERROR("Bad interpolator");
  if (!dom) VPB_ERROR("Bad grid");
  VPB_CUDA(cudaMemsetAsync(d_en, 0, sizeof(double), ctx().stream));
  if (np == 0) return;
  const float qdt_2mc = (float)(0.5 * q_m * dom->d.dt / dom->d.cvac);
  energy_p_kernel<<<grid_for(np, 256), 256, 0, ctx().stream>>>(PView(d_p, dom->d.p_plane), np, qdt_2mc, d_f, dom->d.fi_bytes, d_en);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

void vpb_accumulate_rho_p(vpb_domain_t *dom, vpb_field_t *d_f, const vpb_particle_t *d_p, int np) {
  if (!d_f) VPB_ERROR("Bad field");
  if (!d_p) VPB_ERROR("Bad particle array");
  if (np < 0) VPB_ERROR("Bad number of particles");
  if (!dom) VPB_ERROR("Bad grid");
  if (np == 0) return;
  const DomainDev &g = dom->d;
  const float r8V = (float)(0.125 * g.rdx * g.rdy * g.rdz);             // rho_p.c:37
  rho_p_kernel<<<grid_for(np, 256), 256, 0, ctx().stream>>>(d_f, PView(d_p, g.p_plane), np, r8V, g);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

// in_place: d_in holds component planes, d_out is scratch and the sorted planes return to d_in.
// lookahead > 0: look-ahead grouping (SortAhead), in-place only.
static void sort_particles(vpb_domain_t *dom, const vpb_particle_t *d_in, vpb_particle_t *d_out, int np, int *d_partition, bool in_place,
                           int lookahead) {
  if (!dom) VPB_ERROR("Bad grid");
  if (!d_partition) VPB_ERROR("Bad partition");
  if (np < 0) VPB_ERROR("Bad number of particles");
  Context &c = ctx();
  ProfScope prof(1);
  const int nv = dom->d.nv, nv1 = nv + 1;
  // scratch: cursor[nv1] | perm[np] | perm_sorted[np] (also the look-ahead keys until the rank pass) | scan scratch
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t off_perm = al((size_t)nv1 * 4), off_perm2 = off_perm + al((size_t)np * 4 + 4),
               off_big = off_perm2 + al((size_t)np * 4 + 4), off_scan = off_big + al((size_t)(kRankBigCap + 1) * 4);
  char *s = (char *)scratch(off_scan + scan_scratch_bytes(nv1));
  int *cursor = (int *)s, *perm = (int *)(s + off_perm), *perm2 = (int *)(s + off_perm2), *big = (int *)(s + off_big);
  const DomainDev &gd = dom->d;
  SortAhead ahead;
  ahead.L = lookahead > 0 ? lookahead : 0;
  ahead.sx = gd.sx; ahead.sy = gd.sy; ahead.nx = gd.nx; ahead.ny = gd.ny; ahead.nz = gd.nz;
  ahead.kx = 2.f * ahead.L * gd.cvac * gd.dt * gd.rdx;
  ahead.ky = 2.f * ahead.L * gd.cvac * gd.dt * gd.rdy;
  ahead.kz = 2.f * ahead.L * gd.cvac * gd.dt * gd.rdz;
  // in place (planes): the transposition to records doubles as the histogram pass and always leaves the keys
  const bool fused = in_place && np > 0;
  int *keys = (ahead.L || fused) ? perm2 : nullptr;
  VPB_CUDA(cudaMemsetAsync(cursor, 0, (size_t)nv1 * 4, c.stream));
  if (fused) {
    if (!d_in || !d_out) VPB_ERROR("Bad particle array");
    planes_to_records_kernel<1><<<grid_for(np, 256), 256, 0, c.stream>>>(PView(d_in, dom->d.p_plane), reinterpret_cast<float4 *>(d_out), np,
                                                                         cursor, ahead, keys);
  } else if (np > 0) {
    sort_hist_kernel<<<grid_for(np, 256), 256, 0, c.stream>>>(PView(d_in, dom->d.p_plane), np, cursor, ahead, keys);
  }
  exclusive_scan_i32(cursor, d_partition, nv1, s + off_scan, c.stream);   // partition[nv] = np (sort_p.c:54-59)
  count_launch(1 + scan_launches(nv1));
  if (np == 0) return;
  if (!d_in || !d_out) VPB_ERROR("Bad particle array");
  VPB_CUDA(cudaMemcpyAsync(cursor, d_partition, (size_t)nv1 * 4, cudaMemcpyDeviceToDevice, c.stream));
  sort_claim_kernel<<<grid_for(np, 256), 256, 0, c.stream>>>(PView(d_in, dom->d.p_plane), np, cursor, perm, keys);
  // The reference's out-of-place sort is stable (sort_p.c:74): slots claimed by atomics are re-ranked by source index
  // within every voxel.  A look-ahead grouping is not the reference's order to begin with, so it keeps the claim order
  // (the particles of a group in arbitrary order; 15 ms less per 2^30 particles).
  if (ahead.L) perm2 = perm;
  else {
    VPB_CUDA(cudaMemsetAsync(big, 0, sizeof(int), c.stream));
    sort_rank_kernel<0><<<grid_for((long)nv * 32, 256), 256, 0, c.stream>>>(d_partition, nv, perm, perm2, big);
    sort_rank_big_kernel<0><<<c.sm_count, 256, 0, c.stream>>>(d_partition, perm, perm2, big);   // returns at once when nothing is listed
    count_launch();
  }
  if (in_place) {   // vpb_sort_p_planes: d_out is scratch, the sorted planes return to d_in
    sort_gather_records_to_planes_kernel<<<grid_for(np, 256), 256, 0, c.stream>>>(reinterpret_cast<const float4 *>(d_out),
                                                                                  PView(d_in, dom->d.p_plane), np, perm2);
    count_launch();
  } else if (dom->d.p_plane > 0) {
    sort_gather_planes_kernel<<<dim3(grid_for(np, 256), 9), 256, 0, c.stream>>>(reinterpret_cast<const float *>(d_in), reinterpret_cast<float *>(d_out),
                                                                                dom->d.p_plane, dom->d.p_plane, np, perm2);
  } else {
    sort_gather_kernel<<<grid_for(3L * np, 256), 256, 0, c.stream>>>(d_in, d_out, np, perm2);
  }
  count_launch(4);
  VPB_CUDA(cudaGetLastError());
}

void vpb_sort_p(vpb_domain_t *dom, const vpb_particle_t *d_in, vpb_particle_t *d_out, int np, int *d_partition) {
  sort_particles(dom, d_in, d_out, np, d_partition, false, 0);
}

// Stable counting sort of a component-plane array IN PLACE: d_p holds the sorted planes on return, d_tmp (same
// capacity) is scratch.  Same permutation as vpb_sort_p.
void vpb_sort_p_planes(vpb_domain_t *dom, vpb_particle_t *d_p, vpb_particle_t *d_tmp, int np, int *d_partition) {
  if (!dom) VPB_ERROR("Bad grid");
  if (dom->d.p_plane <= 0) VPB_ERROR("the domain keeps its particles in the reference layout: use vpb_sort_p");
  sort_particles(dom, d_p, d_tmp, np, d_partition, true, 0);
}

// Same, grouping the particles by the voxel they will be in `lookahead` steps from now (see SortAhead): d_partition
// then describes those groups, not the particles' current voxels.
void vpb_sort_p_planes_ahead(vpb_domain_t *dom, vpb_particle_t *d_p, vpb_particle_t *d_tmp, int np, int *d_partition, int lookahead) {
  if (!dom) VPB_ERROR("Bad grid");
  if (dom->d.p_plane <= 0) VPB_ERROR("the domain keeps its particles in the reference layout: use vpb_sort_p");
  if (lookahead < 0) VPB_ERROR("Bad look-ahead");
  sort_particles(dom, d_p, d_tmp, np, d_partition, true, lookahead);
}

// Particle layout of a domain's device-resident species arrays (include/vpic_b200.h "Device particle layout")
void vpb_domain_set_particle_layout(vpb_domain_t *dom, long plane) {
  if (!dom) VPB_ERROR("Bad grid");
  if (plane < 0 || (plane & 63)) VPB_ERROR("particle plane stride must be a non-negative multiple of 64 (got %ld)", plane);
  dom->d.p_plane = plane;
}

long vpb_domain_particle_layout(const vpb_domain_t *dom) {
  if (!dom) VPB_ERROR("Bad grid");
  return dom->d.p_plane;
}

// to_planes != 0: d_src is particle_t[np], d_dst the domain's component planes; else the other way round
void vpb_particle_convert(vpb_domain_t *dom, vpb_particle_t *d_dst, const vpb_particle_t *d_src, long np, int to_planes) {
  if (!dom) VPB_ERROR("Bad grid");
  if (!d_dst || !d_src) VPB_ERROR("Bad particle array");
  if (dom->d.p_plane <= 0) VPB_ERROR("the domain keeps its particles in the reference layout");
  if (np <= 0) return;
  const PView planes(to_planes ? (const void *)d_dst : (const void *)d_src, dom->d.p_plane), aos(to_planes ? (const void *)d_src : (const void *)d_dst, 0);
  particle_convert_kernel<<<grid_for(np, 256), 256, 0, ctx().stream>>>(to_planes ? planes : aos, to_planes ? aos : planes, np);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

}  // extern "C"
