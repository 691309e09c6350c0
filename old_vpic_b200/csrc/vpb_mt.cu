// vpb_mt.cu -- SURVEY.md 8(f)4: the particle load of a deck, batched on the device, from the REFERENCE'S OWN
// random-number stream.
//
// A deck loads its plasma one particle at a time on the host: seed_rand(seed), then per particle a few
// uniform_rand / maxwellian_rand and one inject_particle (src/vpic/vpic.hxx:491-505, src/vpic/misc.cxx:16-105);
// at 10^9 particles per GPU that serial loop takes minutes.  Here the same stream and the same arithmetic run on
// the device:
//   * the generator is the reference's (src/util/mtrand/mtrand.c:16-62): MT19937 with seed_mt_rng's seeding.  One
//     CTA advances the 624-word state in shared memory -- the recurrence splits into three phases of 227, 227 and
//     170 independent words -- and appends tempered words to a stream buffer;
//   * mt_drand is drand53_o of a word pair (mtrand_conv.h:58); mt_drandn is the 256-layer ziggurat over word PAIRS
//     (mtrand.c:395-438): one pair when the trapezoid test accepts (98.8 %), one or two more pairs per rejection
//     round otherwise.  How many words a deviate consumes therefore depends on the words themselves, and the
//     position of record r in the stream on every record before it.  The stream is PARSED in parallel: (1) for
//     every pair position the length a normal deviate starting there would have, (2) for every position the
//     length of a whole record (the deck's token sequence, e.g. UUUNNNNNN) starting there, (3) per chunk of 1024
//     positions and per possible entry offset the exit offset into the next chunk and the number of records, (4) a
//     serial walk over the chunks only (one thread, tables in shared memory), (5) the record starts, (6) one
//     thread per record evaluates its deviates;
//   * the layer table is rebuilt at start-up by the construction that made the reference's (make_zig.c: bisection
//     on the tail start in long double, inverse density through a double sqrt);
//   * the two libm calls of the rejection branch decide acceptance only (a 1-ulp difference between the device's
//     and the host's exp flips a decision with probability ~1e-16); the tail layer's VALUE, R - log(u)/R (2.6e-4 of
//     the deviates), is recomputed on the host with the host's log and patched in, so that every double equals
//     what the reference computes on the same machine;
//   * inject_particle's placement arithmetic (misc.cxx:45-89, double precision, far-wall rules) runs per particle,
//     and the particles that belong to the local domain are appended in stream order by a stable compaction.
// The generator state can be taken over from and handed back to a host program in the reference's own format
// (get_mt_rng_state / set_mt_rng_state, mtrand.c:74-124).
#include <math.h>
#include <string.h>
#include <vector>
#include "vpb_common.cuh"
#include "vpb_pview.cuh"
#include "vpb_scan.cuh"

namespace vpb {

constexpr int kMtN = 624, kMtM = 397;
constexpr int kChunk = 1024;       // pair positions per chunk of the parse
constexpr int kEntries = 64;       // entry offsets tried per chunk: a record may reach this far into the next chunk
constexpr int kMaxProg = 32;

__host__ __device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v) {
  return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((0u - (v & 1u)) & 0x9908b0dfu);
}
__host__ __device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
static uint32_t mt_untemper(uint32_t y) {
  y ^= y >> 18;
  y ^= (y << 15) & 0xefc60000u;
  uint32_t t = y;                                  // y ^= (y << 7) & mask, inverted 7 bits at a time
  for (int k = 0; k < 5; k++) t = y ^ ((t << 7) & 0x9d2c5680u);
  y = t;
  t = y;                                           // y ^= y >> 11
  for (int k = 0; k < 3; k++) t = y ^ (t >> 11);
  return t;
}

// nblocks x 624 tempered words appended to out; state (624 raw words) advanced in place
__global__ void __launch_bounds__(256) mt_generate_kernel(uint32_t *__restrict__ state, uint32_t *__restrict__ out, long nblocks) {
  __shared__ uint32_t s[kMtN];
  const int t = threadIdx.x;
  for (int i = t; i < kMtN; i += 256) s[i] = state[i];
  __syncthreads();
  for (long b = 0; b < nblocks; b++) {
    uint32_t v = 0;
    if (t < 227) v = s[t + kMtM] ^ mt_twist(s[t], s[t + 1]);                                  // words 0..226 (mtrand.c:31)
    __syncthreads();
    if (t < 227) s[t] = v;
    __syncthreads();
    if (t < 227) v = s[t] ^ mt_twist(s[227 + t], s[228 + t]);                                 // words 227..453 (:32)
    __syncthreads();
    if (t < 227) s[227 + t] = v;
    __syncthreads();
    if (t < 170) v = s[227 + t] ^ mt_twist(s[454 + t], t == 169 ? s[0] : s[455 + t]);         // words 454..623 (:32-33)
    __syncthreads();
    if (t < 170) s[454 + t] = v;
    __syncthreads();
    for (int i = t; i < kMtN; i += 256) out[b * kMtN + i] = mt_temper(s[i]);
  }
  for (int i = t; i < kMtN; i += 256) state[i] = s[i];
}

struct ZigTable {
  const double *x, *y;     // 257 entries each (device)
  double r, scale;
};

__device__ __forceinline__ double d53_o(uint32_t a, uint32_t b) { return ((a >> 5) * 67108864. + (b >> 6) + 1.5) * (1. / 9007199254740994.); }
__device__ __forceinline__ double d53_c(uint32_t a, uint32_t b) { return ((a >> 5) * 67108864. + ((b >> 6) + (b & 1u))) * (1. / 9007199254740992.); }
__device__ __forceinline__ double d53_c1(uint32_t a, uint32_t b) { return ((a >> 5) * 67108864. + ((b >> 6) + 1u)) * (1. / 9007199254740992.); }

// mt_drandn (mtrand.c:395-438) reading pairs w[2p], w[2p+1] from position p on; at most npairs positions exist.
// Returns the number of pairs consumed (0: the stream ends before the deviate is complete); *val receives the
// deviate, *tail the position of the pair a tail value was made from (-1: not a tail value).
__device__ __forceinline__ int zig_normal(const uint32_t *__restrict__ w, long p, long npairs, const ZigTable &Z, double *val, long *tail) {
  long q = p;
  for (;;) {
    if (q >= npairs) return 0;
    uint32_t a = w[2 * q], b = w[2 * q + 1];
    q++;
    const uint32_t s = a & 1u, i = (a & 0x1feu) >> 1;
    const double j = 4294967296. * b + (double)((a & 0xfffff800u) + ((a & 0x400u) << 1));
    double x = j * (Z.scale * Z.x[i + 1]);
    long tl = -1;
    bool ok = x < Z.x[i];
    if (!ok) {
      if (q >= npairs) return 0;
      a = w[2 * q]; b = w[2 * q + 1];
      q++;
      double y = d53_c(a, b);
      if (i != 255) y = Z.y[i] + (Z.y[i + 1] - Z.y[i]) * y;
      else {
        if (q >= npairs) return 0;
        a = w[2 * q]; b = w[2 * q + 1];
        tl = q;
        q++;
        x = Z.r - (1. / Z.r) * log(d53_c1(a, b));
        y *= exp(-Z.r * (x - 0.5 * Z.r));
      }
      ok = y < exp(-0.5 * x * x);
    }
    if (ok) {
      if (val) *val = s ? -x : x;
      if (tail) *tail = tl;
      return (int)(q - p);
    }
  }
}

// (1) length in pairs of a normal deviate starting at every position (0: runs off the end, 255: longer than 254)
__global__ void __launch_bounds__(256) mt_normal_len_kernel(const uint32_t *__restrict__ w, long npairs, const ZigTable Z, unsigned char *__restrict__ nlen) {
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += (long)gridDim.x * blockDim.x) {
    const int n = zig_normal(w, p, npairs, Z, nullptr, nullptr);
    nlen[p] = (unsigned char)(n > 254 ? 255 : n);
  }
}

struct Program {
  int len;
  char tok[kMaxProg];     // 'U' or 'N'
};

// (2) length in pairs of a whole record starting at every position (0xffff: incomplete)
__global__ void __launch_bounds__(256) mt_record_len_kernel(const unsigned char *__restrict__ nlen, long npairs, const Program P,
                                                            unsigned short *__restrict__ jump) {
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += (long)gridDim.x * blockDim.x) {
    long q = p;
    bool ok = true;
    for (int t = 0; t < P.len && ok; t++) {
      if (q >= npairs) { ok = false; break; }
      if (P.tok[t] == 'U') q++;
      else {
        const int n = nlen[q];
        if (n == 0 || n == 255) ok = false; else q += n;
      }
    }
    jump[p] = (ok && q - p < 0xffff) ? (unsigned short)(q - p) : (unsigned short)0xffff;
  }
}

// (3) one CTA per chunk, one thread per entry offset: walk the chunk record by record
__global__ void __launch_bounds__(kEntries) mt_chunk_walk_kernel(const unsigned short *__restrict__ jump, long npairs,
                                                                 unsigned short *__restrict__ exit_off, unsigned short *__restrict__ count) {
  __shared__ unsigned short sj[kChunk];
  const long c = blockIdx.x, base = c * kChunk;
  for (int i = threadIdx.x; i < kChunk; i += kEntries) sj[i] = base + i < npairs ? jump[base + i] : (unsigned short)0xffff;
  __syncthreads();
  int pos = threadIdx.x, n = 0;
  bool dead = false;
  while (pos < kChunk) {
    const unsigned short j = sj[pos];
    if (j == 0xffff) { dead = true; break; }
    pos += j;
    n++;
  }
  // dead: the stream ends inside this chunk on this path; otherwise the next record starts pos - kChunk into the
  // next chunk (0xfffe: further than the entry offsets tried -- a record of more than 64 pairs)
  exit_off[c * kEntries + threadIdx.x] = dead ? (unsigned short)0xffff : (pos - kChunk < kEntries ? (unsigned short)(pos - kChunk) : (unsigned short)0xfffe);
  count[c * kEntries + threadIdx.x] = (unsigned short)n;
}

// (4) the only serial part: which entry offset each chunk is really entered at, and how many records precede it.
// One CTA; the two tables go through shared memory a tile of chunks at a time, thread 0 walks.
// out[0] = complete records in the stream, out[1] = error flag (record longer than kEntries pairs)
constexpr int kScanTile = 128;
__global__ void __launch_bounds__(256) mt_chunk_scan_kernel(const unsigned short *__restrict__ exit_off, const unsigned short *__restrict__ count,
                                                            long nchunks, unsigned char *__restrict__ entry, long *__restrict__ first_record,
                                                            long *__restrict__ out) {
  __shared__ unsigned short se[kScanTile * kEntries], sc[kScanTile * kEntries];
  __shared__ int s_e, s_stop;
  __shared__ long s_base;
  if (threadIdx.x == 0) { s_e = 0; s_base = 0; s_stop = 0; }
  __syncthreads();
  for (long c0 = 0; c0 < nchunks; c0 += kScanTile) {
    const int nt = (int)min((long)kScanTile, nchunks - c0);
    for (int i = threadIdx.x; i < nt * kEntries; i += blockDim.x) { se[i] = exit_off[c0 * kEntries + i]; sc[i] = count[c0 * kEntries + i]; }
    __syncthreads();
    if (threadIdx.x == 0 && !s_stop) {
      int e = s_e;
      long base = s_base;
      for (int k = 0; k < nt; k++) {
        entry[c0 + k] = (unsigned char)e;
        first_record[c0 + k] = base;
        base += sc[k * kEntries + e];
        const unsigned short x = se[k * kEntries + e];
        if (x >= 0xfffe) {
          s_stop = 1;
          if (x == 0xfffe) out[1] = 1;
          for (long r = c0 + k + 1; r < nchunks; r++) { entry[r] = 255; first_record[r] = base; }   // never entered
          break;
        }
        e = x;
      }
      s_e = e;
      s_base = base;
    }
    __syncthreads();
    if (s_stop) break;
  }
  if (threadIdx.x == 0) out[0] = s_base;
}

// (5) record starts: thread per chunk walks from its real entry.  The thread that meets record n_want-1 also leaves
// the position where record n_want would start (= the new cursor).
__global__ void __launch_bounds__(128) mt_record_starts_kernel(const unsigned short *__restrict__ jump, long npairs, long nchunks,
                                                               const unsigned char *__restrict__ entry, const long *__restrict__ first_record,
                                                               long n_want, unsigned int *__restrict__ starts, long *__restrict__ end_pos) {
  const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunks || entry[c] == 255) return;
  long pos = c * kChunk + entry[c], r = first_record[c];
  const long lim = min((c + 1) * (long)kChunk, npairs);
  while (pos < lim && r < n_want) {
    const unsigned short j = jump[pos];
    if (j == 0xffff) break;
    starts[r] = (unsigned int)pos;
    pos += j;
    if (r == n_want - 1) *end_pos = pos;
    r++;
  }
}

// (6) the deviates of every record; tail values are listed for the host (see the top of the file)
__global__ void __launch_bounds__(256) mt_eval_kernel(const uint32_t *__restrict__ w, long npairs, const ZigTable Z, const Program P,
                                                      const unsigned int *__restrict__ starts, long n, double *__restrict__ out,
                                                      unsigned int *__restrict__ n_tail, long *__restrict__ tail_where, unsigned int *__restrict__ tail_pair,
                                                      unsigned int tail_cap) {
  for (long r = (long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long)gridDim.x * blockDim.x) {
    long q = starts[r];
    for (int t = 0; t < P.len; t++) {
      double v;
      if (P.tok[t] == 'U') { v = d53_o(w[2 * q], w[2 * q + 1]); q++; }
      else {
        long tl;
        q += zig_normal(w, q, npairs, Z, &v, &tl);
        if (tl >= 0) {     // where it went, the two words it was made from, its sign
          const unsigned int k = atomicAdd(n_tail, 1u);
          if (k < tail_cap) {
            tail_where[k] = (r * P.len + t) | (v < 0 ? (1L << 62) : 0L);
            tail_pair[2 * k] = w[2 * tl];
            tail_pair[2 * k + 1] = w[2 * tl + 1];
          }
        }
      }
      out[r * P.len + t] = v;
    }
  }
}

__global__ void mt_patch_kernel(double *__restrict__ out, const long *__restrict__ where, const double *__restrict__ val, unsigned int n) {
  const unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[where[k]] = val[k];
}

// ---- batched inject_particle (misc.cxx:16-105, age = 0, update_rhob = 0) ---------------------------------------
struct LoadMap {
  int col[6];            // columns of the draw table: three uniform deviates (position), three normal deviates (momentum)
  double lo[3], hi[3];   // uniform_rand(lo, hi) = lo*(1-d) + hi*d   (vpic.hxx:497-500)
  double dev[3];         // maxwellian_rand(dev) = dev * n           (vpic.hxx:503-505)
};

struct GridBox {
  double x0, y0, z0, x1, y1, z1;
  int nx, ny, nz;
  int far_shared[3];     // bc[BOUNDARY(1,0,0)] >= 0 etc.: a particle exactly on the far wall belongs to the neighbour
};

__device__ __forceinline__ bool place(double &x, int &ix, double x0, double x1, int nx, int far_shared) {
  if ((x < x0) | (x > x1) | ((x == x1) & (far_shared != 0))) return false;     // misc.cxx:37-39
  x = ((double)nx) * ((x - x0) / (x1 - x0));                                     // :53-59
  ix = (int)x;
  x -= (double)ix;
  x = (x + x) - 1;
  if (ix == nx) { x = 1; ix = nx - 1; }
  ix++;
  return true;
}

// pass 0: keep[k] = does particle k belong to this rank; pass 1: write the kept ones at np0 + slot[k]
template <int PASS>
__global__ void __launch_bounds__(256) inject_table_kernel(const double *__restrict__ tab, int stride, long n, const LoadMap M, const GridBox G,
                                                           int *__restrict__ keep, const int *__restrict__ slot, const PView p, int np0, int max_np,
                                                           float q, long tag, long tag_step, int *__restrict__ overflow) {
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x) {
    const double *row = tab + k * (long)stride;
    double d = row[M.col[0]];
    double x = M.lo[0] * (1 - d) + M.hi[0] * d;
    d = row[M.col[1]];
    double y = M.lo[1] * (1 - d) + M.hi[1] * d;
    d = row[M.col[2]];
    double z = M.lo[2] * (1 - d) + M.hi[2] * d;
    int ix, iy, iz;
    const bool in = place(x, ix, G.x0, G.x1, G.nx, G.far_shared[0]) & place(y, iy, G.y0, G.y1, G.ny, G.far_shared[1]) &
                    place(z, iz, G.z0, G.z1, G.nz, G.far_shared[2]);
    if (PASS == 0) { keep[k] = in ? 1 : 0; continue; }
    if (!in) continue;
    const long pos = (long)np0 + slot[k];
    if (pos >= max_np) { atomicAdd(overflow, 1); continue; }                        // "No room to inject particle" (:43)
    const double ux = M.dev[0] * row[M.col[3]], uy = M.dev[1] * row[M.col[4]], uz = M.dev[2] * row[M.col[5]];
    p.set_pos(pos, make_float4((float)x, (float)y, (float)z, __int_as_float(ix + (G.nx + 2) * (iy + (G.ny + 2) * iz))));
    p.set_mom(pos, make_float4((float)ux, (float)uy, (float)uz, q));
    const longlong2 tags = make_longlong2(tag + k * tag_step, 0);
    p.set_tag(pos, *reinterpret_cast<const float4 *>(&tags));
  }
}

}  // namespace vpb

using namespace vpb;

// ---- host side --------------------------------------------------------------------------------------------------
struct vpb_mt {
  uint32_t *d_state = nullptr;     // 624 raw words: the state the next block is generated from
  uint32_t *d_words = nullptr;     // tempered words, whole 624-word blocks from position 0
  size_t cap_words = 0, n_valid = 0, cur = 0;
  double *d_zig = nullptr;         // zig_x[257] | zig_y[257]
  double zig_r = 0;
  uint32_t h_state[kMtN];          // host copy of the seed / imported state until the first block is generated
  double *d_table = nullptr;       // deviate table of the load loops (grow-only, freed with the generator)
  size_t table_cap = 0;
};

// make_zig.c:9-62.  The inverse density goes through a double sqrt there, which is part of what the table is.
static long double zig_pdf(long double x) { return expl(-0.5l * x * x); }
static long double zig_pdf_inv(long double y) { return (y <= 0 || y >= 1) ? 0 : (long double)sqrt((double)(-2.0l * logl(y))); }
static long double zig_build(long double *x, long double *y, int N, long double r) {
  const long double v = zig_pdf(r) / r + r * zig_pdf(r);
  x[N] = v / zig_pdf(r);
  y[N] = zig_pdf(x[N]);
  x[N - 1] = r;
  y[N - 1] = zig_pdf(x[N - 1]);
  for (int n = N - 2; n > 0; n--) {
    x[n] = zig_pdf_inv(y[n + 1] + v / x[n + 1]);
    y[n] = zig_pdf(x[n]);
  }
  x[0] = 0;
  y[0] = zig_pdf(x[0]);
  return v - (x[1] - x[0]) * (y[0] - y[1]);
}

static void zig_table_host(double *x_out, double *y_out, double *r_out) {
  static double zx[257], zy[257], zr;
  static bool ready = false;
  if (!ready) {
    long double x[257], y[257], a = 0, b = 10, r;
    for (;;) {
      r = 0.5 * (a + b);
      if (r == a || r == b) break;
      const long double dv = zig_build(x, y, 256, r);
      if (dv == 0) break;
      if (dv > 0) a = r; else b = r;
    }
    for (int n = 0; n <= 256; n++) { zx[n] = (double)x[n]; zy[n] = (double)y[n]; }
    zr = (double)r;
    ready = true;
  }
  memcpy(x_out, zx, sizeof(zx));
  memcpy(y_out, zy, sizeof(zy));
  *r_out = zr;
}

static double h_d53_c1(uint32_t a, uint32_t b) { return ((a >> 5) * 67108864. + ((b >> 6) + 1u)) * (1. / 9007199254740992.); }

static void mt_upload_state(vpb_mt *m) {
  VPB_CUDA(cudaMemcpyAsync(m->d_state, m->h_state, sizeof(m->h_state), cudaMemcpyHostToDevice, ctx().stream));
  VPB_CUDA(cudaStreamSynchronize(ctx().stream));
}

// make sure at least `want` words follow the cursor; whole blocks before the cursor's block are dropped
static void mt_ensure(vpb_mt *m, size_t want) {
  cudaStream_t st = ctx().stream;
  if (m->n_valid - m->cur >= want) return;
  const size_t drop = (m->cur / kMtN) * kMtN, keep = m->n_valid - drop;
  const size_t more = ((want - (m->n_valid - m->cur)) + kMtN - 1) / kMtN;
  const size_t need = keep + more * kMtN;
  if (need > m->cap_words || drop) {
    uint32_t *nw = m->d_words;
    size_t ncap = m->cap_words;
    if (need > m->cap_words) {
      ncap = need + need / 8 + 4 * kMtN;
      VPB_CUDA(cudaMalloc(&nw, ncap * sizeof(uint32_t)));
    }
    if (keep) {
      if (nw != m->d_words) VPB_CUDA(cudaMemcpyAsync(nw, m->d_words + drop, keep * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
      else if (drop) {   // same buffer: through scratch (the ranges may overlap)
        void *tmp = scratch(keep * sizeof(uint32_t));
        VPB_CUDA(cudaMemcpyAsync(tmp, m->d_words + drop, keep * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        VPB_CUDA(cudaMemcpyAsync(nw, tmp, keep * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
      }
    }
    if (nw != m->d_words) {
      VPB_CUDA(cudaStreamSynchronize(st));
      if (m->d_words) cudaFree(m->d_words);
      m->d_words = nw;
      m->cap_words = ncap;
    }
    m->cur -= drop;
    m->n_valid = keep;
  }
  mt_generate_kernel<<<1, 256, 0, st>>>(m->d_state, m->d_words + m->n_valid, (long)more);
  count_launch();
  m->n_valid += more * kMtN;
}

extern "C" {

// seed_mt_rng (mtrand.c:54-62)
vpb_mt_t *vpb_mt_create(unsigned int seed) {
  ctx();
  vpb_mt *m = new vpb_mt;
  m->h_state[0] = seed ^ 0x900df00cu;
  for (int j = 1; j < kMtN; j++) m->h_state[j] = 1812433253u * (m->h_state[j - 1] ^ (m->h_state[j - 1] >> 30)) + (uint32_t)j;
  VPB_CUDA(cudaMalloc(&m->d_state, sizeof(m->h_state)));
  mt_upload_state(m);
  double tab[514];
  zig_table_host(tab, tab + 257, &m->zig_r);
  VPB_CUDA(cudaMalloc(&m->d_zig, sizeof(tab)));
  VPB_CUDA(cudaMemcpy(m->d_zig, tab, sizeof(tab), cudaMemcpyHostToDevice));
  return m;
}

void vpb_mt_destroy(vpb_mt_t *m) {
  if (!m) return;
  cudaStreamSynchronize(ctx().stream);
  cudaFree(m->d_state);
  cudaFree(m->d_words);
  cudaFree(m->d_zig);
  cudaFree(m->d_table);
  delete m;
}

// the layer table the device uses (host computation, no GPU needed): x[257], y[257], *r
void vpb_mt_ziggurat_table(double *x, double *y, double *r) { zig_table_host(x, y, r); }

// Generator state in the format of get_mt_rng_state / set_mt_rng_state (mtrand.c:74-124): 4 bytes `next`, then 624
// little-endian words; 4*(624+1) bytes.
void vpb_mt_set_state(vpb_mt_t *m, const void *s2500) {
  if (!m || !s2500) VPB_ERROR("Bad args");
  const unsigned char *s = (const unsigned char *)s2500;
  auto rd = [&](int k) { return (uint32_t)s[4 * k] | ((uint32_t)s[4 * k + 1] << 8) | ((uint32_t)s[4 * k + 2] << 16) | ((uint32_t)s[4 * k + 3] << 24); };
  uint32_t next = rd(0);
  if (next > (uint32_t)kMtN) next = kMtN;
  for (int j = 0; j < kMtN; j++) m->h_state[j] = rd(1 + j);
  mt_upload_state(m);
  // the unread words of the current block, tempered, become the head of the stream buffer (padded in front so that
  // the buffer still starts on a block boundary)
  std::vector<uint32_t> w(kMtN);
  for (int j = 0; j < kMtN; j++) w[j] = mt_temper(m->h_state[j]);
  if (m->cap_words < (size_t)4 * kMtN) {
    if (m->d_words) cudaFree(m->d_words);
    m->cap_words = 64 * kMtN;
    VPB_CUDA(cudaMalloc(&m->d_words, m->cap_words * sizeof(uint32_t)));
  }
  VPB_CUDA(cudaMemcpy(m->d_words, w.data(), kMtN * sizeof(uint32_t), cudaMemcpyHostToDevice));
  m->n_valid = kMtN;
  m->cur = next;
}

void vpb_mt_get_state(vpb_mt_t *m, void *s2500) {
  if (!m || !s2500) VPB_ERROR("Bad args");
  cudaStream_t st = ctx().stream;
  uint32_t state[kMtN], next;
  if (m->cur >= m->n_valid) {      // everything generated so far is consumed: the raw state is the live one
    VPB_CUDA(cudaMemcpyAsync(state, m->d_state, sizeof(state), cudaMemcpyDeviceToHost, st));
    VPB_CUDA(cudaStreamSynchronize(st));
    next = kMtN;
  } else {                          // the cursor's block, un-tempered
    const size_t b0 = (m->cur / kMtN) * kMtN;
    VPB_CUDA(cudaMemcpyAsync(state, m->d_words + b0, sizeof(state), cudaMemcpyDeviceToHost, st));
    VPB_CUDA(cudaStreamSynchronize(st));
    for (int j = 0; j < kMtN; j++) state[j] = mt_untemper(state[j]);
    next = (uint32_t)(m->cur - b0);
  }
  unsigned char *s = (unsigned char *)s2500;
  auto wr = [&](int k, uint32_t v) { s[4 * k] = v & 0xff; s[4 * k + 1] = (v >> 8) & 0xff; s[4 * k + 2] = (v >> 16) & 0xff; s[4 * k + 3] = (v >> 24) & 0xff; };
  wr(0, next);
  for (int j = 0; j < kMtN; j++) wr(1 + j, state[j]);
}

// n raw 32-bit words (mt_urand), for tests of the generator itself
void vpb_mt_words(vpb_mt_t *m, unsigned int *d_out, long n) {
  if (!m || n < 0) VPB_ERROR("Bad args");
  if (n == 0) return;
  mt_ensure(m, (size_t)n);
  VPB_CUDA(cudaMemcpyAsync(d_out, m->d_words + m->cur, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx().stream));
  m->cur += (size_t)n;
}

// n records of the token program prog ('U' = mt_drand, 'N' = mt_drandn) in stream order: d_out[r*len + t].  Advances
// the generator exactly as the n*len host calls would.  Synchronises.
void vpb_mt_draw(vpb_mt_t *m, const char *prog, long n, double *d_out) {
  if (!m || !prog || n < 0 || !d_out) VPB_ERROR("Bad args");
  Program P;
  P.len = (int)strlen(prog);
  if (P.len < 1 || P.len > kMaxProg) VPB_ERROR("token program of %d tokens (1..%d)", P.len, kMaxProg);
  for (int t = 0; t < P.len; t++) {
    if (prog[t] != 'U' && prog[t] != 'N') VPB_ERROR("token '%c' (U = mt_drand, N = mt_drandn)", prog[t]);
    P.tok[t] = prog[t];
  }
  Context &c = ctx();
  cudaStream_t st = c.stream;
  ZigTable Z;
  Z.x = m->d_zig; Z.y = m->d_zig + 257; Z.r = m->zig_r; Z.scale = 1. / 1.8446744073709551616e+19;
  const long kBatch = 1L << 24;                      // records per parse: bounds the scratch (about 110 bytes a token pair)
  long done = 0;
  double extra = 1.03;
  while (done < n) {
    const long nb = n - done < kBatch ? n - done : kBatch;
    const size_t want = (size_t)((double)nb * P.len * 2 * extra) + 16384;
    mt_ensure(m, want);
    const long npairs = (long)((m->n_valid - m->cur) / 2);
    const long nchunks = (npairs + kChunk - 1) / kChunk;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const unsigned int tail_cap = (unsigned int)(nb * P.len / 1000 + 4096);
    const size_t o_jump = al((size_t)npairs), o_exit = o_jump + al((size_t)npairs * 2), o_cnt = o_exit + al((size_t)nchunks * kEntries * 2),
                 o_entry = o_cnt + al((size_t)nchunks * kEntries * 2), o_first = o_entry + al((size_t)nchunks), o_starts = o_first + al((size_t)nchunks * 8),
                 o_out = o_starts + al((size_t)nb * 4), o_tw = o_out + 256, o_tp = o_tw + al((size_t)tail_cap * 8), o_tv = o_tp + al((size_t)tail_cap * 8),
                 o_end = o_tv + al((size_t)tail_cap * 8);
    char *s = (char *)scratch(o_end + 256);
    unsigned char *nlen = (unsigned char *)s, *entry = (unsigned char *)(s + o_entry);
    unsigned short *jump = (unsigned short *)(s + o_jump), *exit_off = (unsigned short *)(s + o_exit), *count = (unsigned short *)(s + o_cnt);
    long *first = (long *)(s + o_first), *outv = (long *)(s + o_out);      // outv: [0] records, [1] error, [2] end position, [3] tails
    unsigned int *starts = (unsigned int *)(s + o_starts), *tail_pair = (unsigned int *)(s + o_tp);
    long *tail_where = (long *)(s + o_tw);
    double *tail_val = (double *)(s + o_tv);
    const uint32_t *w = m->d_words + m->cur;
    VPB_CUDA(cudaMemsetAsync(outv, 0, 4 * sizeof(long), st));
    mt_normal_len_kernel<<<c.sm_count * 8, 256, 0, st>>>(w, npairs, Z, nlen);
    mt_record_len_kernel<<<c.sm_count * 8, 256, 0, st>>>(nlen, npairs, P, jump);
    mt_chunk_walk_kernel<<<(unsigned int)nchunks, kEntries, 0, st>>>(jump, npairs, exit_off, count);
    mt_chunk_scan_kernel<<<1, 256, 0, st>>>(exit_off, count, nchunks, entry, first, outv);
    mt_record_starts_kernel<<<(unsigned int)((nchunks + 127) / 128), 128, 0, st>>>(jump, npairs, nchunks, entry, first, nb, starts, outv + 2);
    count_launch(5);
    long h[4];
    VPB_CUDA(cudaMemcpyAsync(h, outv, sizeof(h), cudaMemcpyDeviceToHost, st));
    VPB_CUDA(cudaStreamSynchronize(st));
    if (h[1]) VPB_ERROR("vpb_mt_draw: a record of more than %d word pairs (a run of ziggurat rejections this long has probability < 1e-50)", kEntries);
    const long got = h[0] < nb ? h[0] : nb;
    if (got == 0) { extra *= 1.5; continue; }        // (tiny n with an unlucky start) more words next time round
    if (got < nb) {
      // fewer complete records than asked for: take what is there, the loop generates more
      VPB_CUDA(cudaMemsetAsync(outv + 2, 0, sizeof(long), st));
      mt_record_starts_kernel<<<(unsigned int)((nchunks + 127) / 128), 128, 0, st>>>(jump, npairs, nchunks, entry, first, got, starts, outv + 2);
      count_launch();
    }
    double *out = d_out + done * P.len;
    mt_eval_kernel<<<c.sm_count * 8, 256, 0, st>>>(w, npairs, Z, P, starts, got, out, (unsigned int *)(outv + 3), tail_where, tail_pair, tail_cap);
    count_launch();
    VPB_CUDA(cudaMemcpyAsync(h, outv, sizeof(h), cudaMemcpyDeviceToHost, st));
    VPB_CUDA(cudaStreamSynchronize(st));
    const unsigned int ntail = (unsigned int)(h[3] & 0xffffffffL);
    if (ntail > tail_cap) VPB_ERROR("vpb_mt_draw: %u tail deviates in %ld records (room for %u)", ntail, got, tail_cap);
    if (ntail) {
      // the tail layer's value with the HOST's log, as the reference computes it (mtrand.c:430)
      std::vector<uint32_t> ww(2 * (size_t)ntail);
      std::vector<long> tw(ntail);
      std::vector<double> tv(ntail);
      VPB_CUDA(cudaMemcpyAsync(ww.data(), tail_pair, 2 * (size_t)ntail * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      VPB_CUDA(cudaMemcpyAsync(tw.data(), tail_where, ntail * sizeof(long), cudaMemcpyDeviceToHost, st));
      VPB_CUDA(cudaStreamSynchronize(st));
      const double R = m->zig_r;
      for (unsigned int k = 0; k < ntail; k++) {
        const double x = R - (1. / R) * log(h_d53_c1(ww[2 * k], ww[2 * k + 1]));
        tv[k] = (tw[k] >> 62) & 1 ? -x : x;
        tw[k] &= ~(1L << 62);
      }
      VPB_CUDA(cudaMemcpyAsync(tail_where, tw.data(), ntail * sizeof(long), cudaMemcpyHostToDevice, st));
      VPB_CUDA(cudaMemcpyAsync(tail_val, tv.data(), ntail * sizeof(double), cudaMemcpyHostToDevice, st));
      mt_patch_kernel<<<(ntail + 255) / 256, 256, 0, st>>>(out, tail_where, tail_val, ntail);
      count_launch();
      VPB_CUDA(cudaStreamSynchronize(st));
    }
    m->cur += 2 * (size_t)h[2];
    done += got;
    if (got < nb) extra *= 1.25;
  }
  VPB_CUDA(cudaGetLastError());
}

// n calls of inject_particle(sp, x, y, z, ux, uy, uz, q, tag, 0, 0) (misc.cxx:16-105) in order, the arguments taken from
// row k of a table of deviates: x = lo*(1-t[col0]) + hi*t[col0] ..., ux = dev0 * t[col3] ...  Particles outside the
// local domain (or on a far wall shared with a neighbour) are skipped as the reference skips them; the others are
// appended in order; row k carries the tag `tag + k*tag_step`.  Returns the new particle count.  d_p is in the domain's
// particle layout.
int vpb_inject_from_draws(vpb_domain_t *dom, vpb_particle_t *d_p, int np, int max_np, const double *d_table, int stride, long n,
                          const int col[6], const double lo[3], const double hi[3], const double dev[3], double q, long tag, long tag_step) {
  if (!dom || !dom->host_grid) VPB_ERROR("Bad grid");
  if (!d_p || !d_table || n < 0 || np < 0 || np > max_np) VPB_ERROR("Bad args");
  if (n == 0) return np;
  if (n > 0x7fffffffL) VPB_ERROR("at most 2^31-1 particles per call");
  Context &c = ctx();
  cudaStream_t st = c.stream;
  const vpb_grid_t *g = dom->host_grid;
  LoadMap M;
  GridBox G;
  for (int k = 0; k < 6; k++) {
    if (col[k] < 0 || col[k] >= stride) VPB_ERROR("column %d outside the table (stride %d)", col[k], stride);
    M.col[k] = col[k];
  }
  for (int k = 0; k < 3; k++) { M.lo[k] = lo[k]; M.hi[k] = hi[k]; M.dev[k] = dev[k]; }
  G.x0 = (double)g->x0; G.y0 = (double)g->y0; G.z0 = (double)g->z0;
  G.x1 = (double)g->x1; G.y1 = (double)g->y1; G.z1 = (double)g->z1;
  G.nx = g->nx; G.ny = g->ny; G.nz = g->nz;
  G.far_shared[0] = g->bc[VPB_BOUNDARY(1, 0, 0)] >= 0;
  G.far_shared[1] = g->bc[VPB_BOUNDARY(0, 1, 0)] >= 0;
  G.far_shared[2] = g->bc[VPB_BOUNDARY(0, 0, 1)] >= 0;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t o_slot = al((size_t)(n + 1) * 4), o_flag = o_slot + al((size_t)(n + 1) * 4), o_scan = o_flag + 256;
  char *s = (char *)scratch(o_scan + scan_scratch_bytes(n + 1));
  int *keep = (int *)s, *slot = (int *)(s + o_slot), *flag = (int *)(s + o_flag);
  const PView pv(d_p, dom->d.p_plane);
  VPB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  VPB_CUDA(cudaMemsetAsync(keep + n, 0, sizeof(int), st));
  inject_table_kernel<0><<<c.sm_count * 8, 256, 0, st>>>(d_table, stride, n, M, G, keep, nullptr, pv, np, max_np, (float)q, tag, tag_step, flag);
  exclusive_scan_i32(keep, slot, (int)(n + 1), s + o_scan, st);
  inject_table_kernel<1><<<c.sm_count * 8, 256, 0, st>>>(d_table, stride, n, M, G, nullptr, slot, pv, np, max_np, (float)q, tag, tag_step, flag);
  count_launch(2 + scan_launches(n + 1));
  int h[2];
  VPB_CUDA(cudaMemcpyAsync(&h[0], slot + n, sizeof(int), cudaMemcpyDeviceToHost, st));
  VPB_CUDA(cudaMemcpyAsync(&h[1], flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  VPB_CUDA(cudaStreamSynchronize(st));
  if (h[1]) VPB_ERROR("No room to inject particle (%d of %d do not fit max_np=%d)", h[1], h[0], max_np);
  VPB_CUDA(cudaGetLastError());
  return np + h[0];
}

// The load loop of a thermal deck (SURVEY.md 8d, configs[0]/[3] recipe; oracle/decks/thermal_c1.cxx): per iteration one
// position from three uniform_rand(lo, hi) and two co-located particles, each with three maxwellian_rand.  A deck
// writes the three deviates as ARGUMENTS of inject_particle, whose evaluation order is the compiler's: g++ on x86-64
// goes right to left, so the first deviate drawn is uz (args_right_to_left = 1); 0 = ux first.
// np[2] (in/out): particle counts of the two arrays.  Iteration k tags both its particles tag0 + k*tag_step (decks pass 0 or
// the loop counter).  Returns the number of iterations.
long vpb_load_pairs_mt(vpb_domain_t *dom, vpb_mt_t *rng, long n, const double lo[3], const double hi[3], double vth_a, double vth_b, double q_a,
                       double q_b, vpb_particle_t *d_a, int max_a, vpb_particle_t *d_b, int max_b, int np[2], int args_right_to_left,
                       long tag0, long tag_step) {
  if (!dom || !rng || n < 0 || !np) VPB_ERROR("Bad args");
  const long kBatch = 1L << 23;
  const int o = args_right_to_left ? 1 : 0;
  const int col_a[6] = {0, 1, 2, o ? 5 : 3, 4, o ? 3 : 5}, col_b[6] = {0, 1, 2, o ? 8 : 6, 7, o ? 6 : 8};
  const double dev_a[3] = {vth_a, vth_a, vth_a}, dev_b[3] = {vth_b, vth_b, vth_b};
  const long nb_max = n < kBatch ? n : kBatch;
  if ((size_t)nb_max * 9 > rng->table_cap) {
    VPB_CUDA(cudaStreamSynchronize(ctx().stream));
    cudaFree(rng->d_table);
    rng->table_cap = (size_t)nb_max * 9;
    VPB_CUDA(cudaMalloc(&rng->d_table, rng->table_cap * sizeof(double)));
  }
  double *tab = rng->d_table;
  for (long done = 0; done < n;) {
    const long nb = n - done < kBatch ? n - done : kBatch;
    vpb_mt_draw(rng, "UUUNNNNNN", nb, tab);
    np[0] = vpb_inject_from_draws(dom, d_a, np[0], max_a, tab, 9, nb, col_a, lo, hi, dev_a, q_a, tag0 + done * tag_step, tag_step);
    np[1] = vpb_inject_from_draws(dom, d_b, np[1], max_b, tab, 9, nb, col_b, lo, hi, dev_b, q_b, tag0 + done * tag_step, tag_step);
    done += nb;
  }
  return n;
}

}  // extern "C"
