// vpb_boundary.cu -- K4: boundary_p, the particle-migration step
// (src/species_advance/standard/boundary_p.c:9-71,77-505) on device arrays, with
// NCCL send/recv over NVLink in place of the reference's MPI ports.
//
// One call = one round of the reference's boundary_p for all species of a rank:
//   1. every mover left by advance_p (or by a previous round's injection) sits on
//      a domain face.  classify: absorbing face -> accumulate_rhob and drop; face
//      shared with another rank -> becomes a 48-byte particle_injector_t (the
//      reference's wire record, normal coordinate sign-flipped, voxel index
//      rebased to the receiver) in that face's send buffer; anything else ->
//      absorbed with a warning, as the reference does (boundary_p.c:312-315).
//      Send order per face is the reference's: species in list order, movers in
//      decreasing particle index (boundary_p.c:194-201).
//   2. the removed particles' slots are back-filled from the tail of the array.
//   3. per-face counts, then payloads, are exchanged (one NCCL group each).
//   4. received particles are appended in the reference's order (faces as the
//      reference's receive loop visits them, each buffer back to front,
//      boundary_p.c:457-497), finish their move with move_p, and those that hit
//      yet another face become the movers of the next round.
// Two forms of steps 3-4.  The reference's own protocol (counts first, then payloads of exactly that size; the
// host reads the counts to size the messages and the injection launches: three stream synchronisations per
// round) serves the reference-named boundary_p(), whose arrays may have to GROW on the host before the arrivals
// are appended.  The device-resident driver's rounds (vpb_boundary_p_round) use ONE fixed-capacity message per
// face instead: record 0 is a header carrying the count, the records follow, the capacity is what both sides
// derived from the counts of earlier rounds (both sides of a face know both directions' counts, so they always
// agree); everything downstream of the exchange takes its sizes from the headers ON THE DEVICE, and the host
// reads all counters back once at the end of the round.  A face whose count exceeds the capacity sends the rest
// in a second, exactly sized message after that read-back (both sides see the overflow in the header).
// Difference from the reference, documented in DESIGN.md: the serial back-fill
// loop (r[0] = p0[--np]) makes the final ORDER of the surviving particles depend
// on the serial visiting order; here the k-th highest hole takes the k-th highest
// tail survivor.  The multiset of particles per species is identical.
#include <vector>
#include "vpb_comm.cuh"
#include "vpb_move_p.cuh"
#include "vpb_pview.cuh"

namespace vpb {

constexpr int kBins = 8;          // 0..6 real bins, 7 = "none"
constexpr int kRankThreads = 256;

// Stable multi-bin ranking: rank[k] = base[code] + number of earlier elements (in visiting order) with the same
// code.  reverse!=0 visits k = n-1..0.  base[] (device, 7 ints) is read at entry and updated at exit, so
// consecutive calls continue the numbering (species after species).  Three small launches: every block counts the
// bins of its segment of the visiting order, one thread per bin turns the counts into segment bases, every block
// ranks its segment.  (One block walking a million movers took milliseconds per call at 256^3 cells per GPU.)
constexpr int kRankSeg = 4096;     // elements per block

__device__ __forceinline__ int rank_code_at(const unsigned char *__restrict__ code, int n, int reverse, int r) {
  return r < n ? code[reverse ? n - 1 - r : r] : 7;
}

// n_dev != nullptr: the element count lives on the device (fused migration rounds); n is then only the launch bound
__global__ void __launch_bounds__(kRankThreads) rank_bins_count_kernel(const unsigned char *__restrict__ code, int n, int reverse,
                                                                       int *__restrict__ tab, const int *__restrict__ n_dev) {
  __shared__ int cnt[kBins];
  if (n_dev) n = min(n, *n_dev);
  const int tid = threadIdx.x;
  if (tid < kBins) cnt[tid] = 0;
  __syncthreads();
  const int r0 = blockIdx.x * kRankSeg, r1 = min(n, r0 + kRankSeg);
  int mine[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int r = r0 + tid; r < r1; r += kRankThreads) {
    const int c = rank_code_at(code, n, reverse, r);
#pragma unroll
    for (int b = 0; b < 7; b++) mine[b] += (c == b);
  }
#pragma unroll
  for (int b = 0; b < 7; b++) {
    int v = mine[b];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0 && v) atomicAdd(&cnt[b], v);
  }
  __syncthreads();
  if (tid < kBins) tab[blockIdx.x * kBins + tid] = tid < 7 ? cnt[tid] : 0;
}

__global__ void rank_bins_scan_kernel(int *__restrict__ tab, int nblocks, int *__restrict__ base_io) {
  const int b = threadIdx.x;
  if (b >= 7) return;
  int run = base_io[b];
  for (int k = 0; k < nblocks; k++) {
    const int t = tab[k * kBins + b];
    tab[k * kBins + b] = run;
    run += t;
  }
  base_io[b] = run;
}

__global__ void __launch_bounds__(kRankThreads) rank_bins_rank_kernel(const unsigned char *__restrict__ code, int n, int reverse,
                                                                      int *__restrict__ rank, const int *__restrict__ tab,
                                                                      const int *__restrict__ n_dev) {
  __shared__ int base[kBins];
  if (n_dev) n = min(n, *n_dev);
  __shared__ int wcnt[kRankThreads / 32][kBins];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid < kBins) base[tid] = tab[blockIdx.x * kBins + tid];
  __syncthreads();
  const int seg0 = blockIdx.x * kRankSeg, seg1 = min(n, seg0 + kRankSeg);
  for (int r0 = seg0; r0 < seg1; r0 += kRankThreads) {
    const int r = r0 + tid;
    const int k = reverse ? n - 1 - r : r;
    const int c = r < seg1 ? rank_code_at(code, n, reverse, r) : 7;
    int mine = 0;
#pragma unroll
    for (int b = 0; b < 7; b++) {
      const unsigned m = __ballot_sync(0xffffffffu, c == b);
      if (lane == 0) wcnt[w][b] = __popc(m);
      if (c == b) mine = __popc(m & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (c < 7) {
      int before = 0;
      for (int j = 0; j < w; j++) before += wcnt[j][c];
      rank[k] = base[c] + before + mine;
    }
    __syncthreads();
    if (tid < 7) {
      int t = 0;
      for (int j = 0; j < kRankThreads / 32; j++) t += wcnt[j][tid];
      base[tid] += t;
    }
    __syncthreads();
  }
}

static int *g_rank_tab = nullptr;
static int g_rank_tab_blocks = 0;

static void rank_bins(const unsigned char *code, int n, int reverse, int *rank, int *base_io, cudaStream_t st,
                      const int *n_dev = nullptr) {
  if (n <= 0) return;
  const int nblocks = (n + kRankSeg - 1) / kRankSeg;
  if (nblocks > g_rank_tab_blocks) {
    if (g_rank_tab) { VPB_CUDA(cudaStreamSynchronize(st)); cudaFree(g_rank_tab); }
    g_rank_tab_blocks = nblocks + nblocks / 2 + 64;
    VPB_CUDA(cudaMalloc(&g_rank_tab, (size_t)g_rank_tab_blocks * kBins * sizeof(int)));
  }
  rank_bins_count_kernel<<<nblocks, kRankThreads, 0, st>>>(code, n, reverse, g_rank_tab, n_dev);
  rank_bins_scan_kernel<<<1, 32, 0, st>>>(g_rank_tab, nblocks, base_io);
  rank_bins_rank_kernel<<<nblocks, kRankThreads, 0, st>>>(code, n, reverse, rank, g_rank_tab, n_dev);
  count_launch(3);
}

// boundary_p.c:9-71 on the device: trilinear deposit of a removed particle's charge to rhob,
// surface nodes weighted twice
__device__ void accumulate_rhob_dev(vpb_field_t *__restrict__ f, float dx, float dy, float dz, int vi, float q, const DomainDev &g) {
  float w0 = (float)(0.125 * q * g.rdx * g.rdy * g.rdz), w1, w2, w3, w4, w5, w6, w7, t;
  t = dx; t *= w0; w1 = w0 + t; w0 -= t;
  t = dy; w3 = 1 + t; w2 = w0 * w3; w3 *= w1; t = 1 - t; w0 *= t; w1 *= t;
  t = dz; w7 = 1 + t; w4 = w0 * w7; w5 = w1 * w7; w6 = w2 * w7; w7 *= w3;
  t = 1 - t; w0 *= t; w1 *= t; w2 *= t; w3 *= t;
  int i = vi, j = i / g.sx;
  i -= j * g.sx;
  const int k = j / g.sy;
  j -= k * g.sy;
  if (i == 1) { w0 += w0; w2 += w2; w4 += w4; w6 += w6; }
  if (i == g.nx) { w1 += w1; w3 += w3; w5 += w5; w7 += w7; }
  if (j == 1) { w0 += w0; w1 += w1; w4 += w4; w5 += w5; }
  if (j == g.ny) { w2 += w2; w3 += w3; w6 += w6; w7 += w7; }
  if (k == 1) { w0 += w0; w1 += w1; w2 += w2; w3 += w3; }
  if (k == g.nz) { w4 += w4; w5 += w5; w6 += w6; w7 += w7; }
  float *rhob = &FCOMP(f, g, vi, 11);                                      // rhob = component 11
  const size_t X = 4 * (size_t)g.fqv, Y = X * (size_t)g.sx, Z = X * (size_t)g.sxy;
  red_add(rhob, w0); red_add(rhob + X, w1); red_add(rhob + Y, w2); red_add(rhob + X + Y, w3);
  red_add(rhob + Z, w4); red_add(rhob + Z + X, w5); red_add(rhob + Z + Y, w6); red_add(rhob + Z + Y + X, w7);
}

struct FaceInfo {
  int64_t rbase[6];    // range[] of the rank across each face (0 if not shared remotely)
  int64_t rangem;      // range[nproc]
  int nb;              // grid->nb: custom particle-boundary handlers of the deck (neighbor code -3-k, k < nb)
  int handled;         // != 0: the caller has already run those handlers on the host for this round's movers
};

// boundary_p.c:203-316: which face did mover k end on, and what happens there.
// code: 0..5 = send through that face, 6 = removed locally (absorbed).
__global__ void __launch_bounds__(256) classify_kernel(const PView p, const vpb_particle_mover_t *__restrict__ pm,
                                                       int nm, unsigned char *__restrict__ code, vpb_field_t *__restrict__ f,
                                                       const DomainDev g, const FaceInfo fi, int *__restrict__ n_unknown) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nm) return;
  const int pi = pm[k].i;
  // the hole filling below (and the reference's, boundary_p.c:168-176,243-247) needs the movers in ascending particle
  // order; advance_p and inject_particle produce them that way, a list built by other host code might not
  if (k + 1 < nm && pm[k + 1].i <= pi) atomicAdd(n_unknown + 1, 1);
  const float4 r0 = p.pos(pi);
  const float4 r1 = p.mom(pi);
  const int vi = __float_as_int(r0.w);
  const float pos[3] = {r0.x, r0.y, r0.z}, u[3] = {r1.x, r1.y, r1.z};
  int c = -1;
  bool unknown = true, handler = false;
#pragma unroll
  for (int face = 0; face < 6 && c < 0; face++) {
    const int ax = face % 3;
    const bool hit = face < 3 ? (pos[ax] == -1.f && u[ax] < 0) : (pos[ax] == 1.f && u[ax] > 0);
    if (!hit) continue;
    const int64_t nn = g.nbr64[6 * (size_t)vi + face];
    if (nn == vpb_absorb_particles) { c = 6; unknown = false; }
    else if ((nn >= 0 && nn < g.rangel) || (nn > g.rangeh && nn <= fi.rangem)) { c = face; unknown = false; }
    else if (nn <= -3 && -nn - 3 < (int64_t)fi.nb) {
      // boundary_p.c:271-277: the deck's handler -nn-3, a HOST callback.  The reference-named boundary_p() has run it
      // for this mover before coming here (vpb_dropin.cu); what is left is to destroy the particle ("Particle is
      // destroyed after it is handled"), without the rhob deposit of an absorption.  Anyone else must not get here.
      c = 6; unknown = false; handler = true;
      if (!fi.handled) atomicAdd(n_unknown + 5, 1);
    }
  }
  if (c < 0) c = 6;
  if (c == 6 && !handler) {
    accumulate_rhob_dev(f, r0.x, r0.y, r0.z, vi, r1.w, g);
    if (unknown) atomicAdd(n_unknown, 1);
  }
  code[k] = (unsigned char)c;
}

// Where the records of a face go.  Reference protocol: slot = rank in the face's buffer.  Fused protocol: the buffer
// starts with a header record and holds at most cap[face] injectors; the rest of an over-full face goes to a common
// overflow list (faces one after the other, each in rank order) that leaves in a second message.
struct PackPlan {
  int first;                 // 0, or 1 when record 0 of a buffer is the header
  int cap[6];                // injectors a buffer may hold
  float4 *overflow;          // nullptr: no capacity limit
  const int *count;          // final per-face counts (device), for the overflow offsets
};

// write the injector records of one species (boundary_p.c:250-263)
__global__ void __launch_bounds__(256) pack_injectors_kernel(const PView p, const vpb_particle_mover_t *__restrict__ pm,
                                                             int nm, const unsigned char *__restrict__ code, const int *__restrict__ rank,
                                                             int sp_id, const DomainDev g, const FaceInfo fi, float4 *const *__restrict__ sendbuf,
                                                             const PackPlan plan) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nm) return;
  const int face = code[k];
  if (face >= 6) return;
  const float4 m = reinterpret_cast<const float4 *>(pm)[k];
  const int pi = __float_as_int(m.w);
  float4 r0 = p.pos(pi);
  const float4 r1 = p.mom(pi);
  const int ax = face % 3;
  if (ax == 0) r0.x = -r0.x; else if (ax == 1) r0.y = -r0.y; else r0.z = -r0.z;
  const int64_t nn = g.nbr64[6 * (size_t)__float_as_int(r0.w) + face];
  r0.w = __int_as_float((int)(nn - fi.rbase[face]));
  const int r = rank[k];
  float4 *o;
  if (plan.overflow == nullptr || r < plan.cap[face]) {
    o = sendbuf[face] + 3 * (size_t)(plan.first + r);
  } else {
    long base = 0;
    for (int f = 0; f < face; f++) base += max(0, plan.count[f] - plan.cap[f]);
    o = plan.overflow + 3 * (size_t)(base + (r - plan.cap[face]));
  }
  o[0] = r0;
  o[1] = r1;
  o[2] = make_float4(m.x, m.y, m.z, __int_as_float(sp_id));
}

constexpr int kMigMagic = 0x76706221;   // header tag of a fused migration message

// record 0 of every fused message: {count, capacity, round, magic}
__global__ void write_headers_kernel(float4 *const *__restrict__ sendbuf, const int *__restrict__ count, const PackPlan plan, int round,
                                     unsigned remote_mask) {
  const int f = threadIdx.x;
  if (f >= 6 || !((remote_mask >> f) & 1u)) return;
  sendbuf[f][0] = make_float4(__int_as_float(count[f]), __int_as_float(plan.cap[f]), __int_as_float(round), __int_as_float(kMigMagic));
}

// removal: tailflag[j]=1 if particle np'+j is a mover; *nh = number of movers below np'
__global__ void __launch_bounds__(256) mark_tail_kernel(const vpb_particle_mover_t *__restrict__ pm, int nm, int np_new,
                                                        unsigned char *__restrict__ tail_code, int *__restrict__ nh) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nm) return;
  const int i = pm[k].i;
  if (i >= np_new) tail_code[i - np_new] = 7;   // 7 = not ranked
  const int inext = (k + 1 < nm) ? pm[k + 1].i : 0x7fffffff;
  if (i < np_new && inext >= np_new) *nh = k + 1;
}

// survivor j of the tail (rank r among survivors, counted from the top) fills hole pm[nh-1-r]
__global__ void __launch_bounds__(256) backfill_kernel(const PView p, const vpb_particle_mover_t *__restrict__ pm,
                                                       int nm, int np_new, const unsigned char *__restrict__ tail_code,
                                                       const int *__restrict__ rank, const int *__restrict__ nh) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;   // 3 float4 per particle
  const int j = (int)(t / 3), piece = (int)(t - 3L * j);
  if (j >= nm || tail_code[j] != 0) return;
  const int hole = pm[*nh - 1 - rank[j]].i;
  p.set_quad(hole, piece, p.quad(np_new + j, piece));
}

// received buffers -> one list in the reference's injection order (each buffer back to front)
__global__ void __launch_bounds__(256) gather_injectors_kernel(float4 *__restrict__ list, const float4 *__restrict__ buf, int n, int off) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (int)(t / 3), piece = (int)(t - 3L * j);
  if (j >= n) return;
  list[3 * (size_t)(off + j) + piece] = buf[3 * (size_t)(n - 1 - j) + piece];
}

// The same for fused messages: the counts are in the headers.  Faces are visited in the reference's receive order
// (3,4,5,0,1,2), each buffer back to front.  Thread 0 also leaves the counts where the host will read them:
// dc[8+f] = count announced by face f's header, dc[14] = injectors in the list, dc[37] = malformed headers.
struct RecvPlan {
  const float4 *buf[6];      // nullptr: face not shared with another rank
  int cap[6];
  int round;
};

__global__ void __launch_bounds__(256) gather_fused_kernel(float4 *__restrict__ list, const RecvPlan R, int *__restrict__ dc) {
  const int order[6] = {3, 4, 5, 0, 1, 2};
  int n[6], total = 0, bad = 0;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const int f = order[k];
    n[k] = 0;
    if (R.buf[f]) {
      const float4 h = R.buf[f][0];
      const int cnt = __float_as_int(h.x);
      if (__float_as_int(h.w) != kMigMagic || __float_as_int(h.y) != R.cap[f] || __float_as_int(h.z) != R.round || cnt < 0) bad++;
      else n[k] = min(cnt, R.cap[f]);
      if (blockIdx.x == 0 && threadIdx.x == 0) dc[8 + f] = cnt;
    }
    total += n[k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { dc[14] = total; dc[37] = bad; }
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (int)(t / 3), piece = (int)(t - 3L * j);
  if (j >= total) return;
  int k = 0, off = 0;
  while (j >= off + n[k]) { off += n[k]; k++; }
  const int f = order[k];
  list[3 * (size_t)j + piece] = R.buf[f][3 * (size_t)(1 + (n[k] - 1 - (j - off))) + piece];
}

struct SpeciesTable {
  int n;
  int id[7];
  vpb_particle_t *p[7];
  vpb_particle_mover_t *pm[7];
  int np[7];
  int max_np[7], max_nm[7];   // capacities, checked on the device by the fused rounds
  long plane;       // particle layout of the domain (vpb_pview.cuh)
};

__global__ void __launch_bounds__(256) species_code_kernel(const float4 *__restrict__ list, int n, const SpeciesTable T,
                                                           unsigned char *__restrict__ code, const int *__restrict__ n_dev) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, *n_dev);
  if (q >= n) return;
  const int id = __float_as_int(list[3 * (size_t)q + 2].w);
  int c = 7;
  for (int s = 0; s < T.n; s++) if (T.id[s] == id) c = s;
  code[q] = (unsigned char)c;
}

// boundary_p.c:478-496: append, then finish the move.  Leaves the remaining displacement and the
// particle's new index in list[q] and marks unresolved movers in code2.
__global__ void __launch_bounds__(128) inject_kernel(float4 *__restrict__ list, int n, const unsigned char *__restrict__ code,
                                                     const int *__restrict__ rank, const SpeciesTable T, float *__restrict__ a0,
                                                     const int32_t *__restrict__ nbr, unsigned char *__restrict__ code2,
                                                     const int *__restrict__ n_dev, int *__restrict__ overflow) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, *n_dev);
  if (q >= n) return;
  const int s = code[q];
  if (s >= 7) { code2[q] = 7; return; }
  // fused rounds: the host has not seen the arrival counts yet, so the capacity test is here (the host raises the
  // error after its read-back); nothing is written past the end of an array
  if (overflow && T.np[s] + rank[q] >= T.max_np[s]) { atomicAdd(overflow, 1); code2[q] = 7; return; }
  const float4 a = list[3 * (size_t)q], b = list[3 * (size_t)q + 1], c = list[3 * (size_t)q + 2];
  Mover m;
  m.dx = a.x; m.dy = a.y; m.dz = a.z; m.i = __float_as_int(a.w);
  m.ux = b.x; m.uy = b.y; m.uz = b.z; m.q = b.w;
  m.dispx = c.x; m.dispy = c.y; m.dispz = c.z;
  const int pos = T.np[s] + rank[q];
  const int unresolved = move_p_dev(m, a0, nbr);
  const PView pv(T.p[s], T.plane);                          // tags are left as they were (boundary_p.c:488-491)
  pv.set_pos(pos, make_float4(m.dx, m.dy, m.dz, __int_as_float(m.i)));
  pv.set_mom(pos, make_float4(m.ux, m.uy, m.uz, m.q));
  list[3 * (size_t)q + 2] = make_float4(m.dispx, m.dispy, m.dispz, __int_as_float(pos));
  code2[q] = unresolved ? (unsigned char)s : (unsigned char)7;
}

__global__ void __launch_bounds__(256) compact_movers_kernel(const float4 *__restrict__ list, int n, const unsigned char *__restrict__ code2,
                                                             const int *__restrict__ rank2, const SpeciesTable T,
                                                             const int *__restrict__ n_dev, int *__restrict__ overflow) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, *n_dev);
  if (q >= n) return;
  const int s = code2[q];
  if (s >= 7) return;
  if (overflow && rank2[q] >= T.max_nm[s]) { atomicAdd(overflow, 1); return; }
  reinterpret_cast<float4 *>(T.pm[s])[rank2[q]] = list[3 * (size_t)q + 2];
}

// single-mover / single-particle entry points (host code of the reference calls these one at a time:
// inject_particle, misc.cxx:102; custom boundary handlers)
__global__ void move_p_one_kernel(vpb_particle_t *p, vpb_particle_mover_t *pm, float *a0, const int32_t *nbr, int *result) {
  Mover s;
  const int k = pm->i;
  const float4 a = reinterpret_cast<const float4 *>(p + k)[0], b = reinterpret_cast<const float4 *>(p + k)[1];
  s.dx = a.x; s.dy = a.y; s.dz = a.z; s.i = __float_as_int(a.w);
  s.ux = b.x; s.uy = b.y; s.uz = b.z; s.q = b.w;
  s.dispx = pm->dispx; s.dispy = pm->dispy; s.dispz = pm->dispz;
  *result = move_p_dev(s, a0, nbr);
  reinterpret_cast<float4 *>(p + k)[0] = make_float4(s.dx, s.dy, s.dz, __int_as_float(s.i));
  reinterpret_cast<float4 *>(p + k)[1] = make_float4(s.ux, s.uy, s.uz, s.q);
  pm->dispx = s.dispx; pm->dispy = s.dispy; pm->dispz = s.dispz;
}

__global__ void accumulate_rhob_one_kernel(vpb_field_t *f, const vpb_particle_t *p, const DomainDev g) {
  accumulate_rhob_dev(f, p->dx, p->dy, p->dz, p->i, p->q, g);
}

struct BoundaryBuffers {
  float4 *send[6] = {}, *recv[6] = {};
  size_t cap[6] = {};          // records (injectors + the fused header)
  float4 *xrecv = nullptr;     // second messages of over-full faces (fused protocol)
  size_t xcap = 0;
  float4 *local = nullptr;     // injectors made by the deck's custom handlers on the host
  size_t lcap = 0;
  float4 **d_send_table = nullptr;
  // [0..6] send bins (faces 0..5, 6 = absorbed), [8..13] counts received, [14] injectors in the arrival list,
  // [16..22] arrivals per species, [24..30] new movers per species, [32] nh, [33] unknown interactions,
  // [34] mover-order violations, [35] arrivals beyond max_np, [36] new movers beyond max_nm, [37] bad headers,
  // [38] movers on a custom-handler face whose handler nobody ran, [40..47] tail bins
  int *d_counts = nullptr;
};
static BoundaryBuffers g_bb;

static void ensure_cap(int face, size_t n) {
  if (g_bb.cap[face] >= n) return;
  cudaStream_t st = ctx().stream;
  VPB_CUDA(cudaStreamSynchronize(st));
  if (g_bb.send[face]) { cudaFree(g_bb.send[face]); cudaFree(g_bb.recv[face]); }
  const size_t want = n + n / 4 + 1024;
  VPB_CUDA(cudaMalloc(&g_bb.send[face], want * 48));
  VPB_CUDA(cudaMalloc(&g_bb.recv[face], want * 48));
  g_bb.cap[face] = want;
  if (!g_bb.d_send_table) VPB_CUDA(cudaMalloc(&g_bb.d_send_table, 6 * sizeof(float4 *)));
  VPB_CUDA(cudaMemcpy(g_bb.d_send_table, g_bb.send, 6 * sizeof(float4 *), cudaMemcpyHostToDevice));
}

static inline int blocks(long n, int tb) { return (int)((n + tb - 1) / tb); }
static inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

static const int kFaceBound[6] = {VPB_BOUNDARY(-1, 0, 0), VPB_BOUNDARY(0, -1, 0), VPB_BOUNDARY(0, 0, -1),
                                  VPB_BOUNDARY(1, 0, 0),  VPB_BOUNDARY(0, 1, 0),  VPB_BOUNDARY(0, 0, 1)};
static const int kRecvOrder[6] = {3, 4, 5, 0, 1, 2};   // the reference's receive loop: what arrived from +x first

// Injectors the deck's custom boundary handlers made on the host for this round (boundary_p.c's cmlist): set by the
// reference-named boundary_p() around its call, injected after the received buffers (boundary_p.c:457-461, face 6)
static const vpb_particle_injector_t *g_local_inj = nullptr;
static int g_local_inj_n = 0;
static bool g_local_inj_set = false;

struct Faces {
  bool remote[6], any_remote = false;
  int peer[6];
  unsigned mask = 0;
  FaceInfo fi;
};

static Faces faces_of(const vpb_domain_t *dom) {
  const DomainDev &g = dom->d;
  Faces F;
  for (int f = 0; f < 6; f++) {
    const int b = g.bc[kFaceBound[f]];
    F.remote[f] = b >= 0 && b < g.nproc && b != g.rank;   // SHARED_REMOTELY (boundary_p.c:103-104)
    F.peer[f] = F.remote[f] ? b : -1;
    F.fi.rbase[f] = F.remote[f] ? dom->range[b] : 0;
    F.any_remote |= F.remote[f];
    if (F.remote[f]) F.mask |= 1u << f;
  }
  F.fi.rangem = dom->range[g.nproc];
  F.fi.nb = dom->host_grid ? dom->host_grid->nb : 0;
  F.fi.handled = g_local_inj_set ? 1 : 0;
  return F;
}

struct MoverScratch {
  unsigned char *code, *tcode;
  int *rank, *trank;
  float4 *overflow;
};

// classify every species' movers and rank them per face (dc[0..6] = final bin counts)
static void classify_movers(vpb_domain_t *dom, vpb_species_state_t *sp, int n_sp, vpb_field_t *d_f, const Faces &F, const MoverScratch &M,
                            int *dc, cudaStream_t st) {
  const DomainDev &g = dom->d;
  long off = 0;
  for (int s = 0; s < n_sp; s++) {
    const int nm = sp[s].nm;
    if (nm) {
      classify_kernel<<<blocks(nm, 256), 256, 0, st>>>(PView(sp[s].p, g.p_plane), sp[s].pm, nm, M.code + off, d_f, g, F.fi, dc + 33);
      rank_bins(M.code + off, nm, 1, M.rank + off, dc, st);
      count_launch(1);
    }
    off += nm;
  }
}

// injector records into the send buffers, then the removal of all movers' particles: holes below np' are filled from
// the tail.  np and nm of every species are updated.
static void pack_and_remove(vpb_domain_t *dom, vpb_species_state_t *sp, int n_sp, const Faces &F, const MoverScratch &M, const PackPlan &plan,
                            bool sends, int *dc, cudaStream_t st) {
  const DomainDev &g = dom->d;
  long off = 0;
  for (int s = 0; s < n_sp; s++) {
    const int nm = sp[s].nm;
    if (nm) {
      if (sends) {
        pack_injectors_kernel<<<blocks(nm, 256), 256, 0, st>>>(PView(sp[s].p, g.p_plane), sp[s].pm, nm, M.code + off, M.rank + off, sp[s].id, g,
                                                                F.fi, g_bb.d_send_table, plan);
        count_launch();
      }
      const int np_new = sp[s].np - nm;
      if (np_new < 0) VPB_ERROR("more movers than particles");
      VPB_CUDA(cudaMemsetAsync(M.tcode, 0, (size_t)nm, st));
      VPB_CUDA(cudaMemsetAsync(dc + 32, 0, sizeof(int), st));
      VPB_CUDA(cudaMemsetAsync(dc + 40, 0, 8 * sizeof(int), st));
      mark_tail_kernel<<<blocks(nm, 256), 256, 0, st>>>(sp[s].pm, nm, np_new, M.tcode, dc + 32);
      rank_bins(M.tcode, nm, 1, M.trank, dc + 40, st);
      backfill_kernel<<<blocks(3L * nm, 256), 256, 0, st>>>(PView(sp[s].p, g.p_plane), sp[s].pm, nm, np_new, M.tcode, M.trank, dc + 32);
      count_launch(2);
      sp[s].np = np_new;
    }
    off += nm;
    sp[s].nm = 0;
  }
}

static void check_mover_flags(const int *h, int rank) {
  if (h[38])
    VPB_ERROR("boundary_p: %d movers ended on cell faces bound to the deck's custom particle-boundary handlers.  Those are host "
              "callbacks (boundary_p.c:271-277): only the reference-named boundary_p(), which runs them on the host first, "
              "can process such a grid", h[38]);
  if (h[33]) VPB_WARNING("Unknown boundary interaction ... using absorption (%d particles, rank=%d)", h[33], rank);
  if (h[34])
    VPB_ERROR("boundary_p: a mover list is not in ascending particle order (%d inversions); removing its particles would "
              "overwrite live ones (boundary_p.c:168-176 relies on the same order)", h[34]);
}

static SpeciesTable species_table(const vpb_domain_t *dom, const vpb_species_state_t *sp, int n_sp) {
  SpeciesTable T;
  T.n = n_sp;
  T.plane = dom->d.p_plane;
  for (int s = 0; s < 7; s++) { T.id[s] = -1; T.p[s] = nullptr; T.pm[s] = nullptr; T.np[s] = T.max_np[s] = T.max_nm[s] = 0; }
  for (int s = 0; s < n_sp; s++) {
    T.id[s] = sp[s].id; T.p[s] = sp[s].p; T.pm[s] = sp[s].pm; T.np[s] = sp[s].np; T.max_np[s] = sp[s].max_np; T.max_nm[s] = sp[s].max_nm;
  }
  return T;
}

static vpb_grow_hook_t g_grow_hook = nullptr;
static void *g_grow_user = nullptr;

// Injection with counts the host knows (boundary_p.c:388-497): nr[f] injectors in rbuf[f].  Two read-backs: arrivals
// per species (the arrays may have to grow first), new movers per species.
static void inject_known(vpb_domain_t *dom, vpb_species_state_t *sp, int n_sp, vpb_accumulator_t *d_a, const float4 *const rbuf[6],
                         const int nr[6], bool append_movers, const float4 *local = nullptr, int n_local = 0) {
  Context &c = ctx();
  cudaStream_t st = c.stream;
  const DomainDev &g = dom->d;
  int *dc = g_bb.d_counts;
  long n_in = n_local;
  for (int f = 0; f < 6; f++) n_in += nr[f];
  if (n_in == 0) { VPB_CUDA(cudaStreamSynchronize(st)); return; }
  if (!d_a) VPB_ERROR("Bad accumulator");
  const size_t o_list = 0, o_c1 = al256((size_t)n_in * 48), o_r1 = o_c1 + al256((size_t)n_in), o_c2 = o_r1 + al256((size_t)n_in * 4),
               o_r2 = o_c2 + al256((size_t)n_in), o_fin = o_r2 + al256((size_t)n_in * 4);
  char *s2 = (char *)scratch(o_fin + 256);
  float4 *list = (float4 *)(s2 + o_list);
  unsigned char *c1 = (unsigned char *)(s2 + o_c1), *c2 = (unsigned char *)(s2 + o_c2);
  int *r1 = (int *)(s2 + o_r1), *r2 = (int *)(s2 + o_r2);
  int off = 0;
  for (int k = 0; k < 6; k++) {
    const int f = kRecvOrder[k];
    if (!nr[f]) continue;
    gather_injectors_kernel<<<blocks(3L * nr[f], 256), 256, 0, st>>>(list, rbuf[f], nr[f], off);
    count_launch();
    off += nr[f];
  }
  if (n_local) {   // the handlers' own injectors come last (face 6 of boundary_p.c:457-461), back to front like the others
    gather_injectors_kernel<<<blocks(3L * n_local, 256), 256, 0, st>>>(list, local, n_local, off);
    count_launch();
    off += n_local;
  }
  SpeciesTable T = species_table(dom, sp, n_sp);
  const int n = (int)n_in;
  VPB_CUDA(cudaMemsetAsync(dc + 16, 0, 8 * sizeof(int), st));
  if (append_movers) {
    // a second batch of the same round (overflow of a fused message): new movers go behind those of the first batch
    for (int s = 0; s < n_sp; s++) c.h_pinned_i[48 + s] = sp[s].nm;
    VPB_CUDA(cudaMemcpyAsync(dc + 24, c.h_pinned_i + 48, 8 * sizeof(int), cudaMemcpyHostToDevice, st));
  } else {
    VPB_CUDA(cudaMemsetAsync(dc + 24, 0, 8 * sizeof(int), st));
  }
  species_code_kernel<<<blocks(n, 256), 256, 0, st>>>(list, n, T, c1, nullptr);
  rank_bins(c1, n, 0, r1, dc + 16, st);
  VPB_CUDA(cudaMemcpyAsync(c.h_pinned_i, dc + 16, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
  VPB_CUDA(cudaStreamSynchronize(st));
  int cnt[7];
  for (int s = 0; s < n_sp; s++) {
    cnt[s] = c.h_pinned_i[s];
    const int need_nm = (append_movers ? sp[s].nm : 0) + cnt[s];
    if (sp[s].np + cnt[s] > sp[s].max_np || need_nm > sp[s].max_nm) {
      // boundary_p.c:416-447 grows the arrays here; their owner does it through the hook
      const bool grown = g_grow_hook && g.p_plane == 0 && g_grow_hook(g_grow_user, s, sp[s].np + cnt[s], need_nm, &sp[s]) &&
                         sp[s].np + cnt[s] <= sp[s].max_np && need_nm <= sp[s].max_nm;
      if (!grown) {
        if (sp[s].np + cnt[s] > sp[s].max_np)
          VPB_ERROR("species %d: %d particles + %d arrivals exceed max_np=%d (the reference would grow the array by 31%%, "
                    "boundary_p.c:416-430; size max_np with head-room)", sp[s].id, sp[s].np, cnt[s], sp[s].max_np);
        VPB_ERROR("species %d: %d arrivals exceed max_nm=%d", sp[s].id, need_nm, sp[s].max_nm);
      }
      T.p[s] = sp[s].p;
      T.pm[s] = sp[s].pm;
    }
  }
  inject_kernel<<<blocks(n, 128), 128, 0, st>>>(list, n, c1, r1, T, reinterpret_cast<float *>(d_a), g.nbr, c2, nullptr, nullptr);
  rank_bins(c2, n, 0, r2, dc + 24, st);
  compact_movers_kernel<<<blocks(n, 256), 256, 0, st>>>(list, n, c2, r2, T, nullptr, nullptr);
  count_launch(3);
  VPB_CUDA(cudaMemcpyAsync(c.h_pinned_i, dc + 24, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
  VPB_CUDA(cudaStreamSynchronize(st));
  for (int s = 0; s < n_sp; s++) {
    sp[s].np += cnt[s];
    sp[s].nm = c.h_pinned_i[s];
  }
  VPB_CUDA(cudaGetLastError());
}

static long check_species(const vpb_domain_t *dom, const vpb_species_state_t *sp, int n_sp) {
  const DomainDev &g = dom->d;
  long total = 0;
  for (int s = 0; s < n_sp; s++) {
    if (sp[s].nm < 0 || sp[s].nm > sp[s].max_nm) VPB_ERROR("Bad mover count");
    if (g.p_plane > 0 && sp[s].max_np > g.p_plane) VPB_ERROR("species %d: max_np exceeds the domain's particle plane stride", sp[s].id);
    total += sp[s].nm;
  }
  return total;
}

static MoverScratch mover_scratch(long total, int maxnm, bool with_overflow) {
  const size_t o_rank = al256((size_t)total), o_tcode = o_rank + al256((size_t)total * 4), o_trank = o_tcode + al256((size_t)maxnm),
               o_ovf = o_trank + al256((size_t)maxnm * 4), o_end = o_ovf + (with_overflow ? al256((size_t)total * 48) : 0);
  char *scr = (char *)scratch(o_end + 256);
  MoverScratch M;
  M.code = (unsigned char *)scr;
  M.rank = (int *)(scr + o_rank);
  M.tcode = (unsigned char *)(scr + o_tcode);
  M.trank = (int *)(scr + o_trank);
  M.overflow = with_overflow ? (float4 *)(scr + o_ovf) : nullptr;
  return M;
}

// The reference's protocol: counts, then payloads of exactly that size.  ns_out/nr_out (may be null): what went
// through every face, for the capacities of later fused rounds.
static void boundary_round_exact(vpb_domain_t *dom, vpb_species_state_t *sp, int n_sp, vpb_field_t *d_f, vpb_accumulator_t *d_a,
                                 int *ns_out, int *nr_out) {
  Context &c = ctx();
  cudaStream_t st = c.stream;
  const DomainDev &g = dom->d;
  const Faces F = faces_of(dom);
  if (!g_bb.d_counts) VPB_CUDA(cudaMalloc(&g_bb.d_counts, 64 * sizeof(int)));
  int *dc = g_bb.d_counts;
  VPB_CUDA(cudaMemsetAsync(dc, 0, 64 * sizeof(int), st));
  const long total = check_species(dom, sp, n_sp);
  if (total == 0 && !F.any_remote) return;
  if (total > 0 && !d_f) VPB_ERROR("Bad field");
  int maxnm = 0;
  for (int s = 0; s < n_sp; s++) maxnm = sp[s].nm > maxnm ? sp[s].nm : maxnm;
  const MoverScratch M = mover_scratch(total, maxnm, false);

  int ns[6] = {0, 0, 0, 0, 0, 0}, nr[6] = {0, 0, 0, 0, 0, 0};
  Xfer x[12];
  int nx = 0;
  if (total > 0) classify_movers(dom, sp, n_sp, d_f, F, M, dc, st);
  // per-face counts first (boundary_p.c:330-365): both sides then know every payload size
  if (F.any_remote) {
    for (int f = 0; f < 6; f++)
      if (F.remote[f]) x[nx++] = {dc + f, sizeof(int), F.peer[f], nullptr, 0, -1};
    for (int k = 0; k < 6; k++) {
      const int f = kRecvOrder[k];
      if (F.remote[f]) x[nx++] = {nullptr, 0, -1, dc + 8 + f, sizeof(int), F.peer[f]};
    }
    comm_exchange(x, nx);
  }
  VPB_CUDA(cudaMemcpyAsync(c.h_pinned_i, dc, 40 * sizeof(int), cudaMemcpyDeviceToHost, st));
  VPB_CUDA(cudaStreamSynchronize(st));
  for (int f = 0; f < 6; f++) { ns[f] = c.h_pinned_i[f]; nr[f] = F.remote[f] ? c.h_pinned_i[8 + f] : 0; }
  check_mover_flags(c.h_pinned_i, g.rank);
  for (int f = 0; f < 6; f++) {
    if (ns[f] && !F.remote[f]) VPB_ERROR("movers classified for face %d which is not shared with another rank", f);
    const int need = ns[f] > nr[f] ? ns[f] : nr[f];
    if (need) ensure_cap(f, (size_t)need);
    if (ns_out) ns_out[f] = ns[f];
    if (nr_out) nr_out[f] = nr[f];
  }
  if (total > 0) {
    bool sends = false;
    for (int f = 0; f < 6; f++) sends |= ns[f] > 0;
    PackPlan plan;
    plan.first = 0; plan.overflow = nullptr; plan.count = nullptr;
    for (int f = 0; f < 6; f++) plan.cap[f] = 0x7fffffff;
    pack_and_remove(dom, sp, n_sp, F, M, plan, sends, dc, st);
  }
  // payloads (boundary_p.c:369-384)
  if (F.any_remote) {
    nx = 0;
    for (int f = 0; f < 6; f++)
      if (F.remote[f] && ns[f]) x[nx++] = {g_bb.send[f], (size_t)ns[f] * 48, F.peer[f], nullptr, 0, -1};
    for (int k = 0; k < 6; k++) {
      const int f = kRecvOrder[k];
      if (F.remote[f] && nr[f]) x[nx++] = {nullptr, 0, -1, g_bb.recv[f], (size_t)nr[f] * 48, F.peer[f]};
    }
    if (nx) comm_exchange(x, nx);
  }
  const float4 *local = nullptr;
  if (g_local_inj_n > 0) {
    if ((size_t)g_local_inj_n > g_bb.lcap) {
      VPB_CUDA(cudaStreamSynchronize(st));
      if (g_bb.local) cudaFree(g_bb.local);
      g_bb.lcap = (size_t)g_local_inj_n + (size_t)g_local_inj_n / 4 + 1024;
      VPB_CUDA(cudaMalloc(&g_bb.local, g_bb.lcap * 48));
    }
    VPB_CUDA(cudaMemcpyAsync(g_bb.local, g_local_inj, (size_t)g_local_inj_n * 48, cudaMemcpyHostToDevice, st));
    local = g_bb.local;
  }
  inject_known(dom, sp, n_sp, d_a, g_bb.recv, nr, false, local, g_local_inj_n);
}

// capacity both sides of a face derive from the counts they both know
static int capacity_for(int ns, int nr) {
  const int m = ns > nr ? ns : nr;
  int cap = m + m / 2 + 64;
  const int lim = tuning("boundary.cap_max", 0);     // tests: force the second message
  if (lim > 0 && cap > lim) cap = lim;
  return cap;
}

// One fixed-capacity message per face, sizes from the headers on the device, one read-back (see the top of the file).
static void boundary_round_fused(vpb_domain_t *dom, vpb_species_state_t *sp, int n_sp, vpb_field_t *d_f, vpb_accumulator_t *d_a, int slot) {
  Context &c = ctx();
  cudaStream_t st = c.stream;
  const DomainDev &g = dom->d;
  const Faces F = faces_of(dom);
  int *cap = dom->mig_cap[slot];
  if (!g_bb.d_counts) VPB_CUDA(cudaMalloc(&g_bb.d_counts, 64 * sizeof(int)));
  int *dc = g_bb.d_counts;
  const long total = check_species(dom, sp, n_sp);
  if (total > 0 && !d_f) VPB_ERROR("Bad field");
  if (!d_a) VPB_ERROR("Bad accumulator");
  long cap_in = 0;
  for (int f = 0; f < 6; f++)
    if (F.remote[f]) { ensure_cap(f, (size_t)cap[f] + 1); cap_in += cap[f]; }
  VPB_CUDA(cudaMemsetAsync(dc, 0, 64 * sizeof(int), st));
  // one scratch block: mover arrays | overflow list | arrival arrays sized by the capacities
  int maxnm = 0;
  for (int s = 0; s < n_sp; s++) maxnm = sp[s].nm > maxnm ? sp[s].nm : maxnm;
  const size_t o_rank = al256((size_t)total), o_tcode = o_rank + al256((size_t)total * 4), o_trank = o_tcode + al256((size_t)maxnm),
               o_ovf = o_trank + al256((size_t)maxnm * 4), o_list = o_ovf + al256((size_t)total * 48),
               o_c1 = o_list + al256((size_t)cap_in * 48), o_r1 = o_c1 + al256((size_t)cap_in), o_c2 = o_r1 + al256((size_t)cap_in * 4),
               o_r2 = o_c2 + al256((size_t)cap_in), o_end = o_r2 + al256((size_t)cap_in * 4);
  char *scr = (char *)scratch(o_end + 256);
  MoverScratch M;
  M.code = (unsigned char *)scr; M.rank = (int *)(scr + o_rank); M.tcode = (unsigned char *)(scr + o_tcode); M.trank = (int *)(scr + o_trank);
  M.overflow = (float4 *)(scr + o_ovf);
  float4 *list = (float4 *)(scr + o_list);
  unsigned char *c1 = (unsigned char *)(scr + o_c1), *c2 = (unsigned char *)(scr + o_c2);
  int *r1 = (int *)(scr + o_r1), *r2 = (int *)(scr + o_r2);

  PackPlan plan;
  plan.first = 1; plan.overflow = M.overflow; plan.count = dc;
  for (int f = 0; f < 6; f++) plan.cap[f] = F.remote[f] ? cap[f] : 0;
  if (total > 0) classify_movers(dom, sp, n_sp, d_f, F, M, dc, st);
  write_headers_kernel<<<1, 32, 0, st>>>(g_bb.d_send_table, dc, plan, slot, F.mask);
  count_launch();
  if (total > 0) pack_and_remove(dom, sp, n_sp, F, M, plan, true, dc, st);
  Xfer x[12];
  int nx = 0;
  for (int f = 0; f < 6; f++)
    if (F.remote[f]) x[nx++] = {g_bb.send[f], ((size_t)cap[f] + 1) * 48, F.peer[f], nullptr, 0, -1};
  for (int k = 0; k < 6; k++) {
    const int f = kRecvOrder[k];
    if (F.remote[f]) x[nx++] = {nullptr, 0, -1, g_bb.recv[f], ((size_t)cap[f] + 1) * 48, F.peer[f]};
  }
  comm_exchange(x, nx);

  RecvPlan R;
  R.round = slot;
  for (int f = 0; f < 6; f++) { R.buf[f] = F.remote[f] ? g_bb.recv[f] : nullptr; R.cap[f] = F.remote[f] ? cap[f] : 0; }
  SpeciesTable T = species_table(dom, sp, n_sp);
  const int nmax = (int)cap_in;
  if (nmax > 0) {
    gather_fused_kernel<<<blocks(3L * nmax, 256), 256, 0, st>>>(list, R, dc);
    species_code_kernel<<<blocks(nmax, 256), 256, 0, st>>>(list, nmax, T, c1, dc + 14);
    rank_bins(c1, nmax, 0, r1, dc + 16, st, dc + 14);
    inject_kernel<<<blocks(nmax, 128), 128, 0, st>>>(list, nmax, c1, r1, T, reinterpret_cast<float *>(d_a), g.nbr, c2, dc + 14, dc + 35);
    rank_bins(c2, nmax, 0, r2, dc + 24, st, dc + 14);
    compact_movers_kernel<<<blocks(nmax, 256), 256, 0, st>>>(list, nmax, c2, r2, T, dc + 14, dc + 36);
    count_launch(4);
  }
  VPB_CUDA(cudaMemcpyAsync(c.h_pinned_i, dc, 48 * sizeof(int), cudaMemcpyDeviceToHost, st));
  VPB_CUDA(cudaStreamSynchronize(st));
  VPB_CUDA(cudaGetLastError());
  int h[48];
  memcpy(h, c.h_pinned_i, sizeof(h));
  check_mover_flags(h, g.rank);
  if (h[37]) VPB_ERROR("boundary_p: %d fused migration messages arrived with a header that does not match this rank's round %d / capacities "
                       "(ranks out of step?)", h[37], slot);
  int ns[6], nr[6];
  bool over = false;
  for (int f = 0; f < 6; f++) {
    ns[f] = h[f];
    nr[f] = F.remote[f] ? h[8 + f] : 0;
    if (ns[f] && !F.remote[f]) VPB_ERROR("movers classified for face %d which is not shared with another rank", f);
    over |= F.remote[f] && (ns[f] > cap[f] || nr[f] > cap[f]);
  }
  for (int s = 0; s < n_sp; s++) {
    const int cnt = h[16 + s];
    if (h[35] && sp[s].np + cnt > sp[s].max_np)
      VPB_ERROR("species %d: %d particles + %d arrivals exceed max_np=%d (size max_np with head-room)", sp[s].id, sp[s].np, cnt, sp[s].max_np);
    if (h[36] && h[24 + s] > sp[s].max_nm) VPB_ERROR("species %d: %d new movers exceed max_nm=%d", sp[s].id, h[24 + s], sp[s].max_nm);
    sp[s].np += cnt;
    sp[s].nm = h[24 + s];
  }
  if (h[35] || h[36]) VPB_ERROR("boundary_p: arrivals beyond a species' capacity (%d particles, %d movers)", h[35], h[36]);
  if (over) {
    // the rest of the over-full faces, exactly sized: both sides read the same two counts
    size_t need = 0;
    int extra[6];
    for (int f = 0; f < 6; f++) { extra[f] = F.remote[f] && nr[f] > cap[f] ? nr[f] - cap[f] : 0; need += (size_t)extra[f]; }
    if (need > g_bb.xcap) {
      VPB_CUDA(cudaStreamSynchronize(st));
      if (g_bb.xrecv) cudaFree(g_bb.xrecv);
      g_bb.xcap = need + need / 4 + 1024;
      VPB_CUDA(cudaMalloc(&g_bb.xrecv, g_bb.xcap * 48));
    }
    const float4 *xb[6];
    nx = 0;
    size_t soff = 0, roff = 0;
    for (int f = 0; f < 6; f++) {
      const int more = F.remote[f] && ns[f] > cap[f] ? ns[f] - cap[f] : 0;
      if (more) x[nx++] = {M.overflow + 3 * soff, (size_t)more * 48, F.peer[f], nullptr, 0, -1};
      soff += (size_t)more;
    }
    for (int f = 0; f < 6; f++) { xb[f] = g_bb.xrecv + 3 * roff; roff += (size_t)extra[f]; }
    for (int k = 0; k < 6; k++) {
      const int f = kRecvOrder[k];
      if (extra[f]) x[nx++] = {nullptr, 0, -1, const_cast<float4 *>(xb[f]), (size_t)extra[f] * 48, F.peer[f]};
    }
    if (nx) comm_exchange(x, nx);
    inject_known(dom, sp, n_sp, d_a, xb, extra, true);
  }
  // grow a capacity before it is reached (never shrink: a message slot costs nothing but its transfer time)
  for (int f = 0; f < 6; f++) {
    const int m = ns[f] > nr[f] ? ns[f] : nr[f];
    if (F.remote[f] && m + m / 4 > cap[f]) cap[f] = capacity_for(ns[f], nr[f]);
  }
}

}  // namespace vpb

using namespace vpb;

extern "C" {

void vpb_move_p_one(vpb_domain_t *dom, vpb_particle_t *d_p, vpb_particle_mover_t *d_pm, vpb_accumulator_t *d_a, int *d_result) {
  move_p_one_kernel<<<1, 1, 0, ctx().stream>>>(d_p, d_pm, reinterpret_cast<float *>(d_a), dom->d.nbr, d_result);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

void vpb_accumulate_rhob_one(vpb_domain_t *dom, vpb_field_t *d_f, const vpb_particle_t *d_particle) {
  accumulate_rhob_one_kernel<<<1, 1, 0, ctx().stream>>>(d_f, d_particle, dom->d);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

// One round of boundary_p over n_sp species (at most 7).  sp[s].nm is the number of movers in
// sp[s].pm (ascending particle index); on return np/nm are updated.  Synchronises the stream.
// round >= 0 (the caller's rounds of one step, advance.cxx:94-96, numbered from 0) and tuning boundary.fused != 0:
// fixed-capacity fused messages once a first exact round has told both sides of every face what passes through it;
// every rank of the job must use the same numbering.  Otherwise the reference's exact two-message protocol, which
// is the default: measured on 2 and 4 B200s the two cost the same (60.84 / 60.72 and 60.95 / 60.96 ms per step,
// profiles/r2r, r2s) -- what a rank spends in here is mostly the wait for the slowest rank's advance_p at the first
// message, not read-backs or message count.
void vpb_boundary_p_round(vpb_domain_t *dom, vpb_species_state_t *sp, int n_sp, vpb_field_t *d_f, vpb_accumulator_t *d_a, int round) {
  if (!dom) VPB_ERROR("Bad grid");
  if (n_sp < 0 || n_sp > 7) VPB_ERROR("boundary_p handles at most 7 species per call (got %d)", n_sp);
  if (n_sp && !sp) VPB_ERROR("Bad species");
  const int mode = tuning("boundary.fused", 0);
  bool any_remote = false;
  for (int f = 0; f < 6; f++) {
    const int b = dom->d.bc[kFaceBound[f]];
    any_remote |= b >= 0 && b < dom->d.nproc && b != dom->d.rank;
  }
  if (round < 0 || mode == 0 || !any_remote || g_local_inj_set) { boundary_round_exact(dom, sp, n_sp, d_f, d_a, nullptr, nullptr); return; }
  const int slot = round < 3 ? round : 2;
  if (!dom->mig_cap_valid[slot]) {
    int ns[6], nr[6];
    for (int f = 0; f < 6; f++) ns[f] = nr[f] = 0;
    boundary_round_exact(dom, sp, n_sp, d_f, d_a, ns, nr);
    for (int f = 0; f < 6; f++) dom->mig_cap[slot][f] = capacity_for(ns[f], nr[f]);
    dom->mig_cap_valid[slot] = true;
    return;
  }
  boundary_round_fused(dom, sp, n_sp, d_f, d_a, slot);
}

void vpb_boundary_p(vpb_domain_t *dom, vpb_species_state_t *sp, int n_sp, vpb_field_t *d_f, vpb_accumulator_t *d_a) {
  // boundary.fused = 2 (tests): the reference-named entry point through the fused rounds as well, the round taken
  // from the call count (advance.cxx makes three calls a step); arrays cannot grow on that path
  static long calls = 0;
  const int round = tuning("boundary.fused", 0) == 2 ? (int)(calls++ % 3) : -1;
  vpb_boundary_p_round(dom, sp, n_sp, d_f, d_a, round);
}

// The reference-named boundary_p() brackets its call with these: `inj[0..n)` are the injectors the deck's custom
// handlers made on the host for this round's movers (n may be 0); NULL/-1 ends the bracket.
void vpb_boundary_set_local_injectors(const vpb_particle_injector_t *inj, int n) {
  g_local_inj = n > 0 ? inj : nullptr;
  g_local_inj_n = n > 0 ? n : 0;
  g_local_inj_set = n >= 0;
}

void vpb_boundary_set_grow_hook(vpb_grow_hook_t hook, void *user) {
  g_grow_hook = hook;
  g_grow_user = user;
}

}  // extern "C"
