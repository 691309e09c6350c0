// vpb_dropin.cu -- layer (A) of include/vpic_b200.h: the reference's own entry
// points (same names, prototypes, error behaviour), implemented on the device
// kernels.  These are the symbols that replace the reference's translation units
// at link time.
//
// Pointer handling: each argument array is classified with
// cudaPointerGetAttributes.  Device or managed memory (vpb_dev_alloc,
// vpb_malloc_managed, util_malloc_aligned below) is used in place; plain host
// memory is staged into a cached device buffer before the kernel and copied back
// after it (that path is what bench.py reports as `e2e`).  Every entry point
// returns with the stream drained, because the caller is host code that will
// read the arrays next.
#include <math.h>
#include <unordered_map>
#include <vector>
#include "vpb_advance_p.cuh"
#include "vpb_common.cuh"

namespace vpb {

enum { RD = 1, WR = 2, RW = 3 };

struct StageBuf { void *dev = nullptr; size_t cap = 0; };
static std::unordered_map<const void *, StageBuf> g_stage;
static std::unordered_map<const void *, int> g_nmat;          // material_coefficient_t* -> count
static std::unordered_map<const void *, vpb_domain_t *> g_domains;
static int g_world_nproc = 1;
// traversal hints: particle array (caller's pointer) -> device copy of partition[] from our last sort_p of it
struct PartHint { int *dev = nullptr; size_t n = 0; int np = 0; };
static std::unordered_map<const void *, PartHint> g_part_hint;
static size_t g_h2d_total = 0, g_d2h_total = 0;   // bytes moved by the staging path (bench.py e2e accounting)
// Managed arrays already brought to the device once: allocation -> bytes prefetched so far.  Measured on a B200
// (profiles/r2a_deck_trace.txt): cudaMemPrefetchAsync costs 140 us of host time per call even when every page is
// already resident -- 1444 calls = 203 ms of a 248 ms run -- so an array is prefetched when it is first seen (one DMA
// instead of a storm of GPU page faults: 3-9 ms instead of 36-41 ms for a 400 MB species) and never again; pages the
// host program touches later come back through ordinary demand migration.  dropin.prefetch: 0 never, 1 first sight
// (default), 2 every call.
static std::unordered_map<const void *, size_t> g_prefetched;

static size_t nvox(const vpb_grid_t *g) { return (size_t)(g->nx + 2) * (g->ny + 2) * (g->nz + 2); }

struct Acq { void *host; void *dev; size_t bytes; int mode; };

struct Residency {
  std::vector<Acq> staged;
  size_t h2d = 0, d2h = 0;

  void *get(const void *host, size_t bytes, int mode) {
    if (!host || bytes == 0) return const_cast<void *>(host);
    Context &c = ctx();
    cudaPointerAttributes at;
    cudaError_t e;
    { TraceScope _t("  ptr_attributes"); e = cudaPointerGetAttributes(&at, host); }
    if (e != cudaSuccess) { cudaGetLastError(); at.type = cudaMemoryTypeUnregistered; }
    if (at.type == cudaMemoryTypeDevice) return const_cast<void *>(host);
    if (at.type == cudaMemoryTypeManaged) {
      const int policy = tuning("dropin.prefetch", 1);
      if (policy == 0) return const_cast<void *>(host);
      if (policy == 1) {
        size_t &done = g_prefetched[host];
        if (done >= bytes) return const_cast<void *>(host);
        done = bytes;
      }
      TraceScope _t("  prefetch");
      cudaMemPrefetchAsync(host, bytes, c.device, c.stream);
      cudaGetLastError();
      return const_cast<void *>(host);
    }
    StageBuf &sb = g_stage[host];
    if (sb.cap < bytes) {
      if (sb.dev) { VPB_CUDA(cudaStreamSynchronize(c.stream)); VPB_CUDA(cudaFree(sb.dev)); }
      VPB_CUDA(cudaMalloc(&sb.dev, bytes));
      sb.cap = bytes;
    }
    if (mode & RD) { VPB_CUDA(cudaMemcpyAsync(sb.dev, host, bytes, cudaMemcpyHostToDevice, c.stream)); h2d += bytes; }
    staged.push_back({const_cast<void *>(host), sb.dev, bytes, mode});
    return sb.dev;
  }

  // A field array: like get(), and when the array had to be staged anyway (plain host memory) and the tuning
  // "dropin.planar" is on, the staged copy is re-laid out planar for the duration of the call (the layout the
  // device-resident driver uses; tests run the whole layer-A suite through it with VPB_DROPIN_PLANAR=1).
  struct PlanarAcq { vpb_domain_t *dom; void *aos, *planar; int mode; };
  std::vector<PlanarAcq> planar;
  void *get_field(vpb_domain_t *dom, const void *host, size_t bytes, int mode) {
    void *d = get(host, bytes, mode);
    if (d == host || !host || !tuning("dropin.planar", 0)) return d;
    Context &c = ctx();
    for (auto &pa : planar)
      if (pa.aos == d) return pa.planar;
    vpb_domain_set_field_layout(dom, 1);
    void *pl = nullptr;
    VPB_CUDA(cudaMallocAsync(&pl, vpb_field_bytes(dom), c.stream));
    VPB_CUDA(cudaMemsetAsync(pl, 0, vpb_field_bytes(dom), c.stream));
    if (mode & RD) vpb_field_convert(dom, (vpb_field_t *)pl, (const vpb_field_t *)d, 1);
    planar.push_back({dom, d, pl, mode});
    return pl;
  }

  // A particle array: like get(), and with the tuning "dropin.particle_planes" on, a staged array is re-laid out
  // as component planes for the duration of the call (the layout the device-resident driver uses), so that the
  // layer-A parity suite also exercises the two-particles-per-lane advance_p and the plane accessors
  // (VPB_DROPIN_PARTICLE_PLANES=1).
  struct PlaneAcq { vpb_domain_t *dom; void *aos, *planes; long np; int mode; };
  std::vector<PlaneAcq> pplanes;
  void *get_particles(vpb_domain_t *dom, const void *host, long np, int mode) {
    void *d = get(host, (size_t)np * sizeof(vpb_particle_t), mode);
    if (d == host || !host || np <= 0 || !tuning("dropin.particle_planes", 0)) return d;
    Context &c = ctx();
    const long plane = (np + 63) / 64 * 64;
    void *pl = nullptr;
    VPB_CUDA(cudaMallocAsync(&pl, (size_t)plane * sizeof(vpb_particle_t), c.stream));
    VPB_CUDA(cudaMemsetAsync(pl, 0, (size_t)plane * sizeof(vpb_particle_t), c.stream));
    vpb_domain_set_particle_layout(dom, plane);
    vpb_particle_convert(dom, (vpb_particle_t *)pl, (const vpb_particle_t *)d, np, 1);
    pplanes.push_back({dom, d, pl, np, mode});
    return pl;
  }

  // copy results back and drain the stream
  void finish() {
    Context &c = ctx();
    for (auto &pa : pplanes) {
      if (pa.mode & WR) vpb_particle_convert(pa.dom, (vpb_particle_t *)pa.aos, (const vpb_particle_t *)pa.planes, pa.np, 0);
      VPB_CUDA(cudaFreeAsync(pa.planes, c.stream));
      vpb_domain_set_particle_layout(pa.dom, 0);
    }
    pplanes.clear();
    for (auto &pa : planar) {
      if (pa.mode & WR) vpb_field_convert(pa.dom, (vpb_field_t *)pa.aos, (const vpb_field_t *)pa.planar, 0);
      VPB_CUDA(cudaFreeAsync(pa.planar, c.stream));
      vpb_domain_set_field_layout(pa.dom, 0);
    }
    planar.clear();
    for (auto &a : staged)
      if (a.mode & WR) { VPB_CUDA(cudaMemcpyAsync(a.host, a.dev, a.bytes, cudaMemcpyDeviceToHost, c.stream)); d2h += a.bytes; }
    { TraceScope _t("  finish_sync"); VPB_CUDA(cudaStreamSynchronize(c.stream)); }
    staged.clear();
    g_h2d_total += h2d; g_d2h_total += d2h;
    h2d = d2h = 0;
  }
};

// Cheap fingerprint of the neighbor table (a grid_t may be edited -- set_domain_particle_bc rewrites the outward
// entries of a whole face, grid/ops.c -- or its address reused).  It runs on every layer-A call and the table is
// managed memory read by the host here, so it samples instead of walking: the outward entry of five interior cells
// spread over each of the six faces, plus the two ends of the table (was: 4096 words, 26 us per call and 17 % of a
// 64^3 step, profiles/r2a_deck_trace.txt).
static uint64_t neighbor_print(const vpb_grid_t *g) {
  if (!g->neighbor) return 0;
  const int n[3] = {g->nx, g->ny, g->nz};
  const long sx = g->nx + 2, sxy = sx * (g->ny + 2);
  uint64_t h = 1469598103934665603ull ^ (uint64_t)(uintptr_t)g->neighbor;
  for (int face = 0; face < 6; face++) {
    const int axis = face % 3, hi = face / 3;          // neighbor[6*v + face]: faces 0,1,2 = -x,-y,-z; 3,4,5 = +x,+y,+z
    for (int k = 0; k < 5; k++) {
      int c[3];
      for (int a = 0; a < 3; a++) c[a] = 1 + (int)(((long)(n[a] - 1) * ((k * (a + 2)) % 5)) / 4);
      c[axis] = hi ? n[axis] : 1;
      const long v = c[0] + sx * c[1] + sxy * c[2];
      h = (h ^ (uint64_t)g->neighbor[6 * v + face]) * 1099511628211ull + (uint64_t)v;
    }
  }
  const size_t last = 6 * nvox(g) - 1;
  return (h ^ (uint64_t)g->neighbor[0]) * 1099511628211ull ^ (uint64_t)g->neighbor[last];
}
static std::unordered_map<const void *, uint64_t> g_domain_print;

// A host program that is the reference itself never calls vpb_comm_init: bring the communicator up through its own
// message layer the first time a grid with an mp handle shows up.  The bootstrap is collective, so it is also tried
// from new_field_advance / new_interpolator / new_accumulators, which every rank calls from finalize_field_advance
// (vpic.hxx:373-384) before any rank-dependent work such as inject_particle can reach a hot-path entry point.
static void ensure_comm(const vpb_grid_t *g) {
  static bool boot_tried = false;
  if (boot_tried || !g || !g->mp || g_world_nproc != 1) return;
  boot_tried = true;
  const int n = vpb_comm_autoboot(g->mp);
  if (n > 1) g_world_nproc = n;
}

static vpb_domain_t *domain_of(const vpb_grid_t *g) {
  if (!g) VPB_ERROR("Bad grid");
  TraceScope _t("  domain_of");
  auto it = g_domains.find(g);
  if (it != g_domains.end() && g_domain_print[g] != neighbor_print(g)) {
    vpb_domain_destroy(it->second);
    g_domains.erase(it);
    it = g_domains.end();
  }
  if (it != g_domains.end()) {
    const DomainDev &d = it->second->d;
    bool same = d.nx == g->nx && d.ny == g->ny && d.nz == g->nz && d.dt == g->dt && d.cvac == g->cvac && d.eps0 == g->eps0 &&
                d.damp == g->damp && d.rdx == g->rdx && d.rdy == g->rdy && d.rdz == g->rdz && d.rangel == g->rangel;
    for (int i = 0; i < 27 && same; i++) same = d.bc[i] == g->bc[i];
    if (same) return it->second;
    vpb_domain_destroy(it->second);
    g_domains.erase(it);
  }
  ensure_comm(g);
  // the centre entry of bc[] is this rank (grid_structors.c:22, ops.c:47)
  vpb_domain_t *dom = vpb_domain_create(g, g->bc[VPB_BOUNDARY(0, 0, 0)], g_world_nproc);
  g_domains[g] = dom;
  g_domain_print[g] = neighbor_print(g);
  return dom;
}

static int nmat_of(const vpb_material_coefficient_t *m) {
  auto it = g_nmat.find(m);
  if (it == g_nmat.end())
    VPB_ERROR("material coefficient array %p was not created by new_material_coefficients; "
              "call vpb_register_material_coefficients(m, n) first", (const void *)m);
  return it->second;
}

}  // namespace vpb

using namespace vpb;

extern "C" {

// ---------------------------------------------------------------------------
// bookkeeping the reference keeps implicitly
// ---------------------------------------------------------------------------
void vpb_set_world(int nproc) { g_world_nproc = nproc < 1 ? 1 : nproc; }
void vpb_register_material_coefficients(const vpb_material_coefficient_t *m, int n_mat) { g_nmat[m] = n_mat; }
void vpb_grid_changed(const vpb_grid_t *g) {
  auto it = g_domains.find(g);
  if (it != g_domains.end()) { vpb_domain_destroy(it->second); g_domains.erase(it); }
}
void vpb_staging_release(void) {
  VPB_CUDA(cudaStreamSynchronize(ctx().stream));
  for (auto &kv : g_stage) if (kv.second.dev) cudaFree(kv.second.dev);
  g_stage.clear();
}
vpb_domain_t *vpb_domain_of_grid(const vpb_grid_t *g) { return domain_of(g); }

// ---------------------------------------------------------------------------
// util_malloc_aligned / util_free_aligned (util.c:46-91): managed memory, so every
// array the reference allocates is directly usable by the kernels AND by the deck
// ---------------------------------------------------------------------------
void util_malloc_aligned(const char *err_fmt, void *mem_ref, size_t n, size_t a) {
  // util.c:46-78: err_fmt has exactly two %lu (bytes, alignment); a must be a power of two.  cudaMallocManaged returns
  // at least 256-byte alignment; the reference asks for <= 128.
  if (!err_fmt) err_fmt = "malloc aligned failed (n=%lu, a=%lu)";
  char **mem = (char **)mem_ref;
  if (!mem || a == 0 || (a & (a - 1)) != 0) {
    fprintf(stderr, "Error at %s(%i):\n\t", __FILE__, __LINE__);
    fprintf(stderr, err_fmt, (unsigned long)n, (unsigned long)a);
    fprintf(stderr, "\n");
    exit(1);
  }
  if (n == 0) { *mem = nullptr; return; }
  void *p = nullptr;
  ctx();
  cudaError_t e = cudaMallocManaged(&p, n);
  if (e != cudaSuccess) {
    fprintf(stderr, "Error at %s(%i):\n\t", __FILE__, __LINE__);
    fprintf(stderr, err_fmt, (unsigned long)n, (unsigned long)a);
    fprintf(stderr, "\n");
    exit(1);
  }
  *mem = (char *)p;
}

void util_free_aligned(void *mem_ref) {
  char **mem = (char **)mem_ref;
  if (!mem || !*mem) return;
  cudaStreamSynchronize(ctx().stream);
  g_prefetched.erase(*mem);
  cudaFree(*mem);
  *mem = nullptr;
}

// ---------------------------------------------------------------------------
// species_advance (spa.h)
// ---------------------------------------------------------------------------
// Host particle arrays larger than this many particles are streamed through the device in pieces:
// H2D of piece i+1, the kernel on piece i and D2H of piece i-1 run concurrently (three streams, three
// rotating device buffers), so a call costs about max(PCIe in, PCIe out, kernel) instead of their sum.
struct PipeBuffers {
  vpb_particle_t *buf[3] = {nullptr, nullptr, nullptr};
  size_t cap = 0;   // particles per buffer
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[3], ev_k[3], ev_free[3], ev_aux;
  bool ready = false;
};
static PipeBuffers g_pipe;

static void pipe_prepare(size_t piece) {
  PipeBuffers &P = g_pipe;
  if (!P.ready) {
    VPB_CUDA(cudaStreamCreateWithFlags(&P.s_in, cudaStreamNonBlocking));
    VPB_CUDA(cudaStreamCreateWithFlags(&P.s_out, cudaStreamNonBlocking));
    for (int i = 0; i < 3; i++) {
      VPB_CUDA(cudaEventCreateWithFlags(&P.ev_in[i], cudaEventDisableTiming));
      VPB_CUDA(cudaEventCreateWithFlags(&P.ev_k[i], cudaEventDisableTiming));
      VPB_CUDA(cudaEventCreateWithFlags(&P.ev_free[i], cudaEventDisableTiming));
    }
    VPB_CUDA(cudaEventCreateWithFlags(&P.ev_aux, cudaEventDisableTiming));
    P.ready = true;
  }
  if (P.cap < piece) {
    VPB_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < 3; i++) {
      if (P.buf[i]) cudaFree(P.buf[i]);
      VPB_CUDA(cudaMalloc(&P.buf[i], piece * sizeof(vpb_particle_t)));
    }
    P.cap = piece;
  }
}

static bool is_plain_host(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return at.type == cudaMemoryTypeUnregistered || at.type == cudaMemoryTypeHost;
}

int advance_p(vpb_particle_t *p0, int np, const float q_m, vpb_particle_mover_t *pm, int max_nm, vpb_accumulator_t *a0,
              const vpb_interpolator_t *f0, const vpb_grid_t *g) {
  TraceScope _ts("advance_p");
  if (!p0) VPB_ERROR("Bad particle array");
  if (np < 0) VPB_ERROR("Bad number of particles");
  if (!pm) VPB_ERROR("Bad particle mover");
  if (max_nm < 0) VPB_ERROR("Bad number of movers");
  if (!a0) VPB_ERROR("Bad accumulator");
  if (!f0) VPB_ERROR("Bad interpolator");
  if (!g) VPB_ERROR("Bad grid");
  vpb_domain_t *dom = domain_of(g);
  Context &c = ctx();
  Residency r;
  const int piece = tuning("dropin.piece", 4 << 20) & ~31;   // particles per piece
  const bool streamed = is_plain_host(p0) && np > 2 * piece && tuning("dropin.pipeline", 1) && !tuning("dropin.particle_planes", 0);
  int *d_out = nullptr;
  VPB_CUDA(cudaMallocAsync(&d_out, 2 * sizeof(int), c.stream));
  // movers: staged without copying max_nm records back; only the nm that exist are returned
  vpb_particle_mover_t *dpm = pm;
  const bool pm_host = is_plain_host(pm);
  if (pm_host) {
    StageBuf &sb = g_stage[pm];
    const size_t bytes = (size_t)max_nm * sizeof(*pm) + 16;
    if (sb.cap < bytes) {
      if (sb.dev) { VPB_CUDA(cudaStreamSynchronize(c.stream)); VPB_CUDA(cudaFree(sb.dev)); }
      VPB_CUDA(cudaMalloc(&sb.dev, bytes));
      sb.cap = bytes;
    }
    dpm = (vpb_particle_mover_t *)sb.dev;
  }
  if (!streamed) {
    vpb_particle_t *dp = (vpb_particle_t *)r.get_particles(dom, p0, np, RW);
    vpb_accumulator_t *da = (vpb_accumulator_t *)r.get(a0, nvox(g) * sizeof(*a0), RW);
    const vpb_interpolator_t *df = (const vpb_interpolator_t *)r.get(f0, nvox(g) * sizeof(*f0), RD);
    const int *d_part = nullptr;
    auto ph = g_part_hint.find(p0);
    if (ph != g_part_hint.end() && ph->second.np == np && ph->second.n == nvox(g) + 1) d_part = ph->second.dev;
    vpb_advance_p_ordered(dom, dp, np, q_m, dpm, max_nm, da, df, d_out, d_part);
  } else {
    PipeBuffers &P = g_pipe;
    pipe_prepare((size_t)piece);
    // field-sized arrays go over whole, on the input stream, before the first piece
    vpb_accumulator_t *da = (vpb_accumulator_t *)r.get(a0, nvox(g) * sizeof(*a0), RW);
    const vpb_interpolator_t *df = (const vpb_interpolator_t *)r.get(f0, nvox(g) * sizeof(*f0), RD);
    AdvanceJob J;
    advance_p_begin(dom, np, q_m, max_nm, da, df, J, c.stream);
    VPB_CUDA(cudaEventRecord(P.ev_aux, c.stream));
    VPB_CUDA(cudaStreamWaitEvent(P.s_in, P.ev_aux, 0));    // buffers of a previous call are idle by now
    VPB_CUDA(cudaStreamWaitEvent(P.s_out, P.ev_aux, 0));
    const int npieces = (np + piece - 1) / piece;
    // advance_p reads and writes only the 32 hot bytes of a 48-byte record (the tags are neither used nor changed): with
    // dropin.hot_only the copies are 2-D, 32 bytes of every 48 (a third less over PCIe if the copy engine keeps its
    // rate on 32-byte rows -- to be measured; default off)
    const bool hot_only = tuning("dropin.hot_only", 0) != 0;
    for (int i = 0; i < npieces; i++) {
      const int b = i % 3, k0 = i * piece, k1 = (k0 + piece < np) ? k0 + piece : np;
      size_t bytes = (size_t)(k1 - k0) * sizeof(vpb_particle_t);
      if (i >= 3) VPB_CUDA(cudaStreamWaitEvent(P.s_in, P.ev_free[b], 0));
      if (hot_only)
        VPB_CUDA(cudaMemcpy2DAsync(P.buf[b], sizeof(vpb_particle_t), p0 + k0, sizeof(vpb_particle_t), 32, (size_t)(k1 - k0),
                                   cudaMemcpyHostToDevice, P.s_in));
      else
        VPB_CUDA(cudaMemcpyAsync(P.buf[b], p0 + k0, bytes, cudaMemcpyHostToDevice, P.s_in));
      VPB_CUDA(cudaEventRecord(P.ev_in[b], P.s_in));
      VPB_CUDA(cudaStreamWaitEvent(c.stream, P.ev_in[b], 0));
      advance_p_range(J, P.buf[b] - k0, k0, k1, nullptr, c.stream);
      VPB_CUDA(cudaEventRecord(P.ev_k[b], c.stream));
      VPB_CUDA(cudaStreamWaitEvent(P.s_out, P.ev_k[b], 0));
      if (hot_only) {
        VPB_CUDA(cudaMemcpy2DAsync(p0 + k0, sizeof(vpb_particle_t), P.buf[b], sizeof(vpb_particle_t), 32, (size_t)(k1 - k0),
                                   cudaMemcpyDeviceToHost, P.s_out));
        bytes = (size_t)(k1 - k0) * 32;
      } else {
        VPB_CUDA(cudaMemcpyAsync(p0 + k0, P.buf[b], bytes, cudaMemcpyDeviceToHost, P.s_out));
      }
      VPB_CUDA(cudaEventRecord(P.ev_free[b], P.s_out));
      g_h2d_total += bytes; g_d2h_total += bytes;
    }
    advance_p_end(J, dpm, d_out, c.stream);
    VPB_CUDA(cudaStreamSynchronize(P.s_out));
  }
  VPB_CUDA(cudaMemcpyAsync(c.h_pinned_i, d_out, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  VPB_CUDA(cudaFreeAsync(d_out, c.stream));
  r.finish();
  const int nm = c.h_pinned_i[0];
  if (pm_host && nm > 0) {
    VPB_CUDA(cudaMemcpyAsync(pm, dpm, (size_t)nm * sizeof(*pm), cudaMemcpyDeviceToHost, c.stream));
    VPB_CUDA(cudaStreamSynchronize(c.stream));
    g_d2h_total += (size_t)nm * sizeof(*pm);
  }
  const int ignored = nm >= max_nm ? vpb_advance_p_ignored() : 0;
  if (ignored) VPB_WARNING("advance_p ran out of storage for %d movers", ignored);   // advance_p.cxx:463-465
  return nm;
}

void center_p(vpb_particle_t *p0, int np, const float q_m, const vpb_interpolator_t *f0, const vpb_grid_t *g) {
  TraceScope _ts("center_p");
  if (np < 0) VPB_ERROR("Bad number of particles");
  if (!f0) VPB_ERROR("Bad interpolator");
  if (!g) VPB_ERROR("Bad grid");
  vpb_domain_t *dom = domain_of(g);
  Residency r;
  vpb_particle_t *dp = (vpb_particle_t *)r.get_particles(dom, p0, np, RW);
  const vpb_interpolator_t *df = (const vpb_interpolator_t *)r.get(f0, nvox(g) * sizeof(*f0), RD);
  vpb_center_p(dom, dp, np, q_m, df);
  r.finish();
}

void uncenter_p(vpb_particle_t *p0, int np, const float q_m, const vpb_interpolator_t *f0, const vpb_grid_t *g) {
  TraceScope _ts("uncenter_p");
  if (np < 0) VPB_ERROR("Bad number of particles");
  if (!f0) VPB_ERROR("Bad interpolator");
  if (!g) VPB_ERROR("Bad grid");
  vpb_domain_t *dom = domain_of(g);
  Residency r;
  vpb_particle_t *dp = (vpb_particle_t *)r.get_particles(dom, p0, np, RW);
  const vpb_interpolator_t *df = (const vpb_interpolator_t *)r.get(f0, nvox(g) * sizeof(*f0), RD);
  vpb_uncenter_p(dom, dp, np, q_m, df);
  r.finish();
}

double energy_p(const vpb_particle_t *p0, int np, float q_m, const vpb_interpolator_t *f0, const vpb_grid_t *g) {
  TraceScope _ts("energy_p");
  if (np < 0) VPB_ERROR("Bad number of particles");
  if (!f0) VPB_ERROR("Bad interpolator");
  if (!g) VPB_ERROR("Bad grid");
  vpb_domain_t *dom = domain_of(g);
  Context &c = ctx();
  Residency r;
  const vpb_particle_t *dp = (const vpb_particle_t *)r.get_particles(dom, p0, np, RD);
  const vpb_interpolator_t *df = (const vpb_interpolator_t *)r.get(f0, nvox(g) * sizeof(*f0), RD);
  double *d_en = nullptr;
  VPB_CUDA(cudaMallocAsync(&d_en, sizeof(double), c.stream));
  vpb_energy_p(dom, dp, np, q_m, df, d_en);
  vpb_comm_allsum_d(d_en, 1);   // mp_allsum_d (energy_p.cxx:155)
  VPB_CUDA(cudaMemcpyAsync(c.h_pinned_d, d_en, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  VPB_CUDA(cudaFreeAsync(d_en, c.stream));
  r.finish();
  return (double)g->cvac * (double)g->cvac * c.h_pinned_d[0] / (double)q_m;   // energy_p.cxx:156
}

void accumulate_rho_p(vpb_field_t *f, const vpb_particle_t *p0, int np, const vpb_grid_t *g) {
  TraceScope _ts("accumulate_rho_p");
  if (!f) VPB_ERROR("Bad field");
  if (!p0) VPB_ERROR("Bad particle array");
  if (np < 0) VPB_ERROR("Bad number of particles");
  if (!g) VPB_ERROR("Bad grid");
  vpb_domain_t *dom = domain_of(g);
  Residency r;
  vpb_field_t *df = (vpb_field_t *)r.get_field(dom, f, nvox(g) * sizeof(*f), RW);
  const vpb_particle_t *dp = (const vpb_particle_t *)r.get_particles(dom, p0, np, RD);
  vpb_accumulate_rho_p(dom, df, dp, np);
  r.finish();
}

// move_p.c:20-136 for ONE mover.  Host code of the reference calls this per injected particle
// (misc.cxx:102); it costs a kernel launch here and is meant for set-up paths only.
int move_p(vpb_particle_t *p0, vpb_particle_mover_t *m, vpb_accumulator_t *a0, const vpb_grid_t *g) {
  TraceScope _ts("move_p");
  vpb_domain_t *dom = domain_of(g);
  Context &c = ctx();
  Residency r;
  // only the one particle the mover names is touched; stage just that record when p0 is host memory
  vpb_particle_mover_t hm = *m;
  vpb_particle_t *pk = p0 + hm.i;
  cudaPointerAttributes at;
  bool on_device = cudaPointerGetAttributes(&at, p0) == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
  cudaGetLastError();
  vpb_particle_t *dp = on_device ? p0 : ((vpb_particle_t *)r.get(pk, sizeof(*pk), RW)) - hm.i;
  vpb_accumulator_t *da = (vpb_accumulator_t *)r.get(a0, nvox(g) * sizeof(*a0), RW);
  char *d_tmp = nullptr;
  VPB_CUDA(cudaMallocAsync(&d_tmp, 64, c.stream));
  VPB_CUDA(cudaMemcpyAsync(d_tmp, &hm, sizeof(hm), cudaMemcpyHostToDevice, c.stream));
  vpb_move_p_one(dom, dp, (vpb_particle_mover_t *)d_tmp, da, (int *)(d_tmp + 32));
  VPB_CUDA(cudaMemcpyAsync(c.h_pinned_i, d_tmp, 48, cudaMemcpyDeviceToHost, c.stream));
  VPB_CUDA(cudaFreeAsync(d_tmp, c.stream));
  r.finish();
  memcpy(m, c.h_pinned_i, sizeof(*m));
  return c.h_pinned_i[8];
}

// boundary_p.c:9-71 for ONE particle
void accumulate_rhob(vpb_field_t *f0, const vpb_particle_t *p, const vpb_grid_t *g) {
  TraceScope _ts("accumulate_rhob");
  if (!f0 || !p) VPB_ERROR("Bad field or particle");
  vpb_domain_t *dom = domain_of(g);
  Context &c = ctx();
  Residency r;
  vpb_field_t *df = (vpb_field_t *)r.get_field(dom, f0, nvox(g) * sizeof(*f0), RW);
  vpb_particle_t hp = *p, *d_one = nullptr;
  VPB_CUDA(cudaMallocAsync(&d_one, sizeof(hp), c.stream));
  VPB_CUDA(cudaMemcpyAsync(d_one, &hp, sizeof(hp), cudaMemcpyHostToDevice, c.stream));
  vpb_accumulate_rhob_one(dom, df, d_one);
  VPB_CUDA(cudaFreeAsync(d_one, c.stream));
  r.finish();
}

// boundary_p.c:416-447: "Resizing local %s particle storage": n + n/4 + n/16, copy, free the old array.  Only arrays that
// came from util_malloc_aligned (managed memory) can be replaced this way; anything else keeps the overflow error.
static bool is_managed(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeManaged;
}

static int grow_species(void *user, int index, int need_np, int need_nm, vpb_species_state_t *st) {
  vpb_species_t *sp = (*(std::vector<vpb_species_t *> *)user)[index];
  Context &c = ctx();
  if (need_np > sp->max_np) {
    if (!is_managed(sp->p)) return 0;
    int n = need_np;
    n = n + (n >> 2) + (n >> 4);
    VPB_WARNING("Resizing local %s particle storage from %i to %i", sp->name, sp->max_np, n);
    vpb_particle_t *new_p = nullptr;
    util_malloc_aligned("MALLOC_ALIGNED( new_p, (%lu bytes), 128 (%lu bytes) ) failed", &new_p, (size_t)n * sizeof(*new_p), 128);
    VPB_CUDA(cudaMemcpyAsync(new_p, st->p, (size_t)st->np * sizeof(*new_p), cudaMemcpyDefault, c.stream));
    VPB_CUDA(cudaStreamSynchronize(c.stream));
    auto ph = g_part_hint.find(sp->p);          // the traversal hint belongs to the old array
    if (ph != g_part_hint.end()) {
      if (ph->second.dev) cudaFree(ph->second.dev);
      g_part_hint.erase(ph);
    }
    util_free_aligned(&sp->p);
    sp->p = new_p; sp->max_np = n;
    st->p = new_p; st->max_np = n;
  }
  if (need_nm > sp->max_nm) {
    if (!is_managed(sp->pm)) return 0;
    int n = need_nm;
    n = n + (n >> 2) + (n >> 4);
    VPB_WARNING("Resizing local %s mover storage from %i to %i", sp->name, sp->max_nm, n);
    vpb_particle_mover_t *new_pm = nullptr;
    util_malloc_aligned("MALLOC_ALIGNED( new_pm, (%lu bytes), 128 (%lu bytes) ) failed", &new_pm, (size_t)n * sizeof(*new_pm), 128);
    util_free_aligned(&sp->pm);      // no movers are pending at this point of the round (boundary_p.c:323-324)
    sp->pm = new_pm; sp->max_nm = n;
    st->pm = new_pm; st->max_nm = n;
  }
  return 1;
}

// boundary_p.c:77-505: one round over the species list.  rng belongs to the deck's custom boundary
// handlers (host callbacks), which are called from here (below).
void boundary_p(vpb_species_t *sp_list, vpb_field_t *f0, vpb_accumulator_t *a0, const vpb_grid_t *g, void *rng) {
  TraceScope _ts("boundary_p");
  if (!g) VPB_ERROR("Bad grid");
  vpb_domain_t *dom = domain_of(g);
  Residency r;
  std::vector<vpb_species_t *> list;
  for (vpb_species_t *sp = sp_list; sp; sp = sp->next) list.push_back(sp);
  if (list.size() > 7) VPB_ERROR("boundary_p handles at most 7 species per call");
  {
    // nothing pending anywhere and no face shared with another rank: the reference's loops all run zero times
    // (a one-rank periodic deck calls this three times a step)
    long pending = 0;
    for (vpb_species_t *sp : list) pending += sp->nm;
    bool remote = false;
    static const int fb[6] = {VPB_BOUNDARY(-1, 0, 0), VPB_BOUNDARY(0, -1, 0), VPB_BOUNDARY(0, 0, -1),
                              VPB_BOUNDARY(1, 0, 0),  VPB_BOUNDARY(0, 1, 0),  VPB_BOUNDARY(0, 0, 1)};
    for (int f = 0; f < 6; f++) remote |= g->bc[fb[f]] >= 0 && g->bc[fb[f]] < dom->d.nproc && g->bc[fb[f]] != dom->d.rank;
    if (pending == 0 && !remote) {
      for (vpb_species_t *sp : list)
        if (sp->id < 0 || sp->id >= 64) VPB_ERROR("Invalid sp->id");   // boundary_p.c:396
      return;
    }
  }
  // boundary_p.c:271-277: movers that ended on a face bound to one of the deck's custom handlers (neighbor code -3-k,
  // k < grid->nb).  The handlers are the host program's code (parameters, host RNG, calls back into accumulate_rhob ...):
  // they run HERE, on the host, in the order the reference visits the movers -- species in list order, each species'
  // movers from the last to the first, faces tested in the order of the reference's TEST_FACE sequence -- so that the
  // handlers' random-number draws and their injector list (cmlist) come out the same.  The device then destroys those
  // particles without the rhob deposit of an absorption and injects the list after the received buffers.
  const bool with_handlers = g->nb > 0 && g->boundary != nullptr && dom->n_handler_faces > 0;
  std::vector<vpb_particle_injector_t> cm;
  int ncm = 0;
  if (with_handlers) {
    if (!g->neighbor) VPB_ERROR("Bad grid");
    VPB_CUDA(cudaStreamSynchronize(ctx().stream));
    size_t total = 0;
    for (vpb_species_t *sp : list) total += (size_t)(sp->nm > 0 ? sp->nm : 0);
    cm.resize(total + 1);
    vpb_particle_injector_t *cmp = cm.data();
    const vpb_boundary_t *bt = (const vpb_boundary_t *)g->boundary;
    const int64_t rangem = dom->range[dom->d.nproc];
    for (vpb_species_t *sp : list) {
      for (int k = sp->nm - 1; k >= 0; k--) {
        vpb_particle_t *pr = sp->p + sp->pm[k].i;
        const float pos[3] = {pr->dx, pr->dy, pr->dz}, u[3] = {pr->ux, pr->uy, pr->uz};
        for (int face = 0; face < 6; face++) {
          const int ax = face % 3;
          const bool hit = face < 3 ? (pos[ax] == -1.f && u[ax] < 0) : (pos[ax] == 1.f && u[ax] > 0);
          if (!hit) continue;
          const int64_t nn = g->neighbor[6 * (size_t)pr->i + face];
          if (nn == vpb_absorb_particles) break;                                                    // :207-212
          if ((nn >= 0 && nn < g->rangel) || (nn > g->rangeh && nn <= rangem)) break;               // :218-267
          const int64_t h = -nn - 3;
          if (h >= 0 && h < (int64_t)g->nb) {                                                       // :272-277
            bt[h].handler(const_cast<char *>(bt[h].params), pr, sp->pm + k, f0, a0, g, sp, &cmp, rng, face);
            break;
          }
        }
      }
    }
    ncm = (int)(cmp - cm.data());
    if ((size_t)ncm > total) VPB_ERROR("custom boundary handlers made %d injectors for %zu movers", ncm, total);
  }
  std::vector<vpb_species_state_t> st(list.size());
  for (size_t s = 0; s < list.size(); s++) {
    vpb_species_t *sp = list[s];
    if (sp->id < 0 || sp->id >= 64) VPB_ERROR("Invalid sp->id");   // boundary_p.c:396
    st[s].p = (vpb_particle_t *)r.get(sp->p, (size_t)sp->max_np * sizeof(vpb_particle_t), RW);
    st[s].pm = (vpb_particle_mover_t *)r.get(sp->pm, (size_t)sp->max_nm * sizeof(vpb_particle_mover_t), RW);
    st[s].np = sp->np; st[s].max_np = sp->max_np; st[s].nm = sp->nm; st[s].max_nm = sp->max_nm; st[s].id = sp->id;
  }
  vpb_field_t *df = f0 ? (vpb_field_t *)r.get_field(dom, f0, nvox(g) * sizeof(*f0), RW) : nullptr;
  vpb_accumulator_t *da = a0 ? (vpb_accumulator_t *)r.get(a0, nvox(g) * sizeof(*a0), RW) : nullptr;
  vpb_boundary_set_grow_hook(grow_species, &list);
  if (with_handlers) vpb_boundary_set_local_injectors(cm.data(), ncm);
  vpb_boundary_p(dom, st.data(), (int)st.size(), df, da);
  if (with_handlers) vpb_boundary_set_local_injectors(nullptr, -1);
  vpb_boundary_set_grow_hook(nullptr, nullptr);
  for (size_t s = 0; s < list.size(); s++) { list[s]->np = st[s].np; list[s]->nm = st[s].nm; }
  r.finish();
}

// One spare particle array per capacity: what an out-of-place sort writes into.  The array it read from becomes the
// next spare (the reference allocates a new array and frees the old one on every such sort, sort_p.c:69-77; managed
// allocations and frees synchronise the device, so the pair is recycled instead).  Species of equal capacity share it.
static std::unordered_map<size_t, vpb_particle_t *> g_sort_spare;

void sort_p(vpb_species_t *sp, const vpb_grid_t *g) {
  TraceScope _ts("sort_p");
  if (!sp) VPB_ERROR("Bad species");
  if (!g) VPB_ERROR("Bad grid");
  vpb_domain_t *dom = domain_of(g);
  Context &c = ctx();
  const size_t nv1 = nvox(g) + 1;
  if (!sp->partition) util_malloc_aligned("MALLOC_ALIGNED( sp->partition, (%lu bytes), 128 (%lu bytes) ) failed", &sp->partition,
                                          nv1 * sizeof(int), 128);   // sort_p.c:32
  if (sp->np == 0) return;   // sort_p.c:35
  Residency r;
  int *dpart = (int *)r.get(sp->partition, nv1 * sizeof(int), WR);
  const vpb_particle_t *old_p = sp->p;
  if (sp->sort_out_of_place && is_managed(sp->p)) {
    // sort_p.c:63-77: a new array of max_np records receives the particles in order and replaces sp->p
    const size_t cap = (size_t)sp->max_np * sizeof(vpb_particle_t);
    vpb_particle_t *new_p = nullptr;
    auto sparep = g_sort_spare.find(cap);
    if (sparep != g_sort_spare.end() && sparep->second) { new_p = sparep->second; sparep->second = nullptr; }
    else util_malloc_aligned("MALLOC_ALIGNED( new_p, (%lu bytes), 128 (%lu bytes) ) failed", &new_p, cap, 128);
    vpb_particle_t *dp = (vpb_particle_t *)r.get(sp->p, (size_t)sp->np * sizeof(vpb_particle_t), RD);
    vpb_particle_t *dn = (vpb_particle_t *)r.get(new_p, (size_t)sp->np * sizeof(vpb_particle_t), WR);
    vpb_sort_p(dom, dp, dn, sp->np, dpart);
    g_sort_spare[cap] = sp->p;
    sp->p = new_p;
  } else {
    // in place (the reference's cycle sort, sort_p.c:79-101, leaves the particles grouped by voxel in the SAME array; so
    // does this: a copy goes to scratch and the stable sort writes back into the caller's array), or an array this
    // library cannot replace (plain host memory: staged)
    vpb_particle_t *dp = (vpb_particle_t *)r.get(sp->p, (size_t)sp->np * sizeof(vpb_particle_t), RW);
    vpb_particle_t *tmp = nullptr;
    VPB_CUDA(cudaMallocAsync(&tmp, (size_t)sp->np * sizeof(vpb_particle_t), c.stream));
    VPB_CUDA(cudaMemcpyAsync(tmp, dp, (size_t)sp->np * sizeof(vpb_particle_t), cudaMemcpyDeviceToDevice, c.stream));
    vpb_sort_p(dom, tmp, dp, sp->np, dpart);
    VPB_CUDA(cudaFreeAsync(tmp, c.stream));
  }
  // remember the layout for advance_p's traversal (keyed by the caller's array)
  if (old_p != sp->p) {
    auto stale = g_part_hint.find(old_p);
    if (stale != g_part_hint.end()) { g_part_hint[sp->p] = stale->second; g_part_hint.erase(stale); }
  }
  PartHint &h = g_part_hint[sp->p];
  if (h.n != nv1) {
    if (h.dev) cudaFree(h.dev);
    VPB_CUDA(cudaMalloc(&h.dev, nv1 * sizeof(int)));
    h.n = nv1;
  }
  VPB_CUDA(cudaMemcpyAsync(h.dev, dpart, nv1 * sizeof(int), cudaMemcpyDeviceToDevice, c.stream));
  h.np = sp->np;
  r.finish();
}

// ---------------------------------------------------------------------------
// sf_interface (sf_interface.h)
// ---------------------------------------------------------------------------
vpb_interpolator_t *new_interpolator(vpb_grid_t *g) {
  if (!g) VPB_ERROR("Invalid grid.");
  if (g->nx < 1 || g->ny < 1 || g->nz < 1) VPB_ERROR("Invalid grid resolution.");
  ensure_comm(g);
  return (vpb_interpolator_t *)vpb_malloc_managed(nvox(g) * sizeof(vpb_interpolator_t));
}
void delete_interpolator(vpb_interpolator_t *fi) { util_free_aligned(&fi); }

// One replica: the device needs no per-pipeline copies (sf_interface.c:65-72 sizes 1+n_pipeline of them)
vpb_accumulator_t *new_accumulators(vpb_grid_t *g) {
  if (!g) VPB_ERROR("Bad grid.");
  if (g->nx < 1 || g->ny < 1 || g->nz < 1) VPB_ERROR("Bad resolution.");
  ensure_comm(g);
  return (vpb_accumulator_t *)vpb_malloc_managed(((nvox(g) + 1) & ~(size_t)1) * sizeof(vpb_accumulator_t));
}
void delete_accumulators(vpb_accumulator_t *a) { util_free_aligned(&a); }

// src/field_advance/field_advance.c:3-28: bind a grid, a material list and a method table
vpb_field_advance_t *new_field_advance(vpb_grid_t *g, vpb_material_t *m_list, vpb_field_advance_methods_t *fam) {
  if (!g || !m_list || !fam) VPB_ERROR("Bad args");
  ensure_comm(g);
  vpb_field_advance_t *fa = (vpb_field_advance_t *)calloc(1, sizeof(*fa));
  if (!fa) VPB_ERROR("Could not allocate field_advance_t");
  fa->method[0] = fam[0];
  fa->f = fa->method->new_field(g);
  fa->m = fa->method->new_material_coefficients(g, m_list);
  fa->g = g;
  return fa;
}

void delete_field_advance(vpb_field_advance_t *fa) {
  if (!fa) return;   // do-nothing request
  fa->method->delete_material_coefficients(fa->m);
  fa->method->delete_field(fa->f);
  free(fa);
}

// src/sf_interface/sf_structors.c: hydro_t[nv], zero-filled
vpb_hydro_t *new_hydro(vpb_grid_t *g) {
  if (!g) VPB_ERROR("Bad grid.");
  if (g->nx < 1 || g->ny < 1 || g->nz < 1) VPB_ERROR("Bad resolution.");
  return (vpb_hydro_t *)vpb_malloc_managed(nvox(g) * sizeof(vpb_hydro_t));
}
void delete_hydro(vpb_hydro_t *h) { util_free_aligned(&h); }

void clear_hydro(vpb_hydro_t *h, const vpb_grid_t *g) {
  TraceScope _ts("clear_hydro");
  if (!h) VPB_ERROR("Bad hydro");
  if (!g) VPB_ERROR("Bad grid");
  vpb_domain_t *dom = domain_of(g);
  Residency r;
  vpb_hydro_t *dh = (vpb_hydro_t *)r.get(h, nvox(g) * sizeof(*h), WR);
  vpb_clear_hydro(dom, dh);
  r.finish();
}

// hydro_p.c:24-161
void accumulate_hydro_p(vpb_hydro_t *h0, const vpb_particle_t *p0, int n, float q_m, const vpb_interpolator_t *f0,
                        const vpb_grid_t *g) {
  TraceScope _ts("accumulate_hydro_p");
  if (!h0) VPB_ERROR("Bad hydro");
  if (!p0) VPB_ERROR("Bad particle array");
  if (n < 0) VPB_ERROR("Bad number of particles");
  if (!f0) VPB_ERROR("Bad field");
  if (!g) VPB_ERROR("Bad grid");
  vpb_domain_t *dom = domain_of(g);
  Residency r;
  vpb_hydro_t *dh = (vpb_hydro_t *)r.get(h0, nvox(g) * sizeof(*h0), RW);
  const vpb_particle_t *dp = (const vpb_particle_t *)r.get_particles(dom, p0, n, RD);
  const vpb_interpolator_t *df = (const vpb_interpolator_t *)r.get(f0, nvox(g) * sizeof(*f0), RD);
  vpb_accumulate_hydro_p(dom, dh, dp, n, q_m, df);
  r.finish();
}

// sf_interface/hydro.c:30-141, :146-184
void synchronize_hydro(vpb_hydro_t *h, const vpb_grid_t *g) {
  TraceScope _ts("synchronize_hydro");
  if (!h) VPB_ERROR("Bad hydro");
  if (!g) VPB_ERROR("Bad grid");
  vpb_domain_t *dom = domain_of(g);
  Residency r;
  vpb_hydro_t *dh = (vpb_hydro_t *)r.get(h, nvox(g) * sizeof(*h), RW);
  vpb_synchronize_hydro(dom, dh);
  r.finish();
}

void local_adjust_hydro(vpb_hydro_t *h, const vpb_grid_t *g) {
  TraceScope _ts("local_adjust_hydro");
  if (!h) VPB_ERROR("Bad hydro");
  if (!g) VPB_ERROR("Bad grid");
  vpb_domain_t *dom = domain_of(g);
  Residency r;
  vpb_hydro_t *dh = (vpb_hydro_t *)r.get(h, nvox(g) * sizeof(*h), RW);
  vpb_local_adjust_hydro(dom, dh);
  r.finish();
}

void load_interpolator(vpb_interpolator_t *fi, const vpb_field_t *f, const vpb_grid_t *g) {
  TraceScope _ts("load_interpolator");
  if (!fi) VPB_ERROR("Bad interpolator");
  if (!f) VPB_ERROR("Bad field");
  if (!g) VPB_ERROR("Bad grid");
  vpb_domain_t *dom = domain_of(g);
  Residency r;
  // RW, not WR: ghost voxels and _pad of the caller's array are left as they were
  vpb_interpolator_t *dfi = (vpb_interpolator_t *)r.get(fi, nvox(g) * sizeof(*fi), RW);
  const vpb_field_t *df = (const vpb_field_t *)r.get_field(dom, f, nvox(g) * sizeof(*f), RD);
  vpb_load_interpolator(dom, dfi, df);
  r.finish();
}

void clear_accumulators(vpb_accumulator_t *a, const vpb_grid_t *g) {
  TraceScope _ts("clear_accumulators");
  if (!a) VPB_ERROR("Bad accumulator");
  if (!g) VPB_ERROR("Bad grid");
  vpb_domain_t *dom = domain_of(g);
  Residency r;
  vpb_accumulator_t *da = (vpb_accumulator_t *)r.get(a, nvox(g) * sizeof(*a), WR);
  vpb_clear_accumulators(dom, da);
  r.finish();
}

// the device accumulates into one replica; nothing to reduce (reduce_accumulators.cxx:144-165)
void reduce_accumulators(vpb_accumulator_t *a, const vpb_grid_t *g) {
  if (!a) VPB_ERROR("Bad accumulator");
  if (!g) VPB_ERROR("Bad grid");
}

void unload_accumulator(vpb_field_t *f, const vpb_accumulator_t *a, const vpb_grid_t *g) {
  TraceScope _ts("unload_accumulator");
  if (!f) VPB_ERROR("Bad field");
  if (!a) VPB_ERROR("Bad accumulator");
  if (!g) VPB_ERROR("Bad grid");
  vpb_domain_t *dom = domain_of(g);
  Residency r;
  vpb_field_t *df = (vpb_field_t *)r.get_field(dom, f, nvox(g) * sizeof(*f), RW);
  const vpb_accumulator_t *da = (const vpb_accumulator_t *)r.get(a, nvox(g) * sizeof(*a), RD);
  vpb_unload_accumulator(dom, df, da);
  r.finish();
}

// ---------------------------------------------------------------------------
// field advance vtables (field_advance.h:185-302, sfa.c:4-52, vacuum/vfa.c)
// ---------------------------------------------------------------------------
static vpb_field_t *fa_new_field(vpb_grid_t *g) {
  if (!g) VPB_ERROR("Bad grid.");
  if (g->nx < 1 || g->ny < 1 || g->nz < 1) VPB_ERROR("Bad resolution.");
  ensure_comm(g);
  return (vpb_field_t *)vpb_malloc_managed(nvox(g) * sizeof(vpb_field_t));
}
static void fa_delete_field(vpb_field_t *f) { util_free_aligned(&f); }

static float minf(float a, float b) { return a < b ? a : b; }

// sfa.c:80-171: host-side set-up, restated (same expressions and types)
static vpb_material_coefficient_t *fa_new_material_coefficients(vpb_grid_t *g, vpb_material_t *m_list) {
  if (!g) VPB_ERROR("Invalid grid.");
  if (!m_list) VPB_ERROR("Empty material list.");
  float ax = g->nx > 1 ? g->cvac * g->dt * g->rdx : 0; ax *= ax;
  float ay = g->ny > 1 ? g->cvac * g->dt * g->rdy : 0; ay *= ay;
  float az = g->nz > 1 ? g->cvac * g->dt * g->rdz : 0; az *= az;
  int n_mat = 0;
  for (const vpb_material_t *m = m_list; m; m = m->next) {
    const float cg2 = ax / minf(m->epsy * m->muz, m->epsz * m->muy) + ay / minf(m->epsz * m->mux, m->epsx * m->muz) +
                      az / minf(m->epsx * m->muy, m->epsy * m->mux);
    if (cg2 >= 1) VPB_WARNING("Material \"%s\" Courant condition estimate = %e", m->name, sqrt(cg2));
    if (m->zetax != 0 || m->zetay != 0 || m->zetaz != 0)
      VPB_WARNING("Standard field advance does not support magnetic conductivity yet.");
    n_mat++;
  }
  vpb_material_coefficient_t *mc0 = (vpb_material_coefficient_t *)vpb_malloc_managed((size_t)n_mat * sizeof(*mc0));
  for (const vpb_material_t *m = m_list; m; m = m->next) {
    vpb_material_coefficient_t *mc = mc0 + m->id;
    const float eps[3] = {m->epsx, m->epsy, m->epsz}, sig[3] = {m->sigmax, m->sigmay, m->sigmaz};
    float a[3], *dd = &mc->decayx;
    for (int X = 0; X < 3; X++) {
      a[X] = (sig[X] * g->dt) / (eps[X] * g->eps0);
      dd[2 * X] = exp(-a[X]);
      if (a[X] == 0) dd[2 * X + 1] = 1. / eps[X];
      else if (dd[2 * X] == 0) dd[2 * X + 1] = 0;
      else dd[2 * X + 1] = 2. * exp(-0.5 * a[X]) * sinh(0.5 * a[X]) / (a[X] * eps[X]);
    }
    mc->rmux = 1. / m->mux; mc->rmuy = 1. / m->muy; mc->rmuz = 1. / m->muz;
    mc->nonconductive = (a[0] == 0 && a[1] == 0 && a[2] == 0) ? 1. : 0.;
    mc->epsx = m->epsx; mc->epsy = m->epsy; mc->epsz = m->epsz;
  }
  g_nmat[mc0] = n_mat;
  return mc0;
}
static void fa_delete_material_coefficients(vpb_material_coefficient_t *mc) {
  g_nmat.erase(mc);
  util_free_aligned(&mc);
}

// vacuum/vfa.c:56-79: accepts only trivial materials, keeps no coefficients
static vpb_material_coefficient_t *vfa_new_material_coefficients(vpb_grid_t *g, vpb_material_t *m_list) {
  if (!g) VPB_ERROR("Invalid grid.");
  if (g->damp != 0) VPB_ERROR("Vacuum field advance does not support TCA radiation damping");
  for (const vpb_material_t *m = m_list; m; m = m->next)
    if (m->epsx != 1 || m->epsy != 1 || m->epsz != 1 || m->mux != 1 || m->muy != 1 || m->muz != 1 || m->sigmax != 0 ||
        m->sigmay != 0 || m->sigmaz != 0 || m->zetax != 0 || m->zetay != 0 || m->zetaz != 0)
      VPB_ERROR("Material %s is not supported by vacuum (hint) field advance", m->name);
  return nullptr;
}
static void vfa_delete_material_coefficients(vpb_material_coefficient_t *mc) { (void)mc; }

struct FieldCall {
  vpb_domain_t *dom;
  Residency r;
  vpb_field_t *df;
  const vpb_material_coefficient_t *dm = nullptr;
  int n_mat = 1;
  FieldCall(vpb_field_t *f, const vpb_grid_t *g, int mode, const vpb_material_coefficient_t *m = nullptr, bool need_m = false) {
    if (!f) VPB_ERROR("Bad field");
    if (need_m && !m) VPB_ERROR("Bad material coefficients");
    if (!g) VPB_ERROR("Bad grid");
    dom = domain_of(g);
    df = (vpb_field_t *)r.get_field(dom, f, nvox(g) * sizeof(*f), mode);
    if (m) {
      n_mat = nmat_of(m);
      dm = (const vpb_material_coefficient_t *)r.get(m, (size_t)n_mat * sizeof(*m), RD);
    }
  }
};

static double read_back(double *d, int n, double *out) {
  Context &c = ctx();
  VPB_CUDA(cudaMemcpyAsync(c.h_pinned_d, d, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  VPB_CUDA(cudaStreamSynchronize(c.stream));
  for (int i = 0; i < n; i++) out[i] = c.h_pinned_d[i];
  return out[0];
}
static double *dev_doubles(int n) {
  double *d = nullptr;
  VPB_CUDA(cudaMallocAsync(&d, n * sizeof(double), ctx().stream));
  return d;
}
static void free_doubles(double *d) { VPB_CUDA(cudaFreeAsync(d, ctx().stream)); }

static void fa_advance_b(vpb_field_t *f, const vpb_grid_t *g, float frac) {
  TraceScope _ts("advance_b");
  FieldCall k(f, g, RW);
  vpb_advance_b(k.dom, k.df, frac);
  k.r.finish();
}
static void fa_advance_e(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  TraceScope _ts("advance_e");
  FieldCall k(f, g, RW, m, true);
  vpb_advance_e(k.dom, k.df, k.dm, k.n_mat, 0);
  k.r.finish();
}
static void vfa_advance_e(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  TraceScope _ts("vadvance_e");
  (void)m;
  FieldCall k(f, g, RW);
  vpb_advance_e(k.dom, k.df, nullptr, 1, 1);
  k.r.finish();
}
static void energy_common(double *en, const vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g, bool need_m) {
  TraceScope _ts("energy");
  if (!en) VPB_ERROR("Bad energy");
  FieldCall k(const_cast<vpb_field_t *>(f), g, RD, m, need_m);
  double *d = dev_doubles(6);
  vpb_energy_f(k.dom, k.df, k.dm, k.n_mat, d);
  vpb_comm_allsum_d(d, 6);
  double loc[6];
  read_back(d, 6, loc);
  free_doubles(d);
  const double v0 = 0.5 * g->eps0 * g->dx * g->dy * g->dz;   // energy_f.c:170
  for (int i = 0; i < 6; i++) en[i] = loc[i] * v0;
  k.r.finish();
}
static void fa_energy_f(double *en, const vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  energy_common(en, f, m, g, true);
}
static void vfa_energy_f(double *en, const vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  (void)m;
  energy_common(en, f, nullptr, g, false);
}
static void fa_clear_jf(vpb_field_t *f, const vpb_grid_t *g) { TraceScope _ts("clear_jf"); FieldCall k(f, g, RW); vpb_clear_jf(k.dom, k.df); k.r.finish(); }
static void fa_synchronize_jf(vpb_field_t *f, const vpb_grid_t *g) { TraceScope _ts("synchronize_jf"); FieldCall k(f, g, RW); vpb_synchronize_jf(k.dom, k.df); k.r.finish(); }
static void fa_clear_rhof(vpb_field_t *f, const vpb_grid_t *g) { TraceScope _ts("clear_rhof"); FieldCall k(f, g, RW); vpb_clear_rhof(k.dom, k.df); k.r.finish(); }
static void fa_synchronize_rho(vpb_field_t *f, const vpb_grid_t *g) { TraceScope _ts("synchronize_rho"); FieldCall k(f, g, RW); vpb_synchronize_rho(k.dom, k.df); k.r.finish(); }
static void fa_compute_rhob(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  TraceScope _ts("compute_rhob");
  FieldCall k(f, g, RW, m, true); vpb_compute_rhob(k.dom, k.df, k.dm, k.n_mat); k.r.finish();
}
static void vfa_compute_rhob(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  TraceScope _ts("compute_rhob");
  (void)m; FieldCall k(f, g, RW); vpb_compute_rhob(k.dom, k.df, nullptr, 1); k.r.finish();
}
static void fa_compute_curl_b(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  TraceScope _ts("compute_curl_b");
  FieldCall k(f, g, RW, m, true); vpb_compute_curl_b(k.dom, k.df, k.dm, k.n_mat); k.r.finish();
}
static void vfa_compute_curl_b(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  TraceScope _ts("compute_curl_b");
  (void)m; FieldCall k(f, g, RW); vpb_compute_curl_b(k.dom, k.df, nullptr, 1); k.r.finish();
}
static double fa_synchronize_tang_e_norm_b(vpb_field_t *f, const vpb_grid_t *g) {
  TraceScope _ts("synchronize_tang_e_norm_b");
  FieldCall k(f, g, RW);
  double *d = dev_doubles(1), out[1];
  vpb_synchronize_tang_e_norm_b(k.dom, k.df, d);
  vpb_comm_allsum_d(d, 1);   // remote.c:412
  read_back(d, 1, out);
  free_doubles(d);
  k.r.finish();
  return out[0];
}
static void fa_compute_div_e_err(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  TraceScope _ts("compute_div_e_err");
  FieldCall k(f, g, RW, m, true); vpb_compute_div_e_err(k.dom, k.df, k.dm, k.n_mat); k.r.finish();
}
static void vfa_compute_div_e_err(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  TraceScope _ts("compute_div_e_err");
  (void)m; FieldCall k(f, g, RW); vpb_compute_div_e_err(k.dom, k.df, nullptr, 1); k.r.finish();
}
static double rms_common(vpb_field_t *f, const vpb_grid_t *g, int which) {
  TraceScope _ts("rms");
  FieldCall k(f, g, RD);
  double *d = dev_doubles(2), loc[2];
  if (which == 0) vpb_compute_rms_div_e_err(k.dom, k.df, d); else vpb_compute_rms_div_b_err(k.dom, k.df, d);
  read_back(d, 1, loc);
  // compute_rms_div_e_err.c:150-160: local[0]=err*dV, local[1]=nx*ny*nz*dV, allsum, eps0*sqrt(ratio)
  double local[2] = {loc[0] * g->dx * g->dy * g->dz, (double)(g->nx * g->ny * g->nz * g->dx * g->dy * g->dz)};
  Context &c = ctx();
  VPB_CUDA(cudaMemcpyAsync(d, local, sizeof(local), cudaMemcpyHostToDevice, c.stream));
  VPB_CUDA(cudaStreamSynchronize(c.stream));
  vpb_comm_allsum_d(d, 2);
  read_back(d, 2, loc);
  free_doubles(d);
  k.r.finish();
  return g->eps0 * sqrt(loc[0] / loc[1]);
}
static double fa_compute_rms_div_e_err(vpb_field_t *f, const vpb_grid_t *g) { return rms_common(f, g, 0); }
static double fa_compute_rms_div_b_err(vpb_field_t *f, const vpb_grid_t *g) { return rms_common(f, g, 1); }
static void fa_clean_div_e(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  TraceScope _ts("clean_div_e");
  FieldCall k(f, g, RW, m, true); vpb_clean_div_e(k.dom, k.df, k.dm, k.n_mat); k.r.finish();
}
static void vfa_clean_div_e(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  TraceScope _ts("clean_div_e");
  (void)m; FieldCall k(f, g, RW); vpb_clean_div_e(k.dom, k.df, nullptr, 1); k.r.finish();
}
static void fa_compute_div_b_err(vpb_field_t *f, const vpb_grid_t *g) { TraceScope _ts("compute_div_b_err"); FieldCall k(f, g, RW); vpb_compute_div_b_err(k.dom, k.df); k.r.finish(); }
static void fa_clean_div_b(vpb_field_t *f, const vpb_grid_t *g) { TraceScope _ts("clean_div_b"); FieldCall k(f, g, RW); vpb_clean_div_b(k.dom, k.df); k.r.finish(); }

#define VPB_STANDARD_TABLE                                                                                              \
  { { fa_new_field, fa_delete_field, fa_new_material_coefficients, fa_delete_material_coefficients, fa_advance_b,       \
      fa_advance_e, fa_energy_f, fa_clear_jf, fa_synchronize_jf, fa_clear_rhof, fa_synchronize_rho, fa_compute_rhob,    \
      fa_compute_curl_b, fa_synchronize_tang_e_norm_b, fa_compute_div_e_err, fa_compute_rms_div_e_err, fa_clean_div_e,  \
      fa_compute_div_b_err, fa_compute_rms_div_b_err, fa_clean_div_b } }
#define VPB_VACUUM_TABLE                                                                                                \
  { { fa_new_field, fa_delete_field, vfa_new_material_coefficients, vfa_delete_material_coefficients, fa_advance_b,     \
      vfa_advance_e, vfa_energy_f, fa_clear_jf, fa_synchronize_jf, fa_clear_rhof, fa_synchronize_rho, vfa_compute_rhob, \
      vfa_compute_curl_b, fa_synchronize_tang_e_norm_b, vfa_compute_div_e_err, fa_compute_rms_div_e_err,                \
      vfa_clean_div_e, fa_compute_div_b_err, fa_compute_rms_div_b_err, fa_clean_div_b } }

// The four names a deck can reach through standard_field_advance / vacuum_field_advance
// (field_advance.h:318-345); the "_v4" ones are what a V4 build of the deck names.
vpb_field_advance_methods_t _standard_field_advance[1] = VPB_STANDARD_TABLE;
vpb_field_advance_methods_t _standard_v4_field_advance[1] = VPB_STANDARD_TABLE;
vpb_field_advance_methods_t _vacuum_field_advance[1] = VPB_VACUUM_TABLE;
vpb_field_advance_methods_t _vacuum_v4_field_advance[1] = VPB_VACUUM_TABLE;

// For callers that cannot read a data symbol (ctypes): 0 standard, 1 vacuum, 2 standard_v4, 3 vacuum_v4
vpb_field_advance_methods_t *vpb_field_advance_table(int which) {
  switch (which) {
  case 0: return _standard_field_advance;
  case 1: return _vacuum_field_advance;
  case 2: return _standard_v4_field_advance;
  case 3: return _vacuum_v4_field_advance;
  }
  return nullptr;
}

// ---------------------------------------------------------------------------
// Deck-side particle diagnostics on the device (vpb_diag.cu), for host / managed arrays: what a deck calls INSTEAD of
// its host loops over sp->p (decks/trecon-part/energy.cxx:90-176, tracer.cxx:125-160).  Only the results cross PCIe.
// ---------------------------------------------------------------------------
void vpb_deck_energy_spectrum(const vpb_particle_t *p0, int np, double dke, int nex, float *dist, double eminp, double emaxp, int nbin,
                              float *edist, const vpb_grid_t *g) {
  TraceScope _ts("deck_energy_spectrum");
  if (!g) VPB_ERROR("Bad grid");
  vpb_domain_t *dom = domain_of(g);
  Residency r;
  const vpb_particle_t *dp = (const vpb_particle_t *)r.get(p0, (size_t)np * sizeof(*p0), RD);
  float *dd = dist ? (float *)r.get(dist, (size_t)nex * nvox(g) * sizeof(float), WR) : nullptr;
  float *de = edist ? (float *)r.get(edist, (size_t)nbin * sizeof(float), WR) : nullptr;
  vpb_energy_spectrum(dom, dp, np, dke, nex, dd, eminp, emaxp, nbin, de);
  r.finish();
}

void vpb_deck_tracer_records(const vpb_particle_t *p0, int np, const vpb_field_t *f, float *out13, int field_of_first, const vpb_grid_t *g) {
  TraceScope _ts("deck_tracer_records");
  if (!g) VPB_ERROR("Bad grid");
  if (np <= 0) return;
  vpb_domain_t *dom = domain_of(g);
  Residency r;
  const vpb_particle_t *dp = (const vpb_particle_t *)r.get(p0, (size_t)np * sizeof(*p0), RD);
  const vpb_field_t *df = (const vpb_field_t *)r.get(f, nvox(g) * sizeof(*f), RD);
  float *dout = (float *)r.get(out13, (size_t)np * 13 * sizeof(float), WR);
  vpb_tracer_records(dom, dp, np, df, g->x0, g->y0, g->z0, field_of_first, dout);
  r.finish();
}

// bytes moved by the staging path since the last call (bench.py e2e accounting)
void vpb_staging_bytes(size_t *h2d, size_t *d2h) { *h2d = g_h2d_total; *d2h = g_d2h_total; g_h2d_total = g_d2h_total = 0; }

}  // extern "C"
