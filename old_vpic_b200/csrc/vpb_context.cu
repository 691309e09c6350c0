// vpb_context.cu -- process/device context, memory plumbing and the device
// mirror of grid_t.  One process drives one GPU (one rank per GPU).
#include <map>
#include <string>
#include <string.h>
#include <time.h>
#include <vector>
#include "vpb_common.cuh"

namespace vpb {

static Context g_ctx;
static bool g_ready = false;
static std::map<std::string, int> g_tuning;

Context &ctx() {
  if (!g_ready) vpb_init(-1);
  return g_ctx;
}

void *scratch(size_t bytes) {
  Context &c = ctx();
  if (bytes > c.scratch_bytes) {
    // Stream-ordered: earlier kernels may still be reading the old block.
    if (c.scratch) VPB_CUDA(cudaFreeAsync(c.scratch, c.stream));
    size_t want = bytes + (bytes >> 2) + 4096;
    VPB_CUDA(cudaMallocAsync(&c.scratch, want, c.stream));
    c.scratch_bytes = want;
  }
  return c.scratch;
}

int tuning(const char *name, int dflt) {
  auto it = g_tuning.find(name);
  if (it != g_tuning.end()) return it->second;
  std::string env = std::string("VPB_") + name;
  for (auto &ch : env) ch = (ch == '.') ? '_' : (char)toupper(ch);
  const char *e = getenv(env.c_str());
  int v = e ? atoi(e) : dflt;
  g_tuning[name] = v;
  return v;
}

// ---- host wall-clock trace of entry points (vpb_common.cuh: TraceScope) ----
bool g_trace_on = false;
struct TraceRec { long calls = 0; double seconds = 0, worst = 0; };
static std::map<std::string, TraceRec> g_trace;
static std::vector<std::string> g_trace_order;
double trace_now() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
void trace_add(const char *label, double s) {
  auto it = g_trace.find(label);
  if (it == g_trace.end()) { g_trace_order.push_back(label); it = g_trace.emplace(label, TraceRec()).first; }
  it->second.calls++;
  it->second.seconds += s;
  if (s > it->second.worst) it->second.worst = s;
}
static void trace_report_to(FILE *fp) {
  fprintf(fp, "vpb trace: %-34s %10s %12s %12s %12s\n", "label", "calls", "total ms", "us/call", "worst ms");
  for (auto &name : g_trace_order) {
    const TraceRec &r = g_trace[name];
    fprintf(fp, "vpb trace: %-34s %10ld %12.3f %12.2f %12.3f\n", name.c_str(), r.calls, 1e3 * r.seconds,
            r.calls ? 1e6 * r.seconds / (double)r.calls : 0.0, 1e3 * r.worst);
  }
  fflush(fp);
}
static void trace_atexit() {
  if (!g_trace_on || g_trace.empty()) return;
  const char *fn = getenv("VPB_TRACE_FILE");
  FILE *fp = fn ? fopen(fn, "a") : nullptr;
  trace_report_to(fp ? fp : stderr);
  if (fp) fclose(fp);
}

struct ProfRec { cudaEvent_t a, b; int cls; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;       // pool of event pairs
static size_t g_prof_used = 0;

bool prof_enabled() { return g_prof_on; }

int prof_begin(int cls) {
  if (!g_prof_on) return -1;
  if (g_prof_used == g_prof.size()) {
    ProfRec r;
    VPB_CUDA(cudaEventCreate(&r.a));
    VPB_CUDA(cudaEventCreate(&r.b));
    g_prof.push_back(r);
  }
  const int idx = (int)g_prof_used++;
  g_prof[idx].cls = cls;
  VPB_CUDA(cudaEventRecord(g_prof[idx].a, ctx().stream));
  return idx;
}
void prof_end(int idx) {
  if (idx < 0 || !g_prof_on || (size_t)idx >= g_prof_used) return;
  VPB_CUDA(cudaEventRecord(g_prof[idx].b, ctx().stream));
}

}  // namespace vpb

using namespace vpb;

// individual durations (ms) of one kernel class in launch order; returns how many were written
extern "C" int vpb_prof_list(int cls, float *out_ms, int max) {
  VPB_CUDA(cudaStreamSynchronize(ctx().stream));
  int n = 0;
  for (size_t i = 0; i < g_prof_used && n < max; i++)
    if (g_prof[i].cls == cls) VPB_CUDA(cudaEventElapsedTime(&out_ms[n++], g_prof[i].a, g_prof[i].b));
  return n;
}

extern "C" void vpb_trace_enable(int on) { g_trace_on = on != 0; }
extern "C" void vpb_trace_reset(void) { g_trace.clear(); g_trace_order.clear(); }
extern "C" void vpb_trace_report(void) { trace_report_to(stderr); }
// seconds and calls of one label since the last reset (0 when the label never ran)
extern "C" double vpb_trace_get(const char *label, long *calls) {
  auto it = g_trace.find(label);
  if (calls) *calls = it == g_trace.end() ? 0 : it->second.calls;
  return it == g_trace.end() ? 0.0 : it->second.seconds;
}

extern "C" void vpb_prof_enable(int on) { g_prof_on = on != 0; g_prof_used = 0; }

// Sum of the recorded durations of one kernel class since the last collect of
// ANY class with reset!=0; synchronises the stream.
extern "C" void vpb_prof_collect(int cls, double *total_ms, int *count, int reset) {
  VPB_CUDA(cudaStreamSynchronize(ctx().stream));
  double t = 0;
  int n = 0;
  for (size_t i = 0; i < g_prof_used; i++)
    if (g_prof[i].cls == cls) {
      float ms = 0;
      VPB_CUDA(cudaEventElapsedTime(&ms, g_prof[i].a, g_prof[i].b));
      t += ms;
      n++;
    }
  if (total_ms) *total_ms = t;
  if (count) *count = n;
  if (reset) g_prof_used = 0;
}

extern "C" {

static int g_l2_fetch = 0;
extern "C" int vpb_l2_fetch_granularity(void) { return g_l2_fetch; }

int vpb_init(int device_ordinal) {
  if (g_ready) {
    if (device_ordinal >= 0 && device_ordinal != g_ctx.device)
      VPB_ERROR("vpb_init(%d) after the process was bound to device %d", device_ordinal, g_ctx.device);
    return 0;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    VPB_ERROR("no CUDA device is usable (%s); libvpic_b200 has no CPU fallback", cudaGetErrorString(e));
  if (device_ordinal < 0) {
    // the launcher's local rank: torchrun, Open MPI, MVAPICH2, Intel MPI / MPICH (hydra), Slurm, the test shim
    static const char *const vars[] = {"LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID",
                                       "SLURM_LOCALID", "VPIC_SHIM_RANK", nullptr};
    const char *lr = nullptr;
    for (int i = 0; vars[i] && !lr; i++) lr = getenv(vars[i]);
    device_ordinal = lr ? atoi(lr) % n : 0;
  }
  if (device_ordinal >= n) VPB_ERROR("device %d requested, %d present", device_ordinal, n);
  VPB_CUDA(cudaSetDevice(device_ordinal));
  cudaDeviceProp prop;
  VPB_CUDA(cudaGetDeviceProperties(&prop, device_ordinal));
  if (prop.major < 10)
    VPB_ERROR("device %d is sm_%d%d; libvpic_b200 is built for sm_100a only", device_ordinal, prop.major, prop.minor);
  g_ctx.device = device_ordinal;
  g_ctx.sm_count = prop.multiProcessorCount;
  VPB_CUDA(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
  for (int i = 0; i < 16; i++) {
    VPB_CUDA(cudaEventCreate(&g_ctx.ev_start[i]));
    VPB_CUDA(cudaEventCreate(&g_ctx.ev_stop[i]));
  }
  VPB_CUDA(cudaMallocHost(&g_ctx.h_pinned_i, 64 * sizeof(int)));
  VPB_CUDA(cudaMallocHost(&g_ctx.h_pinned_d, 64 * sizeof(double)));
  // L2 fetch granularity (bytes fetched from DRAM per missing sector: 32, 64 or 128).  The gathers of this library
  // (interpolator records, accumulator REDs, the sort's record gather) use a fraction of a 128-byte line per miss.
  {
    const int gran = tuning("l2.fetch_granularity", 0);
    if (gran == 32 || gran == 64 || gran == 128) {
      VPB_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran));
    }
    size_t got = 0;
    if (cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity) == cudaSuccess) g_l2_fetch = (int)got;
    cudaGetLastError();
  }
  // keep freed stream-ordered blocks in the pool instead of returning them to the OS
  cudaMemPool_t pool;
  VPB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device_ordinal));
  uint64_t thresh = UINT64_MAX;
  VPB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  g_ready = true;
  if (getenv("VPB_TRACE") && atoi(getenv("VPB_TRACE"))) { g_trace_on = true; atexit(trace_atexit); }
  return 0;
}

void vpb_shutdown(void) {
  if (!g_ready) return;
  cudaStreamSynchronize(g_ctx.stream);
  if (g_ctx.scratch) cudaFreeAsync(g_ctx.scratch, g_ctx.stream);
  cudaStreamSynchronize(g_ctx.stream);
  for (int i = 0; i < 16; i++) { cudaEventDestroy(g_ctx.ev_start[i]); cudaEventDestroy(g_ctx.ev_stop[i]); }
  cudaFreeHost(g_ctx.h_pinned_i);
  cudaFreeHost(g_ctx.h_pinned_d);
  cudaStreamDestroy(g_ctx.stream);
  g_ctx = Context();
  g_ready = false;
}

int vpb_device_sm_count(void) { return ctx().sm_count; }

void *vpb_dev_alloc(size_t bytes) {
  void *d = nullptr;
  if (bytes == 0) bytes = 16;
  VPB_CUDA(cudaMalloc(&d, bytes));
  VPB_CUDA(cudaMemsetAsync(d, 0, bytes, ctx().stream));
  return d;
}
void vpb_dev_free(void *d) {
  if (!d) return;
  VPB_CUDA(cudaStreamSynchronize(ctx().stream));
  VPB_CUDA(cudaFree(d));
}
void *vpb_malloc_managed(size_t bytes) {
  void *d = nullptr;
  if (bytes == 0) bytes = 16;
  VPB_CUDA(cudaMallocManaged(&d, bytes));
  VPB_CUDA(cudaMemAdvise(d, bytes, cudaMemAdviseSetPreferredLocation, ctx().device));
  VPB_CUDA(cudaMemsetAsync(d, 0, bytes, ctx().stream));
  VPB_CUDA(cudaStreamSynchronize(ctx().stream));
  return d;
}
void *vpb_host_alloc_pinned(size_t bytes) {
  void *h = nullptr;
  ctx();
  VPB_CUDA(cudaMallocHost(&h, bytes ? bytes : 16));
  return h;
}
void vpb_host_free_pinned(void *h) { if (h) VPB_CUDA(cudaFreeHost(h)); }
void vpb_h2d(void *d, const void *h, size_t bytes) { VPB_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx().stream)); }
void vpb_d2h(void *h, const void *d, size_t bytes) { VPB_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx().stream)); }
void vpb_d2d(void *dst, const void *src, size_t bytes) { VPB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx().stream)); }
void vpb_memset(void *d, int byte, size_t bytes) { VPB_CUDA(cudaMemsetAsync(d, byte, bytes, ctx().stream)); }
void vpb_sync(void) { VPB_CUDA(cudaStreamSynchronize(ctx().stream)); }
void *vpb_stream(void) { return (void *)ctx().stream; }

void vpb_timer_start(int s) { VPB_CUDA(cudaEventRecord(ctx().ev_start[s & 15], ctx().stream)); }
void vpb_timer_stop(int s) { VPB_CUDA(cudaEventRecord(ctx().ev_stop[s & 15], ctx().stream)); }
float vpb_timer_ms(int s) {
  float ms = 0;
  VPB_CUDA(cudaEventSynchronize(ctx().ev_stop[s & 15]));
  VPB_CUDA(cudaEventElapsedTime(&ms, ctx().ev_start[s & 15], ctx().ev_stop[s & 15]));
  return ms;
}

long vpb_launch_count(int reset) {
  long n = ctx().launches;
  if (reset) ctx().launches = 0;
  return n;
}

void vpb_set_tuning(const char *name, int value) { g_tuning[name] = value; }
// read-only: a knob that was never set or looked up reports 0 and is NOT pinned to 0 by asking
int vpb_get_tuning(const char *name) {
  auto it = g_tuning.find(name);
  if (it != g_tuning.end()) return it->second;
  std::string env = std::string("VPB_") + name;
  for (auto &ch : env) ch = (ch == '.') ? '_' : (char)toupper(ch);
  const char *e = getenv(env.c_str());
  return e ? atoi(e) : 0;
}

// ---------------------------------------------------------------------------
// Domain: device mirror of grid_t
// ---------------------------------------------------------------------------
vpb_domain_t *vpb_domain_create(const vpb_grid_t *g, int rank, int nproc) {
  if (!g) VPB_ERROR("Bad grid");
  if (g->nx < 1 || g->ny < 1 || g->nz < 1) VPB_ERROR("Bad local grid size");
  Context &c = ctx();
  vpb_domain *dom = new vpb_domain();
  DomainDev &d = dom->d;
  d.nx = g->nx; d.ny = g->ny; d.nz = g->nz;
  d.sx = g->nx + 2; d.sy = g->ny + 2; d.sz = g->nz + 2;
  d.sxy = d.sx * d.sy;
  long nv = (long)d.sxy * d.sz;
  // the reference's own limit (grid.h:132-135): 6*voxel must fit an int.  A grid without a neighbor
  // table (field-only use; particles cannot move on it) only needs the voxel index itself to fit.
  if (g->neighbor && 6 * nv > 0x7fffffffL) VPB_ERROR("local domain of %ld voxels exceeds the 2^31/6 voxel limit of grid_t", nv);
  if (nv > 0x7fffffffL) VPB_ERROR("local domain of %ld voxels exceeds 2^31", nv);
  d.nv = (int)nv;
  d.dt = g->dt; d.cvac = g->cvac; d.eps0 = g->eps0; d.damp = g->damp;
  d.dx = g->dx; d.dy = g->dy; d.dz = g->dz;
  d.rdx = g->rdx; d.rdy = g->rdy; d.rdz = g->rdz;
  for (int i = 0; i < 27; i++) d.bc[i] = g->bc[i];
  d.rank = rank; d.nproc = nproc;
  d.rangel = g->rangel; d.rangeh = g->rangeh;
  d.fi_bytes = 80;
  d.p_plane = 0;
  d.fqv = 5; d.fqq = 1;    // the reference's 80-byte field_t until vpb_domain_set_field_layout says otherwise
  dom->host_grid = g;
  dom->range.assign((size_t)nproc + 1, 0);
  for (int r = 0; r <= nproc; r++) dom->range[r] = g->range ? g->range[r] : (int64_t)r * nv;
  if (g->neighbor) {
    // compress: local ids to int32, everything else to a negative code
    std::vector<int32_t> nb((size_t)6 * nv);
    size_t n_handler = 0;
    for (size_t k = 0; k < nb.size(); k++) {
      int64_t n = g->neighbor[k];
      if (n >= g->rangel && n <= g->rangeh) nb[k] = (int32_t)(n - g->rangel);
      else if (n == vpb_reflect_particles) nb[k] = -1;
      else if (n == vpb_absorb_particles) nb[k] = -2;
      else if (n < 0) {
        // boundary_p.c:271-277: -3-k selects the deck's k-th custom particle-boundary handler, a HOST callback with user
        // parameters and the host RNG.  The device cannot run it: move_p leaves such a particle as a mover, and the
        // reference-named boundary_p() calls the handler on the host before the device removes the particle
        // (vpb_dropin.cu).  Callers without that step (layer B, the device-resident driver) are refused when they get
        // there -- absorbing the particle instead would add a rhob deposit the handler never makes.
        if (-n - 3 < (int64_t)g->nb) n_handler++;
        nb[k] = (n > -0x40000000L) ? (int32_t)n : -3;               // any other code: "unknown boundary interaction"
      }
      else nb[k] = INT32_MIN;                                        // owned by another rank
    }
    dom->n_handler_faces = n_handler;
    VPB_CUDA(cudaMalloc(&dom->nbr, nb.size() * sizeof(int32_t)));
    VPB_CUDA(cudaMemcpyAsync(dom->nbr, nb.data(), nb.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c.stream));
    VPB_CUDA(cudaMalloc(&dom->nbr64, nb.size() * sizeof(int64_t)));
    VPB_CUDA(cudaMemcpyAsync(dom->nbr64, g->neighbor, nb.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c.stream));
    VPB_CUDA(cudaStreamSynchronize(c.stream));
  }
  d.nbr = dom->nbr;
  d.nbr64 = dom->nbr64;
  return dom;
}

void vpb_domain_destroy(vpb_domain_t *dom) {
  if (!dom) return;
  VPB_CUDA(cudaStreamSynchronize(ctx().stream));
  sort_group_forget(dom);
  if (dom->nbr) cudaFree(dom->nbr);
  if (dom->nbr64) cudaFree(dom->nbr64);
  for (int f = 0; f < 6; f++) {
    if (dom->face_send[f]) cudaFree(dom->face_send[f]);
    if (dom->face_recv[f]) cudaFree(dom->face_recv[f]);
  }
  delete dom;
}

long vpb_domain_nvoxel(const vpb_domain_t *dom) { return dom ? dom->d.nv : 0; }

}  // extern "C"
