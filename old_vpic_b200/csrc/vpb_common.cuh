// vpb_common.cuh -- shared device/host helpers for libvpic_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../include/vpic_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libvpic_b200 is written for sm_100a (B200) only"
#endif

// Reference error conventions (src/util/util_base.h:201-219): ERROR prints and
// exits, WARNING prints and continues.
#define VPB_ERROR(...)                                                         \
  do {                                                                         \
    fprintf(stderr, "Error at %s(%i):\n\t", __FILE__, __LINE__);               \
    fprintf(stderr, __VA_ARGS__);                                              \
    fprintf(stderr, "\n");                                                     \
    fflush(stderr);                                                            \
    exit(1);                                                                   \
  } while (0)
#define VPB_WARNING(...)                                                       \
  do {                                                                         \
    fprintf(stderr, "Warning at %s(%i):\n\t", __FILE__, __LINE__);             \
    fprintf(stderr, __VA_ARGS__);                                              \
    fprintf(stderr, "\n");                                                     \
    fflush(stderr);                                                            \
  } while (0)
#define VPB_CUDA(call)                                                         \
  do {                                                                         \
    cudaError_t _e = (call);                                                   \
    if (_e != cudaSuccess)                                                     \
      VPB_ERROR("CUDA failure %s: %s (no CPU fallback exists)", #call, cudaGetErrorString(_e)); \
  } while (0)

struct vpb_domain;

namespace vpb {

// Device mirror of the parts of grid_t the kernels read (grid.h:112-167).
struct DomainDev {
  int nx, ny, nz;        // interior cells
  int sx, sy, sz;        // nx+2, ny+2, nz+2
  int sxy;               // sx*sy
  int nv;                // sx*sy*sz voxels including ghosts
  float dt, cvac, eps0, damp;
  float dx, dy, dz, rdx, rdy, rdz;
  int bc[27];
  int rank, nproc;
  // neighbor table compressed to int32 (grid.h:145-150): >=0 local voxel index,
  // -1 reflect_particles, any other negative value = "move_p cannot resolve"
  // (absorb / custom handler / voxel owned by another rank).
  const int32_t *nbr;
  // original 64-bit table, only read by migration (boundary_p.c:304-309)
  const int64_t *nbr64;
  int64_t rangel, rangeh;
  // field_t addressing in 16-byte quads {e,div_e | cb,div_b | tca,rhob | jf,rhof | material ids}: quad q of voxel
  // v sits at float4 index v*fqv + q*fqq.  Reference layout (80-byte AoS): fqv=5, fqq=1.  Planar layout (five
  // planes of nvp quads, device-resident runs): fqv=1, fqq=nvp.  See DESIGN.md "field layout".
  long fqv, fqq;
  // interpolator record stride in bytes: 80 = the reference's interpolator_t; 96 = the same 72 bytes padded to
  // three aligned 32-byte sectors, which advance_p gathers with two 256-bit loads and one 64-bit load
  int fi_bytes;
  // particle arrays of device-resident runs: 0 = the reference's 48-byte particle_t records; > 0 = component
  // planes of p_plane words each (vpb_pview.cuh), every species array of the domain with that capacity
  long p_plane;
};

#define FQ(f, g, v, q) (reinterpret_cast<float4 *>(f) + ((size_t)(v) * (size_t)(g).fqv + (size_t)(q) * (size_t)(g).fqq))
#define CFQ(f, g, v, q) (reinterpret_cast<const float4 *>(f) + ((size_t)(v) * (size_t)(g).fqv + (size_t)(q) * (size_t)(g).fqq))
// float component c (0..19, the reference's member order) of voxel v
#define FCOMP(f, g, v, c) \
  (reinterpret_cast<float *>(f)[4 * ((size_t)(v) * (size_t)(g).fqv + (size_t)((c) >> 2) * (size_t)(g).fqq) + ((c) & 3)])

struct Context {
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_start[16] = {}, ev_stop[16] = {};
  long launches = 0;
  // scratch owned by the context (grown on demand)
  void *scratch = nullptr;
  size_t scratch_bytes = 0;
  int *h_pinned_i = nullptr;       // 64 pinned ints for small read-backs
  double *h_pinned_d = nullptr;    // 64 pinned doubles
};

Context &ctx();                       // exits loudly if no device
void *scratch(size_t bytes);          // stream-ordered scratch of at least `bytes`
int tuning(const char *name, int dflt);

inline void count_launch(int n = 1) { ctx().launches += n; }
void sort_group_forget(const vpb_domain *dom);   // vpb_sort_group.cu: drop the order tables cached for a domain

// Optional per-kernel timing (CUDA events on the library stream around selected
// launches); off unless vpb_prof_enable(1).  Classes: 0 advance_p, 1 sort_p,
// 2 advance_b, 3 advance_e, 4 load_interpolator, 5 unload_accumulator, 6 other (compute_curl_b),
// 7 boundary_p (particle migration), 8 halo (synchronize_jf and the ghost exchanges of advance_e),
// 9 divergence cleaning and shared-face synchronisation.  Scopes may nest (7-9 are set by the step driver).
int prof_begin(int cls);
void prof_end(int idx);
bool prof_enabled();
struct ProfScope {
  int idx;
  explicit ProfScope(int c) : idx(prof_begin(c)) {}
  ~ProfScope() { prof_end(idx); }
};

// Host wall-clock accounting of the layer-A entry points (VPB_TRACE=1 or vpb_trace_enable): calls and seconds
// per label, nested labels ("  prefetch", "  sync") counted inside their entry point.  Written to stderr (or to
// the file VPB_TRACE_FILE names) at exit or by vpb_trace_report.  Costs two clock reads per scope when on.
extern bool g_trace_on;
void trace_add(const char *label, double seconds);
double trace_now();
struct TraceScope {
  const char *label;
  double t0;
  explicit TraceScope(const char *l) : label(l), t0(g_trace_on ? trace_now() : 0.0) {}
  ~TraceScope() { if (g_trace_on) trace_add(label, trace_now() - t0); }
};

}  // namespace vpb

struct vpb_domain {
  vpb::DomainDev d;       // by-value kernel argument
  int32_t *nbr = nullptr;  // device allocations owned by the domain
  int64_t *nbr64 = nullptr;
  const vpb_grid_t *host_grid = nullptr;
  std::vector<int64_t> range;   // copy of grid_t.range[0..nproc] (global voxel-id range of every rank)
  // per-face message buffers (vpb_faces.cu), sized for the largest message kind
  float *face_send[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float *face_recv[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t face_cap[6] = {0, 0, 0, 0, 0, 0};
  size_t n_handler_faces = 0;   // cell faces bound to the deck's custom particle-boundary handlers (host callbacks)
  // particle migration (vpb_boundary.cu): injectors a fused message of round 0, 1, 2+ may carry through each face;
  // both sides of a face derive the same number from the counts of earlier rounds
  int mig_cap[3][6] = {{0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}};
  bool mig_cap_valid[3] = {false, false, false};
};

// ---------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------
#ifdef __CUDACC__
namespace vpb {

__device__ __forceinline__ float4 ldg4(const void *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float2 ldg2(const void *p) { return __ldg(reinterpret_cast<const float2 *>(p)); }

// Vector reduction to global memory: one REDG.E.ADD.F32x4 (sm_90+), 16-B aligned.
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add(float *addr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

// L2 eviction policies (createpolicy + .L2::cache_hint): the particle stream is touched once per
// kernel and must not push the interpolator/accumulator working set out of the 126 MB L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ld_hint4(const void *p, uint64_t pol) {
  float4 r;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ float4 ldg_hint4(const void *p, uint64_t pol) {
  float4 r;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ float2 ldg_hint2(const void *p, uint64_t pol) {
  float2 r;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(r.x), "=f"(r.y) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void st_hint4(void *p, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_add_v4_hint(float *addr, float a, float b, float c, float d, uint64_t pol) {
  asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d), "l"(pol) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace vpb
#endif
