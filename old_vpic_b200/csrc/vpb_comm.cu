// vpb_comm.cu -- NCCL transport (see vpb_comm.cuh).  libnccl is resolved at run
// time with dlopen so that the library also loads in processes that never go
// multi-GPU; the unique id is created here and distributed by the launcher
// (bench.py broadcasts it with torch.distributed, INTEGRATION.md shows the MPI
// equivalent).
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "vpb_comm.cuh"

namespace vpb {

typedef void *nccl_comm_t;
struct nccl_uid { char internal[128]; };
enum { NCCL_INT8 = 0, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

struct Nccl {
  void *h = nullptr;
  int (*GetUniqueId)(nccl_uid *) = nullptr;
  int (*CommInitRank)(nccl_comm_t *, int, nccl_uid, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};

static Nccl g_nccl;
static nccl_comm_t g_comm = nullptr;
static int g_rank = 0, g_nproc = 1;

// Second transport: the host program's own message layer (vpb_mp_transport.hpp).  Chosen by vpb_comm_autoboot when two
// ranks of the job share a GPU (NCCL refuses that), or by the tuning comm.transport = 2.  Field halos are a few
// hundred KB per face and the injector messages a few MB, so this costs latency, not bandwidth; it is the
// compatibility path, NCCL over NVLink is the product path.
static MpLayer g_mp;
static bool g_use_mp = false;

static void mp_exchange_dev(const Xfer *x, int n) {
  cudaStream_t st = ctx().stream;
  auto copy = [st](void *dst, const void *src, size_t bytes) { VPB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st)); };
  auto sync = [st]() { VPB_CUDA(cudaStreamSynchronize(st)); };
  if (!mp_exchange(g_mp, x, n, g_rank, g_nproc, copy, copy, sync))
    VPB_ERROR("exchange of %d transfers does not fit the mp layer (%d ports, messages below 2 GB)", n, kMpPorts);
}

static void load_nccl() {
  if (g_nccl.h) return;
  const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; names[i] && !g_nccl.h; i++) g_nccl.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!g_nccl.h) VPB_ERROR("cannot dlopen libnccl.so.2 (%s); multi-GPU runs need NCCL", dlerror());
#define SYM(field, name)                                                   \
  *(void **)(&g_nccl.field) = dlsym(g_nccl.h, name);                       \
  if (!g_nccl.field) VPB_ERROR("libnccl lacks %s", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(AllReduce, "ncclAllReduce");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
}

#define VPB_NCCL(call)                                                                      \
  do {                                                                                      \
    int _r = (call);                                                                        \
    if (_r != 0) VPB_ERROR("NCCL failure %s: %s", #call, g_nccl.GetErrorString(_r));        \
  } while (0)

int comm_rank() { return g_rank; }
int comm_nproc() { return g_nproc; }
bool comm_is_multi() { return g_comm != nullptr || g_use_mp; }
bool comm_capturable() { return !g_use_mp; }

void comm_exchange(const Xfer *x, int n) {
  cudaStream_t st = ctx().stream;
  bool any_remote = false;
  for (int i = 0; i < n; i++)
    if ((x[i].send_peer >= 0 && x[i].send_peer != g_rank) || (x[i].recv_peer >= 0 && x[i].recv_peer != g_rank)) any_remote = true;
  if (any_remote && g_use_mp) { mp_exchange_dev(x, n); return; }
  if (any_remote && !g_comm) VPB_ERROR("face shared with another rank but vpb_comm_init was not called");
  if (any_remote) VPB_NCCL(g_nccl.GroupStart());
  for (int i = 0; i < n; i++)
    if (x[i].send_peer >= 0 && x[i].send_peer != g_rank)
      VPB_NCCL(g_nccl.Send(x[i].send, x[i].send_bytes, NCCL_INT8, x[i].send_peer, g_comm, st));
  for (int i = 0; i < n; i++)
    if (x[i].recv_peer >= 0 && x[i].recv_peer != g_rank)
      VPB_NCCL(g_nccl.Recv(x[i].recv, x[i].recv_bytes, NCCL_INT8, x[i].recv_peer, g_comm, st));
  if (any_remote) VPB_NCCL(g_nccl.GroupEnd());
}

void comm_allsum_d(double *d_buf, int n) {
  if (g_use_mp && g_nproc > 1) {   // mp_allsum_d itself, on a host copy
    if (n > 64) VPB_ERROR("allsum of %d doubles", n);
    double loc[64], glob[64];
    cudaStream_t st = ctx().stream;
    VPB_CUDA(cudaMemcpyAsync(loc, d_buf, sizeof(double) * n, cudaMemcpyDefault, st));
    VPB_CUDA(cudaStreamSynchronize(st));
    g_mp.allsum_d(loc, glob, n, g_mp.h);
    VPB_CUDA(cudaMemcpyAsync(d_buf, glob, sizeof(double) * n, cudaMemcpyDefault, st));
    VPB_CUDA(cudaStreamSynchronize(st));
    return;
  }
  if (!g_comm || g_nproc == 1) return;
  VPB_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)n, NCCL_FLOAT64, NCCL_SUM, g_comm, ctx().stream));
}

}  // namespace vpb

using namespace vpb;

extern "C" {

// 128-byte NCCL unique id, created on one rank and handed to all the others by the launcher
void vpb_comm_unique_id(void *out128) {
  load_nccl();
  nccl_uid id;
  VPB_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
}

void vpb_comm_init(int rank, int nproc, const void *uid128) {
  if (g_comm) VPB_ERROR("vpb_comm_init called twice");
  if (nproc < 1 || rank < 0 || rank >= nproc) VPB_ERROR("Bad rank/nproc");
  ctx();
  g_rank = rank;
  g_nproc = nproc;
  if (nproc == 1) return;
  load_nccl();
  nccl_uid id;
  memcpy(&id, uid128, sizeof(id));
  VPB_NCCL(g_nccl.CommInitRank(&g_comm, nproc, id, rank));
}

void vpb_comm_finalize(void) {
  if (g_comm) {
    cudaStreamSynchronize(ctx().stream);
    g_nccl.CommDestroy(g_comm);
    g_comm = nullptr;
  }
  g_use_mp = false;
  g_rank = 0;
  g_nproc = 1;
}

// Bootstrap without a line of host code: when the host program is the reference itself (link-time
// substitution) its message layer is in the process -- the *_cxx functions of src/util/mp/mp.hxx:22-143 -- and every
// grid_t carries its handle (grid.h:113).  Rank and size come from mp_rank_cxx / mp_nproc_cxx and the NCCL unique id
// travels from rank 0 through ONE mp_allgather_i_cxx of 32 ints.  The executable must export those symbols
// (-rdynamic).  Returns the world size, or 0 when the reference's message layer is not there or the run has one rank.
int vpb_comm_autoboot(void *mp_handle) {
  if (g_comm || g_use_mp) return g_nproc;
  if (!mp_handle) return 0;
  MpLayer M;
  if (!M.load(mp_handle, false)) return 0;
  const int rank = M.rank_of(mp_handle), nproc = M.nproc_of(mp_handle);
  if (nproc <= 1) return 0;
  // Which GPU does every rank sit on?  Two ranks on one GPU cannot be in one NCCL communicator: such a job runs its
  // exchanges through the host program's message layer instead (tuning comm.transport: 0 decide here, 1 NCCL, 2 mp).
  cudaDeviceProp prop;
  VPB_CUDA(cudaGetDeviceProperties(&prop, ctx().device));
  int id[4];
  memcpy(id, &prop.uuid, sizeof(id));
  std::vector<int> ids(4 * (size_t)nproc);
  M.allgather_i(id, ids.data(), 4, mp_handle);
  bool shared = false;
  for (int a = 0; a < nproc && !shared; a++)
    for (int b = a + 1; b < nproc && !shared; b++) shared = memcmp(&ids[4 * a], &ids[4 * b], sizeof(id)) == 0;
  const int want = tuning("comm.transport", 0);
  if (want == 2 || (want == 0 && shared)) {
    if (!g_mp.load(mp_handle, true)) VPB_ERROR("ranks share a GPU and the host program's mp_*_cxx layer is incomplete");
    g_rank = rank;
    g_nproc = nproc;
    g_use_mp = true;
    return nproc;
  }
  static_assert(sizeof(nccl_uid) == 32 * sizeof(int), "unique id is 32 ints");
  int mine[32];
  memset(mine, 0, sizeof(mine));
  if (rank == 0) vpb_comm_unique_id(mine);
  std::vector<int> all(32 * (size_t)nproc);
  M.allgather_i(mine, all.data(), 32, mp_handle);
  vpb_comm_init(rank, nproc, all.data());      // rank 0's block comes first
  return nproc;
}

int vpb_comm_rank(void) { return g_rank; }
int vpb_comm_nproc(void) { return g_nproc; }

// in-place sum over ranks of n device doubles (mp_allsum_d)
void vpb_comm_allsum_d(double *d_buf, int n) { comm_allsum_d(d_buf, n); }

}  // extern "C"
