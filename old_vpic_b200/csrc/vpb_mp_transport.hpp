// vpb_mp_transport.hpp -- the library's second rank-to-rank transport: the HOST PROGRAM'S OWN message layer (the
// reference's mp_*_cxx functions, src/util/mp/mp.hxx:22-143, found in the process by name), messages staged through
// its port buffers.  Plain C++ (no CUDA) so that tests/test_mp_transport.py can drive exactly this code on CPU ranks
// against the reference's mp_dmp over oracle/mpi_shim; vpb_comm.cu supplies device<->host copies.
#pragma once
#include <dlfcn.h>
#include <stddef.h>
#include <vector>

namespace vpb {

struct Xfer {
  const void *send; size_t send_bytes; int send_peer;   // send_peer < 0: nothing to send
  void *recv; size_t recv_bytes; int recv_peer;          // recv_peer < 0: nothing to receive
};

constexpr int kMpPorts = 27;   // NUM_BUF of util/mp/dmp/mp_t.h:27

struct MpLayer {
  void *h = nullptr;
  int (*rank_of)(void *) = nullptr;
  int (*nproc_of)(void *) = nullptr;
  void (*allgather_i)(int *, int *, int, void *) = nullptr;
  void (*size_recv)(int, int, void *) = nullptr;
  void (*size_send)(int, int, void *) = nullptr;
  void *(*recv_buffer)(int, void *) = nullptr;
  void *(*send_buffer)(int, void *) = nullptr;
  void (*begin_recv)(int, int, int, int, void *) = nullptr;
  void (*begin_send)(int, int, int, int, void *) = nullptr;
  void (*end_recv)(int, void *) = nullptr;
  void (*end_send)(int, void *) = nullptr;
  void (*allsum_d)(double *, double *, int, void *) = nullptr;

  // rank/size/allgather are enough for the NCCL bootstrap; `full` asks for the message functions as well
  bool load(void *handle, bool full) {
#define VPB_MPSYM(field, name)                                    \
  *(void **)(&field) = dlsym(RTLD_DEFAULT, name);                 \
  if (!field) return false
    VPB_MPSYM(rank_of, "mp_rank_cxx");
    VPB_MPSYM(nproc_of, "mp_nproc_cxx");
    VPB_MPSYM(allgather_i, "mp_allgather_i_cxx");
    if (full) {
      VPB_MPSYM(size_recv, "mp_size_recv_buffer_cxx");
      VPB_MPSYM(size_send, "mp_size_send_buffer_cxx");
      VPB_MPSYM(recv_buffer, "mp_recv_buffer_cxx");
      VPB_MPSYM(send_buffer, "mp_send_buffer_cxx");
      VPB_MPSYM(begin_recv, "mp_begin_recv_cxx");
      VPB_MPSYM(begin_send, "mp_begin_send_cxx");
      VPB_MPSYM(end_recv, "mp_end_recv_cxx");
      VPB_MPSYM(end_send, "mp_end_send_cxx");
      VPB_MPSYM(allsum_d, "mp_allsum_d_cxx");
    }
#undef VPB_MPSYM
    h = handle;
    return true;
  }
};

// x[i] travels through port i.  Between one pair of ranks the k-th send matches the k-th receive (the rule NCCL
// applies inside a group), expressed as the message tag.  to_host(dst, src, bytes) / to_device(dst, src, bytes) enqueue
// copies, sync() waits for them.  Returns false when the list does not fit the ports.
template <class ToHost, class ToDevice, class Sync>
bool mp_exchange(MpLayer &M, const Xfer *x, int n, int rank, int nproc, ToHost to_host, ToDevice to_device, Sync sync) {
  if (n > kMpPorts) return false;
  auto remote = [&](int peer, size_t bytes) { return peer >= 0 && peer != rank && bytes > 0; };
  for (int i = 0; i < n; i++)
    if (remote(x[i].send_peer, x[i].send_bytes)) {
      if (x[i].send_bytes > 0x7fffffffu) return false;
      M.size_send(i, (int)x[i].send_bytes, M.h);
      to_host(M.send_buffer(i, M.h), x[i].send, x[i].send_bytes);
    }
  sync();
  std::vector<int> rcount(nproc, 0), scount(nproc, 0);
  for (int i = 0; i < n; i++)
    if (remote(x[i].recv_peer, x[i].recv_bytes)) {
      M.size_recv(i, (int)x[i].recv_bytes, M.h);
      M.begin_recv(i, (int)x[i].recv_bytes, x[i].recv_peer, rcount[x[i].recv_peer]++, M.h);
    }
  for (int i = 0; i < n; i++)
    if (remote(x[i].send_peer, x[i].send_bytes))
      M.begin_send(i, (int)x[i].send_bytes, x[i].send_peer, scount[x[i].send_peer]++, M.h);
  for (int i = 0; i < n; i++)
    if (remote(x[i].recv_peer, x[i].recv_bytes)) {
      M.end_recv(i, M.h);
      to_device(x[i].recv, M.recv_buffer(i, M.h), x[i].recv_bytes);
    }
  for (int i = 0; i < n; i++)
    if (remote(x[i].send_peer, x[i].send_bytes)) M.end_send(i, M.h);
  sync();   // the port buffers are reused by the next exchange
  return true;
}

}  // namespace vpb
