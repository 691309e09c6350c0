// vpb_comm.cuh -- rank-to-rank transport for the face messages and the particle
// migration: NCCL send/recv over NVLink, one rank per GPU.  Replaces the
// reference's MPI layer for this path (src/util/mp/dmp/mp_dmp.c:225-295 and the
// port helpers of src/grid/grid_comm.c).  A face shared with the rank itself
// (periodic along an axis with one rank) is a device copy, not a message.
#pragma once
#include "vpb_common.cuh"
#include "vpb_mp_transport.hpp"   // struct Xfer and the host-staged second transport

namespace vpb {

int comm_rank();
int comm_nproc();
bool comm_is_multi();
bool comm_capturable();   // exchanges are stream work only (NCCL or none): a CUDA graph may hold them

// Posts every send and receive of `x[0..n)` as one NCCL group on the library
// stream (or, for jobs whose ranks share a GPU, through the host program's message layer: vpb_mp_transport.hpp).  Messages between the same pair of ranks are matched in posting order,
// so callers list sends by face 0..5 and receives by face 3,4,5,0,1,2 (a message
// sent through face F arrives through the peer's face (F+3)%6).
void comm_exchange(const Xfer *x, int n);

// sum over ranks, in place, of n doubles on the device (mp_allsum_d, mp_dmp.c:299-311)
void comm_allsum_d(double *d_buf, int n);

}  // namespace vpb
