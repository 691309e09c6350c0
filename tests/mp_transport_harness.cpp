// TEST INFRASTRUCTURE.  Drives old_vpic_b200/csrc/vpb_mp_transport.hpp -- the library's host-staged transport, the very
// code vpb_comm.cu runs with device copies -- on CPU ranks against the reference's own message layer (mp_dmp over
// oracle/mpi_shim's shared-memory MPI).  Built and launched by tests/test_mp_transport.py.
//   usage: harness gpx gpy gpz     (ranks = gpx*gpy*gpz, rank = ix + gpx*(iy + gpy*iz), periodic box)
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "vpb_mp_transport.hpp"

extern "C" {
void mp_init_cxx(int argc, char **argv);
void *new_mp_cxx(void);
void mp_barrier_cxx(void *h);
void mp_finalize_cxx(void *h);
}

using vpb::Xfer;

static uint32_t word(int rank, int face, int round, size_t k) {
  return (uint32_t)(rank * 2654435761u) ^ (uint32_t)(face * 40503u) ^ (uint32_t)(round * 69069u) ^ (uint32_t)(k * 2246822519u);
}

int main(int argc, char **argv) {
  if (argc < 4) return 2;
  const int gp[3] = {atoi(argv[1]), atoi(argv[2]), atoi(argv[3])};
  mp_init_cxx(argc, argv);
  void *h = new_mp_cxx();
  vpb::MpLayer M;
  if (!M.load(h, true)) { fprintf(stderr, "mp layer not found in the process\n"); return 3; }
  const int rank = M.rank_of(h), nproc = M.nproc_of(h);
  if (nproc != gp[0] * gp[1] * gp[2]) { fprintf(stderr, "topology does not match the job\n"); return 4; }
  const int ix[3] = {rank % gp[0], (rank / gp[0]) % gp[1], rank / (gp[0] * gp[1])};
  int peer[6];
  for (int f = 0; f < 6; f++) {
    int j[3] = {ix[0], ix[1], ix[2]};
    const int a = f % 3;
    j[a] = (j[a] + (f < 3 ? -1 : 1) + gp[a]) % gp[a];
    peer[f] = j[0] + gp[0] * (j[1] + gp[1] * j[2]);
  }
  // the allgather the NCCL bootstrap uses
  std::vector<int> all(4 * (size_t)nproc);
  int mine[4] = {rank, 7 * rank, -rank, 1};
  M.allgather_i(mine, all.data(), 4, h);
  for (int r = 0; r < nproc; r++)
    if (all[4 * r] != r || all[4 * r + 1] != 7 * r || all[4 * r + 2] != -r || all[4 * r + 3] != 1) { fprintf(stderr, "allgather\n"); return 5; }

  static const int rorder[6] = {3, 4, 5, 0, 1, 2};
  long checked = 0;
  for (int round = 0; round < 12; round++) {
    // message sizes as the library has them: face planes of a few KB, injector messages up to MBs, sometimes nothing;
    // the size of the message through face F depends on (round, axis) only, so both ends agree on it
    size_t words[6];
    for (int f = 0; f < 6; f++) {
      const int a = f % 3;
      words[f] = (round % 4 == 3 && a == 1) ? 0 : (size_t)(1 + 37 * (a + 1) * (round + 1)) * (round == 7 ? 1024 : 16);
    }
    std::vector<std::vector<uint32_t> > out(6), in(6);
    Xfer x[12];
    int n = 0;
    for (int f = 0; f < 6; f++) {          // sends by face 0..5
      out[f].resize(words[f] + 1);
      for (size_t k = 0; k < words[f]; k++) out[f][k] = word(rank, f, round, k);
      Xfer t = {out[f].data(), 4 * words[f], peer[f], NULL, 0, -1};
      x[n++] = t;
    }
    for (int q = 0; q < 6; q++) {          // receives by face 3,4,5,0,1,2
      const int g = rorder[q];
      in[g].assign(words[g] + 1, 0xdeadbeefu);
      Xfer t = {NULL, 0, -1, in[g].data(), 4 * words[g], peer[g]};
      x[n++] = t;
    }
    auto copy = [](void *d, const void *s, size_t b) { memcpy(d, s, b); };
    auto sync = []() {};
    if (!vpb::mp_exchange(M, x, n, rank, nproc, copy, copy, sync)) { fprintf(stderr, "exchange refused\n"); return 6; }
    for (int g = 0; g < 6; g++) {
      if (peer[g] == rank) {               // a face shared with the rank itself is not a message (the library copies on the device)
        if (in[g][0] != 0xdeadbeefu && words[g]) { fprintf(stderr, "self face was written\n"); return 7; }
        continue;
      }
      for (size_t k = 0; k < words[g]; k++)
        if (in[g][k] != word(peer[g], (g + 3) % 6, round, k)) {
          fprintf(stderr, "rank %d round %d face %d word %zu: got %08x\n", rank, round, g, k, in[g][k]);
          return 8;
        }
      if (in[g][words[g]] != 0xdeadbeefu) { fprintf(stderr, "overrun\n"); return 9; }
      checked += (long)words[g];
    }
    double loc[3] = {1.0 + rank, 0.5 * round, (double)checked}, glob[3];
    M.allsum_d(loc, glob, 3, h);
    if (glob[0] != nproc + 0.5 * nproc * (nproc - 1) || glob[1] != 0.5 * round * nproc) { fprintf(stderr, "allsum\n"); return 10; }
  }
  mp_barrier_cxx(h);
  printf("MP_TRANSPORT_OK rank=%d of %d, %ld words checked\n", rank, nproc, checked);
  mp_finalize_cxx(h);
  return 0;
}
