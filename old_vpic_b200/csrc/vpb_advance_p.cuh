// vpb_advance_p.cuh -- internal three-part form of advance_p (see vpb_advance_p.cu)
#pragma once
#include "vpb_common.cuh"

namespace vpb {

struct AdvanceArgs {
  vpb_particle_t *p;
  int np;
  int nchunks;                    // ceil(np/32)
  float qdt_2mc, cdt_dx, cdt_dy, cdt_dz;
  float *a;                       // accumulator_t[nv] viewed as float[12*nv]
  const vpb_interpolator_t *f;
  int fi_bytes;                   // interpolator record stride (80 or 96, DomainDev::fi_bytes)
  const int32_t *nbr;
  vpb_particle_mover_t *tmp_pm;   // unordered staging, capacity max_nm
  int max_nm;
  int *counters;                  // [0] staged movers  [1] movers ignored (overflow)
  unsigned *bitmap;               // one bit per particle: has an unresolved mover
  // traversal: work item -> range of chunks.  With a partition[] from the last sort the items are the
  // x-rows of voxels visited y-blocked / z-inner (see work_range); without, 64-chunk slabs in order.
  const int *partition;           // first particle of each voxel at the last sort (NULL: linear)
  int nwork, by, sx, sy, sz, nv;
  int chunk_lo, chunk_hi;         // linear mode: the chunk range of this launch
};

struct AdvanceJob {
  AdvanceArgs A;
  int *word_off;
  void *scan_tmp;
  int nwords;
};

void advance_p_begin(vpb_domain_t *dom, int np, float q_m, int max_nm, vpb_accumulator_t *d_a, const vpb_interpolator_t *d_f,
                     AdvanceJob &J, cudaStream_t st);
void advance_p_range(AdvanceJob &J, vpb_particle_t *d_base, int k0, int k1, const int *d_partition, cudaStream_t st);
// component-plane arrays (vpb_advance_p_pair.cu): the whole array in one launch
void advance_p_pair_launch(AdvanceJob &J, float *d_planes, long plane, cudaStream_t st);
void advance_p_end(AdvanceJob &J, vpb_particle_mover_t *d_pm, int *d_nm, cudaStream_t st);

}  // namespace vpb
