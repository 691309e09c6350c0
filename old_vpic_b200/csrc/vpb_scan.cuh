// vpb_scan.cuh -- device-wide exclusive prefix sum of int32 (reduce / scan /
// downsweep over 4096-element blocks, recursing on the block sums).  Used for the
// voxel histogram -> partition[] of sort_p (sort_p.c:54-59) and for ordering
// per-tile mover counts in advance_p.
#pragma once
#include "vpb_common.cuh"

namespace vpb {

constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanBlock = kScanThreads * kScanItems;  // 4096

// block_sums[b] = sum of in[b*4096 .. )
// skip (may be NULL): device int; when it reads 0 the kernel has nothing to do and returns
static __global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const int *__restrict__ in, int *__restrict__ block_sums, int n,
                                                                          const int *__restrict__ skip) {
  __shared__ int ws[kScanThreads / 32];
  if (skip && *skip == 0) return;
  const int base = blockIdx.x * kScanBlock;
  int s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; j++) {
    const int i = base + j * kScanThreads + threadIdx.x;
    if (i < n) s += in[i];
  }
  s = __reduce_add_sync(0xffffffffu, s);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int j = 0; j < kScanThreads / 32; j++) t += ws[j];
    block_sums[blockIdx.x] = t;
  }
}

// out[i] = block_off[b] + exclusive prefix of in within block b (in may alias out)
static __global__ void __launch_bounds__(kScanThreads) scan_block_kernel(const int *in, int *out, const int *__restrict__ block_off, int n,
                                                                         const int *__restrict__ skip) {
  __shared__ int ws[kScanThreads / 32];
  if (skip && *skip == 0) return;
  const int base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;  // each thread owns 8 consecutive items
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; j++) {
    v[j] = (base + j < n) ? in[base + j] : 0;
    s += v[j];
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) ws[w] = incl;
  __syncthreads();
  int wbase = block_off ? block_off[blockIdx.x] : 0;
  for (int j = 0; j < w; j++) wbase += ws[j];
  int run = wbase + incl - s;
#pragma unroll
  for (int j = 0; j < kScanItems; j++) {
    if (base + j < n) out[base + j] = run;
    run += v[j];
  }
}

inline size_t scan_scratch_bytes(long n) {
  size_t b = 0;
  while (n > kScanBlock) {
    n = (n + kScanBlock - 1) / kScanBlock;
    b += ((size_t)n * sizeof(int) + 255) & ~(size_t)255;
  }
  return b + 256;
}

inline int scan_launches(long n) {
  int l = 1;
  while (n > kScanBlock) { n = (n + kScanBlock - 1) / kScanBlock; l += 2; }
  return l;
}

// in may alias out.  tmp must hold scan_scratch_bytes(n).
inline void exclusive_scan_i32(const int *in, int *out, int n, void *tmp, cudaStream_t st, const int *skip = nullptr) {
  if (n <= 0) return;
  const int nb = (n + kScanBlock - 1) / kScanBlock;
  if (nb == 1) {
    scan_block_kernel<<<1, kScanThreads, 0, st>>>(in, out, nullptr, n, skip);
    return;
  }
  int *sums = (int *)tmp;
  void *next = (char *)tmp + (((size_t)nb * sizeof(int) + 255) & ~(size_t)255);
  scan_reduce_kernel<<<nb, kScanThreads, 0, st>>>(in, sums, n, skip);
  exclusive_scan_i32(sums, sums, nb, next, st, skip);
  scan_block_kernel<<<nb, kScanThreads, 0, st>>>(in, out, sums, n, skip);
}

}  // namespace vpb
