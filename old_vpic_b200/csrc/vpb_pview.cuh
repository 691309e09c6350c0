// vpb_pview.cuh -- one accessor for the two device layouts of a particle array.
//
//  * plane == 0: the reference's particle_t[] (species_advance.h:28-34), 48-byte AoS records of three
//    16-byte quads {dx,dy,dz,i | ux,uy,uz,q | tag,tag2}.  This is what every layer-A entry point sees.
//  * plane  > 0: component planes (device-resident runs, DESIGN.md "particle layout"): eight planes of
//    `plane` 4-byte words -- dx, dy, dz, i, ux, uy, uz, q -- followed by one plane of `plane` 16-byte tag
//    pairs.  Same 48 bytes per particle, same allocation size for a capacity of `plane` particles.  A lane of
//    advance_p that owns two consecutive particles reads each component as ONE aligned 64-bit word that is
//    already the register pair a packed f32x2 instruction wants, and only the six words that change
//    (dx,dy,dz,ux,uy,uz) are written back: 56 bytes of particle traffic per advance instead of 96.
//    `plane` must be a multiple of 64.
#pragma once
#include "vpb_common.cuh"

namespace vpb {

struct PView {
  float *b;
  long plane;
  __host__ __device__ PView() : b(nullptr), plane(0) {}
  __host__ __device__ PView(const void *p, long pl) : b(reinterpret_cast<float *>(const_cast<void *>(p))), plane(pl) {}

#ifdef __CUDACC__
  __device__ __forceinline__ float *comp(int c) const { return b + (size_t)c * (size_t)plane; }
  __device__ __forceinline__ float4 pos(long k) const {        // dx, dy, dz, i
    if (!plane) return reinterpret_cast<const float4 *>(b)[3 * k];
    return make_float4(b[k], b[plane + k], b[2 * plane + k], b[3 * plane + k]);
  }
  __device__ __forceinline__ float4 mom(long k) const {        // ux, uy, uz, q
    if (!plane) return reinterpret_cast<const float4 *>(b)[3 * k + 1];
    return make_float4(b[4 * plane + k], b[5 * plane + k], b[6 * plane + k], b[7 * plane + k]);
  }
  __device__ __forceinline__ float4 tag(long k) const {        // tag, tag2 (two int64 as raw bits)
    if (!plane) return reinterpret_cast<const float4 *>(b)[3 * k + 2];
    return reinterpret_cast<const float4 *>(b + 8 * plane)[k];
  }
  __device__ __forceinline__ int voxel(long k) const {
    return __float_as_int(plane ? b[3 * plane + k] : b[12 * k + 3]);
  }
  __device__ __forceinline__ void set_pos(long k, float4 v) const {
    if (!plane) { reinterpret_cast<float4 *>(b)[3 * k] = v; return; }
    b[k] = v.x; b[plane + k] = v.y; b[2 * plane + k] = v.z; b[3 * plane + k] = v.w;
  }
  __device__ __forceinline__ void set_mom(long k, float4 v) const {
    if (!plane) { reinterpret_cast<float4 *>(b)[3 * k + 1] = v; return; }
    b[4 * plane + k] = v.x; b[5 * plane + k] = v.y; b[6 * plane + k] = v.z; b[7 * plane + k] = v.w;
  }
  __device__ __forceinline__ void set_tag(long k, float4 v) const {
    if (!plane) { reinterpret_cast<float4 *>(b)[3 * k + 2] = v; return; }
    reinterpret_cast<float4 *>(b + 8 * plane)[k] = v;
  }
  // quad 0,1,2 of particle k
  __device__ __forceinline__ float4 quad(long k, int piece) const { return piece == 0 ? pos(k) : (piece == 1 ? mom(k) : tag(k)); }
  __device__ __forceinline__ void set_quad(long k, int piece, float4 v) const {
    if (piece == 0) set_pos(k, v); else if (piece == 1) set_mom(k, v); else set_tag(k, v);
  }
#endif
};

}  // namespace vpb
