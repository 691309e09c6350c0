// vpb_move_p.cuh -- device move_p (src/species_advance/standard/move_p.c:20-136)
// and the quadrant-current arithmetic it shares with advance_p; used by the push
// (vpb_advance_p.cu) and by particle injection (vpb_boundary.cu).
#pragma once
#include "vpb_common.cuh"

namespace vpb {

// One quadrant set of the charge-conserving deposit (advance_p.cxx:136-155 /
// move_p.c:73-92): X is the current direction, Y,Z the transverse ones.
__device__ __forceinline__ void accumulate_j(float q, float uX, float dY, float dZ, float v5, float &o0, float &o1,
                                             float &o2, float &o3) {
  float v0, v1, v2, v3, v4;
  v4 = q * uX;
  v1 = v4 * dY;
  v0 = v4 - v1;
  v1 += v4;
  v4 = 1.0f + dZ;
  v2 = v0 * v4;
  v3 = v1 * v4;
  v4 = 1.0f - dZ;
  v0 *= v4;
  v1 *= v4;
  v0 += v5;
  v1 -= v5;
  v2 -= v5;
  v3 += v5;
  o0 = v0; o1 = v1; o2 = v2; o3 = v3;
}

struct Mover {
  float dx, dy, dz;
  int i;
  float ux, uy, uz, q;
  float dispx, dispy, dispz;
};

// move_p.c:20-136 on registers.  Returns 1 if the mover is still in use.
static __device__ __forceinline__ int move_p_dev(Mover &s, float *__restrict__ a0, const int32_t *__restrict__ nbr) {
  for (;;) {
    float s_midx = s.dx, s_midy = s.dy, s_midz = s.dz;
    float s_dispx = s.dispx, s_dispy = s.dispy, s_dispz = s.dispz;
    const float dirx = (s_dispx > 0) ? 1.0f : -1.0f;
    const float diry = (s_dispy > 0) ? 1.0f : -1.0f;
    const float dirz = (s_dispz > 0) ? 1.0f : -1.0f;
    const float big = (float)3.4e38;
    float v0 = (s_dispx == 0) ? big : (dirx - s_midx) / s_dispx;
    float v1 = (s_dispy == 0) ? big : (diry - s_midy) / s_dispy;
    float v2 = (s_dispz == 0) ? big : (dirz - s_midz) / s_dispz;
    float v3 = 2.0f;
    int type = 3;
    if (v0 < v3) { v3 = v0; type = 0; }
    if (v1 < v3) { v3 = v1; type = 1; }
    if (v2 < v3) { v3 = v2; type = 2; }
    v3 *= 0.5f;
    s_dispx *= v3; s_dispy *= v3; s_dispz *= v3;
    s_midx += s_dispx; s_midy += s_dispy; s_midz += s_dispz;
    // the reference multiplies by the DOUBLE constant (1./3.) here (move_p.c:71)
    const float v5 = (float)((double)(((s.q * s_dispx) * s_dispy) * s_dispz) * (1. / 3.));
    float *a = a0 + 12 * (size_t)s.i;
    float o0, o1, o2, o3;
    accumulate_j(s.q, s_dispx, s_midy, s_midz, v5, o0, o1, o2, o3);
    red_add_v4(a, o0, o1, o2, o3);
    accumulate_j(s.q, s_dispy, s_midz, s_midx, v5, o0, o1, o2, o3);
    red_add_v4(a + 4, o0, o1, o2, o3);
    accumulate_j(s.q, s_dispz, s_midx, s_midy, v5, o0, o1, o2, o3);
    red_add_v4(a + 8, o0, o1, o2, o3);
    s.dispx -= s_dispx; s.dispy -= s_dispy; s.dispz -= s_dispz;
    s.dx += s_dispx + s_dispx; s.dy += s_dispy + s_dispy; s.dz += s_dispz + s_dispz;
    if (type == 3) return 0;
    const float dir = (type == 0) ? dirx : (type == 1) ? diry : dirz;
    const int n = __ldg(nbr + 6 * (size_t)s.i + ((dir > 0) ? 3 : 0) + type);
    if (n < 0) {  // hit a boundary: put the particle exactly on it
      if (type == 0) s.dx = dir; else if (type == 1) s.dy = dir; else s.dz = dir;
      if (n != -1) return 1;  // only reflection is resolved locally
      if (type == 0) { s.ux = -s.ux; s.dispx = -s.dispx; }
      else if (type == 1) { s.uy = -s.uy; s.dispy = -s.dispy; }
      else { s.uz = -s.uz; s.dispz = -s.dispz; }
    } else {
      s.i = n;
      if (type == 0) s.dx = -dir; else if (type == 1) s.dy = -dir; else s.dz = -dir;
    }
  }
}


}  // namespace vpb
