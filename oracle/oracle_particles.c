/* TEST INFRASTRUCTURE -- see vpic_oracle.h.  Particle-side restatements.
 * Build with -O2 -ffp-contract=off (no FMA): the reference's x86-64 build has
 * none either (cray-haswell.conf: -O2 -mfpmath=sse, no -march). */
#include "vpic_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* The 12 quadrant currents of one straight streak through one voxel
 * (advance_p.cxx:127-162, move_p.c:66-96).  q: charge, h[3]: half displacement,
 * m[3]: streak midpoint, corr = q*hx*hy*hz/3 as computed by the caller. */
static void streak_currents(float out[12], float q, const float h[3], const float m[3], float corr) {
  for (int c = 0; c < 3; c++) {
    const int Y = (c + 1) % 3, Z = (c + 2) % 3;
    float v0, v1, v2, v3, v4;
    v4 = q * h[c];
    v1 = v4 * m[Y];
    v0 = v4 - v1;
    v1 += v4;
    v4 = 1 + m[Z];
    v2 = v0 * v4;
    v3 = v1 * v4;
    v4 = 1 - m[Z];
    v0 *= v4;
    v1 *= v4;
    v0 += corr;
    v1 -= corr;
    v2 -= corr;
    v3 += corr;
    out[4 * c + 0] = v0; out[4 * c + 1] = v1; out[4 * c + 2] = v2; out[4 * c + 3] = v3;
  }
}

int orc_move_p(vpb_particle_t *p0, vpb_particle_mover_t *pm, vpb_accumulator_t *a0, const vpb_grid_t *g) {
  vpb_particle_t *p = p0 + pm->i;
  float *r = &p->dx, *u = &p->ux, *rem = &pm->dispx;
  for (;;) {
    float mid[3], h[3], dir[3], t[3], j[12];
    for (int c = 0; c < 3; c++) {
      mid[c] = r[c];
      h[c] = rem[c];
      dir[c] = (h[c] > 0) ? 1 : -1;
      /* twice the fractional distance to the face along c (move_p.c:49-51) */
      t[c] = (h[c] == 0) ? (float)3.4e38 : (dir[c] - mid[c]) / h[c];
    }
    float frac = 2;
    int type = 3;
    for (int c = 0; c < 3; c++)
      if (t[c] < frac) { frac = t[c]; type = c; }
    frac *= 0.5f;
    for (int c = 0; c < 3; c++) { h[c] *= frac; mid[c] += h[c]; }
    /* double constant 1./3. in the reference (move_p.c:71) */
    const float corr = (float)((double)(p->q * h[0] * h[1] * h[2]) * (1. / 3.));
    streak_currents(j, p->q, h, mid, corr);
    float *a = (float *)(a0 + p->i);
    for (int c = 0; c < 12; c++) a[c] += j[c];
    for (int c = 0; c < 3; c++) { rem[c] -= h[c]; r[c] += h[c] + h[c]; }
    if (type == 3) return 0;
    const float d = dir[type];
    const int64_t nb = g->neighbor[6 * (int64_t)p->i + ((d > 0) ? 3 : 0) + type];
    if (nb < g->rangel || nb > g->rangeh) {
      r[type] = d;
      if (nb != vpb_reflect_particles) return 1;
      u[type] = -u[type];
      rem[type] = -rem[type];
    } else {
      p->i = (int32_t)(nb - g->rangel);
      r[type] = -d;
    }
  }
}

/* Half-step electric kick and the field at the particle (advance_p.cxx:73-83). */
static void gather(const vpb_interpolator_t *f, float qdt_2mc, float dx, float dy, float dz, float ha[3], float cb[3]) {
  ha[0] = qdt_2mc * ((f->ex + dy * f->dexdy) + dz * (f->dexdz + dy * f->d2exdydz));
  ha[1] = qdt_2mc * ((f->ey + dz * f->deydz) + dx * (f->deydx + dz * f->d2eydzdx));
  ha[2] = qdt_2mc * ((f->ez + dx * f->dezdx) + dy * (f->dezdy + dx * f->d2ezdxdy));
  cb[0] = f->cbx + dx * f->dcbxdx;
  cb[1] = f->cby + dy * f->dcbydy;
  cb[2] = f->cbz + dz * f->dcbzdz;
}

/* Boris rotation with the reference's 6th-order tan correction
 * (advance_p.cxx:90-102); k = qdt_2mc (full) or qdt_4mc (half, center_p). */
static void boris(float u[3], const float cb[3], float k) {
  const float one = 1.f, one_third = (float)(1. / 3.), two_fifteenths = (float)(2. / 15.);
  float v0, v1, v2, v3, v4;
  v0 = k / sqrtf(one + (u[0] * u[0] + (u[1] * u[1] + u[2] * u[2])));
  v1 = cb[0] * cb[0] + (cb[1] * cb[1] + cb[2] * cb[2]);
  v2 = (v0 * v0) * v1;
  v3 = v0 * (one + v2 * (one_third + v2 * two_fifteenths));
  v4 = v3 / (one + v1 * (v3 * v3));
  v4 += v4;
  v0 = u[0] + v3 * (u[1] * cb[2] - u[2] * cb[1]);
  v1 = u[1] + v3 * (u[2] * cb[0] - u[0] * cb[2]);
  v2 = u[2] + v3 * (u[0] * cb[1] - u[1] * cb[0]);
  u[0] += v4 * (v1 * cb[2] - v2 * cb[1]);
  u[1] += v4 * (v2 * cb[0] - v0 * cb[2]);
  u[2] += v4 * (v0 * cb[1] - v1 * cb[0]);
}

int orc_advance_p(vpb_particle_t *p0, int np, float q_m, vpb_particle_mover_t *pm, int max_nm, vpb_accumulator_t *a0,
                  const vpb_interpolator_t *f0, const vpb_grid_t *g) {
  const float qdt_2mc = 0.5 * q_m * g->dt / g->cvac; /* advance_p.cxx:425-428 */
  const float cdt[3] = {g->cvac * g->dt * g->rdx, g->cvac * g->dt * g->rdy, g->cvac * g->dt * g->rdz};
  const float one = 1.f, one_third = (float)(1. / 3.);
  int nm = 0;
  for (int k = 0; k < np; k++) {
    vpb_particle_t *p = p0 + k;
    float ha[3], cb[3], u[3] = {p->ux, p->uy, p->uz}, h[3], mid[3], end[3];
    const float r[3] = {p->dx, p->dy, p->dz};
    gather(f0 + p->i, qdt_2mc, r[0], r[1], r[2], ha, cb);
    for (int c = 0; c < 3; c++) u[c] += ha[c];
    boris(u, cb, qdt_2mc);
    for (int c = 0; c < 3; c++) u[c] += ha[c];
    p->ux = u[0]; p->uy = u[1]; p->uz = u[2];
    const float rg = one / sqrtf(one + (u[0] * u[0] + (u[1] * u[1] + u[2] * u[2])));
    for (int c = 0; c < 3; c++) {
      h[c] = u[c] * cdt[c];
      h[c] *= rg;
      mid[c] = r[c] + h[c];
      end[c] = mid[c] + h[c];
    }
    if (end[0] <= one && end[1] <= one && end[2] <= one && -end[0] <= one && -end[1] <= one && -end[2] <= one) {
      float j[12];
      p->dx = end[0]; p->dy = end[1]; p->dz = end[2];
      const float corr = p->q * h[0] * h[1] * h[2] * one_third;
      streak_currents(j, p->q, h, mid, corr);
      float *a = (float *)(a0 + p->i);
      for (int c = 0; c < 12; c++) a[c] += j[c];
    } else {
      vpb_particle_mover_t m = {h[0], h[1], h[2], k};
      if (orc_move_p(p0, &m, a0, g) && nm < max_nm) pm[nm++] = m;
    }
  }
  return nm;
}

void orc_center_p(vpb_particle_t *p0, int np, float q_m, const vpb_interpolator_t *f0, const vpb_grid_t *g) {
  const float qdt_2mc = 0.5 * q_m * g->dt / g->cvac;
  const float qdt_4mc = 0.5 * qdt_2mc; /* center_p.cxx:14-15: half-angle rotation */
  for (int k = 0; k < np; k++) {
    vpb_particle_t *p = p0 + k;
    float ha[3], cb[3], u[3] = {p->ux, p->uy, p->uz};
    gather(f0 + p->i, qdt_2mc, p->dx, p->dy, p->dz, ha, cb);
    for (int c = 0; c < 3; c++) u[c] += ha[c];
    boris(u, cb, qdt_4mc);
    p->ux = u[0]; p->uy = u[1]; p->uz = u[2];
  }
}

/* uncenter_p.cxx:14-15: both constants are negated (runs the half step backwards),
 * rotation first, then the (negative) half kick (uncenter_p.cxx:49-66). */
void orc_uncenter_p(vpb_particle_t *p0, int np, float q_m, const vpb_interpolator_t *f0, const vpb_grid_t *g) {
  const float qdt_2mc_fwd = 0.5 * q_m * g->dt / g->cvac;
  const float qdt_2mc = -qdt_2mc_fwd;
  const float qdt_4mc = -0.5 * qdt_2mc_fwd;
  for (int k = 0; k < np; k++) {
    vpb_particle_t *p = p0 + k;
    float ha[3], cb[3], u[3] = {p->ux, p->uy, p->uz};
    gather(f0 + p->i, qdt_2mc, p->dx, p->dy, p->dz, ha, cb);
    boris(u, cb, qdt_4mc);
    for (int c = 0; c < 3; c++) u[c] += ha[c];
    p->ux = u[0]; p->uy = u[1]; p->uz = u[2];
  }
}

double orc_energy_p(const vpb_particle_t *p0, int np, float q_m, const vpb_interpolator_t *f0, const vpb_grid_t *g) {
  const float qdt_2mc = 0.5 * q_m * g->dt / g->cvac;
  const float one = 1.f;
  double en = 0;
  for (int k = 0; k < np; k++) {
    const vpb_particle_t *p = p0 + k;
    const vpb_interpolator_t *f = f0 + p->i;
    const float dx = p->dx, dy = p->dy, dz = p->dz;
    float v0 = p->ux + qdt_2mc * ((f->ex + dy * f->dexdy) + dz * (f->dexdz + dy * f->d2exdydz));
    float v1 = p->uy + qdt_2mc * ((f->ey + dz * f->deydz) + dx * (f->deydx + dz * f->d2eydzdx));
    float v2 = p->uz + qdt_2mc * ((f->ez + dx * f->dezdx) + dy * (f->dezdy + dx * f->d2ezdxdy));
    v0 = v0 * v0 + v1 * v1 + v2 * v2;
    v0 /= (float)sqrt(one + v0) + one;
    en += (double)v0 * (double)p->q;
  }
  return (double)g->cvac * (double)g->cvac * en / (double)q_m;
}

/* The 8 trilinear node weights of a particle, w = charge/8V already folded in
 * (rho_p.c:43-65, boundary_p.c:19-43).  Order: (x-,y-,z-),(x+,y-,z-),(x-,y+,z-),
 * (x+,y+,z-), then the same four at z+. */
static void node_weights(float w[8], float w0, float x, float y, float z) {
  float t, w1, w2, w3, w4, w5, w6, w7;
  t = x;
  t *= w0;
  w1 = w0 + t;
  w0 -= t;
  t = y;
  w3 = 1 + t;
  w2 = w0 * w3;
  w3 *= w1;
  t = 1 - t;
  w0 *= t;
  w1 *= t;
  t = z;
  w7 = 1 + t;
  w4 = w0 * w7;
  w5 = w1 * w7;
  w6 = w2 * w7;
  w7 *= w3;
  t = 1 - t;
  w0 *= t;
  w1 *= t;
  w2 *= t;
  w3 *= t;
  w[0] = w0; w[1] = w1; w[2] = w2; w[3] = w3; w[4] = w4; w[5] = w5; w[6] = w6; w[7] = w7;
}

void orc_accumulate_rho_p(vpb_field_t *f, const vpb_particle_t *p0, int np, const vpb_grid_t *g) {
  const float r8V = 0.125 * g->rdx * g->rdy * g->rdz;
  const int sx = g->nx + 2, sxy = sx * (g->ny + 2);
  for (int k = 0; k < np; k++) {
    const vpb_particle_t *p = p0 + k;
    float w[8];
    node_weights(w, r8V * p->q, p->dx, p->dy, p->dz);
    for (int n = 0; n < 8; n++) f[p->i + (n & 1) + ((n >> 1) & 1) * sx + (n >> 2) * sxy].rhof += w[n];
  }
}

void orc_accumulate_rhob(vpb_field_t *f, const vpb_particle_t *p, const vpb_grid_t *g) {
  const int sx = g->nx + 2, sy = g->ny + 2, sxy = sx * sy;
  float w[8];
  node_weights(w, (float)(0.125 * p->q * g->rdx * g->rdy * g->rdz), p->dx, p->dy, p->dz);
  const int iz = p->i / sxy, iy = (p->i - iz * sxy) / sx, ix = p->i - iz * sxy - iy * sx;
  const int c[3] = {ix, iy, iz}, n[3] = {g->nx, g->ny, g->nz};
  /* nodes on the domain surface get twice the weight (boundary_p.c:49-60) */
  for (int d = 0; d < 3; d++)
    for (int s = 0; s < 2; s++)
      if (c[d] == (s ? n[d] : 1))
        for (int k = 0; k < 8; k++)
          if (((k >> d) & 1) == s) w[k] += w[k];
  for (int k = 0; k < 8; k++) f[p->i + (k & 1) + ((k >> 1) & 1) * sx + (k >> 2) * sxy].rhob += w[k];
}

void orc_sort_p(const vpb_particle_t *in, vpb_particle_t *out, int np, int *partition, const vpb_grid_t *g) {
  const int nv = (g->nx + 2) * (g->ny + 2) * (g->nz + 2);
  int *next = (int *)calloc((size_t)nv + 1, sizeof(int));
  for (int k = 0; k < np; k++) next[in[k].i]++;
  int run = 0;
  for (int v = 0; v <= nv; v++) {
    partition[v] = run;
    run += next[v];
    next[v] = partition[v];
  }
  for (int k = 0; k < np; k++) out[next[in[k].i]++] = in[k];
  free(next);
}

int orc_boundary_p_pack(vpb_particle_t *p0, int np, const vpb_particle_mover_t *pm, int nm, int sp_id, vpb_field_t *f,
                        const vpb_grid_t *g, int rank, int nproc, vpb_particle_injector_t *out[6], int n_out[6]) {
  static const int fbound[6] = {VPB_BOUNDARY(-1, 0, 0), VPB_BOUNDARY(0, -1, 0), VPB_BOUNDARY(0, 0, -1),
                                VPB_BOUNDARY(1, 0, 0),  VPB_BOUNDARY(0, 1, 0),  VPB_BOUNDARY(0, 0, 1)};
  int64_t rbase[6];
  for (int face = 0; face < 6; face++) {
    const int b = g->bc[fbound[face]];
    rbase[face] = (b >= 0 && b < nproc && b != rank) ? g->range[b] : 0;
    n_out[face] = 0;
  }
  const int64_t rangem = g->range[nproc];
  for (int k = nm - 1; k >= 0; k--) { /* reverse so back-filling keeps lower indices valid (boundary_p.c:168-176) */
    vpb_particle_t *r = p0 + pm[k].i;
    const float pos[3] = {r->dx, r->dy, r->dz}, u[3] = {r->ux, r->uy, r->uz};
    int handled = 0;
    for (int face = 0; face < 6 && !handled; face++) {
      const int ax = face % 3, up = face >= 3;
      if (!(up ? (pos[ax] == 1 && u[ax] > 0) : (pos[ax] == -1 && u[ax] < 0))) continue;
      const int64_t nn = g->neighbor[6 * (int64_t)r->i + face];
      if (nn == vpb_absorb_particles) {
        orc_accumulate_rhob(f, r, g);
        handled = 1;
      } else if ((nn >= 0 && nn < g->rangel) || (nn > g->rangeh && nn <= rangem)) {
        vpb_particle_injector_t *o = out[face] + n_out[face]++;
        o->dx = (ax == 0) ? -r->dx : r->dx;
        o->dy = (ax == 1) ? -r->dy : r->dy;
        o->dz = (ax == 2) ? -r->dz : r->dz;
        o->i = (int32_t)(nn - rbase[face]);
        o->ux = r->ux; o->uy = r->uy; o->uz = r->uz; o->q = r->q;
        o->dispx = pm[k].dispx; o->dispy = pm[k].dispy; o->dispz = pm[k].dispz;
        o->sp_id = sp_id;
        handled = 1;
      }
      /* custom handlers (nn<=-3) are host callbacks: out of scope, fall through */
    }
    if (!handled) orc_accumulate_rhob(f, r, g); /* "unknown boundary interaction ... using absorption" */
    *r = p0[--np];
  }
  return np;
}

int orc_boundary_p_inject(vpb_particle_t *p0, int *np, vpb_particle_mover_t *pm, int nm, const vpb_particle_injector_t *in,
                          int n_in, int sp_id, vpb_accumulator_t *a0, const vpb_grid_t *g) {
  int added = 0;
  for (int k = n_in - 1; k >= 0; k--) {
    if (in[k].sp_id != sp_id) continue;
    vpb_particle_t *p = p0 + *np;
    vpb_particle_mover_t *m = pm + nm + added;
    p->dx = in[k].dx; p->dy = in[k].dy; p->dz = in[k].dz; p->i = in[k].i;
    p->ux = in[k].ux; p->uy = in[k].uy; p->uz = in[k].uz; p->q = in[k].q;
    m->dispx = in[k].dispx; m->dispy = in[k].dispy; m->dispz = in[k].dispz;
    m->i = *np;
    (*np)++;
    added += orc_move_p(p0, m, a0, g);
  }
  return added;
}
