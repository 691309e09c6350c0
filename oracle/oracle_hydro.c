/* TEST INFRASTRUCTURE -- see vpic_oracle.h.  Hydro-moment restatements:
 *   accumulate_hydro_p   src/species_advance/standard/hydro_p.c:24-161
 *   local_adjust_hydro   src/sf_interface/hydro.c:146-184
 *   synchronize_hydro    src/sf_interface/hydro.c:30-141 (single rank: a face whose bc is this rank
 *                        exchanges with the opposite face of the same array)
 * Build flags as for the other oracle files (-O2 -ffp-contract=off -mfpmath=sse).  The reference mixes
 * float variables with double literals in two places; C promotion is reproduced by writing the same
 * operand types. */
#include "vpic_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NCOMP 14   /* jx jy jz rho px py pz ke txx tyy tzz tyz tzx txy */

void orc_clear_hydro(vpb_hydro_t *h, const vpb_grid_t *g) {
  memset(h, 0, (size_t)(g->nx + 2) * (g->ny + 2) * (g->nz + 2) * sizeof(*h));
}

void orc_accumulate_hydro_p(vpb_hydro_t *h0, const vpb_particle_t *p0, int n, float q_m, const vpb_interpolator_t *f0,
                            const vpb_grid_t *g) {
  /* hydro_p.c:48-52 */
  const float qdt_2mc = 0.5 * q_m * g->dt / g->cvac;
  const float qdt_4mc2 = 0.25 * q_m * g->dt / (g->cvac * g->cvac);
  const float c = g->cvac;
  const float r8V = 0.125 * g->rdx * g->rdy * g->rdz;
  const float mc_q = g->cvac / q_m;
  const int sx = g->nx + 2, sxy = sx * (g->ny + 2);
  for (int k = 0; k < n; k++) {
    const vpb_particle_t *p = p0 + k;
    const vpb_interpolator_t *f = f0 + p->i;
    const float x = p->dx, y = p->dy, z = p->dz, q = p->q;
    float u[3] = {p->ux, p->uy, p->uz};
    /* half E kick, B at the particle (hydro_p.c:70-78) */
    u[0] += qdt_2mc * ((f->ex + y * f->dexdy) + z * (f->dexdz + y * f->d2exdydz));
    u[1] += qdt_2mc * ((f->ey + z * f->deydz) + x * (f->deydx + z * f->d2eydzdx));
    u[2] += qdt_2mc * ((f->ez + x * f->dezdx) + y * (f->dezdy + x * f->d2ezdxdy));
    const float b[3] = {f->cbx + x * f->dcbxdx, f->cby + y * f->dcbydy, f->cbz + z * f->dcbzdz};
    /* half Boris rotation and kinetic energy (hydro_p.c:83-101) */
    float ke_mc = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
    float cg = sqrt(1 + ke_mc);              /* gamma: float sum, double sqrt, rounded to float */
    ke_mc *= c / (cg + 1);
    cg = c / cg;                             /* c / gamma */
    float t0 = qdt_4mc2 * cg;
    const float b2 = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
    const float t2 = t0 * t0 * b2;
    const float t3 = t0 * (1 + (1. / 3.) * t2 * (1 + 0.4 * t2));   /* double literals: evaluated in double */
    float t4 = t3 / (1 + b2 * t3 * t3);
    t4 += t4;
    const float up[3] = {u[0] + t3 * (u[1] * b[2] - u[2] * b[1]), u[1] + t3 * (u[2] * b[0] - u[0] * b[2]),
                         u[2] + t3 * (u[0] * b[1] - u[1] * b[0])};
    u[0] += t4 * (up[1] * b[2] - up[2] * b[1]);
    u[1] += t4 * (up[2] * b[0] - up[0] * b[2]);
    u[2] += t4 * (up[0] * b[1] - up[1] * b[0]);
    const float v[3] = {u[0] * cg, u[1] * cg, cg * u[2]};
    /* trilinear node weights of q/8V (hydro_p.c:108-128) */
    float w[8], t;
    w[0] = r8V * q;
    t = x * w[0];
    w[1] = w[0] + t;
    w[0] -= t;
    w[3] = 1 + y;
    w[2] = w[0] * w[3];
    w[3] *= w[1];
    t = 1 - y;
    w[0] *= t;
    w[1] *= t;
    w[7] = 1 + z;
    w[4] = w[0] * w[7];
    w[5] = w[1] * w[7];
    w[6] = w[2] * w[7];
    w[7] *= w[3];
    t = 1 - z;
    w[0] *= t;
    w[1] *= t;
    w[2] *= t;
    w[3] *= t;
    for (int nd = 0; nd < 8; nd++) {   /* hydro_p.c:131-157 */
      vpb_hydro_t *h = h0 + p->i + (nd & 1) + ((nd >> 1) & 1) * sx + (nd >> 2) * sxy;
      float wn = w[nd];
      h->jx += wn * v[0];
      h->jy += wn * v[1];
      h->jz += wn * v[2];
      h->rho += wn;
      wn *= mc_q;
      const float px = wn * u[0], py = wn * u[1], pz = wn * u[2];
      h->px += px;
      h->py += py;
      h->pz += pz;
      h->ke += wn * ke_mc;
      h->txx += px * v[0];
      h->tyy += py * v[1];
      h->tzz += pz * v[2];
      h->tyz += py * v[2];
      h->tzx += pz * v[0];
      h->txy += px * v[1];
    }
  }
}

static int face_bc(const vpb_grid_t *g, int X, int s) {
  int ijk[3] = {0, 0, 0};
  ijk[X] = s;
  return g->bc[VPB_BOUNDARY(ijk[0], ijk[1], ijk[2])];
}

/* node plane X == p, the other two coordinates 1..n+1 */
#define FOR_NODE_PLANE(X, p, body)                                                         \
  do {                                                                                     \
    const int n_[3] = {g->nx, g->ny, g->nz};                                               \
    const int sx_ = g->nx + 2, sy_ = g->ny + 2;                                            \
    int c_[3];                                                                             \
    c_[X] = (p);                                                                           \
    /* loop order of the reference's x/y/z_NODE_LOOP: x fastest, then y, then z */         \
    for (int zz = (X == 2 ? (p) : 1); zz <= (X == 2 ? (p) : n_[2] + 1); zz++)              \
      for (int yy = (X == 1 ? (p) : 1); yy <= (X == 1 ? (p) : n_[1] + 1); yy++)            \
        for (int xx = (X == 0 ? (p) : 1); xx <= (X == 0 ? (p) : n_[0] + 1); xx++) {        \
          float *hv = (float *)(h + xx + sx_ * (yy + (long)sy_ * zz));                     \
          (void)c_;                                                                        \
          body                                                                             \
        }                                                                                  \
  } while (0)

void orc_local_adjust_hydro(vpb_hydro_t *h, const vpb_grid_t *g, int nproc) {
  const int n[3] = {g->nx, g->ny, g->nz};
  for (int s = -1; s <= 1; s += 2)          /* hydro.c:177-182: -x -y -z +x +y +z */
    for (int X = 0; X < 3; X++) {
      const int bc = face_bc(g, X, s);
      if (!(bc < 0 || bc > nproc)) continue;
      const int face = s < 0 ? 1 : n[X] + 1;
      FOR_NODE_PLANE(X, face, { for (int c = 0; c < NCOMP; c++) hv[c] *= 2; });
    }
}

/* one face message (hydro.c:46-67): [cell size along the face normal, 14 moments of every node of the face plane].
 * face 0..5 = -x -y -z +x +y +z; a message packed from face F is consumed through the receiver's face (F+3)%6 */
int orc_hydro_face_floats(int face, const vpb_grid_t *g) {
  const int n[3] = {g->nx, g->ny, g->nz}, X = face % 3;
  return 1 + NCOMP * (n[(X + 1) % 3] + 1) * (n[(X + 2) % 3] + 1);
}

void orc_hydro_face_pack(int face, const vpb_hydro_t *hc, const vpb_grid_t *g, float *buf) {
  vpb_hydro_t *h = (vpb_hydro_t *)hc;
  const int n[3] = {g->nx, g->ny, g->nz}, X = face % 3;
  const float cell[3] = {g->dx, g->dy, g->dz};
  float *q = buf;
  *(q++) = cell[X];
  FOR_NODE_PLANE(X, (face < 3 ? 1 : n[X] + 1), { for (int c = 0; c < NCOMP; c++) *(q++) = hv[c]; });
}

/* hydro.c:69-99: twice weighted sum of my plane and the neighbour's */
void orc_hydro_face_unpack(int face, vpb_hydro_t *h, const vpb_grid_t *g, const float *buf) {
  const int n[3] = {g->nx, g->ny, g->nz}, X = face % 3;
  const float cell[3] = {g->dx, g->dy, g->dz};
  const float *m = buf;
  float rw = *(m++), lw = rw + cell[X];
  rw /= lw;
  lw = cell[X] / lw;
  lw += lw;
  rw += rw;
  FOR_NODE_PLANE(X, (face < 3 ? 1 : n[X] + 1), { for (int c = 0; c < NCOMP; c++) hv[c] = lw * hv[c] + rw * (*(m++)); });
}

void orc_synchronize_hydro(vpb_hydro_t *h, const vpb_grid_t *g, int rank, int nproc) {
  orc_local_adjust_hydro(h, g, nproc);
  for (int X = 0; X < 3; X++) {             /* x faces, then y, then z (hydro.c:107-136) */
    float *msg[2] = {NULL, NULL};           /* [0]: packed from the -X face, [1]: from the +X face */
    for (int s = 0; s < 2; s++) {
      if (face_bc(g, X, s ? 1 : -1) != rank) continue;   /* only self-joined faces exist on one rank */
      msg[s] = (float *)malloc((size_t)orc_hydro_face_floats(X + 3 * s, g) * sizeof(float));
      orc_hydro_face_pack(X + 3 * s, h, g, msg[s]);
    }
    /* what was sent through my -X face comes back through my +X face, which the reference unpacks first */
    if (msg[0]) orc_hydro_face_unpack(X + 3, h, g, msg[0]);
    if (msg[1]) orc_hydro_face_unpack(X, h, g, msg[1]);
    free(msg[0]);
    free(msg[1]);
  }
}
