/* TEST INFRASTRUCTURE -- see vpic_oracle.h.  Field-side restatements
 * (standard field advance: sfa.c, advance_b.c, advance_e.c, local.c, remote.c,
 * energy_f.c, the div-clean files, load_interpolator.cxx, unload_accumulator.cxx).
 *
 * The reference spells every face out through macros with (X,Y,Z) permuted
 * cyclically; here a face is (axis X, side s) and Y=(X+1)%3, Z=(X+2)%3, and a
 * field_t is addressed as 20 floats (ex..rhof = 0..15) plus 8 uint16 material
 * ids, so one loop serves all three orientations.  Evaluation order inside each
 * expression is the reference's.  Build with -ffp-contract=off. */
#include "vpic_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { EX = 0, EY, EZ, DIVE, CBX, CBY, CBZ, DIVB, TCAX, TCAY, TCAZ, RHOB, JFX, JFY, JFZ, RHOF };
enum { EMAT = 0, NMAT = 3, FMAT = 4, CMAT = 7 };

typedef struct dims {
  int n[3];      /* interior cells per axis */
  long st[3];    /* voxel stride per axis */
  long nv;
} dims_t;

static dims_t dims_of(const vpb_grid_t *g) {
  dims_t d;
  d.n[0] = g->nx; d.n[1] = g->ny; d.n[2] = g->nz;
  d.st[0] = 1; d.st[1] = g->nx + 2; d.st[2] = (long)(g->nx + 2) * (g->ny + 2);
  d.nv = d.st[2] * (g->nz + 2);
  return d;
}

#define F(f, v, c) (((float *)(f))[20 * (long)(v) + (c)])
#define CF(f, v, c) (((const float *)(f))[20 * (long)(v) + (c)])
static inline int mat_id(const vpb_field_t *f, long v, int k) { return ((const uint16_t *)&f[v].ematx)[k]; }
static inline long vox(const dims_t *d, int x, int y, int z) { return x + d->st[1] * y + d->st[2] * z; }

/* Inclusive box; loops always run z outer, y middle, x inner (the reference's
 * XYZ_LOOP, local.c:25-28), which also fixes the order of values in a message. */
typedef struct box { int lo[3], hi[3]; } box_t;
#define FOR_BOX(b, x, y, z)                        \
  for (int z = (b).lo[2]; z <= (b).hi[2]; z++)     \
    for (int y = (b).lo[1]; y <= (b).hi[1]; y++)   \
      for (int x = (b).lo[0]; x <= (b).hi[0]; x++)

/* plane X==p, Y in 1..nY+eY, Z in 1..nZ+eZ  (local.c:30-44):
 *   face loop (0,0); node loop (1,1); "YZ edge" = Y-directed edges (0,1);
 *   "ZY edge" = Z-directed edges (1,0). */
static box_t plane(const dims_t *d, int X, int p, int eY, int eZ) {
  const int Y = (X + 1) % 3, Z = (X + 2) % 3;
  box_t b;
  b.lo[X] = b.hi[X] = p;
  b.lo[Y] = 1; b.hi[Y] = d->n[Y] + eY;
  b.lo[Z] = 1; b.hi[Z] = d->n[Z] + eZ;
  return b;
}

static int face_bc(const vpb_grid_t *g, int X, int s) {
  int ijk[3] = {0, 0, 0};
  ijk[X] = s;
  return g->bc[VPB_BOUNDARY(ijk[0], ijk[1], ijk[2])];
}
static int is_local(int bc, int nproc) { return bc < 0 || bc > nproc; }   /* local.c:70 */
static int is_remote(int bc, int nproc) { return bc >= 0 && bc < nproc; } /* grid_comm.c:17 */
static float cell_size(const vpb_grid_t *g, int X) { return X == 0 ? g->dx : X == 1 ? g->dy : g->dz; }
static float rcell(const vpb_grid_t *g, int X) { return X == 0 ? g->rdx : X == 1 ? g->rdy : g->rdz; }

/* The reference visits faces in the order -x,-y,-z,+x,+y,+z. */
#define FOR_FACES(X, s) \
  for (int _f = 0, X = 0, s = -1; _f < 6; _f++, X = _f % 3, s = _f < 3 ? -1 : 1)

/* ---------------------------------------------------------------------- */
/* species <-> field                                                       */
/* ---------------------------------------------------------------------- */

void orc_load_interpolator(vpb_interpolator_t *fi, const vpb_field_t *f, const vpb_grid_t *g) {
  const dims_t d = dims_of(g);
  const float fourth = 0.25f, half = 0.5f;
  box_t b = {{1, 1, 1}, {d.n[0], d.n[1], d.n[2]}};
  FOR_BOX(b, x, y, z) {
    const long v = vox(&d, x, y, z);
    float *o = (float *)(fi + v);
    /* e components: bilinear in the two transverse directions (load_interpolator.cxx:73-103) */
    for (int X = 0; X < 3; X++) {
      const int Y = (X + 1) % 3, Z = (X + 2) % 3;
      const float w0 = CF(f, v, EX + X), w1 = CF(f, v + d.st[Y], EX + X), w2 = CF(f, v + d.st[Z], EX + X),
                  w3 = CF(f, v + d.st[Y] + d.st[Z], EX + X);
      o[4 * X + 0] = fourth * ((w3 + w0) + (w1 + w2));
      o[4 * X + 1] = fourth * ((w3 - w0) + (w1 - w2));
      o[4 * X + 2] = fourth * ((w3 - w0) - (w1 - w2));
      o[4 * X + 3] = fourth * ((w3 + w0) - (w1 + w2));
    }
    /* b components: linear along their own direction (:105-121) */
    for (int X = 0; X < 3; X++) {
      const float w0 = CF(f, v, CBX + X), w1 = CF(f, v + d.st[X], CBX + X);
      o[12 + 2 * X] = half * (w1 + w0);
      o[13 + 2 * X] = half * (w1 - w0);
    }
  }
}

void orc_clear_accumulators(vpb_accumulator_t *a, const vpb_grid_t *g) {
  memset(a, 0, sizeof(vpb_accumulator_t) * (size_t)dims_of(g).nv);
}

void orc_unload_accumulator(vpb_field_t *f, const vpb_accumulator_t *a, const vpb_grid_t *g) {
  const dims_t d = dims_of(g);
  const float c[3] = {0.25 * g->rdy * g->rdz / g->dt, 0.25 * g->rdz * g->rdx / g->dt, 0.25 * g->rdx * g->rdy / g->dt};
  const float *A = (const float *)a;
  box_t b = {{1, 1, 1}, {d.n[0] + 1, d.n[1] + 1, d.n[2] + 1}};
  FOR_BOX(b, x, y, z) {
    const long v = vox(&d, x, y, z);
    for (int X = 0; X < 3; X++) { /* unload_accumulator.cxx:49-51 */
      const int Y = (X + 1) % 3, Z = (X + 2) % 3;
      F(f, v, JFX + X) += c[X] * (A[12 * v + 4 * X] + A[12 * (v - d.st[Y]) + 4 * X + 1] + A[12 * (v - d.st[Z]) + 4 * X + 2] +
                                  A[12 * (v - d.st[Y] - d.st[Z]) + 4 * X + 3]);
    }
  }
}

/* ---------------------------------------------------------------------- */
/* local boundary conditions (local.c)                                     */
/* ---------------------------------------------------------------------- */

void orc_local_ghost_tang_b(vpb_field_t *f, const vpb_grid_t *g, int nproc) {
  const dims_t d = dims_of(g);
  const float cdt[3] = {g->cvac * g->dt * g->rdx, g->cvac * g->dt * g->rdy, g->cvac * g->dt * g->rdz};
  const float higend = (d.n[0] > 1 || d.n[1] > 1 || d.n[2] > 1) ? 1.03527618 : 1.;
  FOR_FACES(X, s) {
    const int bc = face_bc(g, X, s);
    if (!is_local(bc, nproc)) continue;
    const int Y = (X + 1) % 3, Z = (X + 2) % 3;
    const int ghost = s < 0 ? 0 : d.n[X] + 1, face = s < 0 ? 1 : d.n[X] + 1;
    const long in = -s * d.st[X]; /* ghost -> interior neighbour */
    box_t bY = plane(&d, X, ghost, 1, 0), bZ = plane(&d, X, ghost, 0, 1); /* cbY on ZY edges, cbZ on YZ edges */
    if (bc == vpb_pec_fields) {
      FOR_BOX(bY, x, y, z) { long v = vox(&d, x, y, z); F(f, v, CBX + Y) = F(f, v + in, CBX + Y); }
      FOR_BOX(bZ, x, y, z) { long v = vox(&d, x, y, z); F(f, v, CBX + Z) = F(f, v + in, CBX + Z); }
    } else if (bc == vpb_symmetric_fields || bc == vpb_pmc_fields) {
      FOR_BOX(bY, x, y, z) { long v = vox(&d, x, y, z); F(f, v, CBX + Y) = -F(f, v + in, CBX + Y); }
      FOR_BOX(bZ, x, y, z) { long v = vox(&d, x, y, z); F(f, v, CBX + Z) = -F(f, v + in, CBX + Z); }
    } else if (bc == vpb_absorb_fields) { /* 1st-order Higdon, local.c:84-111 */
      float drive = cdt[X] * higend;
      const float decay = (1 - drive) / (1 + drive);
      drive = 2 * drive / (1 + drive);
      const long to_face = (long)(face - ghost) * d.st[X];
      FOR_BOX(bY, x, y, z) {
        const long vg = vox(&d, x, y, z), vh = vg + in, vf = vg + to_face;
        float t1 = cdt[X] * (F(f, vf + in, EX + Z) - F(f, vf, EX + Z));
        t1 = s < 0 ? t1 : -t1;
        float t2 = F(f, vh + d.st[Z], EX + X);
        t2 = cdt[Z] * (t2 - F(f, vh, EX + X));
        F(f, vg, CBX + Y) = decay * F(f, vg, CBX + Y) + drive * F(f, vh, CBX + Y) - t1 + t2;
      }
      FOR_BOX(bZ, x, y, z) {
        const long vg = vox(&d, x, y, z), vh = vg + in, vf = vg + to_face;
        float t1 = cdt[X] * (F(f, vf + in, EX + Y) - F(f, vf, EX + Y));
        t1 = s < 0 ? t1 : -t1;
        float t2 = F(f, vh + d.st[Y], EX + X);
        t2 = cdt[Y] * (t2 - F(f, vh, EX + X));
        F(f, vg, CBX + Z) = decay * F(f, vg, CBX + Z) + drive * F(f, vh, CBX + Z) + t1 - t2;
      }
    }
  }
}

void orc_local_ghost_norm_e(vpb_field_t *f, const vpb_grid_t *g, int nproc) {
  const dims_t d = dims_of(g);
  FOR_FACES(X, s) {
    const int bc = face_bc(g, X, s);
    if (!is_local(bc, nproc)) continue;
    const int ghost = s < 0 ? 0 : d.n[X] + 1;
    const long in = -s * d.st[X];
    box_t b = plane(&d, X, ghost, 1, 1);
    FOR_BOX(b, x, y, z) {
      const long v = vox(&d, x, y, z);
      if (bc == vpb_pec_fields) {
        F(f, v, EX + X) = F(f, v + in, EX + X);
        F(f, v, TCAX + X) = F(f, v + in, TCAX + X);
      } else if (bc == vpb_symmetric_fields || bc == vpb_pmc_fields) {
        F(f, v, EX + X) = -F(f, v + in, EX + X);
        F(f, v, TCAX + X) = -F(f, v + in, TCAX + X);
      } else if (bc == vpb_absorb_fields) {
        F(f, v, EX + X) = 2 * F(f, v + in, EX + X) - F(f, v + 2 * in, EX + X);
        F(f, v, TCAX + X) = 2 * F(f, v + in, TCAX + X) - F(f, v + 2 * in, TCAX + X);
      }
    }
  }
}

void orc_local_ghost_div_b(vpb_field_t *f, const vpb_grid_t *g, int nproc) {
  const dims_t d = dims_of(g);
  FOR_FACES(X, s) {
    const int bc = face_bc(g, X, s);
    if (!is_local(bc, nproc)) continue;
    const int ghost = s < 0 ? 0 : d.n[X] + 1;
    const long in = -s * d.st[X];
    box_t b = plane(&d, X, ghost, 0, 0);
    FOR_BOX(b, x, y, z) {
      const long v = vox(&d, x, y, z);
      if (bc == vpb_pec_fields) F(f, v, DIVB) = F(f, v + in, DIVB);
      else if (bc == vpb_symmetric_fields || bc == vpb_pmc_fields) F(f, v, DIVB) = -F(f, v + in, DIVB);
      else if (bc == vpb_absorb_fields) F(f, v, DIVB) = 0;
    }
  }
}

/* what a local_adjust does to a value on the boundary plane */
enum { KEEP = 0, ZERO, DOUBLE_IT };
static void adjust_plane(vpb_field_t *f, const dims_t *d, box_t b, int comp, int op) {
  if (op == KEEP) return;
  FOR_BOX(b, x, y, z) {
    const long v = vox(d, x, y, z);
    if (op == ZERO) F(f, v, comp) = 0;
    else F(f, v, comp) *= 2;
  }
}
static int bc_slot(int bc) { /* 0 pec, 1 symmetric, 2 pmc, 3 absorb */
  return bc == vpb_pec_fields ? 0 : bc == vpb_symmetric_fields ? 1 : bc == vpb_pmc_fields ? 2 : bc == vpb_absorb_fields ? 3 : -1;
}

void orc_local_adjust_tang_e(vpb_field_t *f, const vpb_grid_t *g, int nproc) {
  const dims_t d = dims_of(g);
  FOR_FACES(X, s) {
    const int bc = face_bc(g, X, s);
    if (!is_local(bc, nproc) || bc != vpb_pec_fields) continue;
    const int Y = (X + 1) % 3, Z = (X + 2) % 3, face = s < 0 ? 1 : d.n[X] + 1;
    adjust_plane(f, &d, plane(&d, X, face, 0, 1), EX + Y, ZERO);
    adjust_plane(f, &d, plane(&d, X, face, 0, 1), TCAX + Y, ZERO);
    adjust_plane(f, &d, plane(&d, X, face, 1, 0), EX + Z, ZERO);
    adjust_plane(f, &d, plane(&d, X, face, 1, 0), TCAX + Z, ZERO);
  }
}

void orc_local_adjust_norm_b(vpb_field_t *f, const vpb_grid_t *g, int nproc) {
  const dims_t d = dims_of(g);
  FOR_FACES(X, s) {
    const int bc = face_bc(g, X, s);
    if (!is_local(bc, nproc) || bc != vpb_symmetric_fields) continue;
    adjust_plane(f, &d, plane(&d, X, s < 0 ? 1 : d.n[X] + 1, 0, 0), CBX + X, ZERO);
  }
}

void orc_local_adjust_div_e(vpb_field_t *f, const vpb_grid_t *g, int nproc) {
  const dims_t d = dims_of(g);
  FOR_FACES(X, s) {
    const int bc = face_bc(g, X, s);
    if (!is_local(bc, nproc) || !(bc == vpb_pec_fields || bc == vpb_absorb_fields)) continue;
    adjust_plane(f, &d, plane(&d, X, s < 0 ? 1 : d.n[X] + 1, 1, 1), DIVE, ZERO);
  }
}

void orc_local_adjust_jf(vpb_field_t *f, const vpb_grid_t *g, int nproc) {
  static const int op[4] = {ZERO, DOUBLE_IT, DOUBLE_IT, DOUBLE_IT};
  const dims_t d = dims_of(g);
  FOR_FACES(X, s) {
    const int bc = face_bc(g, X, s);
    if (!is_local(bc, nproc) || bc_slot(bc) < 0) continue;
    const int Y = (X + 1) % 3, Z = (X + 2) % 3, face = s < 0 ? 1 : d.n[X] + 1;
    adjust_plane(f, &d, plane(&d, X, face, 0, 1), JFX + Y, op[bc_slot(bc)]);
    adjust_plane(f, &d, plane(&d, X, face, 1, 0), JFX + Z, op[bc_slot(bc)]);
  }
}

void orc_local_adjust_rhof(vpb_field_t *f, const vpb_grid_t *g, int nproc) {
  static const int op[4] = {ZERO, DOUBLE_IT, DOUBLE_IT, DOUBLE_IT};
  const dims_t d = dims_of(g);
  FOR_FACES(X, s) {
    const int bc = face_bc(g, X, s);
    if (!is_local(bc, nproc) || bc_slot(bc) < 0) continue;
    adjust_plane(f, &d, plane(&d, X, s < 0 ? 1 : d.n[X] + 1, 1, 1), RHOF, op[bc_slot(bc)]);
  }
}

void orc_local_adjust_rhob(vpb_field_t *f, const vpb_grid_t *g, int nproc) {
  const dims_t d = dims_of(g);
  FOR_FACES(X, s) {
    const int bc = face_bc(g, X, s);
    if (!is_local(bc, nproc) || bc != vpb_pec_fields) continue;
    adjust_plane(f, &d, plane(&d, X, s < 0 ? 1 : d.n[X] + 1, 1, 1), RHOB, ZERO);
  }
}

/* ---------------------------------------------------------------------- */
/* face messages (remote.c)                                                */
/* ---------------------------------------------------------------------- */

int orc_face_message_floats(int kind, int face, const vpb_grid_t *g) {
  const dims_t d = dims_of(g);
  const int X = face % 3, nY = d.n[(X + 1) % 3], nZ = d.n[(X + 2) % 3];
  switch (kind) {
  case ORC_GHOST_TANG_B: case ORC_SYNC_JF: return 1 + nY * (nZ + 1) + nZ * (nY + 1);
  case ORC_GHOST_NORM_E: return 1 + (nY + 1) * (nZ + 1);
  case ORC_GHOST_DIV_B: return 1 + nY * nZ;
  case ORC_SYNC_RHO: return 1 + 2 * (nY + 1) * (nZ + 1);
  case ORC_SYNC_TEB: return 2 * nY * (nZ + 1) + 2 * nZ * (nY + 1) + nY * nZ;
  }
  return 0;
}

int orc_face_pack(int kind, int face, const vpb_field_t *f, const vpb_grid_t *g, float *buf) {
  const dims_t d = dims_of(g);
  const int X = face % 3, s = face < 3 ? -1 : 1, Y = (X + 1) % 3, Z = (X + 2) % 3;
  float *p = buf;
  /* ghost fills read the last interior plane; synchronisations read the shared plane */
  const int inner = s < 0 ? 1 : d.n[X], shared = s < 0 ? 1 : d.n[X] + 1;
  if (kind != ORC_SYNC_TEB) *p++ = cell_size(g, X);
  switch (kind) {
  case ORC_GHOST_TANG_B: { /* remote.c:79-88 */
    box_t b1 = plane(&d, X, inner, 1, 0), b2 = plane(&d, X, inner, 0, 1);
    FOR_BOX(b1, x, y, z) *p++ = CF(f, vox(&d, x, y, z), CBX + Y);
    FOR_BOX(b2, x, y, z) *p++ = CF(f, vox(&d, x, y, z), CBX + Z);
  } break;
  case ORC_GHOST_NORM_E: { /* :153-161 */
    box_t b = plane(&d, X, inner, 1, 1);
    FOR_BOX(b, x, y, z) *p++ = CF(f, vox(&d, x, y, z), EX + X);
  } break;
  case ORC_GHOST_DIV_B: { /* :226-234 */
    box_t b = plane(&d, X, inner, 0, 0);
    FOR_BOX(b, x, y, z) *p++ = CF(f, vox(&d, x, y, z), DIVB);
  } break;
  case ORC_SYNC_JF: { /* :434-445 */
    box_t b1 = plane(&d, X, shared, 0, 1), b2 = plane(&d, X, shared, 1, 0);
    FOR_BOX(b1, x, y, z) *p++ = CF(f, vox(&d, x, y, z), JFX + Y);
    FOR_BOX(b2, x, y, z) *p++ = CF(f, vox(&d, x, y, z), JFX + Z);
  } break;
  case ORC_SYNC_RHO: { /* :553-566 */
    box_t b = plane(&d, X, shared, 1, 1);
    FOR_BOX(b, x, y, z) { const long v = vox(&d, x, y, z); *p++ = CF(f, v, RHOF); *p++ = CF(f, v, RHOB); }
  } break;
  case ORC_SYNC_TEB: { /* :320-339 */
    box_t b0 = plane(&d, X, shared, 0, 0), b1 = plane(&d, X, shared, 0, 1), b2 = plane(&d, X, shared, 1, 0);
    FOR_BOX(b0, x, y, z) *p++ = CF(f, vox(&d, x, y, z), CBX + X);
    FOR_BOX(b1, x, y, z) { const long v = vox(&d, x, y, z); *p++ = CF(f, v, EX + Y); *p++ = CF(f, v, TCAX + Y); }
    FOR_BOX(b2, x, y, z) { const long v = vox(&d, x, y, z); *p++ = CF(f, v, EX + Z); *p++ = CF(f, v, TCAX + Z); }
  } break;
  }
  return (int)(p - buf);
}

double orc_face_unpack(int kind, int face, vpb_field_t *f, const vpb_grid_t *g, const float *buf) {
  const dims_t d = dims_of(g);
  const int X = face % 3, s = face < 3 ? -1 : 1, Y = (X + 1) % 3, Z = (X + 2) % 3;
  const float *p = buf;
  const float dX = cell_size(g, X);
  const int ghost = s < 0 ? 0 : d.n[X] + 1, shared = s < 0 ? 1 : d.n[X] + 1;
  const long in = -s * d.st[X];
  double err = 0;
  if (kind == ORC_GHOST_TANG_B || kind == ORC_GHOST_NORM_E || kind == ORC_GHOST_DIV_B) {
    float lw = *p++;
    const float rw = (2. * dX) / (lw + dX); /* remote.c:108-109 */
    lw = (lw - dX) / (lw + dX);
    if (kind == ORC_GHOST_TANG_B) {
      box_t b1 = plane(&d, X, ghost, 1, 0), b2 = plane(&d, X, ghost, 0, 1);
      FOR_BOX(b1, x, y, z) { const long v = vox(&d, x, y, z); F(f, v, CBX + Y) = rw * (*p++) + lw * F(f, v + in, CBX + Y); }
      FOR_BOX(b2, x, y, z) { const long v = vox(&d, x, y, z); F(f, v, CBX + Z) = rw * (*p++) + lw * F(f, v + in, CBX + Z); }
    } else if (kind == ORC_GHOST_NORM_E) {
      box_t b = plane(&d, X, ghost, 1, 1);
      FOR_BOX(b, x, y, z) { const long v = vox(&d, x, y, z); F(f, v, EX + X) = rw * (*p++) + lw * F(f, v + in, EX + X); }
    } else {
      box_t b = plane(&d, X, ghost, 0, 0);
      FOR_BOX(b, x, y, z) { const long v = vox(&d, x, y, z); F(f, v, DIVB) = rw * (*p++) + lw * F(f, v + in, DIVB); }
    }
  } else if (kind == ORC_SYNC_JF) { /* remote.c:447-465 */
    float rw = *p++, lw = rw + dX;
    rw /= lw;
    lw = dX / lw;
    lw += lw;
    rw += rw;
    box_t b1 = plane(&d, X, shared, 0, 1), b2 = plane(&d, X, shared, 1, 0);
    FOR_BOX(b1, x, y, z) { const long v = vox(&d, x, y, z); F(f, v, JFX + Y) = lw * F(f, v, JFX + Y) + rw * (*p++); }
    FOR_BOX(b2, x, y, z) { const long v = vox(&d, x, y, z); F(f, v, JFX + Z) = lw * F(f, v, JFX + Z) + rw * (*p++); }
  } else if (kind == ORC_SYNC_RHO) { /* :568-584 */
    float hrw = *p++, hlw = hrw + dX;
    hrw /= hlw;
    hlw = dX / hlw;
    const float lw = hlw + hlw, rw = hrw + hrw;
    box_t b = plane(&d, X, shared, 1, 1);
    FOR_BOX(b, x, y, z) {
      const long v = vox(&d, x, y, z);
      F(f, v, RHOF) = lw * F(f, v, RHOF) + rw * (*p++);
      F(f, v, RHOB) = hlw * F(f, v, RHOB) + hrw * (*p++);
    }
  } else if (kind == ORC_SYNC_TEB) { /* :341-372, arithmetic in double */
    double w1, w2;
    box_t b0 = plane(&d, X, shared, 0, 0), b1 = plane(&d, X, shared, 0, 1), b2 = plane(&d, X, shared, 1, 0);
    FOR_BOX(b0, x, y, z) {
      const long v = vox(&d, x, y, z);
      w1 = *p++; w2 = F(f, v, CBX + X); F(f, v, CBX + X) = 0.5 * (w1 + w2); err += (w1 - w2) * (w1 - w2);
    }
    for (int pass = 0; pass < 2; pass++) {
      const int C = pass ? Z : Y;
      box_t b = pass ? b2 : b1;
      FOR_BOX(b, x, y, z) {
        const long v = vox(&d, x, y, z);
        w1 = *p++; w2 = F(f, v, EX + C); F(f, v, EX + C) = 0.5 * (w1 + w2); err += (w1 - w2) * (w1 - w2);
        w1 = *p++; w2 = F(f, v, TCAX + C); F(f, v, TCAX + C) = 0.5 * (w1 + w2);
      }
    }
  }
  return err;
}

/* Every remote face is this rank itself (periodic): the message I send through
 * face F arrives through my opposite face. */
static void self_exchange_all(int kind, vpb_field_t *f, const vpb_grid_t *g, int stage /*0 pack,1 unpack*/, float **bufs) {
  for (int face = 0; face < 6; face++) {
    const int X = face % 3, s = face < 3 ? -1 : 1;
    if (!is_remote(face_bc(g, X, s), 1)) continue;
    if (stage == 0) {
      bufs[face] = (float *)malloc(sizeof(float) * (size_t)orc_face_message_floats(kind, face, g));
      orc_face_pack(kind, face, f, g, bufs[face]);
    }
  }
  if (stage == 1) {
    /* the reference completes receives in port order -x,-y,-z,+x,+y,+z, and port
     * (-1,0,0) carries what arrived from the +x side (grid_comm.c:13-17) */
    static const int order[6] = {3, 4, 5, 0, 1, 2};
    for (int k = 0; k < 6; k++) {
      const int face = order[k], opp = (face + 3) % 6;
      if (!bufs[opp]) continue;
      if (is_remote(face_bc(g, face % 3, face < 3 ? -1 : 1), 1)) orc_face_unpack(kind, face, f, g, bufs[opp]);
    }
    for (int face = 0; face < 6; face++) { free(bufs[face]); bufs[face] = NULL; }
  }
}

/* x pass, then y, then z: edges and corners ride along (remote.c:281-296) */
static double self_sync_passes(int kind, vpb_field_t *f, const vpb_grid_t *g) {
  double err = 0;
  for (int X = 0; X < 3; X++) {
    if (!is_remote(face_bc(g, X, -1), 1) && !is_remote(face_bc(g, X, 1), 1)) continue;
    float *lo = (float *)malloc(sizeof(float) * (size_t)orc_face_message_floats(kind, X, g));
    float *hi = (float *)malloc(sizeof(float) * (size_t)orc_face_message_floats(kind, X + 3, g));
    orc_face_pack(kind, X, f, g, lo);
    orc_face_pack(kind, X + 3, f, g, hi);
    if (is_remote(face_bc(g, X, 1), 1)) err += orc_face_unpack(kind, X + 3, f, g, lo);
    if (is_remote(face_bc(g, X, -1), 1)) err += orc_face_unpack(kind, X, f, g, hi);
    free(lo); free(hi);
  }
  return err;
}

/* ---------------------------------------------------------------------- */
/* field advance                                                           */
/* ---------------------------------------------------------------------- */

void orc_advance_b(vpb_field_t *f, const vpb_grid_t *g, float frac, int nproc) {
  const dims_t d = dims_of(g);
  float p[3];
  for (int X = 0; X < 3; X++) p[X] = (d.n[X] > 1) ? frac * g->cvac * g->dt * rcell(g, X) : 0;
  for (int X = 0; X < 3; X++) { /* cbX lives on X faces: X in 1..nX+1, others 1..n (advance_b.c:117-158) */
    const int Y = (X + 1) % 3, Z = (X + 2) % 3;
    box_t b = {{1, 1, 1}, {d.n[0], d.n[1], d.n[2]}};
    b.hi[X] += 1;
    FOR_BOX(b, x, y, z) {
      const long v = vox(&d, x, y, z);
      F(f, v, CBX + X) -= (p[Y] * (F(f, v + d.st[Y], EX + Z) - F(f, v, EX + Z)) - p[Z] * (F(f, v + d.st[Z], EX + Y) - F(f, v, EX + Y)));
    }
  }
  orc_local_adjust_norm_b(f, g, nproc);
}

void orc_advance_e_update(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g, int vacuum) {
  const dims_t d = dims_of(g);
  const float damp = g->damp, cj = g->dt / g->eps0;
  float p[3];
  for (int X = 0; X < 3; X++) p[X] = (d.n[X] > 1) ? (1 + damp) * g->cvac * g->dt * rcell(g, X) : 0;
  for (int X = 0; X < 3; X++) { /* eX lives on X edges: X in 1..nX, others 1..n+1 (advance_e.c:98-100) */
    const int Y = (X + 1) % 3, Z = (X + 2) % 3;
    box_t b = {{1, 1, 1}, {d.n[0] + 1, d.n[1] + 1, d.n[2] + 1}};
    b.hi[X] -= 1;
    FOR_BOX(b, x, y, z) {
      const long v = vox(&d, x, y, z), vy = v - d.st[Y], vz = v - d.st[Z];
      if (vacuum) { /* vacuum/vfa_advance_e.c:7-9 */
        F(f, v, EX + X) += ((p[Y] * (F(f, v, CBX + Z) - F(f, vy, CBX + Z)) - p[Z] * (F(f, v, CBX + Y) - F(f, vz, CBX + Y))) -
                            cj * F(f, v, JFX + X));
      } else { /* advance_e.c:8-25 */
        const float *rmuZ0 = &m[mat_id(f, v, FMAT + Z)].rmux, *rmuZy = &m[mat_id(f, vy, FMAT + Z)].rmux;
        const float *rmuY0 = &m[mat_id(f, v, FMAT + Y)].rmux, *rmuYz = &m[mat_id(f, vz, FMAT + Y)].rmux;
        const float *dec = &m[mat_id(f, v, EMAT + X)].decayx;
        F(f, v, TCAX + X) = (p[Y] * (F(f, v, CBX + Z) * rmuZ0[Z] - F(f, vy, CBX + Z) * rmuZy[Z]) -
                             p[Z] * (F(f, v, CBX + Y) * rmuY0[Y] - F(f, vz, CBX + Y) * rmuYz[Y])) -
                            damp * F(f, v, TCAX + X);
        F(f, v, EX + X) = dec[2 * X] * F(f, v, EX + X) + dec[2 * X + 1] * (F(f, v, TCAX + X) - cj * F(f, v, JFX + X));
      }
    }
  }
}

void orc_advance_e(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g, int vacuum) {
  float *bufs[6] = {0, 0, 0, 0, 0, 0};
  self_exchange_all(ORC_GHOST_TANG_B, f, g, 0, bufs); /* advance_e.c:114-115 */
  orc_local_ghost_tang_b(f, g, 1);
  self_exchange_all(ORC_GHOST_TANG_B, f, g, 1, bufs); /* :197 */
  orc_advance_e_update(f, m, g, vacuum);
  orc_local_adjust_tang_e(f, g, 1);
}

void orc_curl_b_update(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  const dims_t d = dims_of(g);
  float p[3];
  for (int X = 0; X < 3; X++) p[X] = (d.n[X] > 1) ? g->cvac * g->dt * rcell(g, X) : 0;
  for (int X = 0; X < 3; X++) {
    const int Y = (X + 1) % 3, Z = (X + 2) % 3;
    box_t b = {{1, 1, 1}, {d.n[0] + 1, d.n[1] + 1, d.n[2] + 1}};
    b.hi[X] -= 1;
    FOR_BOX(b, x, y, z) { /* compute_curl_b.c:8-18 */
      const long v = vox(&d, x, y, z), vy = v - d.st[Y], vz = v - d.st[Z];
      const float *rmuZ0 = &m[mat_id(f, v, FMAT + Z)].rmux, *rmuZy = &m[mat_id(f, vy, FMAT + Z)].rmux;
      const float *rmuY0 = &m[mat_id(f, v, FMAT + Y)].rmux, *rmuYz = &m[mat_id(f, vz, FMAT + Y)].rmux;
      F(f, v, TCAX + X) = p[Y] * (F(f, v, CBX + Z) * rmuZ0[Z] - F(f, vy, CBX + Z) * rmuZy[Z]) -
                          p[Z] * (F(f, v, CBX + Y) * rmuY0[Y] - F(f, vz, CBX + Y) * rmuYz[Y]);
    }
  }
}

void orc_compute_curl_b(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  float *bufs[6] = {0, 0, 0, 0, 0, 0};
  self_exchange_all(ORC_GHOST_TANG_B, f, g, 0, bufs);
  orc_local_ghost_tang_b(f, g, 1);
  self_exchange_all(ORC_GHOST_TANG_B, f, g, 1, bufs);
  orc_curl_b_update(f, m, g);
}

void orc_synchronize_jf(vpb_field_t *f, const vpb_grid_t *g) {
  orc_local_adjust_jf(f, g, 1);
  self_sync_passes(ORC_SYNC_JF, f, g);
}

void orc_synchronize_rho(vpb_field_t *f, const vpb_grid_t *g) {
  orc_local_adjust_rhof(f, g, 1);
  orc_local_adjust_rhob(f, g, 1);
  self_sync_passes(ORC_SYNC_RHO, f, g);
}

double orc_synchronize_tang_e_norm_b(vpb_field_t *f, const vpb_grid_t *g) {
  orc_local_adjust_tang_e(f, g, 1);
  orc_local_adjust_norm_b(f, g, 1);
  return self_sync_passes(ORC_SYNC_TEB, f, g);
}

/* rhob_mode 0: div_e_err (compute_div_e_err.c:7-11); 1: rhob (compute_rhob.c:8-12) */
void orc_div_e_err_update(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g, int rhob_mode) {
  const dims_t d = dims_of(g);
  float p[3];
  for (int X = 0; X < 3; X++) p[X] = (d.n[X] > 1) ? (rhob_mode ? g->eps0 * rcell(g, X) : rcell(g, X)) : 0;
  const float cj = 1. / g->eps0;
  box_t b = {{1, 1, 1}, {d.n[0] + 1, d.n[1] + 1, d.n[2] + 1}};
  FOR_BOX(b, x, y, z) {
    const long v = vox(&d, x, y, z);
    float t[3];
    for (int X = 0; X < 3; X++) {
      const long vm = v - d.st[X];
      const float *eps0 = &m[mat_id(f, v, EMAT + X)].epsx, *eps1 = &m[mat_id(f, vm, EMAT + X)].epsx;
      t[X] = p[X] * (eps0[X] * F(f, v, EX + X) - eps1[X] * F(f, vm, EX + X));
    }
    const float nc = m[mat_id(f, v, NMAT)].nonconductive;
    if (rhob_mode) F(f, v, RHOB) = nc * (t[0] + t[1] + t[2] - F(f, v, RHOF));
    else F(f, v, DIVE) = nc * (t[0] + t[1] + t[2] - cj * (F(f, v, RHOF) + F(f, v, RHOB)));
  }
}

void orc_compute_div_e_err(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  float *bufs[6] = {0, 0, 0, 0, 0, 0};
  self_exchange_all(ORC_GHOST_NORM_E, f, g, 0, bufs);
  orc_local_ghost_norm_e(f, g, 1);
  self_exchange_all(ORC_GHOST_NORM_E, f, g, 1, bufs);
  orc_div_e_err_update(f, m, g, 0);
  orc_local_adjust_div_e(f, g, 1);
}

void orc_compute_rhob(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  float *bufs[6] = {0, 0, 0, 0, 0, 0};
  self_exchange_all(ORC_GHOST_NORM_E, f, g, 0, bufs);
  orc_local_ghost_norm_e(f, g, 1);
  self_exchange_all(ORC_GHOST_NORM_E, f, g, 1, bufs);
  orc_div_e_err_update(f, m, g, 1);
  orc_local_adjust_rhob(f, g, 1);
}

static void marder_coeff(const vpb_grid_t *g, float p[3]) { /* clean_div_e.c:39-45 */
  const dims_t d = dims_of(g);
  for (int X = 0; X < 3; X++) p[X] = (d.n[X] > 1) ? rcell(g, X) : 0;
  const float alphadt = 0.3888889 / (p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  for (int X = 0; X < 3; X++) p[X] *= alphadt;
}

void orc_clean_div_e_update(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  const dims_t d = dims_of(g);
  float p[3];
  marder_coeff(g, p);
  for (int X = 0; X < 3; X++) {
    box_t b = {{1, 1, 1}, {d.n[0] + 1, d.n[1] + 1, d.n[2] + 1}};
    b.hi[X] -= 1;
    FOR_BOX(b, x, y, z) { /* clean_div_e.c:6-13 */
      const long v = vox(&d, x, y, z);
      const float *drv = &m[mat_id(f, v, EMAT + X)].decayx;
      F(f, v, EX + X) += drv[2 * X + 1] * p[X] * (F(f, v + d.st[X], DIVE) - F(f, v, DIVE));
    }
  }
}

void orc_clean_div_e(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  orc_clean_div_e_update(f, m, g);
  orc_local_adjust_tang_e(f, g, 1);
}

void orc_compute_div_b_err(vpb_field_t *f, const vpb_grid_t *g) {
  const dims_t d = dims_of(g);
  float p[3];
  for (int X = 0; X < 3; X++) p[X] = (d.n[X] > 1) ? rcell(g, X) : 0;
  box_t b = {{1, 1, 1}, {d.n[0], d.n[1], d.n[2]}};
  FOR_BOX(b, x, y, z) { /* compute_div_b_err.c:44-46 */
    const long v = vox(&d, x, y, z);
    F(f, v, DIVB) = p[0] * (F(f, v + d.st[0], CBX) - F(f, v, CBX)) + p[1] * (F(f, v + d.st[1], CBY) - F(f, v, CBY)) +
                    p[2] * (F(f, v + d.st[2], CBZ) - F(f, v, CBZ));
  }
}

void orc_clean_div_b_update(vpb_field_t *f, const vpb_grid_t *g) {
  const dims_t d = dims_of(g);
  float p[3];
  marder_coeff(g, p);
  for (int X = 0; X < 3; X++) {
    box_t b = {{1, 1, 1}, {d.n[0], d.n[1], d.n[2]}};
    b.hi[X] += 1;
    FOR_BOX(b, x, y, z) { /* clean_div_b.c:6-8 */
      const long v = vox(&d, x, y, z);
      F(f, v, CBX + X) += p[X] * (F(f, v, DIVB) - F(f, v - d.st[X], DIVB));
    }
  }
}

void orc_clean_div_b(vpb_field_t *f, const vpb_grid_t *g) {
  float *bufs[6] = {0, 0, 0, 0, 0, 0};
  self_exchange_all(ORC_GHOST_DIV_B, f, g, 0, bufs);
  orc_local_ghost_div_b(f, g, 1);
  self_exchange_all(ORC_GHOST_DIV_B, f, g, 1, bufs);
  orc_clean_div_b_update(f, g);
  orc_local_adjust_norm_b(f, g, 1);
}

void orc_clear_jf(vpb_field_t *f, const vpb_grid_t *g) {
  const long nv = dims_of(g).nv;
  for (long v = 0; v < nv; v++) f[v].jfx = f[v].jfy = f[v].jfz = 0;
}

void orc_clear_rhof(vpb_field_t *f, const vpb_grid_t *g) {
  const long nv = dims_of(g).nv;
  for (long v = 0; v < nv; v++) f[v].rhof = 0;
}

void orc_energy_f(double en[6], const vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g) {
  const dims_t d = dims_of(g);
  for (int k = 0; k < 6; k++) en[k] = 0;
  box_t b = {{1, 1, 1}, {d.n[0], d.n[1], d.n[2]}};
  FOR_BOX(b, x, y, z) {
    const long v = vox(&d, x, y, z);
    for (int X = 0; X < 3; X++) { /* energy_f.c:51-69: 4 edges, 2 faces per cell */
      const int Y = (X + 1) % 3, Z = (X + 2) % 3;
      const long ve[4] = {v, v + d.st[Y], v + d.st[Z], v + d.st[Y] + d.st[Z]};
      float se[4];
      for (int k = 0; k < 4; k++) {
        const float *eps = &m[mat_id(f, ve[k], EMAT + X)].epsx;
        se[k] = eps[X] * CF(f, ve[k], EX + X) * CF(f, ve[k], EX + X);
      }
      en[X] += 0.25 * (se[0] + se[1] + se[2] + se[3]);
      const long vb[2] = {v, v + d.st[X]};
      float sb[2];
      for (int k = 0; k < 2; k++) {
        const float *rmu = &m[mat_id(f, vb[k], FMAT + X)].rmux;
        sb[k] = rmu[X] * CF(f, vb[k], CBX + X) * CF(f, vb[k], CBX + X);
      }
      en[3 + X] += 0.5 * (sb[0] + sb[1]);
    }
  }
  const double v0 = 0.5 * g->eps0 * g->dx * g->dy * g->dz;
  for (int k = 0; k < 6; k++) en[k] *= v0;
}

void orc_rms_div_e_err_local(double out[2], const vpb_field_t *f, const vpb_grid_t *g) {
  const dims_t d = dims_of(g);
  double err = 0;
  box_t b = {{1, 1, 1}, {d.n[0] + 1, d.n[1] + 1, d.n[2] + 1}};
  FOR_BOX(b, x, y, z) { /* compute_rms_div_e_err.c: interior float product, surface weights 1/2,1/4,1/8 */
    const int c[3] = {x, y, z};
    int nsurf = 0;
    for (int X = 0; X < 3; X++) nsurf += (c[X] == 1 || c[X] == d.n[X] + 1);
    const float e = CF(f, vox(&d, x, y, z), DIVE);
    if (nsurf == 0) err += e * e;
    else err += (nsurf == 1 ? 0.5 : nsurf == 2 ? 0.25 : 0.125) * (double)e * (double)e;
  }
  out[0] = err * g->dx * g->dy * g->dz;
  out[1] = g->nx * g->ny * g->nz * g->dx * g->dy * g->dz;
}

void orc_rms_div_b_err_local(double out[2], const vpb_field_t *f, const vpb_grid_t *g) {
  const dims_t d = dims_of(g);
  double err = 0;
  box_t b = {{1, 1, 1}, {d.n[0], d.n[1], d.n[2]}};
  FOR_BOX(b, x, y, z) { const float e = CF(f, vox(&d, x, y, z), DIVB); err += e * e; }
  out[0] = err * g->dx * g->dy * g->dz;
  out[1] = g->nx * g->ny * g->nz * g->dx * g->dy * g->dz;
}

void orc_material_coefficients(vpb_material_coefficient_t *mc0, const vpb_material_t *m_list, const vpb_grid_t *g) {
  for (const vpb_material_t *m = m_list; m; m = m->next) { /* sfa.c:127-168 */
    vpb_material_coefficient_t *mc = mc0 + m->id;
    const float eps[3] = {m->epsx, m->epsy, m->epsz}, sig[3] = {m->sigmax, m->sigmay, m->sigmaz};
    float a[3], *dd = &mc->decayx;
    for (int X = 0; X < 3; X++) {
      a[X] = (sig[X] * g->dt) / (eps[X] * g->eps0);
      dd[2 * X] = exp(-a[X]);
      if (a[X] == 0) dd[2 * X + 1] = 1. / eps[X];
      else if (dd[2 * X] == 0) dd[2 * X + 1] = 0;
      else dd[2 * X + 1] = 2. * exp(-0.5 * a[X]) * sinh(0.5 * a[X]) / (a[X] * eps[X]);
    }
    mc->rmux = 1. / m->mux; mc->rmuy = 1. / m->muy; mc->rmuz = 1. / m->muz;
    mc->nonconductive = (a[0] == 0 && a[1] == 0 && a[2] == 0) ? 1. : 0.;
    mc->epsx = m->epsx; mc->epsy = m->epsy; mc->epsz = m->epsz;
  }
}
