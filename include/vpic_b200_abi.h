/* vpic_b200_abi.h -- the drop-in boundary, part 1: data layouts.
 *
 * The reference (pdlfs/old-vpic) has no plugin loader: its hot path is a flat set
 * of extern "C" functions over raw arrays, replaced by LINK-TIME SYMBOL
 * SUBSTITUTION.  A replacement library therefore has to agree with
 * the reference on the byte layout of every struct that crosses that boundary.
 * This header re-declares those layouts (prefix vpb_ so it can be included next
 * to the reference's own headers) and pins every size/offset with static
 * asserts; tests/test_abi.py checks the same numbers against the reference
 * compiled from source (oracle/_ref, refh_layout()).
 *
 * Each declaration cites the reference declaration it mirrors (paths relative
 * to the reference root).
 */
#ifndef VPIC_B200_ABI_H
#define VPIC_B200_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define VPB_STATIC_ASSERT(c, m) static_assert(c, m)
extern "C" {
#else
#define VPB_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

/* src/species_advance/species_advance.h:28-34.  48 bytes: the 32 "hot" bytes the
 * push reads and writes, then two 64-bit tags (pdlfs particle-tracking addition)
 * that only sort/migration carry along. */
typedef struct vpb_particle {
  float   dx, dy, dz;   /* cell-local offsets on [-1,1] */
  int32_t i;            /* local voxel index, x-fastest, ghosts included */
  float   ux, uy, uz;   /* gamma*beta */
  float   q;            /* macro-particle charge */
  int64_t tag, tag2;
} vpb_particle_t;
VPB_STATIC_ASSERT(sizeof(vpb_particle_t) == 48, "particle_t is 48 B");
VPB_STATIC_ASSERT(offsetof(vpb_particle_t, i) == 12 && offsetof(vpb_particle_t, ux) == 16 &&
                  offsetof(vpb_particle_t, q) == 28 && offsetof(vpb_particle_t, tag) == 32, "particle_t layout");

/* src/species_advance/species_advance.h:39-42 */
typedef struct vpb_particle_mover {
  float   dispx, dispy, dispz;  /* remaining displacement, cell units */
  int32_t i;                    /* index of the particle in its species array */
} vpb_particle_mover_t;
VPB_STATIC_ASSERT(sizeof(vpb_particle_mover_t) == 16, "particle_mover_t is 16 B");

/* src/species_advance/species_advance.h:48-55 -- migration wire record (no tags) */
typedef struct vpb_particle_injector {
  float   dx, dy, dz;
  int32_t i;
  float   ux, uy, uz, q;
  float   dispx, dispy, dispz;
  int32_t sp_id;
} vpb_particle_injector_t;
VPB_STATIC_ASSERT(sizeof(vpb_particle_injector_t) == 48, "particle_injector_t is 48 B");

/* src/sf_interface/sf_interface.h:45-58 */
typedef struct vpb_interpolator {
  float ex, dexdy, dexdz, d2exdydz;
  float ey, deydz, deydx, d2eydzdx;
  float ez, dezdx, dezdy, d2ezdxdy;
  float cbx, dcbxdx;
  float cby, dcbydy;
  float cbz, dcbzdz;
  float _pad[2];
} vpb_interpolator_t;
VPB_STATIC_ASSERT(sizeof(vpb_interpolator_t) == 80, "interpolator_t is 80 B");

/* src/sf_interface/sf_interface.h:68-77 */
typedef struct vpb_accumulator {
  float jx[4];  /* jx0@(0,-1,-1) jx1@(0,1,-1) jx2@(0,-1,1) jx3@(0,1,1) */
  float jy[4];
  float jz[4];
} vpb_accumulator_t;
VPB_STATIC_ASSERT(sizeof(vpb_accumulator_t) == 48, "accumulator_t is 48 B");

/* src/sf_interface/sf_interface.h:28-38 */
typedef struct vpb_hydro {
  float jx, jy, jz, rho;
  float px, py, pz, ke;
  float txx, tyy, tzz;
  float tyz, tzx, txy;
  float _pad[2];
} vpb_hydro_t;
VPB_STATIC_ASSERT(sizeof(vpb_hydro_t) == 64, "hydro_t is 64 B");

/* src/field_advance/field_advance.h:159-171 (material_id = uint16_t, material.h:40) */
typedef struct vpb_field {
  float    ex, ey, ez, div_e_err;
  float    cbx, cby, cbz, div_b_err;
  float    tcax, tcay, tcaz, rhob;
  float    jfx, jfy, jfz, rhof;
  uint16_t ematx, ematy, ematz, nmat;
  uint16_t fmatx, fmaty, fmatz, cmat;
} vpb_field_t;
VPB_STATIC_ASSERT(sizeof(vpb_field_t) == 80, "field_t is 80 B");
VPB_STATIC_ASSERT(offsetof(vpb_field_t, cbx) == 16 && offsetof(vpb_field_t, tcax) == 32 &&
                  offsetof(vpb_field_t, jfx) == 48 && offsetof(vpb_field_t, ematx) == 64 &&
                  offsetof(vpb_field_t, fmatx) == 72, "field_t layout");

/* src/field_advance/standard/sfa_private.h:24-32 (opaque to callers; 64 B) */
typedef struct vpb_material_coefficient {
  float decayx, drivex;
  float decayy, drivey;
  float decayz, drivez;
  float rmux, rmuy, rmuz;
  float nonconductive;
  float epsx, epsy, epsz;
  float pad[3];
} vpb_material_coefficient_t;
VPB_STATIC_ASSERT(sizeof(vpb_material_coefficient_t) == 64, "material_coefficient_t is 64 B");

/* src/material/material.h:42-50 */
typedef struct vpb_material {
  uint16_t id;
  float epsx, epsy, epsz;
  float mux, muy, muz;
  float sigmax, sigmay, sigmaz;
  float zetax, zetay, zetaz;
  struct vpb_material *next;
  char name[1];
} vpb_material_t;

/* src/grid/grid.h:55,57-66: bc[] index and the boundary-condition codes */
#define VPB_BOUNDARY(i, j, k) (((i) + 1) + 3 * (((j) + 1) + 3 * ((k) + 1)))
enum {
  vpb_pec_fields = -1, vpb_symmetric_fields = -2, vpb_pmc_fields = -3, vpb_absorb_fields = -4,
  vpb_reflect_particles = -1, vpb_absorb_particles = -2
};

/* src/grid/grid.h:112-167.  `mp` is the reference's opaque communications handle
 * (util/mp/mp_handle.h:9); this library never dereferences it. */
typedef struct vpb_grid {
  void    *mp;
  float    dt, cvac, eps0;
  float    damp;
  float    x0, y0, z0;
  float    x1, y1, z1;
  float    dx, dy, dz;
  float    rdx, rdy, rdz;
  int      nx, ny, nz;
  int      bc[27];
  int64_t *range;      /* [nproc+1] global voxel-id range per rank */
  int64_t *neighbor;   /* [6*nvoxel] global voxel id, or <0 particle-bc code */
  int64_t  rangel, rangeh;
  int      nb;
  void    *boundary;   /* custom boundary handlers (host callbacks; grid.h:45-53) */
} vpb_grid_t;
VPB_STATIC_ASSERT(sizeof(vpb_grid_t) == 240, "grid_t is 240 B");
VPB_STATIC_ASSERT(offsetof(vpb_grid_t, dt) == 8 && offsetof(vpb_grid_t, x0) == 24 &&
                  offsetof(vpb_grid_t, dx) == 48 && offsetof(vpb_grid_t, rdx) == 60 &&
                  offsetof(vpb_grid_t, nx) == 72 && offsetof(vpb_grid_t, bc) == 84 &&
                  offsetof(vpb_grid_t, range) == 192 && offsetof(vpb_grid_t, neighbor) == 200 &&
                  offsetof(vpb_grid_t, rangel) == 208 && offsetof(vpb_grid_t, nb) == 224 &&
                  offsetof(vpb_grid_t, boundary) == 232, "grid_t layout");

/* src/grid/grid.h:32-53: the deck's custom particle-boundary handlers, grid_t.boundary[0..nb).  Host callbacks: the
 * reference-named boundary_p() calls them on the host exactly where boundary_p.c:271-277 does. */
struct vpb_species;
typedef void (*vpb_boundary_handler_t)(void *params, vpb_particle_t *r, vpb_particle_mover_t *pm, vpb_field_t *f, vpb_accumulator_t *a,
                                       const vpb_grid_t *g, struct vpb_species *s, vpb_particle_injector_t **ppi, void *rng, int face);
typedef struct vpb_boundary {
  vpb_boundary_handler_t handler;
  char params[1024];     /* MAX_BOUNDARY_DATA_SIZE */
} vpb_boundary_t;
VPB_STATIC_ASSERT(sizeof(vpb_boundary_t) == 1032, "boundary_t is 1032 B");

/* src/species_advance/species_advance.h:61-93 */
typedef struct vpb_species {
  int32_t               id;
  int                   np, max_np;
  vpb_particle_t       *p;
  int                   nm, max_nm;
  vpb_particle_mover_t *pm;
  float                 q_m;
  int                   sort_interval;
  int                   sort_out_of_place;
  int                  *partition;   /* [nvoxel+1] first particle of each voxel after a sort */
  struct vpb_species   *next;
  char                  name[1];
} vpb_species_t;
VPB_STATIC_ASSERT(offsetof(vpb_species_t, p) == 16 && offsetof(vpb_species_t, pm) == 32 &&
                  offsetof(vpb_species_t, q_m) == 40 && offsetof(vpb_species_t, partition) == 56 &&
                  offsetof(vpb_species_t, next) == 64 && offsetof(vpb_species_t, name) == 72, "species_t layout");

/* src/field_advance/field_advance.h:185-302 -- the 20-entry kernel vtable that
 * decks select with finalize_field_advance(standard_field_advance). */
typedef struct vpb_field_advance_methods {
  vpb_field_t *(*new_field)(vpb_grid_t *g);
  void (*delete_field)(vpb_field_t *f);
  vpb_material_coefficient_t *(*new_material_coefficients)(vpb_grid_t *g, vpb_material_t *m_list);
  void (*delete_material_coefficients)(vpb_material_coefficient_t *mc);
  void (*advance_b)(vpb_field_t *f, const vpb_grid_t *g, float frac);
  void (*advance_e)(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g);
  void (*energy_f)(double *energy6, const vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g);
  void (*clear_jf)(vpb_field_t *f, const vpb_grid_t *g);
  void (*synchronize_jf)(vpb_field_t *f, const vpb_grid_t *g);
  void (*clear_rhof)(vpb_field_t *f, const vpb_grid_t *g);
  void (*synchronize_rho)(vpb_field_t *f, const vpb_grid_t *g);
  void (*compute_rhob)(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g);
  void (*compute_curl_b)(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g);
  double (*synchronize_tang_e_norm_b)(vpb_field_t *f, const vpb_grid_t *g);
  void (*compute_div_e_err)(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g);
  double (*compute_rms_div_e_err)(vpb_field_t *f, const vpb_grid_t *g);
  void (*clean_div_e)(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g);
  void (*compute_div_b_err)(vpb_field_t *f, const vpb_grid_t *g);
  double (*compute_rms_div_b_err)(vpb_field_t *f, const vpb_grid_t *g);
  void (*clean_div_b)(vpb_field_t *f, const vpb_grid_t *g);
} vpb_field_advance_methods_t;
VPB_STATIC_ASSERT(sizeof(vpb_field_advance_methods_t) == 160, "field_advance_methods_t has 20 entries");

/* src/field_advance/field_advance.h:307-312 */
typedef struct vpb_field_advance {
  vpb_field_t                 *f;
  vpb_material_coefficient_t  *m;
  vpb_grid_t                  *g;
  vpb_field_advance_methods_t  method[1];
} vpb_field_advance_t;

#ifdef __cplusplus
}
#endif
#endif /* VPIC_B200_ABI_H */
