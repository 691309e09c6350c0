/* TEST INFRASTRUCTURE -- CPU restatement of the reference's hot-path algorithms.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call
 * this; the product (libvpic_b200.so) never does.  Every function restates the
 * SCALAR flavour of the reference (IEEE sqrt/divide, no fused multiply-add;
 * SURVEY.md 8a "numerics") and cites the reference lines it follows.  The
 * restatement is pinned against the reference compiled from source
 * (oracle/_ref/libvpic_ref_scalar.so): tests/test_oracle_vs_ref.py demands
 * bit-identical output, and tests/golden/ holds fixtures generated from the
 * reference by tests/golden/make_golden.py.
 *
 * Structs are the ABI mirrors of include/vpic_b200_abi.h.
 */
#ifndef VPIC_ORACLE_H
#define VPIC_ORACLE_H
#include "../include/vpic_b200_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- particles -------------------------------------------------------- */
/* advance_p.cxx:9-183 (scalar pipeline over all np) + :399-472.  Returns nm. */
int orc_advance_p(vpb_particle_t *p0, int np, float q_m, vpb_particle_mover_t *pm, int max_nm,
                  vpb_accumulator_t *a0, const vpb_interpolator_t *f0, const vpb_grid_t *g);
/* move_p.c:20-136 */
int orc_move_p(vpb_particle_t *p0, vpb_particle_mover_t *pm, vpb_accumulator_t *a0, const vpb_grid_t *g);
/* center_p.cxx:5-70, uncenter_p.cxx:5-72 */
void orc_center_p(vpb_particle_t *p0, int np, float q_m, const vpb_interpolator_t *f0, const vpb_grid_t *g);
void orc_uncenter_p(vpb_particle_t *p0, int np, float q_m, const vpb_interpolator_t *f0, const vpb_grid_t *g);
/* energy_p.cxx:5-48,124-157 (single rank: no allreduce) */
double orc_energy_p(const vpb_particle_t *p0, int np, float q_m, const vpb_interpolator_t *f0, const vpb_grid_t *g);
/* rho_p.c:23-79 */
void orc_accumulate_rho_p(vpb_field_t *f, const vpb_particle_t *p0, int np, const vpb_grid_t *g);
/* oracle_hydro.c: hydro_p.c:24-161, sf_interface/hydro.c:30-184 */
void orc_clear_hydro(vpb_hydro_t *h, const vpb_grid_t *g);
void orc_accumulate_hydro_p(vpb_hydro_t *h0, const vpb_particle_t *p0, int n, float q_m, const vpb_interpolator_t *f0,
                            const vpb_grid_t *g);
void orc_local_adjust_hydro(vpb_hydro_t *h, const vpb_grid_t *g, int nproc);
void orc_synchronize_hydro(vpb_hydro_t *h, const vpb_grid_t *g, int rank, int nproc);
int orc_hydro_face_floats(int face, const vpb_grid_t *g);
void orc_hydro_face_pack(int face, const vpb_hydro_t *h, const vpb_grid_t *g, float *buf);
void orc_hydro_face_unpack(int face, vpb_hydro_t *h, const vpb_grid_t *g, const float *buf);
/* boundary_p.c:9-71 */
void orc_accumulate_rhob(vpb_field_t *f, const vpb_particle_t *p, const vpb_grid_t *g);
/* sort_p.c:16-77 (stable out-of-place counting sort); partition has nv+1 entries */
void orc_sort_p(const vpb_particle_t *in, vpb_particle_t *out, int np, int *partition, const vpb_grid_t *g);
/* boundary_p.c:150-329: classify the movers of one species by face.  For each
 * mover (walked in reverse, with back-fill) either absorbs the particle
 * (accumulate_rhob) or appends a particle_injector_t to out[face].  Returns
 * the new np; n_out[6] receives the injector counts. */
int orc_boundary_p_pack(vpb_particle_t *p0, int np, const vpb_particle_mover_t *pm, int nm, int sp_id,
                        vpb_field_t *f, const vpb_grid_t *g, int rank, int nproc,
                        vpb_particle_injector_t *out[6], int n_out[6]);
/* boundary_p.c:457-497: inject received particles (reverse order) and finish
 * their move.  Returns the number of movers appended to pm; *np is updated. */
int orc_boundary_p_inject(vpb_particle_t *p0, int *np, vpb_particle_mover_t *pm, int nm,
                          const vpb_particle_injector_t *in, int n_in, int sp_id,
                          vpb_accumulator_t *a0, const vpb_grid_t *g);

/* ---- species <-> field ------------------------------------------------ */
void orc_load_interpolator(vpb_interpolator_t *fi, const vpb_field_t *f, const vpb_grid_t *g);
void orc_clear_accumulators(vpb_accumulator_t *a, const vpb_grid_t *g);
void orc_unload_accumulator(vpb_field_t *f, const vpb_accumulator_t *a, const vpb_grid_t *g);

/* ---- field solve (standard field advance).  nproc/rank say which faces are
 * "remote" (0<=bc<nproc); remote faces are served by the pack/unpack pairs
 * below, which the caller wires either to itself (periodic single rank) or to
 * another rank. -------------------------------------------------------- */
void orc_advance_b(vpb_field_t *f, const vpb_grid_t *g, float frac, int nproc);
/* Ghost phases of advance_e are exposed so a test can interleave an exchange:
 *   orc_advance_e = [exchange tang_b] ; local_ghost_tang_b ; update ; local_adjust_tang_e */
void orc_advance_e_update(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g, int vacuum);
void orc_local_ghost_tang_b(vpb_field_t *f, const vpb_grid_t *g, int nproc);
void orc_local_ghost_norm_e(vpb_field_t *f, const vpb_grid_t *g, int nproc);
void orc_local_ghost_div_b(vpb_field_t *f, const vpb_grid_t *g, int nproc);
void orc_local_adjust_tang_e(vpb_field_t *f, const vpb_grid_t *g, int nproc);
void orc_local_adjust_norm_b(vpb_field_t *f, const vpb_grid_t *g, int nproc);
void orc_local_adjust_div_e(vpb_field_t *f, const vpb_grid_t *g, int nproc);
void orc_local_adjust_jf(vpb_field_t *f, const vpb_grid_t *g, int nproc);
void orc_local_adjust_rhof(vpb_field_t *f, const vpb_grid_t *g, int nproc);
void orc_local_adjust_rhob(vpb_field_t *f, const vpb_grid_t *g, int nproc);

/* Face messages (remote.c).  `face` = 0..5 for -x,-y,-z,+x,+y,+z; kind selects
 * the payload.  pack returns the number of floats written (message layout is
 * the reference's, including the leading cell size where it sends one);
 * unpack applies a message that arrived THROUGH face `face` (i.e. sent by the
 * neighbour on that side).  ORC_SYNC_TEB unpack returns the squared
 * desynchronisation it found (remote.c:344-370). */
enum { ORC_GHOST_TANG_B = 0, ORC_GHOST_NORM_E, ORC_GHOST_DIV_B, ORC_SYNC_JF, ORC_SYNC_RHO, ORC_SYNC_TEB };
int orc_face_message_floats(int kind, int face, const vpb_grid_t *g);
int orc_face_pack(int kind, int face, const vpb_field_t *f, const vpb_grid_t *g, float *buf);
double orc_face_unpack(int kind, int face, vpb_field_t *f, const vpb_grid_t *g, const float *buf);

/* Whole single-rank operations (every remote face is the rank itself, i.e.
 * periodic along that axis), composed from the pieces above exactly in the
 * reference's order. */
void orc_advance_e(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g, int vacuum);
void orc_synchronize_jf(vpb_field_t *f, const vpb_grid_t *g);
void orc_synchronize_rho(vpb_field_t *f, const vpb_grid_t *g);
double orc_synchronize_tang_e_norm_b(vpb_field_t *f, const vpb_grid_t *g);
void orc_compute_div_e_err(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g);
void orc_clean_div_e(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g);
void orc_compute_div_b_err(vpb_field_t *f, const vpb_grid_t *g);
void orc_clean_div_b(vpb_field_t *f, const vpb_grid_t *g);
void orc_compute_rhob(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g);
void orc_compute_curl_b(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g);
/* exterior-free pieces, used by the multi-rank tests between exchanges */
void orc_div_e_err_update(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g, int rhob_mode);
void orc_clean_div_e_update(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g);
void orc_clean_div_b_update(vpb_field_t *f, const vpb_grid_t *g);
void orc_curl_b_update(vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g);

void orc_clear_jf(vpb_field_t *f, const vpb_grid_t *g);
void orc_clear_rhof(vpb_field_t *f, const vpb_grid_t *g);
/* local (this rank's) sums: energy_f.c:13-91,168-175; rms pieces
 * compute_rms_div_e_err.c / compute_rms_div_b_err.c before the allreduce */
void orc_energy_f(double en[6], const vpb_field_t *f, const vpb_material_coefficient_t *m, const vpb_grid_t *g);
void orc_rms_div_e_err_local(double out[2], const vpb_field_t *f, const vpb_grid_t *g);
void orc_rms_div_b_err_local(double out[2], const vpb_field_t *f, const vpb_grid_t *g);
/* sfa.c:80-171 */
void orc_material_coefficients(vpb_material_coefficient_t *mc, const vpb_material_t *m_list, const vpb_grid_t *g);

#ifdef __cplusplus
}
#endif
#endif
