/* vpic_b200.h -- the drop-in boundary, part 2: entry points.
 *
 * libvpic_b200.so exports two layers, both plain C ABI (pointers and sizes only):
 *
 *  (A) REFERENCE-NAMED ENTRY POINTS with the reference's exact prototypes
 *      (advance_p, sort_p, load_interpolator, ... and the field-advance vtables).
 *      These are what gets linked INSTEAD of the reference's translation units
 *      (INTEGRATION.md shows the link line).  Pointers are whatever
 *      the caller has: plain host memory is staged host->device->host around the
 *      kernel; memory that came from vpb_malloc_managed()/util_malloc_aligned()
 *      or vpb_dev_alloc() is used in place.
 *
 *  (B) vpb_* DEVICE-RESIDENT ENTRY POINTS: same operations on device pointers,
 *      enqueued on the library's stream without host synchronisation, for callers
 *      that keep the whole state in HBM between steps (bench.py `value`, the
 *      step driver vpb_sim_*).
 *
 * Error behaviour mirrors the reference (util_base.h:213-219): invalid arguments
 * print "Error at file(line): msg" to stderr and exit(1); recoverable conditions
 * print a warning and continue.  There is NO CPU fallback: every entry point
 * fails loudly if no CUDA device is usable.
 */
#ifndef VPIC_B200_H
#define VPIC_B200_H

#include "vpic_b200_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------- */
/* (A) Reference-named entry points.  Compiled with the reference's own names  */
/* unless VPB_NO_REFERENCE_NAMES is defined by the includer (tests that also   */
/* include reference headers).                                                 */
/* ------------------------------------------------------------------------- */
#ifndef VPB_NO_REFERENCE_NAMES

/* src/species_advance/standard/spa.h:57-65 (advance_p.cxx:399-472).
 * Returns the number of movers left in pm[] (particles that hit something
 * move_p cannot resolve locally), in increasing particle-index order. */
int advance_p(vpb_particle_t *p0, int np, const float q_m, vpb_particle_mover_t *pm, int max_nm,
              vpb_accumulator_t *a0, const vpb_interpolator_t *f0, const vpb_grid_t *g);

/* spa.h:43-47 (move_p.c:20-136): finish one mover; 0 = done, 1 = still in use. */
int move_p(vpb_particle_t *p0, vpb_particle_mover_t *m, vpb_accumulator_t *a0, const vpb_grid_t *g);

/* spa.h:23-25 (sort_p.c:16-102): counting sort by voxel; fills sp->partition. */
void sort_p(vpb_species_t *sp, const vpb_grid_t *g);

/* spa.h:72-91 (center_p.cxx:155, uncenter_p.cxx:155) */
void center_p(vpb_particle_t *p0, int np, const float q_m, const vpb_interpolator_t *f0, const vpb_grid_t *g);
void uncenter_p(vpb_particle_t *p0, int np, const float q_m, const vpb_interpolator_t *f0, const vpb_grid_t *g);

/* spa.h:99-104 (energy_p.cxx:124-157) */
double energy_p(const vpb_particle_t *p0, int np, float q_m, const vpb_interpolator_t *f0, const vpb_grid_t *g);

/* spa.h:108-112 (rho_p.c:23-79) */
void accumulate_rho_p(vpb_field_t *f, const vpb_particle_t *p0, int np, const vpb_grid_t *g);

/* spa.h:29-32 (boundary_p.c:9-71) */
void accumulate_rhob(vpb_field_t *f0, const vpb_particle_t *p, const vpb_grid_t *g);

/* spa.h:34-41 (boundary_p.c:77-505).  rng is the reference's mt_rng_t*, only
 * forwarded to custom boundary handlers (none are implemented on the device). */
void boundary_p(vpb_species_t *sp_list, vpb_field_t *f0, vpb_accumulator_t *a0, const vpb_grid_t *g, void *rng);

/* src/sf_interface/sf_interface.h:83-163 */
vpb_interpolator_t *new_interpolator(vpb_grid_t *g);
void delete_interpolator(vpb_interpolator_t *fi);
vpb_accumulator_t *new_accumulators(vpb_grid_t *g);
void delete_accumulators(vpb_accumulator_t *a);
void load_interpolator(vpb_interpolator_t *fi, const vpb_field_t *f, const vpb_grid_t *g);
void clear_accumulators(vpb_accumulator_t *a, const vpb_grid_t *g);
void reduce_accumulators(vpb_accumulator_t *a, const vpb_grid_t *g);
void unload_accumulator(vpb_field_t *f, const vpb_accumulator_t *a, const vpb_grid_t *g);

/* src/sf_interface/sf_interface.h:83-91,140-163 and src/species_advance/standard/spa.h:114-123 (hydro_p.c:24-161,
 * sf_interface/hydro.c:30-184): the 14 hydro moments of a species on the mesh nodes (diagnostic dumps). */
vpb_hydro_t *new_hydro(vpb_grid_t *g);
void delete_hydro(vpb_hydro_t *h);
void clear_hydro(vpb_hydro_t *h, const vpb_grid_t *g);
void accumulate_hydro_p(vpb_hydro_t *h0, const vpb_particle_t *p0, int n, float q_m, const vpb_interpolator_t *f0,
                        const vpb_grid_t *g);
void synchronize_hydro(vpb_hydro_t *h, const vpb_grid_t *g);
void local_adjust_hydro(vpb_hydro_t *h, const vpb_grid_t *g);

/* src/field_advance/field_advance.h:318-345: the vtables decks name through the
 * standard_field_advance / vacuum_field_advance macros. */
extern vpb_field_advance_methods_t _standard_field_advance[1];
extern vpb_field_advance_methods_t _vacuum_field_advance[1];
extern vpb_field_advance_methods_t _standard_v4_field_advance[1];
extern vpb_field_advance_methods_t _vacuum_v4_field_advance[1];

/* src/field_advance/field_advance.h:304-316 (field_advance.c:3-28) */
vpb_field_advance_t *new_field_advance(vpb_grid_t *g, vpb_material_t *m_list, vpb_field_advance_methods_t *fam);
void delete_field_advance(vpb_field_advance_t *fa);

/* src/util/util_base.h:261-280 (util.c:46-91): the reference's single allocation
 * choke-point.  Substituting it puts every large array in CUDA managed memory,
 * so kernels run on deck-visible pointers in place. */
void util_malloc_aligned(const char *err_fmt /* two %lu: bytes, alignment */, void *mem_ref, size_t n, size_t a);
void util_free_aligned(void *mem_ref);

#endif /* VPB_NO_REFERENCE_NAMES */

/* ------------------------------------------------------------------------- */
/* (B) Device-resident layer                                                   */
/* ------------------------------------------------------------------------- */

typedef struct vpb_domain vpb_domain_t;

/* Bind this process to a CUDA device (one process per GPU) and create the
 * library stream.  Idempotent.  Returns 0, or exits loudly if no device. */
int vpb_init(int device_ordinal);
void vpb_shutdown(void);
int vpb_device_sm_count(void);
int vpb_l2_fetch_granularity(void);         /* cudaLimitMaxL2FetchGranularity in effect (tuning l2.fetch_granularity) */

void *vpb_dev_alloc(size_t bytes);           /* cudaMalloc, zero-filled */
void vpb_dev_free(void *d);
void *vpb_malloc_managed(size_t bytes);      /* cudaMallocManaged, zero-filled, preferred location = device */
void *vpb_host_alloc_pinned(size_t bytes);
void vpb_host_free_pinned(void *h);
void vpb_h2d(void *d, const void *h, size_t bytes);   /* async on the library stream */
void vpb_d2h(void *h, const void *d, size_t bytes);
void vpb_d2d(void *dst, const void *src, size_t bytes);
void vpb_memset(void *d, int byte, size_t bytes);
void vpb_sync(void);
void *vpb_stream(void);                      /* cudaStream_t of the library stream */

/* Stream-ordered stopwatch on the library stream (CUDA events). */
void vpb_timer_start(int slot);
void vpb_timer_stop(int slot);
float vpb_timer_ms(int slot);                /* synchronises on the stop event */

/* Per-kernel-class timing with CUDA events on the library stream (off by default).
 * Classes: 0 advance_p, 1 sort_p, 2 advance_b, 3 advance_e, 4 load_interpolator,
 * 5 unload_accumulator, 6 other (compute_curl_b), 7 boundary_p (migration), 8 halo (synchronize_jf, ghost
 * exchanges and local boundary conditions of advance_e), 9 divergence cleaning / shared-face synchronisation. */
void vpb_prof_enable(int on);
void vpb_prof_collect(int cls, double *total_ms, int *count, int reset);
int vpb_prof_list(int cls, float *out_ms, int max);   /* per-launch durations in launch order */
/* Host wall-clock accounting of the layer-A entry points (also switched on by VPB_TRACE=1 in the environment, which
 * prints the table at exit to stderr or to the file VPB_TRACE_FILE names): calls and seconds per entry point. */
void vpb_trace_enable(int on);
void vpb_trace_reset(void);
void vpb_trace_report(void);
double vpb_trace_get(const char *label, long *calls);

/* Count of kernel launches made by this library since the last reset. */
long vpb_launch_count(int reset);

/* Multi-GPU: one rank per GPU, NCCL over NVLink.  Rank 0 creates the 128-byte id,
 * the launcher distributes it (torch.distributed broadcast, MPI_Bcast, a file),
 * every rank calls vpb_comm_init.  Replaces mp_init/new_mp (util/mp/mp.h:14-40). */
void vpb_comm_unique_id(void *out128);
void vpb_comm_init(int rank, int nproc, const void *uid128);
/* The same bootstrap with no host code at all, for host programs that ARE the reference (link-time substitution):
 * rank, size and the id exchange go through the reference's own message layer, found in the process by name
 * (mp_rank_cxx, mp_nproc_cxx, mp_allgather_i_cxx; util/mp/mp.hxx:36-143 -- link the executable with -rdynamic).
 * `mp` is grid_t::mp.  Collective.  Called by the reference-named entry points on first use; returns the world
 * size, 0 if there is no such message layer or only one rank.  Ranks that each have a GPU of their own talk over
 * NCCL; when two ranks of the job share a GPU (which NCCL refuses) the exchanges are staged through the host
 * program's message layer instead (mp_begin_send_cxx ... mp_end_recv_cxx, mp_allsum_d_cxx).  Tuning
 * comm.transport: 0 decide as described, 1 NCCL, 2 the host program's layer. */
int vpb_comm_autoboot(void *mp);
void vpb_comm_finalize(void);
int vpb_comm_rank(void);
int vpb_comm_nproc(void);
void vpb_comm_allsum_d(double *d_buf, int n);   /* mp_allsum_d on device doubles, in place */

/* Bookkeeping the reference keeps implicitly behind its mp handle / structors. */
void vpb_set_world(int nproc);                                   /* ranks in the job, for bc[] "is a rank" tests */
void vpb_register_material_coefficients(const vpb_material_coefficient_t *m, int n_mat);
void vpb_grid_changed(const vpb_grid_t *g);                      /* drop the cached device mirror of g */
vpb_domain_t *vpb_domain_of_grid(const vpb_grid_t *g);           /* cached mirror used by layer (A) */
void vpb_staging_release(void);                                  /* free layer (A)'s cached device buffers */
void vpb_staging_bytes(size_t *h2d, size_t *d2h);                /* bytes staged since the last call */
vpb_field_advance_methods_t *vpb_field_advance_table(int which); /* 0 std, 1 vacuum, 2 std_v4, 3 vacuum_v4 */

/* Tuning knobs (kernel variants measured in profiles/): name -> int. */
void vpb_set_tuning(const char *name, int value);
int vpb_get_tuning(const char *name);

/* A domain is the immutable device mirror of a grid_t (scalars, bc[27], the
 * neighbor table compressed to int32 local ids / particle-bc codes). */
vpb_domain_t *vpb_domain_create(const vpb_grid_t *g, int rank, int nproc);
void vpb_domain_destroy(vpb_domain_t *dom);
long vpb_domain_nvoxel(const vpb_domain_t *dom);

/* advance_p on device arrays.  d_nm (device int, may be NULL) receives the number
 * of movers written to d_pm; the call does not synchronise.  Movers are emitted
 * in increasing particle index (boundary_p.c:168-176 relies on that). */
void vpb_advance_p(vpb_domain_t *dom, vpb_particle_t *d_p, int np, float q_m, vpb_particle_mover_t *d_pm,
                   int max_nm, vpb_accumulator_t *d_a, const vpb_interpolator_t *d_f, int *d_nm);
/* Same, with a traversal hint: d_partition (int[nvoxel+1], may be NULL) is partition[] as the last
 * sort of this array left it.  It only decides the ORDER in which chunks of the array are visited
 * (x-rows, y-blocked, z inner: keeps a voxel's neighbours in L2 once particles have drifted off
 * their sorted voxel); results do not depend on it. */
void vpb_advance_p_ordered(vpb_domain_t *dom, vpb_particle_t *d_p, int np, float q_m, vpb_particle_mover_t *d_pm,
                           int max_nm, vpb_accumulator_t *d_a, const vpb_interpolator_t *d_f, int *d_nm,
                           const int *d_partition);
int vpb_advance_p_ignored(void);   /* movers the last vpb_advance_p dropped because pm[] was full (synchronises) */
void vpb_center_p(vpb_domain_t *dom, vpb_particle_t *d_p, int np, float q_m, const vpb_interpolator_t *d_f);
void vpb_uncenter_p(vpb_domain_t *dom, vpb_particle_t *d_p, int np, float q_m, const vpb_interpolator_t *d_f);
/* d_en: device double[1], receives sum(q*w/(sqrt(1+w)+1)) before the c^2/q_m scale (energy_p.cxx:46) */
void vpb_energy_p(vpb_domain_t *dom, const vpb_particle_t *d_p, int np, float q_m, const vpb_interpolator_t *d_f, double *d_en);
void vpb_accumulate_rho_p(vpb_domain_t *dom, vpb_field_t *d_f, const vpb_particle_t *d_p, int np);

/* Synthetic load for benchmarks / large property tests: ppc particles in every interior
 * voxel (born voxel-sorted), uniform in the cell, Maxwellian momenta of width vth. */
void vpb_load_thermal(vpb_domain_t *dom, vpb_particle_t *d_p, int ppc, float vth, float q,
                      unsigned long long seed, long tag0);
void vpb_copy_positions(vpb_particle_t *d_dst, const vpb_particle_t *d_src, long np);

/* The deck's particle load, batched on the device from the reference's own random-number stream (SURVEY.md 8f-4;
 * csrc/vpb_mt.cu).  Replaces, for a load loop of the usual shape, the serial host loop
 *     seed_rand(s); for(...) inject_particle( sp, uniform_rand(..), .., maxwellian_rand(..), .., q, tag, 0, 0 );
 * (src/vpic/vpic.hxx:491-505, src/vpic/misc.cxx:16-105, src/util/mtrand/mtrand.c) and leaves the same particles. */
typedef struct vpb_mt vpb_mt_t;
/* new_mt_rng(seed) / seed_mt_rng (mtrand.c:39-62) */
vpb_mt_t *vpb_mt_create(unsigned int seed);
void vpb_mt_destroy(vpb_mt_t *rng);
/* the generator state in the byte format of get_mt_rng_state / set_mt_rng_state (mtrand.c:64-124, 4*(624+1) = 2500
 * bytes): a host program hands its generator over with set and takes the stream back where the load left it with get */
void vpb_mt_set_state(vpb_mt_t *rng, const void *state2500);
void vpb_mt_get_state(vpb_mt_t *rng, void *state2500);
/* n words of mt_urand to a device array */
void vpb_mt_words(vpb_mt_t *rng, unsigned int *d_out, long n);
/* n records of the token string prog ('U' = mt_drand, mtrand.c:240; 'N' = mt_drandn, mtrand.c:395-438; at most 32 tokens)
 * in stream order to the device array d_out[n*strlen(prog)]: the doubles the n*strlen(prog) host calls would return,
 * and the generator left where they would leave it. */
void vpb_mt_draw(vpb_mt_t *rng, const char *prog, long n, double *d_out);
/* the ziggurat layer table the device uses, x[257], y[257], *r (host computation: make_zig.c's construction) */
void vpb_mt_ziggurat_table(double *x, double *y, double *r);
/* n calls of inject_particle(sp, x, y, z, ux, uy, uz, q, tag, 0, 0) in order, arguments from row k of the deviate table
 * d_table[n*stride]: x = lo[0]*(1-t[col[0]]) + hi[0]*t[col[0]] (uniform_rand), y, z alike from col[1], col[2];
 * ux = dev[0]*t[col[3]] (maxwellian_rand), uy, uz from col[4], col[5].  Particles the reference would not inject on this
 * rank (outside the local box, or on a far wall shared with a neighbour; misc.cxx:37-39) are skipped, the others appended
 * at d_p[np...] in order (d_p in the domain's particle layout).  Returns the new np; more than max_np is an error.
 * With lo = 0, hi = 1 and dev = 1 the mapping is exact (0*(1-t) + 1*t = t, 1*t = t): the table then holds x, y, z, ux, uy, uz
 * themselves and the call is the batched inject_particle for momenta a deck computed its own way (drifting or
 * relativistic loads such as decks/trecon-part/turbulence.cxx:535-541). */
int vpb_inject_from_draws(vpb_domain_t *dom, vpb_particle_t *d_p, int np, int max_np, const double *d_table, int stride, long n,
                          const int col[6], const double lo[3], const double hi[3], const double dev[3], double q, long tag, long tag_step);
/* The load loop of a thermal deck (BASELINE configs[0]/[3] recipe): n iterations of one position from three
 * uniform_rand(lo, hi) and two co-located particles (charges q_a, q_b) with three maxwellian_rand(vth) each, the deviates
 * written as arguments of inject_particle -- evaluated right to left by g++ (args_right_to_left = 1: the first deviate
 * is uz) or left to right (0).  np[2]: particle counts of the two arrays, in and out.  Iteration k tags its particles
 * tag0 + k*tag_step (oracle/decks/thermal_c1.cxx passes 0, thermal_small.cxx the loop counter). */
long vpb_load_pairs_mt(vpb_domain_t *dom, vpb_mt_t *rng, long n, const double lo[3], const double hi[3], double vth_a, double vth_b, double q_a,
                       double q_b, vpb_particle_t *d_a, int max_a, vpb_particle_t *d_b, int max_b, int np[2], int args_right_to_left,
                       long tag0, long tag_step);
/* Device field layout.  A domain starts out with the reference's 80-byte AoS field_t (what every layer-A entry
 * point uses).  A caller that keeps the field array resident on the device can switch the domain to the PLANAR
 * layout: five planes (e|div_e, cb|div_b, tca|rhob, jf|rhof, material ids) of one 16-byte quad per voxel, so a
 * stencil kernel only moves the quads it uses (advance_b: 48 B per cell instead of the 112 B ncu measures on the
 * AoS array).  All vpb_* field functions follow the domain's current layout; vpb_field_bytes() is the size to
 * allocate, vpb_field_convert() copies between the two layouts (out of place). */
void vpb_domain_set_field_layout(vpb_domain_t *dom, int planar);
int vpb_domain_field_layout(const vpb_domain_t *dom);
size_t vpb_field_bytes(const vpb_domain_t *dom);
void vpb_field_convert(vpb_domain_t *dom, vpb_field_t *d_dst, const vpb_field_t *d_src, int to_planar);

/* Device interpolator layout.  Default: the reference's 80-byte interpolator_t.  wide=1: the same 72 useful
 * bytes in 96-byte records (three aligned 32-byte sectors), which advance_p gathers with two 256-bit loads and
 * one 64-bit load instead of five loads that request the record's sectors five times.  load_interpolator,
 * advance_p, center_p, uncenter_p and energy_p follow the domain's setting. */
void vpb_domain_set_interpolator_layout(vpb_domain_t *dom, int wide);
size_t vpb_interpolator_bytes(const vpb_domain_t *dom);

/* Device particle layout.  Default (plane = 0): the reference's 48-byte particle_t records, what every layer-A entry
 * point uses.  plane > 0 (a multiple of 64): every species array of the domain is nine COMPONENT PLANES in one
 * allocation of plane*48 bytes -- eight planes of `plane` 4-byte words (dx, dy, dz, i, ux, uy, uz, q) followed by one
 * plane of `plane` 16-byte {tag, tag2} pairs -- still passed around as a vpb_particle_t* with capacity max_np <= plane.
 * advance_p then runs the two-particles-per-lane kernel (vpb_advance_p_pair.cu): one 64-bit word per component and
 * lane, packed f32x2 arithmetic, and only the six words that change are written back (56 B of particle traffic per
 * advance instead of the 96 B the 48-byte records impose).  All vpb_* particle functions follow the domain's setting;
 * vpb_particle_convert() copies between the two layouts (out of place). */
void vpb_domain_set_particle_layout(vpb_domain_t *dom, long plane);
long vpb_domain_particle_layout(const vpb_domain_t *dom);
void vpb_particle_convert(vpb_domain_t *dom, vpb_particle_t *d_dst, const vpb_particle_t *d_src, long np, int to_planes);
void vpb_copy_positions_dom(vpb_domain_t *dom, vpb_particle_t *d_dst, const vpb_particle_t *d_src, long np);

/* Synthetic field state: an x-propagating vacuum plane wave (ey, cbz) with `mode` wavelengths across
 * the local nx cells; everything else zero. */
void vpb_load_plane_wave(vpb_domain_t *dom, vpb_field_t *d_f, int mode, float amp);

/* One species' device arrays as boundary_p needs them (the device-side part of species_t). */
typedef struct vpb_species_state {
  vpb_particle_t *p;
  vpb_particle_mover_t *pm;
  int np, max_np;
  int nm, max_nm;     /* nm movers in pm[], ascending particle index (as advance_p leaves them) */
  int id, _pad;       /* species_t.id, carried in particle_injector_t.sp_id */
} vpb_species_state_t;

/* One round of boundary_p (boundary_p.c:77-505) for n_sp<=7 species: absorb / migrate the movers,
 * back-fill, exchange with the face neighbours over NCCL, inject and finish the arrivals.  np and nm
 * of every species are updated; the call synchronises.  advance.cxx:94-96 calls it 3 times a step. */
void vpb_boundary_p(vpb_domain_t *dom, vpb_species_state_t *sp, int n_sp, vpb_field_t *d_f, vpb_accumulator_t *d_a);
/* The same round for callers that number the rounds of a step (0, 1, 2: advance.cxx:94-96), identically on every rank.
 * After one exact round per number, which tells both sides of every face what passes through it, a round costs ONE
 * message per face (fixed capacity, a header record carrying the count) and ONE read-back of all counters:
 * the arrivals are appended, moved and turned into next round's movers with sizes the device reads from the headers.
 * A face that outgrows its capacity sends the rest in a second message and the capacity is raised.  Arrays cannot
 * grow on this path (no hook): arrivals beyond max_np / max_nm are an error.  round < 0 = vpb_boundary_p.
 * Tuning boundary.fused: 0 (default) always the exact protocol, 1 as described (measured on 2 and 4 GPUs: same step time,
 * DESIGN.md section 6). */
void vpb_boundary_p_round(vpb_domain_t *dom, vpb_species_state_t *sp, int n_sp, vpb_field_t *d_f, vpb_accumulator_t *d_a, int round);
/* boundary_p.c:416-447: when the arrivals of a round do not fit, the reference grows the species' arrays by 31 % and
 * warns.  vpb_boundary_p leaves that to the owner of the arrays through this hook (NULL, the default: overflow is an
 * error).  The hook gets the index of the species in the list and the capacities needed; it must update p / max_np
 * (the first np particles preserved) and/or pm / max_nm in *st with device-accessible arrays and return non-zero, or
 * return 0 to refuse.  The reference-named boundary_p() installs one that does what the reference does, for arrays
 * that came from util_malloc_aligned. */
/* boundary_p.c:271-277: a mover that ends on a face bound to one of the deck's custom handlers (neighbor code -3-k) is given
 * to that HOST callback, then destroyed; injectors the callbacks make are injected last (boundary_p.c's cmlist).  The
 * reference-named boundary_p() runs the callbacks and brackets its vpb_boundary_p call with this: inj[0..n) = their
 * injectors (n >= 0 opens the bracket, n < 0 closes it).  Outside such a bracket a mover on a handler face is an error. */
void vpb_boundary_set_local_injectors(const vpb_particle_injector_t *inj, int n);
typedef int (*vpb_grow_hook_t)(void *user, int index, int need_np, int need_nm, vpb_species_state_t *st);
void vpb_boundary_set_grow_hook(vpb_grow_hook_t hook, void *user);

/* single mover / single particle, for the reference's host-side callers (inject_particle, handlers) */
void vpb_move_p_one(vpb_domain_t *dom, vpb_particle_t *d_p, vpb_particle_mover_t *d_pm, vpb_accumulator_t *d_a, int *d_result);
void vpb_accumulate_rhob_one(vpb_domain_t *dom, vpb_field_t *d_f, const vpb_particle_t *d_particle);

/* Stable counting sort by voxel: d_out receives the sorted particles, d_partition
 * (int[nvoxel+1]) the first particle of each voxel (sort_p.c:54-59,74). */
void vpb_sort_p(vpb_domain_t *dom, const vpb_particle_t *d_in, vpb_particle_t *d_out, int np, int *d_partition);
/* Same sort for a component-plane array (Device particle layout), in place: d_p holds the sorted planes on return,
 * d_tmp (same capacity) is scratch.  The permutation is applied to whole 48-byte records staged in d_tmp, which is
 * several times faster than moving the 4-byte plane words through it one by one. */
void vpb_sort_p_planes(vpb_domain_t *dom, vpb_particle_t *d_p, vpb_particle_t *d_tmp, int np, int *d_partition);
/* Look-ahead variant for device-resident runs: particles are grouped by the voxel they WILL be in `lookahead` steps
 * from now at their present velocity (clamped to the local interior); the order inside a group is arbitrary.  The order of a
 * particle array has no physical meaning; this one halves the average distance between a particle and the voxel
 * its neighbours in the array share over a sort interval (lookahead ~ 0.6 sort_interval), which is what advance_p's
 * gathers and REDs pay for.  d_partition describes the groups, not the current voxels.  lookahead = 0 is
 * vpb_sort_p_planes. */
void vpb_sort_p_planes_ahead(vpb_domain_t *dom, vpb_particle_t *d_p, vpb_particle_t *d_tmp, int np, int *d_partition, int lookahead);
/* The device-resident driver's sort (csrc/vpb_sort_group.cu): d_out receives the particles of the component-plane array
 * d_in GROUPED by the voxel they occupy (lookahead == 0) or reach `lookahead` steps from now; the groups follow a
 * brick-Morton order of the voxels, key(x,y,z) = fx[x] + fy[y] + fz[z] (vpb_sort_group_order: tables of n+2 entries
 * for coordinates 0..n+1, host code, returns the number of keys); d_partition (int[keys+1]) = first particle of every
 * group; the order inside a group is arbitrary.  Same particles, same physics as sort_p (sort_p.c:16-77), whose only
 * purpose is that particles of a voxel are neighbours in the array; the reference's ORDER and partition[] are what
 * sort_p / vpb_sort_p / vpb_sort_p_planes deliver. */
/* Deck-side particle diagnostics on the device (csrc/vpb_diag.cu).  vpb_energy_spectrum =
 * decks/trecon-part/energy.cxx:90-176: dist[k*nvoxel + voxel] counts the particles of a voxel in nex kinetic-energy bands
 * of width dke (the last band open-ended), normalised per cell, ghost cells copied from their interior neighbour exactly
 * as that loop does; edist[nbin] is the global spectrum over log-spaced bins between eminp and emaxp.  Either output may
 * be NULL.  vpb_tracer_records = dump_tracers of tracer.cxx:125-160: thirteen floats per particle (q, global x y z,
 * ux uy uz, ex ey ez, cbx cby cbz); field_of_first != 0 reproduces the macro's `field[p->i]` (the voxel of the FIRST
 * particle for every record), 0 uses each particle's own voxel.  Device pointers; the vpb_deck_* forms take the deck's
 * own (host or managed) arrays and return with the results on the host. */
void vpb_energy_spectrum(vpb_domain_t *dom, const vpb_particle_t *d_p, int np, double dke, int nex, float *d_dist, double eminp,
                         double emaxp, int nbin, float *d_edist);
void vpb_tracer_records(vpb_domain_t *dom, const vpb_particle_t *d_p, int np, const vpb_field_t *d_f, float x0, float y0, float z0,
                        int field_of_first, float *d_out);
void vpb_deck_energy_spectrum(const vpb_particle_t *p0, int np, double dke, int nex, float *dist, double eminp, double emaxp, int nbin,
                              float *edist, const vpb_grid_t *g);
void vpb_deck_tracer_records(const vpb_particle_t *p0, int np, const vpb_field_t *f, float *out13, int field_of_first, const vpb_grid_t *g);
void vpb_sort_p_planes_grouped(vpb_domain_t *dom, const vpb_particle_t *d_in, vpb_particle_t *d_out, int np, int *d_partition, int lookahead);
long vpb_sort_group_order(int nx, int ny, int nz, int *fx, int *fy, int *fz);
long vpb_sort_group_keys(vpb_domain_t *dom);

/* Hydro moments on device arrays (vpb_hydro_t[nvoxel], the reference layout; the particle array in the domain's
 * particle layout, the interpolator in the domain's interpolator layout). */
void vpb_clear_hydro(vpb_domain_t *dom, vpb_hydro_t *d_h);
void vpb_accumulate_hydro_p(vpb_domain_t *dom, vpb_hydro_t *d_h, const vpb_particle_t *d_p, int np, float q_m,
                            const vpb_interpolator_t *d_f);
void vpb_local_adjust_hydro(vpb_domain_t *dom, vpb_hydro_t *d_h);
void vpb_synchronize_hydro(vpb_domain_t *dom, vpb_hydro_t *d_h);

void vpb_load_interpolator(vpb_domain_t *dom, vpb_interpolator_t *d_fi, const vpb_field_t *d_f);
void vpb_clear_accumulators(vpb_domain_t *dom, vpb_accumulator_t *d_a);
void vpb_unload_accumulator(vpb_domain_t *dom, vpb_field_t *d_f, const vpb_accumulator_t *d_a);

/* Field solve on device arrays; `vacuum` selects vfa_advance_e (vacuum/vfa_advance_e.c:7-9). */
void vpb_advance_b(vpb_domain_t *dom, vpb_field_t *d_f, float frac);
void vpb_advance_e(vpb_domain_t *dom, vpb_field_t *d_f, const vpb_material_coefficient_t *d_m, int n_mat, int vacuum);
void vpb_clear_jf(vpb_domain_t *dom, vpb_field_t *d_f);
void vpb_clear_rhof(vpb_domain_t *dom, vpb_field_t *d_f);
void vpb_synchronize_jf(vpb_domain_t *dom, vpb_field_t *d_f);
void vpb_synchronize_rho(vpb_domain_t *dom, vpb_field_t *d_f);
/* d_en6: device double[6] = ex,ey,ez,cbx,cby,cbz energies of this rank (energy_f.c:93-179) */
void vpb_energy_f(vpb_domain_t *dom, const vpb_field_t *d_f, const vpb_material_coefficient_t *d_m, int n_mat, double *d_en6);
void vpb_compute_div_e_err(vpb_domain_t *dom, vpb_field_t *d_f, const vpb_material_coefficient_t *d_m, int n_mat);
void vpb_compute_rms_div_e_err(vpb_domain_t *dom, const vpb_field_t *d_f, double *d_out);
void vpb_clean_div_e(vpb_domain_t *dom, vpb_field_t *d_f, const vpb_material_coefficient_t *d_m, int n_mat);
void vpb_compute_div_b_err(vpb_domain_t *dom, vpb_field_t *d_f);
void vpb_compute_rms_div_b_err(vpb_domain_t *dom, const vpb_field_t *d_f, double *d_out);
void vpb_clean_div_b(vpb_domain_t *dom, vpb_field_t *d_f);
void vpb_compute_rhob(vpb_domain_t *dom, vpb_field_t *d_f, const vpb_material_coefficient_t *d_m, int n_mat);
void vpb_compute_curl_b(vpb_domain_t *dom, vpb_field_t *d_f, const vpb_material_coefficient_t *d_m, int n_mat);
void vpb_synchronize_tang_e_norm_b(vpb_domain_t *dom, vpb_field_t *d_f, double *d_err);

/* ------------------------------------------------------------------------- */
/* (C) Time-step driver (vpb_step.cu): the call order of vpic_simulation::advance()      */
/* (src/vpic/advance.cxx:13-244) in host C++ over layer (B), state resident in HBM.     */
/* ------------------------------------------------------------------------- */
typedef struct vpb_sim vpb_sim_t;

/* g is this rank's grid_t (partition_* of the reference, or old_vpic_b200/grid.py); the three layout flags select
 * the device layouts described above (1,1,1 = what bench.py measures).  n_mat vacuum-like materials unless
 * `vacuum` selects vfa_advance_e. */
vpb_sim_t *vpb_sim_create(const vpb_grid_t *g, int rank, int nproc, int n_mat, int vacuum, int field_planar, int wide_interpolator,
                          int particle_planes);
void vpb_sim_destroy(vpb_sim_t *s);
void vpb_sim_set_materials(vpb_sim_t *s, const vpb_material_coefficient_t *m, int n_mat);
/* species in definition order (= species_t.id); max_nm <= 0: the reference's 2*max_np/25 (vpic.hxx:416-420) */
int vpb_sim_define_species(vpb_sim_t *s, const char *name, float q_m, long max_np, long max_nm, int sort_interval);
void vpb_sim_load_thermal(vpb_sim_t *s, int species, int ppc, float vth, float q, unsigned long long seed, long tag0);
/* vpb_load_pairs_mt (the deck's load loop from the reference's random-number stream) into two species of the run */
long vpb_sim_load_pairs_mt(vpb_sim_t *s, vpb_mt_t *rng, int species_a, int species_b, long n, const double lo[3], const double hi[3],
                           double vth_a, double vth_b, double q_a, double q_b, int args_right_to_left, long tag0, long tag_step);
void vpb_sim_set_particles(vpb_sim_t *s, int species, const vpb_particle_t *host, long np);
long vpb_sim_get_particles(vpb_sim_t *s, int species, vpb_particle_t *host, long max);
void vpb_sim_set_fields(vpb_sim_t *s, const vpb_field_t *host);       /* field_t[nvoxel], reference layout */
void vpb_sim_get_fields(vpb_sim_t *s, vpb_field_t *host);
/* vpic_simulation::initialize() after the deck's user_initialization (initialize.cxx:27-95): shared-face synchronisation,
 * one div B clean, curl B, the bound charge density that makes the loaded plasma divergence-consistent (compute_rhob),
 * one div E clean if an error is left, load_interpolator, uncenter_p.  out3 (or NULL): synchronisation error, rms div B
 * error, rms div E error as the reference prints them. */
void vpb_sim_initialize(vpb_sim_t *s, double *out3);
void vpb_sim_set_intervals(vpb_sim_t *s, int clean_div_e_interval, int clean_div_b_interval, int num_comm_round);
/* sync_shared_interval of vpic_simulation (vpic.cxx:14; advance.cxx:199-208): synchronize_tang_e_norm_b every that many
 * steps, 0 = never.  vpb_sim_last_errors: the numbers advance.cxx:160,168,182,190,205 report, latest values --
 * out[0..1] rms div E error before the first / second cleaning pass, out[2..3] the same for div B, out[4] the domain
 * desynchronisation error.  The cleaning passes are skipped when their error is not > 0, as in the reference. */
void vpb_sim_set_sync_shared_interval(vpb_sim_t *s, int interval);
void vpb_sim_last_errors(const vpb_sim_t *s, double *out5);
void vpb_sim_set_sort_lookahead(vpb_sim_t *s, int steps);   /* < 0: 0.6 x sort_interval of each species; 0 (default): off */
void vpb_sim_advance(vpb_sim_t *s, int nsteps);                        /* enqueues nsteps time steps */
/* The deck's five hooks, called where vpic_simulation::advance() calls user_particle_collisions (advance.cxx:67),
 * user_particle_injection (:85), user_current_injection (:123), user_field_injection (:141) and user_diagnostics
 * (:233, after the step counter has advanced).  NULL members are skipped.  Device work is only enqueued when a hook
 * runs: a hook that reads device arrays calls vpb_sync() first. */
typedef struct vpb_sim_callbacks {
  void (*particle_collisions)(void *user, vpb_sim_t *s);
  void (*particle_injection)(void *user, vpb_sim_t *s);
  void (*current_injection)(void *user, vpb_sim_t *s);
  void (*field_injection)(void *user, vpb_sim_t *s);
  void (*diagnostics)(void *user, vpb_sim_t *s);
  void *user;
} vpb_sim_callbacks_t;
void vpb_sim_set_callbacks(vpb_sim_t *s, const vpb_sim_callbacks_t *cb);   /* NULL: none */
void vpb_sim_energies(vpb_sim_t *s, double *out6_plus_nspecies);       /* dump_energies (dump.cxx:37-78), over all ranks */
void vpb_sim_hydro(vpb_sim_t *s, int species, vpb_hydro_t *host);      /* clear + accumulate + synchronize, to the host */
long vpb_sim_step(const vpb_sim_t *s);
/* The field part of a step (advance.cxx:109-147, and :214 on steps without cleaning) is captured once into a CUDA graph
 * and replayed (tuning sim.graph, default 1: runs on one rank; 2: also runs split over ranks, whose segment holds NCCL
 * groups -- untested; off while per-kernel timing is on, with deck hooks inside the segment, or over the host-staged
 * transport).  Returns how many steps replayed it. */
long vpb_sim_graph_replays(const vpb_sim_t *s);
int vpb_sim_num_species(const vpb_sim_t *s);
long vpb_sim_np(vpb_sim_t *s, int species);
vpb_domain_t *vpb_sim_domain(vpb_sim_t *s);
vpb_field_t *vpb_sim_field_array(vpb_sim_t *s);                        /* device arrays, in the domain's layouts */
vpb_interpolator_t *vpb_sim_interpolator_array(vpb_sim_t *s);
vpb_accumulator_t *vpb_sim_accumulator_array(vpb_sim_t *s);
vpb_particle_t *vpb_sim_particle_array(vpb_sim_t *s, int species);

#ifdef __cplusplus
}
#endif
#endif /* VPIC_B200_H */
