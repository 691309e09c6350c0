// vpb_step.cu -- the time-step driver: the call order of vpic_simulation::advance()
// (src/vpic/advance.cxx:13-244) issued against the device-resident layer of this library, with every
// array of the run resident in HBM between steps.  This is host C++ behind a plain C ABI (vpb_sim_*):
// a deck-side caller creates a run from its grid_t, defines species, loads or uploads particles and
// calls vpb_sim_advance(); nothing in the loop touches particle or field data on the host.  Only when a
// face is shared with another rank or absorbs particles does a step read back one small count
// (boundary_p needs the mover counts on the host, as the reference does).
//
// old_vpic_b200/sim.py is the same driver in Python (kept for tests that poke at the pieces);
// tests/test_gpu_step.py runs both on the same input.
#include <math.h>
#include <string>
#include <vector>
#include "vpb_common.cuh"
#include "vpb_comm.cuh"

namespace {

struct Species {
  std::string name;
  float q_m = 0;
  int id = 0;
  int np = 0, max_np = 0, max_nm = 0, sort_interval = 0;
  vpb_particle_t *p = nullptr;
  vpb_particle_mover_t *pm = nullptr;
  int *nm = nullptr;            // device: movers left by advance_p
  int *partition = nullptr;     // device: int[nv+1] from the last sort
};

}  // namespace

struct vpb_sim {
  vpb_domain_t *dom = nullptr;
  const vpb_grid_t *g = nullptr;
  int rank = 0, nproc = 1;
  long nv = 0;
  bool vacuum = false, field_only = false, particle_planes = true;
  int n_mat = 1;
  vpb_field_t *f = nullptr;
  vpb_interpolator_t *fi = nullptr;
  vpb_accumulator_t *a = nullptr;
  vpb_material_coefficient_t *m = nullptr;
  vpb_hydro_t *hydro = nullptr;
  double *scalars = nullptr;     // device double[16]
  std::vector<Species> sp;
  vpb_particle_t *sort_tmp = nullptr;
  long sort_tmp_cap = 0;
  long step = 0;
  int clean_div_e_interval = 0, clean_div_b_interval = 0, num_comm_round = 3;   // vpic.cxx:17
  int sync_shared_interval = 0;                                                 // vpic.cxx:14
  // what the reference prints from advance.cxx:160,168,182,190,205 (last value of each, for callers and tests)
  double div_e_err[2] = {0, 0}, div_b_err[2] = {0, 0}, desync_err = 0;
  int sort_lookahead = 0;        // steps; < 0: 0.6 x the species' sort interval (measured optimum, profiles/README.md)
  int needs_boundary_p = -1;
  vpb_sim_callbacks_t cb = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // The field part of a step (clear_jf ... second advance_b [, load_interpolator]) as an instantiated CUDA graph:
  // [0] without, [1] with load_interpolator.  Captured the second time the segment runs (the first run makes the lazy
  // allocations); every pointer in it belongs to the run, so it is replayed unchanged until the materials change.
  cudaGraphExec_t field_graph[2] = {nullptr, nullptr};
  long field_graph_launches[2] = {0, 0};
  int field_segment_runs = 0;
  long graph_replays = 0;
};

using namespace vpb;

static size_t particle_bytes(const vpb_sim *s, long cap) {
  const long plane = vpb_domain_particle_layout(s->dom);
  return (size_t)(plane > 0 ? plane : cap) * sizeof(vpb_particle_t);
}

static bool needs_boundary_p(vpb_sim *s) {
  if (s->needs_boundary_p >= 0) return s->needs_boundary_p != 0;
  // advance_p can only leave movers where a face is shared with another rank or absorbs particles
  const vpb_grid_t *g = s->g;
  static const int fb[6] = {VPB_BOUNDARY(-1, 0, 0), VPB_BOUNDARY(1, 0, 0), VPB_BOUNDARY(0, -1, 0),
                            VPB_BOUNDARY(0, 1, 0),  VPB_BOUNDARY(0, 0, -1), VPB_BOUNDARY(0, 0, 1)};
  bool need = false;
  for (int k = 0; k < 6; k++) {
    const int bc = g->bc[fb[k]];
    if (bc >= 0 && bc < s->nproc && bc != s->rank) need = true;
  }
  if (!need && g->neighbor) {
    const size_t n = 6 * (size_t)s->nv;
    for (size_t k = 0; k < n && !need; k++)
      if (g->neighbor[k] < 0 && g->neighbor[k] != vpb_reflect_particles) need = true;
  }
  s->needs_boundary_p = need ? 1 : 0;
  return need;
}

static void sort_species(vpb_sim *s, Species &sp) {
  if (s->sort_tmp_cap < sp.max_np) {
    if (s->sort_tmp) vpb_dev_free(s->sort_tmp);
    s->sort_tmp = (vpb_particle_t *)vpb_dev_alloc(particle_bytes(s, sp.max_np));
    s->sort_tmp_cap = sp.max_np;
  }
  const bool planes = vpb_domain_particle_layout(s->dom) > 0;
  const bool grouped = planes && tuning("sort.grouped", 1) != 0;
  if (!sp.partition) {
    const long groups = grouped ? vpb_sort_group_keys(s->dom) : 0;
    sp.partition = (int *)vpb_dev_alloc((size_t)((groups > s->nv ? groups : s->nv) + 1) * sizeof(int));
  }
  if (planes) {
    const int ahead = s->sort_lookahead < 0 ? (3 * sp.sort_interval + 2) / 5 : s->sort_lookahead;
    if (grouped) {   // same particles grouped by voxel, groups in brick-Morton order (vpb_sort_group.cu); out of place
      vpb_sort_p_planes_grouped(s->dom, sp.p, s->sort_tmp, sp.np, sp.partition, ahead);
      std::swap(sp.p, s->sort_tmp);
    } else {
      vpb_sort_p_planes_ahead(s->dom, sp.p, s->sort_tmp, sp.np, sp.partition, ahead);   // sorted planes return to sp.p
    }
  } else {
    vpb_sort_p(s->dom, sp.p, s->sort_tmp, sp.np, sp.partition);
    std::swap(sp.p, s->sort_tmp);                                           // out of place: swap (sort_p.c:76-77)
  }
}

// boundary_p x num_comm_round (advance.cxx:94-103)
static void migrate(vpb_sim *s) {
  if (s->sp.empty() || !needs_boundary_p(s)) return;
  Context &c = ctx();
  const int n = (int)s->sp.size();
  std::vector<vpb_species_state_t> st(n);
  for (int k = 0; k < n; k++)
    VPB_CUDA(cudaMemcpyAsync(c.h_pinned_i + k, s->sp[k].nm, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  VPB_CUDA(cudaStreamSynchronize(c.stream));
  for (int k = 0; k < n; k++) {
    Species &sp = s->sp[k];
    memset(&st[k], 0, sizeof(st[k]));
    st[k].p = sp.p; st[k].pm = sp.pm; st[k].np = sp.np; st[k].max_np = sp.max_np;
    st[k].nm = c.h_pinned_i[k]; st[k].max_nm = sp.max_nm; st[k].id = sp.id;
  }
  for (int r = 0; r < s->num_comm_round; r++) vpb_boundary_p_round(s->dom, st.data(), n, s->f, s->a, r);
  for (int k = 0; k < n; k++) {
    s->sp[k].np = st[k].np;
    if (st[k].nm)   // advance.cxx:98-102
      VPB_WARNING("Ignoring %d unprocessed %s movers (increase num_comm_round)", st[k].nm, s->sp[k].name.c_str());
  }
}

// compute_rms_div_{e,b}_err (compute_rms_div_e_err.c:150-160): local sum x dV and the cell volume total, summed over
// ranks, eps0*sqrt(ratio).  The cleaning passes are conditional on this number (advance.cxx:164,169,186,191), so it is
// read back: one small synchronising copy per pass on cleaning steps only.
static double rms_err(vpb_sim *s, int which) {
  const vpb_grid_t *g = s->g;
  double loc[2];
  if (which == 0) vpb_compute_rms_div_e_err(s->dom, s->f, s->scalars); else vpb_compute_rms_div_b_err(s->dom, s->f, s->scalars);
  if (s->nproc > 1) {
    vpb_d2h(loc, s->scalars, sizeof(double));
    vpb_sync();
    loc[0] *= (double)g->dx * g->dy * g->dz;
    loc[1] = (double)(g->nx * g->ny * g->nz * g->dx * g->dy * g->dz);
    vpb_h2d(s->scalars, loc, 2 * sizeof(double));
    vpb_comm_allsum_d(s->scalars, 2);
    vpb_d2h(loc, s->scalars, 2 * sizeof(double));
    vpb_sync();
  } else {
    vpb_d2h(loc, s->scalars, sizeof(double));
    vpb_sync();
    loc[0] *= (double)g->dx * g->dy * g->dz;
    loc[1] = (double)(g->nx * g->ny * g->nz * g->dx * g->dy * g->dz);
  }
  return g->eps0 * sqrt(loc[0] / loc[1]);
}

static void clean_div_e(vpb_sim *s) {   // advance.cxx:149-173
  vpb_clear_rhof(s->dom, s->f);
  for (Species &sp : s->sp) vpb_accumulate_rho_p(s->dom, s->f, sp.p, sp.np);
  vpb_synchronize_rho(s->dom, s->f);
  vpb_compute_div_e_err(s->dom, s->f, s->m, s->n_mat);
  double err = s->div_e_err[0] = rms_err(s, 0);
  if (err > 0) {
    vpb_clean_div_e(s->dom, s->f, s->m, s->n_mat);
    vpb_compute_div_e_err(s->dom, s->f, s->m, s->n_mat);
    err = s->div_e_err[1] = rms_err(s, 0);
    if (err > 0) vpb_clean_div_e(s->dom, s->f, s->m, s->n_mat);
  }
}

static void clean_div_b(vpb_sim *s) {   // advance.cxx:177-195
  vpb_compute_div_b_err(s->dom, s->f);
  double err = s->div_b_err[0] = rms_err(s, 1);
  if (err > 0) {
    vpb_clean_div_b(s->dom, s->f);
    vpb_compute_div_b_err(s->dom, s->f);
    err = s->div_b_err[1] = rms_err(s, 1);
    if (err > 0) vpb_clean_div_b(s->dom, s->f);
  }
}

static void sync_shared(vpb_sim *s) {   // advance.cxx:199-208
  vpb_synchronize_tang_e_norm_b(s->dom, s->f, s->scalars);
  vpb_comm_allsum_d(s->scalars, 1);     // remote.c:412
  vpb_d2h(&s->desync_err, s->scalars, sizeof(double));
  vpb_sync();
}

// advance.cxx:109-147 (+ :214): everything between particle migration and the divergence cleaning
static void field_segment_launch(vpb_sim *s, bool with_load) {
  vpb_domain_t *dom = s->dom;
  const vpb_sim_callbacks_t &cb = s->cb;
  vpb_clear_jf(dom, s->f);                                                             // :109
  if (!s->sp.empty()) vpb_unload_accumulator(dom, s->f, s->a);                         // :110
  { ProfScope prof(8); vpb_synchronize_jf(dom, s->f); }                                // :112
  if (cb.current_injection) cb.current_injection(cb.user, s);                          // :123 user_current_injection
  vpb_advance_b(dom, s->f, 0.5f);                                                      // :129
  vpb_advance_e(dom, s->f, s->m, s->n_mat, s->vacuum ? 1 : 0);                         // :133
  if (cb.field_injection) cb.field_injection(cb.user, s);                              // :141 user_field_injection
  vpb_advance_b(dom, s->f, 0.5f);                                                      // :147
  if (with_load) vpb_load_interpolator(dom, s->fi, s->f);                              // :214
}

static void drop_field_graphs(vpb_sim *s) {
  for (int k = 0; k < 2; k++)
    if (s->field_graph[k]) { cudaGraphExecDestroy(s->field_graph[k]); s->field_graph[k] = nullptr; }
  s->field_segment_runs = 0;
}

// The segment is ~20 small launches (local boundary conditions, face packs and unpacks, the stencils) whose arguments
// never change: one graph launch instead.  Not while per-kernel timing is on (its events would sit inside the graph),
// not with deck hooks inside the segment, not over the host-staged transport (it synchronises).
static void field_segment(vpb_sim *s, bool with_load) {
  const vpb_sim_callbacks_t &cb = s->cb;
  // Runs split over ranks keep launching kernel by kernel unless sim.graph = 2: the segment then holds NCCL groups, and
  // the first 2-GPU attempt at capturing those did not come back (round 2, call F) -- not pursued, the segment is 1 %
  // of a step at the sizes ranks are split for.
  const int mode = tuning("sim.graph", 1);
  const bool can = mode != 0 && (s->nproc == 1 || mode == 2) && !prof_enabled() && !cb.current_injection &&
                   !cb.field_injection && comm_capturable();
  if (!can || s->field_segment_runs < 1) {
    field_segment_launch(s, with_load);
    if (can) s->field_segment_runs++;
    return;
  }
  Context &c = ctx();
  const int k = with_load ? 1 : 0;
  if (!s->field_graph[k]) {
    cudaGraph_t graph = nullptr;
    const long before = c.launches;
    VPB_CUDA(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
    field_segment_launch(s, with_load);
    VPB_CUDA(cudaStreamEndCapture(c.stream, &graph));
    s->field_graph_launches[k] = c.launches - before;
    c.launches = before;
    VPB_CUDA(cudaGraphInstantiate(&s->field_graph[k], graph, 0));
    VPB_CUDA(cudaGraphDestroy(graph));
  }
  VPB_CUDA(cudaGraphLaunch(s->field_graph[k], c.stream));
  c.launches += s->field_graph_launches[k];
  s->graph_replays++;
}

static void advance_one(vpb_sim *s) {
  vpb_domain_t *dom = s->dom;
  const bool particles = !s->sp.empty();
  const vpb_sim_callbacks_t &cb = s->cb;
  if (particles) vpb_clear_accumulators(dom, s->a);                                   // advance.cxx:38
  for (Species &sp : s->sp)                                                            // :43-51
    if (sp.sort_interval > 0 && s->step % sp.sort_interval == 0) sort_species(s, sp);
  if (cb.particle_collisions) cb.particle_collisions(cb.user, s);                      // :67 user_particle_collisions
  for (Species &sp : s->sp)                                                            // :70-73
    vpb_advance_p_ordered(dom, sp.p, sp.np, sp.q_m, sp.pm, sp.max_nm, s->a, s->fi, sp.nm, sp.partition);
  // reduce_accumulators (:74) is a no-op with one replica
  if (cb.particle_injection) cb.particle_injection(cb.user, s);                        // :85 user_particle_injection
  { ProfScope prof(7); migrate(s); }                                                   // :94-103
  const bool cleaning = (s->clean_div_e_interval > 0 && s->step % s->clean_div_e_interval == 0) ||
                        (s->clean_div_b_interval > 0 && s->step % s->clean_div_b_interval == 0) ||
                        (s->sync_shared_interval > 0 && s->step % s->sync_shared_interval == 0);
  field_segment(s, particles && !cleaning);                                            // :109-147 (+ :214 on plain steps)
  if (s->clean_div_e_interval > 0 && s->step % s->clean_div_e_interval == 0) { ProfScope prof(9); clean_div_e(s); }   // :151
  if (s->clean_div_b_interval > 0 && s->step % s->clean_div_b_interval == 0) { ProfScope prof(9); clean_div_b(s); }   // :177
  if (s->sync_shared_interval > 0 && s->step % s->sync_shared_interval == 0) { ProfScope prof(9); sync_shared(s); }   // :199
  if (particles && cleaning) vpb_load_interpolator(dom, s->fi, s->f);                  // :214
  s->step++;
  if (cb.diagnostics) cb.diagnostics(cb.user, s);                                      // :233 user_diagnostics, after step++
}

extern "C" {

// A run on the calling rank's share `g` of the box.  field_planar / wide_interpolator / particle_planes select the
// device layouts of include/vpic_b200.h (all three on is what bench.py measures).  n_mat vacuum-like materials are
// created unless `vacuum` (vfa_advance_e) is set; vpb_sim_set_materials replaces them.
vpb_sim_t *vpb_sim_create(const vpb_grid_t *g, int rank, int nproc, int n_mat, int vacuum, int field_planar, int wide_interpolator,
                          int particle_planes) {
  if (!g) VPB_ERROR("Bad grid");
  if (n_mat < 1) VPB_ERROR("Bad number of materials");
  vpb_init(-1);
  vpb_sim *s = new vpb_sim;
  s->g = g; s->rank = rank; s->nproc = nproc; s->n_mat = n_mat; s->vacuum = vacuum != 0;
  s->particle_planes = particle_planes != 0;
  s->dom = vpb_domain_create(g, rank, nproc);
  if (s->dom->n_handler_faces)
    VPB_ERROR("grid has %zu cell faces bound to custom particle-boundary handlers (grid->nb = %d): those are host callbacks "
              "(boundary_p.c:271-277) which only the reference-named boundary_p() can run; the device-resident driver has none -- "
              "use reflect_particles / absorb_particles faces", s->dom->n_handler_faces, g->nb);
  s->nv = vpb_domain_nvoxel(s->dom);
  s->field_only = g->neighbor == nullptr;
  vpb_domain_set_field_layout(s->dom, field_planar ? 1 : 0);
  vpb_domain_set_interpolator_layout(s->dom, wide_interpolator ? 1 : 0);
  s->f = (vpb_field_t *)vpb_dev_alloc(vpb_field_bytes(s->dom));
  if (!s->field_only) {
    s->fi = (vpb_interpolator_t *)vpb_dev_alloc(vpb_interpolator_bytes(s->dom));
    s->a = (vpb_accumulator_t *)vpb_dev_alloc((size_t)(s->nv + 1) * sizeof(vpb_accumulator_t));
  }
  s->scalars = (double *)vpb_dev_alloc(16 * sizeof(double));
  if (!s->vacuum) {
    std::vector<vpb_material_coefficient_t> m(n_mat);
    memset(m.data(), 0, m.size() * sizeof(m[0]));
    for (auto &x : m) {
      x.decayx = x.drivex = x.decayy = x.drivey = x.decayz = x.drivez = 1.f;
      x.rmux = x.rmuy = x.rmuz = 1.f;
      x.nonconductive = 1.f;
      x.epsx = x.epsy = x.epsz = 1.f;
    }
    s->m = (vpb_material_coefficient_t *)vpb_dev_alloc(m.size() * sizeof(m[0]));
    vpb_h2d(s->m, m.data(), m.size() * sizeof(m[0]));
    vpb_sync();
  }
  return s;
}

void vpb_sim_set_materials(vpb_sim_t *s, const vpb_material_coefficient_t *m, int n_mat) {
  if (!s || !m || n_mat < 1) VPB_ERROR("Bad args");
  drop_field_graphs(s);
  if (s->m) vpb_dev_free(s->m);
  s->m = (vpb_material_coefficient_t *)vpb_dev_alloc((size_t)n_mat * sizeof(*m));
  vpb_h2d(s->m, m, (size_t)n_mat * sizeof(*m));
  vpb_sync();
  s->n_mat = n_mat;
  s->vacuum = false;
}

void vpb_sim_destroy(vpb_sim_t *s) {
  if (!s) return;
  vpb_sync();
  drop_field_graphs(s);
  for (Species &sp : s->sp) {
    vpb_dev_free(sp.p); vpb_dev_free(sp.pm); vpb_dev_free(sp.nm);
    if (sp.partition) vpb_dev_free(sp.partition);
  }
  if (s->sort_tmp) vpb_dev_free(s->sort_tmp);
  if (s->f) vpb_dev_free(s->f);
  if (s->fi) vpb_dev_free(s->fi);
  if (s->a) vpb_dev_free(s->a);
  if (s->m) vpb_dev_free(s->m);
  if (s->hydro) vpb_dev_free(s->hydro);
  vpb_dev_free(s->scalars);
  vpb_domain_destroy(s->dom);
  delete s;
}

// species_t of the reference minus the host arrays; max_nm <= 0 selects the reference's default 2*max_np/25
// (vpic.hxx:416-420).  Returns the species id (list order = definition order, as boundary_p's send order needs).
int vpb_sim_define_species(vpb_sim_t *s, const char *name, float q_m, long max_np, long max_nm, int sort_interval) {
  if (!s) VPB_ERROR("Bad run");
  if (s->field_only) VPB_ERROR("a field-only grid (no neighbor table) cannot carry particles");
  if (max_np < 1) VPB_ERROR("Bad max_np");
  if (s->sp.size() >= 7) VPB_ERROR("at most 7 species per run");
  if (max_nm <= 0) max_nm = 2 * max_np / 25 > 16 ? 2 * max_np / 25 : 16;
  if (s->particle_planes && s->sp.empty()) vpb_domain_set_particle_layout(s->dom, (max_np + 63) / 64 * 64);
  const long plane = vpb_domain_particle_layout(s->dom);
  if (plane > 0 && max_np > plane) VPB_ERROR("species %s: max_np %ld exceeds the plane stride %ld fixed by the first species", name, max_np, plane);
  Species sp;
  sp.name = name ? name : "";
  sp.q_m = q_m;
  sp.id = (int)s->sp.size();
  sp.max_np = (int)max_np; sp.max_nm = (int)max_nm; sp.sort_interval = sort_interval;
  sp.p = (vpb_particle_t *)vpb_dev_alloc(particle_bytes(s, max_np));
  sp.pm = (vpb_particle_mover_t *)vpb_dev_alloc((size_t)max_nm * sizeof(vpb_particle_mover_t));
  sp.nm = (int *)vpb_dev_alloc(4 * sizeof(int));
  s->sp.push_back(sp);
  return sp.id;
}

static Species &species_of(vpb_sim_t *s, int id) {
  if (!s || id < 0 || id >= (int)s->sp.size()) VPB_ERROR("Bad species");
  return s->sp[id];
}

void vpb_sim_load_thermal(vpb_sim_t *s, int id, int ppc, float vth, float q, unsigned long long seed, long tag0) {
  Species &sp = species_of(s, id);
  const long np = (long)ppc * s->g->nx * s->g->ny * s->g->nz;
  if (np > sp.max_np) VPB_ERROR("species %s: %ld particles exceed max_np %d", sp.name.c_str(), np, sp.max_np);
  sp.np = (int)np;
  vpb_load_thermal(s->dom, sp.p, ppc, vth, q, seed, tag0);
}

// The thermal deck's load loop from the reference's random-number stream (vpb_mt.cu), appended to two species of the run
long vpb_sim_load_pairs_mt(vpb_sim_t *s, vpb_mt_t *rng, int id_a, int id_b, long n, const double lo[3], const double hi[3], double vth_a,
                           double vth_b, double q_a, double q_b, int args_right_to_left, long tag0, long tag_step) {
  Species &a = species_of(s, id_a), &b = species_of(s, id_b);
  int np[2] = {a.np, b.np};
  const long done = vpb_load_pairs_mt(s->dom, rng, n, lo, hi, vth_a, vth_b, q_a, q_b, a.p, a.max_np, b.p, b.max_np, np, args_right_to_left, tag0, tag_step);
  a.np = np[0];
  b.np = np[1];
  return done;
}

// host particle_t[np] -> the species' device array (converted to the domain's layout on the device)
void vpb_sim_set_particles(vpb_sim_t *s, int id, const vpb_particle_t *host, long np) {
  Species &sp = species_of(s, id);
  if (np < 0 || np > sp.max_np) VPB_ERROR("Bad number of particles");
  sp.np = (int)np;
  if (np == 0) return;
  if (vpb_domain_particle_layout(s->dom) > 0) {
    vpb_particle_t *tmp = (vpb_particle_t *)vpb_dev_alloc((size_t)np * sizeof(*host));
    vpb_h2d(tmp, host, (size_t)np * sizeof(*host));
    vpb_particle_convert(s->dom, sp.p, tmp, np, 1);
    vpb_sync();
    vpb_dev_free(tmp);
  } else {
    vpb_h2d(sp.p, host, (size_t)np * sizeof(*host));
    vpb_sync();
  }
}

long vpb_sim_get_particles(vpb_sim_t *s, int id, vpb_particle_t *host, long max) {
  Species &sp = species_of(s, id);
  const long np = sp.np < max ? sp.np : max;
  if (np <= 0) return 0;
  if (vpb_domain_particle_layout(s->dom) > 0) {
    vpb_particle_t *tmp = (vpb_particle_t *)vpb_dev_alloc((size_t)np * sizeof(*host));
    vpb_particle_convert(s->dom, tmp, sp.p, np, 0);
    vpb_d2h(host, tmp, (size_t)np * sizeof(*host));
    vpb_sync();
    vpb_dev_free(tmp);
  } else {
    vpb_d2h(host, sp.p, (size_t)np * sizeof(*host));
    vpb_sync();
  }
  return np;
}

// host field_t[nv] (the reference layout) <-> the device field array
void vpb_sim_set_fields(vpb_sim_t *s, const vpb_field_t *host) {
  if (!s || !host) VPB_ERROR("Bad args");
  const size_t bytes = (size_t)s->nv * sizeof(vpb_field_t);
  if (vpb_domain_field_layout(s->dom)) {
    vpb_field_t *tmp = (vpb_field_t *)vpb_dev_alloc(bytes);
    vpb_h2d(tmp, host, bytes);
    vpb_field_convert(s->dom, s->f, tmp, 1);
    vpb_sync();
    vpb_dev_free(tmp);
  } else {
    vpb_h2d(s->f, host, bytes);
    vpb_sync();
  }
  if (s->fi) vpb_load_interpolator(s->dom, s->fi, s->f);   // initialize.cxx:67
}

// What vpic_simulation::initialize() does between the deck's user_initialization and the first step
// (initialize.cxx:27-95): shared faces made consistent, div B cleaned once, the curl B the damping needs, the BOUND
// charge density chosen so that the loaded plasma is divergence-consistent (rhob = div(eps E) - rhof), one div E clean
// if that left an error, the interpolator, and the momenta moved back half a step (uncenter_p).  Call it after the
// fields and particles of the initial condition are in place.  out3 (may be null): the three numbers the reference
// prints there -- synchronisation error, rms div B error, rms div E error.
void vpb_sim_initialize(vpb_sim_t *s, double *out3) {
  if (!s) VPB_ERROR("Bad args");
  double sync0, div_b, div_e = 0;
  sync_shared(s);                                                                      // :32
  sync0 = s->desync_err;
  vpb_compute_div_b_err(s->dom, s->f);                                                 // :40
  div_b = rms_err(s, 1);                                                               // :41
  vpb_clean_div_b(s->dom, s->f);                                                       // :46 (unconditional there)
  vpb_compute_curl_b(s->dom, s->f, s->m, s->n_mat);                                    // :54
  if (!s->field_only) {
    vpb_clear_rhof(s->dom, s->f);                                                      // :59
    for (Species &sp : s->sp) vpb_accumulate_rho_p(s->dom, s->f, sp.p, sp.np);         // :60-61
    vpb_synchronize_rho(s->dom, s->f);                                                 // :62
    vpb_compute_rhob(s->dom, s->f, s->m, s->n_mat);                                    // :63
    vpb_compute_div_e_err(s->dom, s->f, s->m, s->n_mat);                               // :70
    div_e = rms_err(s, 0);                                                             // :71
    if (div_e > 0) vpb_clean_div_e(s->dom, s->f, s->m, s->n_mat);                      // :75
  }
  sync_shared(s);                                                                      // :80
  if (!s->sp.empty()) {
    vpb_load_interpolator(s->dom, s->fi, s->f);                                        // :90
    for (Species &sp : s->sp) vpb_uncenter_p(s->dom, sp.p, sp.np, sp.q_m, s->fi);      // :92-93
  }
  if (out3) { out3[0] = sync0; out3[1] = div_b; out3[2] = div_e; }
}

void vpb_sim_get_fields(vpb_sim_t *s, vpb_field_t *host) {
  if (!s || !host) VPB_ERROR("Bad args");
  const size_t bytes = (size_t)s->nv * sizeof(vpb_field_t);
  if (vpb_domain_field_layout(s->dom)) {
    vpb_field_t *tmp = (vpb_field_t *)vpb_dev_alloc(bytes);
    vpb_field_convert(s->dom, tmp, s->f, 0);
    vpb_d2h(host, tmp, bytes);
    vpb_sync();
    vpb_dev_free(tmp);
  } else {
    vpb_d2h(host, s->f, bytes);
    vpb_sync();
  }
}

void vpb_sim_set_intervals(vpb_sim_t *s, int clean_div_e_interval, int clean_div_b_interval, int num_comm_round) {
  if (!s) VPB_ERROR("Bad run");
  s->clean_div_e_interval = clean_div_e_interval;
  s->clean_div_b_interval = clean_div_b_interval;
  if (num_comm_round > 0) s->num_comm_round = num_comm_round;
}

// sync_shared_interval (vpic.cxx:14, advance.cxx:199-208): synchronize_tang_e_norm_b every that many steps; 0 = never
void vpb_sim_set_sync_shared_interval(vpb_sim_t *s, int interval) {
  if (!s) VPB_ERROR("Bad run");
  s->sync_shared_interval = interval;
}

// The numbers advance.cxx reports on cleaning / synchronising steps, latest values: out[0..1] rms div E error before the
// first and before the second cleaning pass, out[2..3] the same for div B, out[4] domain desynchronisation error.
void vpb_sim_last_errors(const vpb_sim_t *s, double *out) {
  if (!s || !out) VPB_ERROR("Bad args");
  out[0] = s->div_e_err[0]; out[1] = s->div_e_err[1]; out[2] = s->div_b_err[0]; out[3] = s->div_b_err[1]; out[4] = s->desync_err;
}

void vpb_sim_set_sort_lookahead(vpb_sim_t *s, int steps) {
  if (!s) VPB_ERROR("Bad run");
  s->sort_lookahead = steps;
}

// nsteps time steps; does not synchronise (beyond what migration needs)
void vpb_sim_advance(vpb_sim_t *s, int nsteps) {
  if (!s) VPB_ERROR("Bad run");
  for (int k = 0; k < nsteps; k++) advance_one(s);
}

// The deck's hooks (begin_particle_collisions, begin_particle_injection, begin_current_injection, begin_field_injection,
// begin_diagnostics of deck_wrapper.cxx) at the points advance.cxx:67,85,123,141,233 calls them.  Work is enqueued, not
// finished, when a hook runs: one that reads device arrays calls vpb_sync() first.
void vpb_sim_set_callbacks(vpb_sim_t *s, const vpb_sim_callbacks_t *cb) {
  if (!s) VPB_ERROR("Bad run");
  if (cb) s->cb = *cb;
  else s->cb = vpb_sim_callbacks_t{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  drop_field_graphs(s);
}

// how many steps replayed the captured field segment (0: sim.graph off, per-kernel timing on, hooks inside it, ...)
long vpb_sim_graph_replays(const vpb_sim_t *s) { return s ? s->graph_replays : 0; }

long vpb_sim_step(const vpb_sim_t *s) { return s ? s->step : -1; }
int vpb_sim_num_species(const vpb_sim_t *s) { return s ? (int)s->sp.size() : 0; }
long vpb_sim_np(vpb_sim_t *s, int id) { return species_of(s, id).np; }
vpb_domain_t *vpb_sim_domain(vpb_sim_t *s) { return s ? s->dom : nullptr; }
vpb_field_t *vpb_sim_field_array(vpb_sim_t *s) { return s ? s->f : nullptr; }
vpb_interpolator_t *vpb_sim_interpolator_array(vpb_sim_t *s) { return s ? s->fi : nullptr; }
vpb_accumulator_t *vpb_sim_accumulator_array(vpb_sim_t *s) { return s ? s->a : nullptr; }
vpb_particle_t *vpb_sim_particle_array(vpb_sim_t *s, int id) { return species_of(s, id).p; }

// dump_energies (src/vpic/dump.cxx:37-78): out[0..5] = field energies ex,ey,ez,cbx,cby,cbz, out[6+k] = kinetic
// energy of species k; summed over ranks.  Synchronises.
void vpb_sim_energies(vpb_sim_t *s, double *out) {
  if (!s || !out) VPB_ERROR("Bad args");
  const vpb_grid_t *g = s->g;
  vpb_energy_f(s->dom, s->f, s->m, s->n_mat, s->scalars);
  vpb_comm_allsum_d(s->scalars, 6);
  vpb_d2h(out, s->scalars, 6 * sizeof(double));
  vpb_sync();
  const double scale = 0.5 * g->eps0 * g->dx * g->dy * g->dz;
  for (int k = 0; k < 6; k++) out[k] *= scale;
  for (size_t k = 0; k < s->sp.size(); k++) {
    Species &sp = s->sp[k];
    vpb_energy_p(s->dom, sp.p, sp.np, sp.q_m, s->fi, s->scalars);
    vpb_comm_allsum_d(s->scalars, 1);
    vpb_d2h(out + 6 + k, s->scalars, sizeof(double));
    vpb_sync();
    out[6 + k] *= (double)g->cvac * (double)g->cvac / (double)sp.q_m;   // energy_p.cxx:156
  }
}

// hydro moments of one species as the dump path computes them (clear, accumulate, synchronize): host hydro_t[nv]
void vpb_sim_hydro(vpb_sim_t *s, int id, vpb_hydro_t *host) {
  Species &sp = species_of(s, id);
  if (!host) VPB_ERROR("Bad hydro");
  if (!s->hydro) s->hydro = (vpb_hydro_t *)vpb_dev_alloc((size_t)s->nv * sizeof(vpb_hydro_t));
  vpb_clear_hydro(s->dom, s->hydro);
  vpb_accumulate_hydro_p(s->dom, s->hydro, sp.p, sp.np, sp.q_m, s->fi);
  vpb_synchronize_hydro(s->dom, s->hydro);
  vpb_d2h(host, s->hydro, (size_t)s->nv * sizeof(vpb_hydro_t));
  vpb_sync();
}

}  // extern "C"
