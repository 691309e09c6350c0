// TEST INFRASTRUCTURE -- an input deck written for this repository (not part of the reference): the thermal
// e-/p+ plasma of BASELINE configs[0] scaled down to 16^3 cells x 8 ppc, periodic, split along x over however many ranks the job has
// (every rank draws the same random stream; inject_particle keeps the particles of its own slab).  It uses only the
// reference's deck API (src/vpic/vpic.hxx, deck_wrapper.cxx), so the SAME file builds
//   * against the reference alone            -> oracle/_ref/thermal_small.op        (golden energies), and
//   * against the reference's host objects with the hot path taken from libvpic_b200.so
//                                            -> oracle/_ref/hybrid/thermal_small.b200.op
// and tests/test_gpu_deck.py compares the two "energies" files.
#include <stdlib.h>

begin_globals {
  int dummy;
};

begin_initialization {
  const double L = 16;
  const int n = 16, ppc = 8;
  const double Ne = double(n) * n * n * ppc;

  num_step = 20;
  status_interval = 0;
  sync_shared_interval = 0;
  clean_div_e_interval = 10;
  clean_div_b_interval = 10;

  grid->dt = 0.95 * courant_length( L, L, L, n, n, n );
  grid->cvac = 1;
  grid->eps0 = 1;
  grid->damp = 0;
  define_periodic_grid( 0, 0, 0, L, L, L, n, n, n, int( nproc() ), 1, 1 );

  define_material( "vacuum", 1 );
  finalize_field_advance( standard_field_advance );

  // VPB_DECK_MAXNP (particles per species per rank) makes the arrays tight on purpose: boundary_p then has to grow
  // them when arrivals do not fit (boundary_p.c:416-447, "Resizing local ... particle storage")
  const char * tight = getenv( "VPB_DECK_MAXNP" );
  const double max_np = tight ? atof( tight ) : 1.5 * Ne;
  species_t * electron = define_species( "electron", -1, max_np, -1, 5, 1 );
  species_t * ion      = define_species( "ion",       1, max_np, -1, 5, 1 );

  seed_rand( 7 );
  const double q = L * L * L / Ne;     // plasma frequency 1
  for( int k = 0; k < int(Ne); k++ ) {
    const double x = uniform_rand( 0, L ), y = uniform_rand( 0, L ), z = uniform_rand( 0, L );
    inject_particle( electron, x, y, z, maxwellian_rand( 0.1 ), maxwellian_rand( 0.1 ), maxwellian_rand( 0.1 ), -q, k, 0, 0 );
    inject_particle( ion,      x, y, z, maxwellian_rand( 0.1 ), maxwellian_rand( 0.1 ), maxwellian_rand( 0.1 ),  q, k, 0, 0 );
  }
}

begin_diagnostics {
  dump_energies( "energies", step == 0 ? 0 : 1 );
}

begin_particle_injection { }
begin_current_injection { }
begin_field_injection { }
begin_particle_collisions { }
