// TEST/BENCH INFRASTRUCTURE -- an input deck written for this repository (not part of the reference): BASELINE.json
// configs[0], the thermal e-/p+ plasma at 64^3 cells x 32 particles per cell per species, on the
// reference's deck API.  Built like thermal_small.cxx by oracle/build_hybrid.sh, on the reference alone
// (oracle/_ref/thermal_c1.op, the shipped V4/SSE + pthreads hot path) and on the reference's host objects + libvpic_b200.so
// (oracle/_ref/hybrid/thermal_c1.b200.op); `bench.py --deck-e2e` times both: the same unmodified host program, the
// hot path swapped at link time.  Size and length can be changed from the environment: VPB_DECK_CELLS, VPB_DECK_PPC,
// VPB_DECK_STEPS.  Every step the diagnostics hook reads the step's result back on the host -- dump_energies (6 field
// energies + one kinetic energy per species, appended to `energies`) unless VPB_DECK_ENERGIES=0 -- and appends the wall
// clock to `steps.txt`, so that a caller can time steps W..W+K of the run.
#include <stdlib.h>
#include <stdio.h>
#include <time.h>

begin_globals {
  int dummy;
};

static int env_int( const char * name, int dflt ) {
  const char * e = getenv( name );
  return e ? atoi( e ) : dflt;
}

begin_initialization {
  const int n = env_int( "VPB_DECK_CELLS", 64 ), ppc = env_int( "VPB_DECK_PPC", 32 );
  const double L = n;
  const double Ne = double(n) * n * n * ppc;

  num_step = env_int( "VPB_DECK_STEPS", 40 );
  status_interval = 0;
  sync_shared_interval = 0;
  clean_div_e_interval = 0;
  clean_div_b_interval = 0;

  grid->dt = 0.95 * courant_length( L, L, L, n, n, n );
  grid->cvac = 1;
  grid->eps0 = 1;
  grid->damp = 0;
  define_periodic_grid( 0, 0, 0, L, L, L, n, n, n, int( nproc() ), 1, 1 );

  define_material( "vacuum", 1 );
  finalize_field_advance( standard_field_advance );

  species_t * electron = define_species( "electron", -1, 1.25 * Ne / nproc(), -1, 20, 1 );
  species_t * ion      = define_species( "ion",       1, 1.25 * Ne / nproc(), -1, 20, 1 );

  seed_rand( 7 );
  const double q = L * L * L / Ne;     // plasma frequency 1
  for( double k = 0; k < Ne; k++ ) {
    const double x = uniform_rand( 0, L ), y = uniform_rand( 0, L ), z = uniform_rand( 0, L );
    inject_particle( electron, x, y, z, maxwellian_rand( 0.1 ), maxwellian_rand( 0.1 ), maxwellian_rand( 0.1 ), -q, 0, 0, 0 );
    inject_particle( ion,      x, y, z, maxwellian_rand( 0.1 ), maxwellian_rand( 0.1 ), maxwellian_rand( 0.1 ),  q, 0, 0, 0 );
  }

  // VPB_DECK_DUMP_LOAD=<file>: the two particle arrays exactly as this loop left them (tests/test_oracle_mt.py pins the
  // oracle's restatement of the Mersenne-Twister stream and of inject_particle against them)
  const char * dump = getenv( "VPB_DECK_DUMP_LOAD" );
  if( dump ) {
    FILE * fp = fopen( dump, "wb" );
    if( !fp ) ERROR(( "cannot write %s", dump ));
    const int cnt[2] = { electron->np, ion->np };
    fwrite( cnt, sizeof(int), 2, fp );
    fwrite( electron->p, sizeof(particle_t), electron->np, fp );
    fwrite( ion->p,      sizeof(particle_t), ion->np,      fp );
    fclose( fp );
  }
}

begin_diagnostics {
  const int every = env_int( "VPB_DECK_ENERGIES", 1 );
  if( step == 0 || step == num_step || ( every > 0 && step % every == 0 ) ) dump_energies( "energies", step == 0 ? 0 : 1 );
  if( rank() == 0 ) {
    struct timespec ts;
    clock_gettime( CLOCK_MONOTONIC, &ts );
    FILE * fp = fopen( "steps.txt", step == 0 ? "w" : "a" );
    if( fp ) { fprintf( fp, "%d %.9f\n", int(step), double(ts.tv_sec) + 1e-9 * double(ts.tv_nsec) ); fclose( fp ); }
  }
}

begin_particle_injection { }
begin_current_injection { }
begin_field_injection { }
begin_particle_collisions { }
