// TEST INFRASTRUCTURE -- an input deck written for this repository (not part of the reference).  It pins the
// ORCHESTRATION the history tests rely on (tests/test_gpu_history.py::cpu_history, the call order the C++ driver
// vpb_step.cu follows) to the reference's own vpic_simulation::initialize()/advance(): at step 0 it writes the state
// advance() starts from (every species' particle array after initialize() has uncentred it, and the field array) and
// after every step the eight numbers dump_energies prints, as raw doubles.  tests/test_history_vs_ref_deck.py replays
// the run from that state with the oracle kernels and compares.  12 x 10 x 8 cells, 8 particles per cell per species
// (a multiple of 16 per species: every particle goes through pipeline 0, so float sums are in array order), periodic,
// sort every 5 steps, both divergence cleanings every 5; VPB_PIN_SYNC in the environment sets sync_shared_interval
// (default 0), VPB_PIN_CLEAN the two cleaning intervals (default 5).
#include <stdio.h>
#include <stdlib.h>

begin_globals {
  int dummy;
};

begin_initialization {
  const int nx = 12, ny = 10, nz = 8, ppc = 8;
  const double Ne = double(nx) * ny * nz * ppc;

  num_step = 20;
  status_interval = 0;
  sync_shared_interval = getenv( "VPB_PIN_SYNC" ) ? atoi( getenv( "VPB_PIN_SYNC" ) ) : 0;
  clean_div_e_interval = getenv( "VPB_PIN_CLEAN" ) ? atoi( getenv( "VPB_PIN_CLEAN" ) ) : 5;
  clean_div_b_interval = clean_div_e_interval;

  grid->dt = 0.95 * courant_length( nx, ny, nz, nx, ny, nz );
  grid->cvac = 1;
  grid->eps0 = 1;
  grid->damp = 0;
  define_periodic_grid( 0, 0, 0, nx, ny, nz, nx, ny, nz, 1, 1, 1 );

  define_material( "vacuum", 1 );
  finalize_field_advance( standard_field_advance );

  species_t * electron = define_species( "electron", -1, 1.5 * Ne, -1, 5, 1 );
  species_t * ion      = define_species( "ion",       1, 1.5 * Ne, -1, 5, 1 );

  set_region_field( everywhere, 0.01 * sin( 0.5 * x ), 0, 0.02 * cos( 0.7 * y ), 0, 0.05, 0.03 * sin( 0.4 * x ) );

  seed_rand( 13 );
  const double q = double(nx) * ny * nz / Ne;
  for( int k = 0; k < int(Ne); k++ ) {
    const double x = uniform_rand( 0, nx ), y = uniform_rand( 0, ny ), z = uniform_rand( 0, nz );
    inject_particle( electron, x, y, z, maxwellian_rand( 0.15 ), maxwellian_rand( 0.15 ), maxwellian_rand( 0.15 ), -q, k, 0, 0 );
    inject_particle( ion,      x, y, z, maxwellian_rand( 0.15 ), maxwellian_rand( 0.15 ), maxwellian_rand( 0.15 ),  q, k, 0, 0 );
  }
}

begin_diagnostics {
  if( step == 0 ) {
    FILE * fp = fopen( "state0.bin", "wb" );
    species_t * sp;
    int nsp = 0;
    LIST_FOR_EACH( sp, species_list ) nsp++;
    const int nv = ( grid->nx + 2 ) * ( grid->ny + 2 ) * ( grid->nz + 2 );
    fwrite( &nsp, sizeof(int), 1, fp );
    fwrite( &nv, sizeof(int), 1, fp );
    fwrite( field, sizeof(field_t), nv, fp );
    LIST_FOR_EACH( sp, species_list ) {          // list order = the order advance() visits them in
      fwrite( &sp->np, sizeof(int), 1, fp );
      fwrite( &sp->q_m, sizeof(float), 1, fp );
      fwrite( sp->p, sizeof(particle_t), sp->np, fp );
    }
    fclose( fp );
  }
  {
    double row[16];
    int n = 6;
    species_t * sp;
    field_advance->method->energy_f( row, field, field_advance->m, grid );
    LIST_FOR_EACH( sp, species_list ) row[n++] = energy_p( sp->p, sp->np, sp->q_m, interpolator, grid );
    FILE * fp = fopen( "hist.bin", step == 0 ? "wb" : "ab" );
    fwrite( row, sizeof(double), n, fp );
    fclose( fp );
  }
}

begin_particle_injection { }
begin_current_injection { }
begin_field_injection { }
begin_particle_collisions { }
