// TEST INFRASTRUCTURE -- an input deck written for this repository (not part of the reference): a hot pair plasma in a
// 12 x 10 x 8 box whose six faces absorb fields (first-order Higdon) and particles: every step particles reach a wall,
// boundary_p removes them, accumulates their charge into rhob and back-fills the arrays.  Energies every step and the
// particle counts at the end.  Built like thermal_small.cxx (oracle/build_hybrid.sh).
#include <stdio.h>

begin_globals {
  int dummy;
};

begin_initialization {
  const double Lx = 12, Ly = 10, Lz = 8;
  const int nx = 12, ny = 10, nz = 8, ppc = 16;
  const double Ne = double(nx) * ny * nz * ppc;

  num_step = 20;
  status_interval = 0;
  sync_shared_interval = 0;
  clean_div_e_interval = 10;
  clean_div_b_interval = 10;

  grid->dt = 0.95 * courant_length( Lx, Ly, Lz, nx, ny, nz );
  grid->cvac = 1;
  grid->eps0 = 1;
  grid->damp = 0;
  define_absorbing_grid( 0, 0, 0, Lx, Ly, Lz, nx, ny, nz, int( nproc() ), 1, 1, absorb_particles );   // split along x over the ranks of the job

  define_material( "vacuum", 1 );
  finalize_field_advance( standard_field_advance );

  species_t * electron = define_species( "electron", -1, 1.5 * Ne, -1, 5, 1 );
  species_t * ion      = define_species( "ion",       1, 1.5 * Ne, -1, 5, 1 );

  seed_rand( 13 );
  const double q = Lx * Ly * Lz / Ne;
  for( int k = 0; k < int(Ne); k++ ) {
    const double x = uniform_rand( 0, Lx ), y = uniform_rand( 0, Ly ), z = uniform_rand( 0, Lz );
    inject_particle( electron, x, y, z, maxwellian_rand( 0.3 ), maxwellian_rand( 0.3 ), maxwellian_rand( 0.3 ), -q, k, 0, 0 );
    inject_particle( ion,      x, y, z, maxwellian_rand( 0.1 ), maxwellian_rand( 0.1 ), maxwellian_rand( 0.1 ),  q, k, 0, 0 );
  }
}

begin_diagnostics {
  dump_energies( "energies", step == 0 ? 0 : 1 );
  if( step == num_step ) {
    char name[64];
    if( nproc() > 1 ) sprintf( name, "counts.%d", int( rank() ) );    // one file per rank; the test adds them up
    else              sprintf( name, "counts" );
    FILE * fp = fopen( name, "w" );
    for( species_t * sp = species_list; sp; sp = sp->next ) fprintf( fp, "%s %d\n", sp->name, sp->np );
    fclose( fp );
  }
}

begin_particle_injection { }
begin_current_injection { }
begin_field_injection { }
begin_particle_collisions { }
