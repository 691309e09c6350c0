// TEST INFRASTRUCTURE -- an input deck written for this repository (not part of the reference): the geometry of
// decks/trecon-part scaled down.  2D box 24 x 1 x 16 cells, periodic in x and y, conducting walls that reflect
// particles at the two z faces (turbulence.cxx:262-270), force-free current sheet Bx = b0 tanh(z/w), By = b0 sech(z/w)
// (turbulence.cxx:441-460), a hot pair plasma with the electrons drifting along the sheet, divergence cleaning every
// 10 steps, energies every step and an electron hydro dump at the last step (clear_hydro, accumulate_hydro_p,
// synchronize_hydro through src/vpic/dump.cxx).  Built like thermal_small.cxx: once on the reference alone, once on
// the reference's host objects + libvpic_b200.so (oracle/build_hybrid.sh).
begin_globals {
  int dummy;
};

begin_initialization {
  const double Lx = 24, Ly = 1, Lz = 16, b0 = 0.8, w = 2.0;
  const int nx = 24, ny = 1, nz = 16, ppc = 16;
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
  define_periodic_grid( 0, -0.5 * Ly, -0.5 * Lz, Lx, 0.5 * Ly, 0.5 * Lz, nx, ny, nz, int( nproc() ), 1, 1 );   // split along x: every rank keeps both z walls
  set_domain_field_bc( BOUNDARY( 0, 0, -1 ), pec_fields );
  set_domain_field_bc( BOUNDARY( 0, 0,  1 ), pec_fields );
  set_domain_particle_bc( BOUNDARY( 0, 0, -1 ), reflect_particles );
  set_domain_particle_bc( BOUNDARY( 0, 0,  1 ), reflect_particles );

  define_material( "vacuum", 1 );
  finalize_field_advance( standard_field_advance );

  species_t * electron = define_species( "electron", -1, 1.5 * Ne, -1, 5, 1 );
  species_t * ion      = define_species( "ion",       1, 1.5 * Ne, -1, 5, 1 );

  set_region_field( everywhere, 0, 0, 0, b0 * tanh( z / w ), b0 / cosh( z / w ), 0 );

  seed_rand( 11 );
  const double q = Lx * Ly * Lz / Ne;
  for( int k = 0; k < int(Ne); k++ ) {
    const double x = uniform_rand( 0, Lx ), y = uniform_rand( -0.5 * Ly, 0.5 * Ly ), z = uniform_rand( -0.5 * Lz, 0.5 * Lz );
    const double drift = 0.2 / ( cosh( z / w ) * cosh( z / w ) );
    inject_particle( electron, x, y, z, maxwellian_rand( 0.3 ), maxwellian_rand( 0.3 ) + drift, maxwellian_rand( 0.3 ), -q, k, 0, 0 );
    inject_particle( ion,      x, y, z, maxwellian_rand( 0.3 ), maxwellian_rand( 0.3 ),         maxwellian_rand( 0.3 ),  q, k, 0, 0 );
  }
}

begin_diagnostics {
  dump_energies( "energies", step == 0 ? 0 : 1 );
  if( step == num_step ) dump_hydro( "electron", "ehydro", 0 );
}

begin_particle_injection { }
begin_current_injection { }
begin_field_injection { }
begin_particle_collisions { }
