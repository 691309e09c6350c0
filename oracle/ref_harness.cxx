// TEST INFRASTRUCTURE -- not part of the product.
//
// Flat helpers compiled INTO oracle/_ref/libvpic_ref_*.so next to the unmodified
// reference objects, so that tests/ and bench.py's cpu_baseline leg can drive the
// reference's own C API through ctypes.  It contains no algorithm: it only boots
// the reference's dispatchers the way src/main.cxx:72-80 does, hands out the
// addresses of the reference's field-advance vtables (field_advance.h:318-345)
// and reports the reference's struct layouts so that include/vpic_b200_abi.h can
// be verified against them (tests/test_abi.py).
#include <stddef.h>
#include <vpic.hxx>

#ifndef V4_ACCELERATION  // field_advance.h:334-347 only declares these when V4 is on
extern "C" field_advance_methods_t _standard_v4_field_advance[1];
extern "C" field_advance_methods_t _vacuum_v4_field_advance[1];
#endif

extern "C" {

// src/main.cxx:72-80 -- thread.boot / serial.boot / mp_init, once per process.
int refh_boot( int tpp ) {
  static int booted = 0;
  if( booted ) return thread.n_pipeline;
  thread.boot( tpp, 1 );
  serial.boot( tpp, 1 );
  mp_init( 0, NULL );
  booted = 1;
  return thread.n_pipeline;
}

int refh_n_pipeline( void ) { return thread.n_pipeline; }

// 0: _standard_field_advance  1: _vacuum_field_advance
// 2: _standard_v4_field_advance  3: _vacuum_v4_field_advance
field_advance_methods_t * refh_vtable( int which ) {
  switch( which ) {
  case 0: return _standard_field_advance;
  case 1: return _vacuum_field_advance;
  case 2: return _standard_v4_field_advance;
  case 3: return _vacuum_v4_field_advance;
  }
  return NULL;
}

// Is this build the V4 flavour (advance_p etc. dispatch the V4 pipelines)?
int refh_is_v4( void ) {
#ifdef V4_ACCELERATION
  return 1;
#else
  return 0;
#endif
}

// Struct layout report; order is mirrored by tests/test_abi.py.
int refh_layout( long * out, int max ) {
  long v[] = {
    (long)sizeof(particle_t), (long)offsetof(particle_t,i), (long)offsetof(particle_t,ux),
    (long)offsetof(particle_t,q), (long)offsetof(particle_t,tag), (long)offsetof(particle_t,tag2),
    (long)sizeof(particle_mover_t), (long)offsetof(particle_mover_t,i),
    (long)sizeof(particle_injector_t), (long)offsetof(particle_injector_t,dispx), (long)offsetof(particle_injector_t,sp_id),
    (long)sizeof(interpolator_t), (long)offsetof(interpolator_t,ey), (long)offsetof(interpolator_t,cbx), (long)offsetof(interpolator_t,dcbzdz),
    (long)sizeof(accumulator_t), (long)offsetof(accumulator_t,jy), (long)offsetof(accumulator_t,jz),
    (long)sizeof(field_t), (long)offsetof(field_t,cbx), (long)offsetof(field_t,tcax), (long)offsetof(field_t,rhob),
    (long)offsetof(field_t,jfx), (long)offsetof(field_t,rhof), (long)offsetof(field_t,ematx), (long)offsetof(field_t,fmatx), (long)offsetof(field_t,cmat),
    (long)sizeof(hydro_t),
    (long)sizeof(grid_t), (long)offsetof(grid_t,dt), (long)offsetof(grid_t,damp), (long)offsetof(grid_t,x0),
    (long)offsetof(grid_t,dx), (long)offsetof(grid_t,rdx), (long)offsetof(grid_t,nx), (long)offsetof(grid_t,bc),
    (long)offsetof(grid_t,range), (long)offsetof(grid_t,neighbor), (long)offsetof(grid_t,rangel),
    (long)offsetof(grid_t,rangeh), (long)offsetof(grid_t,nb), (long)offsetof(grid_t,boundary),
    (long)sizeof(species_t), (long)offsetof(species_t,np), (long)offsetof(species_t,max_np), (long)offsetof(species_t,p),
    (long)offsetof(species_t,nm), (long)offsetof(species_t,max_nm), (long)offsetof(species_t,pm), (long)offsetof(species_t,q_m),
    (long)offsetof(species_t,sort_interval), (long)offsetof(species_t,sort_out_of_place),
    (long)offsetof(species_t,partition), (long)offsetof(species_t,next), (long)offsetof(species_t,name),
    (long)sizeof(field_advance_methods_t), (long)offsetof(field_advance_methods_t,advance_b),
    (long)offsetof(field_advance_methods_t,energy_f), (long)offsetof(field_advance_methods_t,clean_div_b),
    (long)sizeof(field_advance_t), (long)offsetof(field_advance_t,method),
    (long)sizeof(material_t), (long)offsetof(material_t,epsx), (long)offsetof(material_t,next), (long)offsetof(material_t,name)
  };
  int n = (int)( sizeof(v)/sizeof(v[0]) );
  for( int i=0; i<n && i<max; i++ ) out[i] = v[i];
  return n;
}

} // extern "C"
