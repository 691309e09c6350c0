/* TEST INFRASTRUCTURE -- not part of the product.
 *
 * Backing implementation of the MPI-1 subset declared in mpi.h (see there for
 * why it exists).  Two transports:
 *
 *  - nproc==1: messages to self are matched by tag in either posting order
 *    (the reference posts recv-then-send in some passes and send-then-recv in
 *    others, src/field_advance/standard/remote.c:378-405).
 *  - nproc>1: each rank is a separate process (forked by tests/ harness).  All
 *    ranks mmap one sparse POSIX-shm file, laid out as
 *        [header][per-rank collective scratch][mailbox slots (src,dst,tag)]
 *    A send copies into its (src,dst,tag) slot (buffered), a recv drains it in
 *    MPI_Wait.  Collectives go through the scratch area and a sense-reversing
 *    barrier and reduce in rank order, so results are deterministic.
 *
 * Environment: VPIC_SHIM_NPROC, VPIC_SHIM_RANK, VPIC_SHIM_SHM (path of the
 * zero-filled shm file, created by the harness), VPIC_SHIM_SLOT_MB (default 4).
 */
#define _GNU_SOURCE
#include "mpi.h"
#include <fcntl.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#define SHIM_MAX_REQ   512
#define SHIM_NTAG      64
#define SHIM_SCRATCH   (1<<16)

typedef struct shim_hdr {
  volatile int barrier_count;
  volatile int barrier_gen;
  int pad[14];
} shim_hdr_t;

typedef struct shim_slot_hdr {
  volatile int full;
  volatile int bytes;
  int pad[14];
} shim_slot_hdr_t;

typedef struct shim_req {
  int kind;          /* 0 free, 1 recv, 2 send */
  int done;
  int peer, tag, bytes, actual;
  void *buf;
} shim_req_t;

static int        g_rank = 0, g_nproc = 1, g_inited = 0;
static char      *g_shm = NULL;
static size_t     g_slot_bytes = 0;  /* payload capacity per slot */
static shim_req_t g_req[SHIM_MAX_REQ];

static size_t dtype_size( MPI_Datatype t ) {
  switch( t ) {
  case MPI_BYTE: case MPI_CHAR: return 1;
  case MPI_INT: case MPI_FLOAT: return 4;
  case MPI_DOUBLE: case MPI_LONG_LONG: case MPI_LONG_LONG_INT: return 8;
  default: fprintf( stderr, "mpi_shim: bad datatype %d\n", t ); abort();
  }
}

static shim_hdr_t *hdr( void ) { return (shim_hdr_t *)g_shm; }
static char *scratch( int r ) { return g_shm + sizeof(shim_hdr_t) + (size_t)r*SHIM_SCRATCH; }
static shim_slot_hdr_t *slot( int src, int dst, int tag ) {
  size_t stride = sizeof(shim_slot_hdr_t) + g_slot_bytes;
  size_t idx = ((size_t)src*g_nproc + dst)*SHIM_NTAG + tag;
  return (shim_slot_hdr_t *)( g_shm + sizeof(shim_hdr_t) + (size_t)g_nproc*SHIM_SCRATCH + idx*stride );
}

static void spin_pause( int *n ) {
  if( ++(*n) > 64 ) { sched_yield(); *n = 0; }
}

static void shim_barrier( void ) {
  if( g_nproc==1 ) return;
  int gen = __atomic_load_n( &hdr()->barrier_gen, __ATOMIC_ACQUIRE );
  int n = __atomic_add_fetch( &hdr()->barrier_count, 1, __ATOMIC_ACQ_REL );
  if( n==g_nproc ) {
    __atomic_store_n( &hdr()->barrier_count, 0, __ATOMIC_RELEASE );
    __atomic_add_fetch( &hdr()->barrier_gen, 1, __ATOMIC_ACQ_REL );
  } else {
    int s = 0;
    while( __atomic_load_n( &hdr()->barrier_gen, __ATOMIC_ACQUIRE )==gen ) spin_pause( &s );
  }
}

int MPI_Init( int *argc, char ***argv ) {
  (void)argc; (void)argv;
  if( g_inited ) return MPI_SUCCESS;
  const char *s;
  if( (s = getenv( "VPIC_SHIM_NPROC" )) ) g_nproc = atoi( s );
  if( (s = getenv( "VPIC_SHIM_RANK"  )) ) g_rank  = atoi( s );
  if( g_nproc<1 || g_rank<0 || g_rank>=g_nproc ) { fprintf( stderr, "mpi_shim: bad rank/nproc\n" ); abort(); }
  if( g_nproc>1 ) {
    size_t mb = 4;
    if( (s = getenv( "VPIC_SHIM_SLOT_MB" )) ) mb = (size_t)atoi( s );
    g_slot_bytes = mb<<20;
    const char *path = getenv( "VPIC_SHIM_SHM" );
    if( !path ) { fprintf( stderr, "mpi_shim: VPIC_SHIM_SHM unset\n" ); abort(); }
    size_t total = sizeof(shim_hdr_t) + (size_t)g_nproc*SHIM_SCRATCH +
                   (size_t)g_nproc*g_nproc*SHIM_NTAG*( sizeof(shim_slot_hdr_t) + g_slot_bytes );
    int fd = open( path, O_RDWR );
    if( fd<0 ) { perror( "mpi_shim: open shm" ); abort(); }
    struct stat st; fstat( fd, &st );
    if( (size_t)st.st_size<total && ftruncate( fd, (off_t)total )!=0 ) { perror( "mpi_shim: ftruncate" ); abort(); }
    g_shm = (char *)mmap( NULL, total, PROT_READ|PROT_WRITE, MAP_SHARED|MAP_NORESERVE, fd, 0 );
    if( g_shm==MAP_FAILED ) { perror( "mpi_shim: mmap" ); abort(); }
    close( fd );
  }
  memset( g_req, 0, sizeof(g_req) );
  g_inited = 1;
  shim_barrier();
  return MPI_SUCCESS;
}

int MPI_Finalize( void ) { shim_barrier(); return MPI_SUCCESS; }
int MPI_Abort( MPI_Comm c, int reason ) { (void)c; fprintf( stderr, "mpi_shim: MPI_Abort(%d)\n", reason ); _exit( reason ? reason : 1 ); }
int MPI_Comm_rank( MPI_Comm c, int *r ) { (void)c; *r = g_rank;  return MPI_SUCCESS; }
int MPI_Comm_size( MPI_Comm c, int *n ) { (void)c; *n = g_nproc; return MPI_SUCCESS; }
double MPI_Wtime( void ) { struct timespec ts; clock_gettime( CLOCK_MONOTONIC, &ts ); return ts.tv_sec + 1e-9*ts.tv_nsec; }
int MPI_Barrier( MPI_Comm c ) { (void)c; shim_barrier(); return MPI_SUCCESS; }

/* ---- collectives ------------------------------------------------------ */

static void reduce_into( void *r, const void *s, int n, MPI_Datatype t, MPI_Op op, int first ) {
  int i;
  if( first ) { memcpy( r, s, (size_t)n*dtype_size( t ) ); return; }
  if( t==MPI_DOUBLE ) { double *a=(double*)r; const double *b=(const double*)s;
    for( i=0;i<n;i++ ) a[i] = op==MPI_SUM ? a[i]+b[i] : (a[i]>b[i]?a[i]:b[i]); }
  else if( t==MPI_INT ) { int *a=(int*)r; const int *b=(const int*)s;
    for( i=0;i<n;i++ ) a[i] = op==MPI_SUM ? a[i]+b[i] : (a[i]>b[i]?a[i]:b[i]); }
  else if( t==MPI_LONG_LONG || t==MPI_LONG_LONG_INT ) { long long *a=(long long*)r; const long long *b=(const long long*)s;
    for( i=0;i<n;i++ ) a[i] = op==MPI_SUM ? a[i]+b[i] : (a[i]>b[i]?a[i]:b[i]); }
  else if( t==MPI_FLOAT ) { float *a=(float*)r; const float *b=(const float*)s;
    for( i=0;i<n;i++ ) a[i] = op==MPI_SUM ? a[i]+b[i] : (a[i]>b[i]?a[i]:b[i]); }
  else { fprintf( stderr, "mpi_shim: reduce on bad type\n" ); abort(); }
}

int MPI_Allreduce( const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c ) {
  (void)c;
  size_t bytes = (size_t)n*dtype_size( t );
  if( g_nproc==1 ) { memmove( r, s, bytes ); return MPI_SUCCESS; }
  if( bytes>SHIM_SCRATCH ) { fprintf( stderr, "mpi_shim: allreduce too big\n" ); abort(); }
  memcpy( scratch( g_rank ), s, bytes );
  shim_barrier();
  for( int k=0; k<g_nproc; k++ ) reduce_into( r, scratch( k ), n, t, op, k==0 );
  shim_barrier();
  return MPI_SUCCESS;
}

int MPI_Reduce( const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c ) {
  (void)c;
  size_t bytes = (size_t)n*dtype_size( t );
  if( g_nproc==1 ) { memmove( r, s, bytes ); return MPI_SUCCESS; }
  if( bytes>SHIM_SCRATCH ) { fprintf( stderr, "mpi_shim: reduce too big\n" ); abort(); }
  memcpy( scratch( g_rank ), s, bytes );
  shim_barrier();
  if( g_rank==root ) for( int k=0; k<g_nproc; k++ ) reduce_into( r, scratch( k ), n, t, op, k==0 );
  shim_barrier();
  return MPI_SUCCESS;
}

int MPI_Allgather( const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c ) {
  (void)c; (void)nr; (void)tr;
  size_t bytes = (size_t)ns*dtype_size( ts );
  if( g_nproc==1 ) { memmove( r, s, bytes ); return MPI_SUCCESS; }
  if( bytes>SHIM_SCRATCH ) { fprintf( stderr, "mpi_shim: allgather too big\n" ); abort(); }
  memcpy( scratch( g_rank ), s, bytes );
  shim_barrier();
  for( int k=0; k<g_nproc; k++ ) memcpy( (char *)r + (size_t)k*bytes, scratch( k ), bytes );
  shim_barrier();
  return MPI_SUCCESS;
}

int MPI_Gather( const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, int root, MPI_Comm c ) {
  (void)c; (void)nr; (void)tr;
  size_t bytes = (size_t)ns*dtype_size( ts );
  if( g_nproc==1 ) { memmove( r, s, bytes ); return MPI_SUCCESS; }
  if( bytes>SHIM_SCRATCH ) { fprintf( stderr, "mpi_shim: gather too big\n" ); abort(); }
  memcpy( scratch( g_rank ), s, bytes );
  shim_barrier();
  if( g_rank==root ) for( int k=0; k<g_nproc; k++ ) memcpy( (char *)r + (size_t)k*bytes, scratch( k ), bytes );
  shim_barrier();
  return MPI_SUCCESS;
}

/* ---- point to point --------------------------------------------------- */

static int new_req( void ) {
  for( int i=1; i<SHIM_MAX_REQ; i++ ) if( g_req[i].kind==0 ) return i;
  fprintf( stderr, "mpi_shim: out of requests\n" ); abort();
}

static void remote_put( const void *buf, int bytes, int dst, int tag ) {
  if( tag<0 || tag>=SHIM_NTAG ) { fprintf( stderr, "mpi_shim: tag %d out of range\n", tag ); abort(); }
  if( (size_t)bytes>g_slot_bytes ) { fprintf( stderr, "mpi_shim: message of %d B exceeds slot\n", bytes ); abort(); }
  shim_slot_hdr_t *sl = slot( g_rank, dst, tag );
  int s = 0;
  while( __atomic_load_n( &sl->full, __ATOMIC_ACQUIRE ) ) spin_pause( &s );
  memcpy( (char *)(sl+1), buf, (size_t)bytes );
  sl->bytes = bytes;
  __atomic_store_n( &sl->full, 1, __ATOMIC_RELEASE );
}

static int remote_get( void *buf, int bytes, int src, int tag ) {
  if( tag<0 || tag>=SHIM_NTAG ) { fprintf( stderr, "mpi_shim: tag %d out of range\n", tag ); abort(); }
  shim_slot_hdr_t *sl = slot( src, g_rank, tag );
  int s = 0;
  while( !__atomic_load_n( &sl->full, __ATOMIC_ACQUIRE ) ) spin_pause( &s );
  int got = sl->bytes;
  memcpy( buf, (char *)(sl+1), (size_t)( got<bytes ? got : bytes ) );
  __atomic_store_n( &sl->full, 0, __ATOMIC_RELEASE );
  return got;
}

int MPI_Irecv( void *buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *req ) {
  (void)c;
  int id = new_req();
  shim_req_t *r = &g_req[id];
  r->kind = 1; r->done = 0; r->peer = src; r->tag = tag; r->buf = buf;
  r->bytes = (int)( (size_t)n*dtype_size( t ) ); r->actual = 0;
  if( src==g_rank ) { /* look for an already-posted matching self send */
    for( int i=1; i<SHIM_MAX_REQ; i++ ) {
      shim_req_t *q = &g_req[i];
      if( q->kind==2 && !q->done && q->peer==g_rank && q->tag==tag ) {
        r->actual = q->bytes;
        memcpy( buf, q->buf, (size_t)( q->bytes<r->bytes ? q->bytes : r->bytes ) );
        r->done = q->done = 1;
        break;
      }
    }
  }
  *req = id;
  return MPI_SUCCESS;
}

int MPI_Issend( const void *buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request *req ) {
  (void)c;
  int id = new_req();
  shim_req_t *r = &g_req[id];
  r->kind = 2; r->done = 0; r->peer = dst; r->tag = tag; r->buf = (void *)buf;
  r->bytes = (int)( (size_t)n*dtype_size( t ) ); r->actual = r->bytes;
  if( dst==g_rank ) {
    for( int i=1; i<SHIM_MAX_REQ; i++ ) {
      shim_req_t *q = &g_req[i];
      if( q->kind==1 && !q->done && q->peer==g_rank && q->tag==tag ) {
        q->actual = r->bytes;
        memcpy( q->buf, buf, (size_t)( r->bytes<q->bytes ? r->bytes : q->bytes ) );
        r->done = q->done = 1;
        break;
      }
    }
  } else {
    remote_put( buf, r->bytes, dst, tag );
    r->done = 1;
  }
  *req = id;
  return MPI_SUCCESS;
}

int MPI_Wait( MPI_Request *req, MPI_Status *st ) {
  int id = *req;
  if( id<=0 || id>=SHIM_MAX_REQ || g_req[id].kind==0 ) return MPI_ERR_REQUEST;
  shim_req_t *r = &g_req[id];
  if( !r->done ) {
    if( r->peer==g_rank ) { fprintf( stderr, "mpi_shim: unmatched self message (tag %d)\n", r->tag ); abort(); }
    if( r->kind==1 ) { r->actual = remote_get( r->buf, r->bytes, r->peer, r->tag ); r->done = 1; }
  }
  if( st ) { st->MPI_SOURCE = r->peer; st->MPI_TAG = r->tag; st->MPI_ERROR = MPI_SUCCESS; st->count_bytes = r->actual; }
  r->kind = 0;
  *req = 0;
  return MPI_SUCCESS;
}

int MPI_Get_count( const MPI_Status *st, MPI_Datatype t, int *count ) {
  *count = (int)( (size_t)st->count_bytes/dtype_size( t ) );
  return MPI_SUCCESS;
}

int MPI_Send( const void *buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c ) {
  MPI_Request rq; MPI_Issend( buf, n, t, dst, tag, c, &rq ); return MPI_Wait( &rq, MPI_STATUS_IGNORE );
}

int MPI_Recv( void *buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st ) {
  MPI_Request rq; MPI_Irecv( buf, n, t, src, tag, c, &rq ); return MPI_Wait( &rq, st );
}
