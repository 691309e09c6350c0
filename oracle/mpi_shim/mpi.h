/* TEST INFRASTRUCTURE -- not part of the product.
 *
 * Minimal MPI-1 subset, just wide enough to build and run the reference
 * (pdlfs/old-vpic) unmodified in a container that has no MPI.  The reference
 * includes <mpi.h> unconditionally (src/util/mp/dmp/mp_t.h:4) and uses the 17
 * entry points below (src/util/mp/dmp/mp_dmp.c:37-400, mp_t.h:7-19).
 *
 * One rank: point-to-point messages to self are matched by tag in either
 * posting order.  N ranks: processes forked by the test harness share a POSIX
 * shm segment (see mpi_shim.c); env VPIC_SHIM_RANK / VPIC_SHIM_NPROC /
 * VPIC_SHIM_SHM select it.  Functional only -- not a performance transport.
 */
#ifndef VPIC_B200_ORACLE_MPI_SHIM_H
#define VPIC_B200_ORACLE_MPI_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR, count_bytes; } MPI_Status;

#define MPI_COMM_WORLD     0
#define MPI_STATUS_IGNORE  ((MPI_Status *)0)

enum { MPI_SUCCESS = 0, MPI_ERR_ARG, MPI_ERR_COMM, MPI_ERR_COUNT, MPI_ERR_OTHER,
       MPI_ERR_RANK, MPI_ERR_REQUEST, MPI_ERR_TAG, MPI_ERR_TYPE };

enum { MPI_BYTE = 1, MPI_CHAR, MPI_INT, MPI_DOUBLE, MPI_LONG_LONG, MPI_LONG_LONG_INT, MPI_FLOAT };
enum { MPI_SUM = 1, MPI_MAX };

int    MPI_Init( int *argc, char ***argv );
int    MPI_Finalize( void );
int    MPI_Abort( MPI_Comm comm, int reason );
int    MPI_Comm_rank( MPI_Comm comm, int *rank );
int    MPI_Comm_size( MPI_Comm comm, int *size );
double MPI_Wtime( void );
int    MPI_Barrier( MPI_Comm comm );
int    MPI_Allreduce( const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm comm );
int    MPI_Reduce( const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm comm );
int    MPI_Allgather( const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm comm );
int    MPI_Gather( const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, int root, MPI_Comm comm );
int    MPI_Send( const void *buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm comm );
int    MPI_Recv( void *buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm comm, MPI_Status *st );
int    MPI_Irecv( void *buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm comm, MPI_Request *req );
int    MPI_Issend( const void *buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm comm, MPI_Request *req );
int    MPI_Wait( MPI_Request *req, MPI_Status *st );
int    MPI_Get_count( const MPI_Status *st, MPI_Datatype t, int *count );

#ifdef __cplusplus
}
#endif
#endif
