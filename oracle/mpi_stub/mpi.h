/* oracle/mpi_stub/mpi.h -- TEST INFRASTRUCTURE ONLY.
 * Single-rank stand-in for <mpi.h> so that the reference's *_MPI.c sources compile in oracle/_ref/
 * without an MPI runtime (none is installed).  With one rank the reference's static row split
 * (/root/reference/sources/main_MIDASPOM_MPI.c:361-372) gives rank 0 every grid row and the
 * Send/Recv gather (:483-505) is never reached by a worker, so Send/Recv abort if called.
 */
#ifndef ORACLE_MPI_STUB_H
#define ORACLE_MPI_STUB_H
#include <stdlib.h>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 1
#define MPI_INT 2
#define MPI_ANY_SOURCE (-1)
#define MPI_SUCCESS 0
static inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return 0; }
static inline int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = 1; return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Send(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; abort(); return 0; }
static inline int MPI_Recv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status *st)
{ (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)st; abort(); return 0; }
#endif
