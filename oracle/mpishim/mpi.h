/* Single-rank stand-in for <mpi.h>, used ONLY to compile the unmodified reference
 * counting binaries (cherryml/counting/_count_transitions.cpp:14,595-597,652,673 use
 * exactly these five calls; no data moves through MPI).  Test infrastructure. */
#ifndef CHERRY_ORACLE_MPI_SHIM_H
#define CHERRY_ORACLE_MPI_SHIM_H
typedef int MPI_Comm;
#define MPI_COMM_WORLD 0
static inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int *n) { (void)c; *n = 1; return 0; }
static inline int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
static inline int MPI_Finalize(void) { return 0; }
#endif
