// TEST INFRASTRUCTURE (oracle shim) -- never linked into the product library.
//
// Mini-MPI: the handful of MPI entry points the reference calls, implemented
// with fork() + one anonymous shared mapping so that the *verbatim* reference sources under
// /root/reference/Source can run at P = 1, 2, 4, 8 ranks on a box with no MPI installation.
// Sums are always taken in rank order, so results are deterministic.
//
//   PNOL_SHIM_NPROCS=<P>   number of ranks MPI_Init forks to (default 1, no fork)
//
// Only MPI_COMM_WORLD, MPI_DOUBLE, MPI_INT and MPI_SUM exist.
#ifndef PNOL_ORACLE_SHIM_MPI_H_
#define PNOL_ORACLE_SHIM_MPI_H_

#include <cstddef>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;

#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 1
#define MPI_INT 2
#define MPI_SUM 1
#define MPI_SUCCESS 0

int MPI_Init(int * argc, char *** argv);
int MPI_Finalize();
int MPI_Comm_size(MPI_Comm comm, int * size);
int MPI_Comm_rank(MPI_Comm comm, int * rank);
int MPI_Barrier(MPI_Comm comm);
int MPI_Allreduce(const void * sendbuf, void * recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Reduce(const void * sendbuf, void * recvbuf, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm comm);
int MPI_Bcast(void * buf, int count, MPI_Datatype dt, int root, MPI_Comm comm);
double MPI_Wtime();

#endif
