/* This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0 (LICENSE at the repository root).
 * It mirrors the interface / host control flow of briandaniel/ParallelNonlinearOptimizationLibrary (MPL-2.0); see NOTICE. */
/*
 * mpi.h -- what is left of MPI for a user translation unit that is compiled against include/pnol instead of the reference's
 * Source/ (the reference's headers #include <mpi.h>, /root/reference/Source/PNOL_Objective.hpp:20, and its example drivers ask
 * for the rank to print once, Source/Examples.cpp:528-531). Put this directory on the include path ONLY when no real MPI is
 * wanted (it lives in its own directory, include/pnol/nompi, for that reason): the "parallel machine" of this library is pnol::Runtime (one process per B200; NCCL below the C-ABI), so the
 * process-level queries answer from the torchrun environment (RANK / WORLD_SIZE, else 0 / 1) and there are deliberately NO data
 * collectives here -- a user call to MPI_Allreduce & co. fails to compile, which marks the call site that has to move to a
 * pnol_* entry point.
 */
#ifndef PNOL_MPI_FACADE_H_
#define PNOL_MPI_FACADE_H_

#include <chrono>
#include <cstdlib>

typedef int MPI_Comm;
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0

inline int pnol_mpi_env_int(const char * name, int fallback)
{
	const char * v = std::getenv(name);
	return (v && *v) ? std::atoi(v) : fallback;
}
inline int MPI_Init(int *, char ***) { return MPI_SUCCESS; }
inline int MPI_Finalize() { return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm, int * size) { *size = pnol_mpi_env_int("WORLD_SIZE", 1); return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm, int * rank) { *rank = pnol_mpi_env_int("RANK", 0); return MPI_SUCCESS; }
inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
inline double MPI_Wtime()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#endif
