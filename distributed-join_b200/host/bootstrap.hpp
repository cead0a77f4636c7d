// bootstrap.hpp -- what the reference takes from MPI outside the data path, without MPI
// (SURVEY.md App. C): rank/size discovery, a byte broadcast for the ncclUniqueId, barrier,
// wall clock, small all-reduces.  Ranks are started by any launcher that exports RANK and
// WORLD_SIZE (torchrun --no-python, mpirun's OMPI_COMM_WORLD_*, or scripts/djrun.sh); the
// rendezvous is a file under $DJ_RENDEZVOUS_DIR (default /tmp) keyed by MASTER_PORT.
#pragma once

#include <cstddef>
#include <cstdint>

struct dj_comm;
class Communicator;

namespace dj_bootstrap {

void init(int* argc, char*** argv);  // MPI_Init
void finalize();                     // MPI_Finalize
int rank();                          // MPI_Comm_rank(MPI_COMM_WORLD)
int size();                          // MPI_Comm_size(MPI_COMM_WORLD)
double wtime();                      // MPI_Wtime

// MPI_Bcast of `bytes` bytes from rank 0 through the rendezvous file (used once, for the
// 128-byte ncclUniqueId, before any communicator exists).
void broadcast_from_root(void* buf, std::size_t bytes, const char* tag);

// After the communicator exists the collectives ride on it (NCCL): MPI_Barrier,
// MPI_Allreduce(MAX / SUM) on one value.
void set_communicator(Communicator* c);
void barrier();
double allreduce_max(double v);
int64_t allreduce_sum(int64_t v);

}  // namespace dj_bootstrap
