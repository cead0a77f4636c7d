#include "communicator.hpp"

#include "bootstrap.hpp"
#include "error.hpp"

// Backend-agnostic all-gather out of the reference's own primitives: every rank sends its words
// to every other rank inside one start()/stop() batch (device staging buffers, as send/recv
// take device memory).
void Communicator::allgather_i64(const int64_t* mine, int n, int64_t* all)
{
  int64_t* d = nullptr;
  CUDA_RT_CALL(cudaMalloc(&d, (size_t)n * (mpi_size + 1) * sizeof(int64_t)));
  CUDA_RT_CALL(cudaMemcpy(d, mine, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice));
  int64_t* d_all = d + n;
  CUDA_RT_CALL(cudaMemcpy(d_all + (size_t)mpi_rank * n, d, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToDevice));
  start();
  for (int r = 0; r < mpi_size; r++) {
    if (r == mpi_rank) continue;
    send(d, n, (int)sizeof(int64_t), r);
    recv(d_all + (size_t)r * n, n, (int)sizeof(int64_t), r);
  }
  stop();
  CUDA_RT_CALL(cudaMemcpy(all, d_all, (size_t)n * mpi_size * sizeof(int64_t), cudaMemcpyDeviceToHost));
  CUDA_RT_CALL(cudaFree(d));
}

void NCCLCommunicator::initialize()
{
  mpi_rank = dj_bootstrap::rank();
  mpi_size = dj_bootstrap::size();
  CUDA_RT_CALL(cudaGetDevice(&current_device));
  // rank 0 draws the ncclUniqueId; the launcher-side rendezvous replaces MPI_Bcast
  // (src/communicator.cpp:803-806)
  unsigned char id[128] = {0};
  if (mpi_rank == 0 && mpi_size > 1) NCCL_CALL(dj_comm_unique_id(id));
  dj_bootstrap::broadcast_from_root(id, sizeof(id), "nccl_id");
  NCCL_CALL(dj_comm_create(mpi_rank, mpi_size, mpi_size > 1 ? id : nullptr, &comm));
  nccl_comm = static_cast<ncclComm_t>(dj_comm_nccl_handle(comm));
  CUDA_RT_CALL(cudaStreamCreateWithFlags(&comm_stream, cudaStreamNonBlocking));
  dj_bootstrap::set_communicator(this);
}

void NCCLCommunicator::start() { NCCL_CALL(dj_comm_group_start(comm)); }

// No staging buffers: the reference copies every message into a freshly allocated 256 B-aligned
// buffer and back (src/communicator.cpp:820-869); NCCL on NVLink does not need that, and
// bucket starts produced by libdj_b200 are 8 B-aligned rows of 256 B-aligned columns.
void NCCLCommunicator::send(const void* buf, int64_t count, int element_size, int dest)
{
  if (count > 0) NCCL_CALL(dj_comm_send(comm, buf, count * element_size, dest, comm_stream));
}

void NCCLCommunicator::recv(void* buf, int64_t count, int element_size, int source)
{
  if (count > 0) NCCL_CALL(dj_comm_recv(comm, buf, count * element_size, source, comm_stream));
}

void NCCLCommunicator::stop()
{
  NCCL_CALL(dj_comm_group_end(comm));
  CUDA_RT_CALL(cudaStreamSynchronize(comm_stream));
}

void NCCLCommunicator::allgather_i64(const int64_t* mine, int n, int64_t* all)
{
  NCCL_CALL(dj_comm_allgather_i64(comm, mine, n, all, comm_stream));
}

void NCCLCommunicator::finalize()
{
  CUDA_RT_CALL(cudaStreamDestroy(comm_stream));
  NCCL_CALL(dj_comm_destroy(comm));
  comm      = nullptr;
  nccl_comm = nullptr;
}
