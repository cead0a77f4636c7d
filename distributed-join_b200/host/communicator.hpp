// communicator.hpp -- point-to-point transport abstraction with the reference's interface
// (src/communicator.hpp:31-90) and its NCCL backend (:326-353), implemented on libdj_b200.
// UCX backends are out of scope (north star: "no UCX, no MPI on the data path").
#pragma once

#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdint>

#include "../../include/dj_b200.h"

class Communicator {
  // Like the reference (src/communicator.hpp:32): no thread-safety guarantee.
 public:
  virtual void initialize() = 0;  // collective over all ranks, at most once per process
  virtual void start()      = 0;  // open a batch of sends/receives (no nesting)
  virtual void stop()       = 0;  // block until every transfer since start() has landed
  virtual void send(const void* buf, int64_t count, int element_size, int dest) = 0;
  virtual void recv(void* buf, int64_t count, int element_size, int source)     = 0;
  virtual void finalize()       = 0;
  virtual bool group_by_batch() = 0;
  virtual ~Communicator()       = default;

  // size exchange without MPI (replaces MPI_Isend/Irecv in communicate_sizes,
  // src/all_to_all_comm.cpp:54-100): all-gather `n` int64 per rank.  Not part of the reference's
  // interface, so it is NOT pure: the default is built from start/send/recv/stop and works for
  // any backend a user derives; NCCLCommunicator overrides it with one collective.
  virtual void allgather_i64(const int64_t* mine, int n, int64_t* all);

  int mpi_rank;
  int mpi_size;
  int current_device;
};

class NCCLCommunicator : public Communicator {
 public:
  void initialize() override;
  void start() override;
  void stop() override;
  void send(const void* buf, int64_t count, int element_size, int dest) override;
  void recv(void* buf, int64_t count, int element_size, int source) override;
  void finalize() override;
  // NCCL >= 2.8 takes many messages per peer in one group (the image has 2.27/2.28), so the
  // whole batch -- every column of a table -- rides in one ncclGroup.  The reference returns
  // false here only because of NCCL 2.7 (src/communicator.hpp:340-342).
  bool group_by_batch() override { return true; }
  void allgather_i64(const int64_t* mine, int n, int64_t* all) override;

  ncclComm_t nccl_comm     = nullptr;  // as in the reference (src/communicator.hpp:346): the raw handle
  cudaStream_t comm_stream = nullptr;  // as in the reference: the stream transfers run on
  dj_comm_t* comm          = nullptr;  // libdj_b200's communicator object (owns nccl_comm)
};
