// all_to_all.cpp -- BASELINE config 3: raw communicator bandwidth sweep in the shape of the
// reference's benchmark/all_to_all.cpp (sizes 1 MB ... 4.096 GB per rank, plus the 8.192 GB the
// baseline asks for; REPEAT exchanges per size; self excluded; "Bandwidth per GPU (GB/s)" =
// size/N*(N-1)*REPEAT/t, unidirectional), on Communicator::{start,send,recv,stop} over NVLink.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include <cuda_profiler_api.h>

#include "../host/bootstrap.hpp"
#include "../host/communicator.hpp"
#include "../host/error.hpp"
#include "../host/setup.hpp"

static int REPEAT                    = 4;
static std::string COMMUNICATOR_NAME = "NCCL";
static int64_t MAX_SIZE              = 8'192'000'000LL;

static double run_all_to_all(int64_t size, Communicator* communicator, bool print)
{
  const int n = communicator->mpi_size, me = communicator->mpi_rank;
  const int64_t per_peer = size / n;
  std::vector<void*> send(n, nullptr), recv(n, nullptr);
  for (int r = 0; r < n; r++) {
    if (r == me) continue;
    CUDA_RT_CALL(cudaMalloc(&send[r], per_peer));
    CUDA_RT_CALL(cudaMalloc(&recv[r], per_peer));
  }
  CUDA_RT_CALL(cudaDeviceSynchronize());
  dj_bootstrap::barrier();
  cudaProfilerStart();
  const double t0 = dj_bootstrap::wtime();
  for (int it = 0; it < REPEAT; it++) {
    communicator->start();
    for (int r = 0; r < n; r++)
      if (r != me) communicator->send(send[r], per_peer, 1, r);
    for (int r = 0; r < n; r++)
      if (r != me) communicator->recv(recv[r], per_peer, 1, r);
    communicator->stop();
  }
  const double local = dj_bootstrap::wtime() - t0;
  cudaProfilerStop();
  const double t = dj_bootstrap::allreduce_max(local);  // slowest rank
  const double gbs = (double)per_peer * (n - 1) * REPEAT / t / 1e9;
  if (print && me == 0)
    std::cout << "Size (MB): " << size / 1e6 << ", Elasped time (s): " << t << ", Bandwidth per GPU (GB/s): " << gbs
              << std::endl;
  for (int r = 0; r < n; r++) {  // every buffer is released (the reference leaks some: SURVEY App. A)
    if (send[r]) CUDA_RT_CALL(cudaFree(send[r]));
    if (recv[r]) CUDA_RT_CALL(cudaFree(recv[r]));
  }
  return gbs;
}

int main(int argc, char* argv[])
{
  dj_bootstrap::init(&argc, &argv);
  set_cuda_device();
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--repeat") && i + 1 < argc) REPEAT = atoi(argv[i + 1]);
    if (!strcmp(argv[i], "--communicator") && i + 1 < argc) COMMUNICATOR_NAME = argv[i + 1];
    if (!strcmp(argv[i], "--max-size") && i + 1 < argc) MAX_SIZE = atoll(argv[i + 1]);
  }
  Communicator* communicator{nullptr};
  registered_memory_resource* registered_mr{nullptr};
  rmm::mr::pool_memory_resource<rmm::mr::device_memory_resource>* pool_mr{nullptr};
  setup_memory_pool_and_communicator(communicator, registered_mr, pool_mr, COMMUNICATOR_NAME, "none", 0);
  if (communicator->mpi_size < 2) {
    if (communicator->mpi_rank == 0) std::cout << "all_to_all needs at least 2 ranks" << std::endl;
  } else {
    run_all_to_all(4'000'000LL, communicator, false);  // warm-up, as in the reference (:184)
    std::vector<double> results;
    for (int64_t size = 1'000'000LL; size <= MAX_SIZE; size *= 2) results.push_back(run_all_to_all(size, communicator, true));
    if (communicator->mpi_rank == 0) {
      std::cout << "{\"benchmark\": \"all_to_all\", \"n_gpus\": " << communicator->mpi_size << ", \"repeat\": " << REPEAT
                << ", \"gbps_per_gpu_by_size\": [";
      for (size_t i = 0; i < results.size(); i++) std::cout << (i ? ", " : "") << results[i];
      std::cout << "]}" << std::endl;
    }
  }
  destroy_memory_pool_and_communicator(communicator, registered_mr, pool_mr, COMMUNICATOR_NAME, "none");
  dj_bootstrap::finalize();
  return 0;
}
