#include "shuffle_on.hpp"

#include <iostream>

#include "bootstrap.hpp"
#include "error.hpp"

std::unique_ptr<cudf::table> shuffle_on(cudf::table_view const& input,
                                        std::vector<cudf::size_type> const& on_columns,
                                        CommunicationGroup comm_group, Communicator* communicator,
                                        std::vector<ColumnCompressionOptions> compression_options,
                                        cudf::hash_id hash_function, uint32_t hash_seed, bool report_timing,
                                        void* preallocated_pinned_buffer)
{
  const int rank = communicator->mpi_rank;
  double t0      = report_timing ? dj_bootstrap::wtime() : 0.0;

  // stage 1: one bucket per group member (libdj_b200 radix hash-partition kernel)
  auto partitioned = cudf::hash_partition(input, on_columns, comm_group.size(), hash_function, hash_seed);
  CUDA_RT_CALL(cudaStreamSynchronize(nullptr));
  std::vector<cudf::size_type>& offsets = partitioned.second;
  offsets.push_back(partitioned.first->num_rows());
  if (report_timing) {
    std::cout << "Rank " << rank << ": Hash partition in shuffle takes " << (dj_bootstrap::wtime() - t0) * 1e3
              << "ms" << std::endl;
    t0 = dj_bootstrap::wtime();
  }

  // stage 2: exchange (own bucket by device copy, the rest over NVLink)
  AllToAllCommunicator exchange(partitioned.first->view(), offsets, comm_group, communicator,
                                std::move(compression_options), true);
  std::unique_ptr<cudf::table> shuffled = exchange.allocate_communicated_table();
  exchange.launch_communication(shuffled->mutable_view(), report_timing, preallocated_pinned_buffer);
  if (report_timing)
    std::cout << "Rank " << rank << ": All-to-all communication in shuffle takes "
              << (dj_bootstrap::wtime() - t0) * 1e3 << "ms" << std::endl;
  return shuffled;
}

std::unique_ptr<cudf::table> shuffle_on(cudf::table_view const& input,
                                        std::vector<cudf::size_type> const& on_columns,
                                        Communicator* communicator,
                                        std::vector<ColumnCompressionOptions> compression_options,
                                        cudf::hash_id hash_function, uint32_t hash_seed, bool report_timing,
                                        void* preallocated_pinned_buffer)
{
  return shuffle_on(input, on_columns, CommunicationGroup(communicator->mpi_size, 1), communicator,
                    std::move(compression_options), hash_function, hash_seed, report_timing,
                    preallocated_pinned_buffer);
}
