#include "shuffle_on.hpp"

#include "error.hpp"
#include "stage_timer.hpp"

std::unique_ptr<cudf::table> shuffle_on(
  cudf::table_view const& input, std::vector<cudf::size_type> const& on_columns, CommunicationGroup comm_group,
  Communicator* communicator, std::vector<ColumnCompressionOptions> compression_options,
  cudf::hash_id hash_function, uint32_t hash_seed, bool report_timing, void* preallocated_pinned_buffer)
{
  StageTimer timer(report_timing, communicator->mpi_rank);

  // stage 1: libdj_b200's radix hash-partition kernel, one bucket per group member.  cuDF hands
  // back the bucket start offsets only; the exchange wants the closing offset as well.
  auto partitioned = cudf::hash_partition(input, on_columns, comm_group.size(), hash_function, hash_seed);
  CUDA_RT_CALL(cudaStreamSynchronize(nullptr));
  std::unique_ptr<cudf::table>& buckets  = partitioned.first;
  std::vector<cudf::size_type>& offsets = partitioned.second;
  offsets.push_back(buckets->num_rows());
  timer.lap("Hash partition in shuffle");

  // stage 2: own bucket by device copy, every other bucket over NVLink, straight into the result
  AllToAllCommunicator exchange(buckets->view(), offsets, comm_group, communicator,
                                std::move(compression_options), /*explicit_copy_to_current_rank=*/true);
  std::unique_ptr<cudf::table> result = exchange.allocate_communicated_table();
  exchange.launch_communication(result->mutable_view(), report_timing, preallocated_pinned_buffer);
  timer.lap("All-to-all communication in shuffle");
  return result;
}
