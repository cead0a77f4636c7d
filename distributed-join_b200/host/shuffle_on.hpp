// shuffle_on.hpp -- the reference's shuffle operator (interface: src/shuffle_on.hpp:44-64).
//
// shuffle_on() hash-partitions `input` on `on_columns` into one bucket per member of the
// communication group and exchanges the buckets, so that afterwards all rows with equal keys
// live on the same rank.  Collective over the group.  Row order of the result is unspecified.
#pragma once

#include <memory>
#include <utility>
#include <vector>

#include "all_to_all_comm.hpp"
#include "communicator.hpp"
#include "compression.hpp"
#include "cudf_shim.hpp"

// Within `comm_group` (reference overload #1).
std::unique_ptr<cudf::table> shuffle_on(
  cudf::table_view const& input, std::vector<cudf::size_type> const& on_columns, CommunicationGroup comm_group,
  Communicator* communicator, std::vector<ColumnCompressionOptions> compression_options,
  cudf::hash_id hash_function = cudf::hash_id::HASH_MURMUR3, uint32_t hash_seed = cudf::DEFAULT_HASH_SEED,
  bool report_timing = false, void* preallocated_pinned_buffer = nullptr);

// Across all ranks with stride 1 (reference overload #2): a thin forwarder.
inline std::unique_ptr<cudf::table> shuffle_on(
  cudf::table_view const& input, std::vector<cudf::size_type> const& on_columns, Communicator* communicator,
  std::vector<ColumnCompressionOptions> compression_options,
  cudf::hash_id hash_function = cudf::hash_id::HASH_MURMUR3, uint32_t hash_seed = cudf::DEFAULT_HASH_SEED,
  bool report_timing = false, void* preallocated_pinned_buffer = nullptr)
{
  CommunicationGroup everyone(communicator->mpi_size, 1);
  return shuffle_on(input, on_columns, everyone, communicator, std::move(compression_options), hash_function,
                    hash_seed, report_timing, preallocated_pinned_buffer);
}
