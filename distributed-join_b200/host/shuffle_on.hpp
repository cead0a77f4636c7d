// shuffle_on.hpp -- reference interface src/shuffle_on.hpp:44-64: hash-partition `input` on
// `on_columns` into one bucket per group member and exchange the buckets.
#pragma once

#include <memory>
#include <vector>

#include "all_to_all_comm.hpp"
#include "communicator.hpp"
#include "compression.hpp"
#include "cudf_shim.hpp"

std::unique_ptr<cudf::table> shuffle_on(cudf::table_view const& input,
                                        std::vector<cudf::size_type> const& on_columns,
                                        CommunicationGroup comm_group, Communicator* communicator,
                                        std::vector<ColumnCompressionOptions> compression_options,
                                        cudf::hash_id hash_function = cudf::hash_id::HASH_MURMUR3,
                                        uint32_t hash_seed          = cudf::DEFAULT_HASH_SEED,
                                        bool report_timing = false, void* preallocated_pinned_buffer = nullptr);

std::unique_ptr<cudf::table> shuffle_on(cudf::table_view const& input,
                                        std::vector<cudf::size_type> const& on_columns,
                                        Communicator* communicator,
                                        std::vector<ColumnCompressionOptions> compression_options,
                                        cudf::hash_id hash_function = cudf::hash_id::HASH_MURMUR3,
                                        uint32_t hash_seed          = cudf::DEFAULT_HASH_SEED,
                                        bool report_timing = false, void* preallocated_pinned_buffer = nullptr);
