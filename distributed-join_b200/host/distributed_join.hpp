// distributed_join.hpp -- reference interface src/distributed_join.hpp:65-76.
//
// Collective inner join of two tables whose rows are spread arbitrarily over the ranks.  The
// result of every rank holds left columns followed by right columns; the global answer is the
// concatenation over ranks, in no particular row order.
#pragma once

#include <cstdint>
#include <memory>
#include <vector>

#include "communicator.hpp"
#include "compression.hpp"
#include "cudf_shim.hpp"

std::unique_ptr<cudf::table> distributed_inner_join(
  cudf::table_view left, cudf::table_view right, std::vector<cudf::size_type> const& left_on,
  std::vector<cudf::size_type> const& right_on, Communicator* communicator,
  std::vector<ColumnCompressionOptions> left_compression_options,
  std::vector<ColumnCompressionOptions> right_compression_options, int over_decom_factor = 1,
  bool report_timing = false, void* preallocated_pinned_buffer = nullptr, int nvlink_domain_size = 1);

// the north star's spelling of the same entry point
inline std::unique_ptr<cudf::table> distributed_join(
  cudf::table_view left, cudf::table_view right, std::vector<cudf::size_type> const& left_on,
  std::vector<cudf::size_type> const& right_on, Communicator* communicator,
  std::vector<ColumnCompressionOptions> left_compression_options,
  std::vector<ColumnCompressionOptions> right_compression_options, int over_decom_factor = 1,
  bool report_timing = false, void* preallocated_pinned_buffer = nullptr, int nvlink_domain_size = 1)
{
  return distributed_inner_join(left, right, left_on, right_on, communicator, std::move(left_compression_options),
                                std::move(right_compression_options), over_decom_factor, report_timing,
                                preallocated_pinned_buffer, nvlink_domain_size);
}
