// all_to_all_comm.hpp -- table-level all-to-all with the reference's interface
// (src/all_to_all_comm.hpp:72-362) for fixed-width columns, over Communicator + libdj_b200.
// String columns and compression are out of scope; no MPI: sizes travel through
// Communicator::allgather_i64, the rank comes from dj_bootstrap instead of MPI_COMM_WORLD.
#pragma once

#include <cassert>
#include <cstdint>
#include <memory>
#include <vector>

#include "bootstrap.hpp"
#include "communicator.hpp"
#include "compression.hpp"
#include "cudf_shim.hpp"

// A set of ranks that exchange data with each other: ranks are cut into grids of `grid_size`
// consecutive ranks and, inside a grid, sampled with spacing `stride`
// (e.g. 16 ranks, grid 8, stride 2 -> {0,2,4,6} {1,3,5,7} {8,10,12,14} {9,11,13,15}).
class CommunicationGroup {
 public:
  CommunicationGroup(int grid_size, int stride = 1) : grid_size(grid_size), stride(stride)
  {
    assert(grid_size % stride == 0 && "group size must be a multiple of stride");
    mpi_rank    = dj_bootstrap::rank();
    group_start = mpi_rank / grid_size * grid_size + mpi_rank % stride;
  }
  int size() const { return grid_size / stride; }
  int get_global_rank(int local_idx) const { return group_start + local_idx * stride; }
  int get_local_idx() const { return (mpi_rank - group_start) / stride; }

 private:
  int mpi_rank;
  int group_start;
  int grid_size;
  int stride;
};

// send_offset[i+1]-send_offset[i] rows go to group member i; recv_offset (resized here) is the
// exclusive prefix sum of what every member sends to this rank.  Collective over the group.
void communicate_sizes(std::vector<int64_t> const& send_offset, std::vector<int64_t>& recv_offset,
                       CommunicationGroup comm_group, Communicator* communicator);
void communicate_sizes(std::vector<cudf::size_type> const& send_offset, std::vector<int64_t>& recv_offset,
                       CommunicationGroup comm_group, Communicator* communicator);

void warmup_all_to_all(Communicator* communicator);

// One column's worth of exchange: element offsets per group member on both sides.
struct AllToAllCommBuffer {
  const void* send_buffer;
  void* recv_buffer;
  std::vector<int64_t> send_offsets;
  std::vector<int64_t> recv_offsets;
  cudf::data_type dtype;
  CompressionMethod compression_method;

  AllToAllCommBuffer(const void* send_buffer, void* recv_buffer, std::vector<int64_t> send_offsets,
                     std::vector<int64_t> recv_offsets, cudf::data_type dtype,
                     CompressionMethod compression_method = CompressionMethod::none)
    : send_buffer(send_buffer), recv_buffer(recv_buffer), send_offsets(std::move(send_offsets)),
      recv_offsets(std::move(recv_offsets)), dtype(dtype), compression_method(compression_method)
  {
  }
};

// Plans only: one AllToAllCommBuffer per (fixed-width) column of `input` / preallocated `output`.
void append_to_all_to_all_comm_buffers(cudf::table_view input, cudf::mutable_table_view output,
                                       std::vector<cudf::size_type> const& send_offsets,
                                       std::vector<int64_t> const& recv_offsets,
                                       std::vector<AllToAllCommBuffer>& all_to_all_comm_buffers,
                                       std::vector<ColumnCompressionOptions> compression_options);

// Executes the plans.  With a group_by_batch() communicator the call must sit between
// communicator->start() and communicator->stop() (as in the reference); otherwise it opens one
// start/stop pair per column itself.
void all_to_all_comm(std::vector<AllToAllCommBuffer>& all_to_all_comm_buffers, CommunicationGroup comm_group,
                     Communicator* communicator, bool include_current_rank = true, bool report_timing = false,
                     void* preallocated_pinned_buffer = nullptr);

// Decompression hook of the reference; nothing to do without compression.
void postprocess_all_to_all_comm(std::vector<AllToAllCommBuffer>& all_to_all_comm_buffers,
                                 CommunicationGroup comm_group, Communicator* communicator,
                                 bool include_current_rank = true, bool report_timing = false);

class AllToAllCommunicator {
 public:
  // Collective over comm_group.  `offsets` (group size + 1 entries) are row ranges of input_table
  // per destination.  With explicit_copy_to_current_rank the self partition is copied
  // device-to-device when the receive table is allocated instead of going through the wire.
  AllToAllCommunicator(cudf::table_view input_table, std::vector<cudf::size_type> offsets,
                       CommunicationGroup comm_group, Communicator* communicator,
                       std::vector<ColumnCompressionOptions> compression_options,
                       bool explicit_copy_to_current_rank = false);
  AllToAllCommunicator(cudf::table_view input_table, std::vector<cudf::size_type> offsets,
                       Communicator* communicator, std::vector<ColumnCompressionOptions> compression_options,
                       bool explicit_copy_to_current_rank = false);

  AllToAllCommunicator(const AllToAllCommunicator&) = delete;
  AllToAllCommunicator& operator=(const AllToAllCommunicator&) = delete;
  AllToAllCommunicator(AllToAllCommunicator&&)                 = default;

  std::unique_ptr<cudf::table> allocate_communicated_table();  // synchronous w.r.t. stream 0
  // Collective; blocks the host until the data has landed.
  void launch_communication(cudf::mutable_table_view communicated_table, bool report_timing = false,
                            void* preallocated_pinned_buffer = nullptr);

 private:
  cudf::table_view input_table;
  CommunicationGroup comm_group;
  Communicator* communicator;
  bool explicit_copy_to_current_rank;
  std::vector<cudf::size_type> send_offsets;
  std::vector<int64_t> recv_offsets;
  std::vector<ColumnCompressionOptions> compression_options;
};
