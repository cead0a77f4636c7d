// distribute_table.cpp -- see distribute_table.hpp.  The reference moves the metadata with
// MPI_Bcast / MPI_Gather (src/distribute_table.cpp:115-140,186-198); here it rides on the
// communicator's all-gather, and the columns travel by Communicator::send / recv exactly as in the
// reference (one start()/stop() batch per column).
#include "distribute_table.hpp"

#include <algorithm>
#include <vector>

#include "error.hpp"

namespace {

int64_t slice_rows(int64_t global_rows, int rank, int world)
{
  return global_rows / world + (rank < global_rows % world ? 1 : 0);
}

int64_t slice_begin(int64_t global_rows, int rank, int world)
{
  return std::min<int64_t>(rank, global_rows % world) + (global_rows / world) * rank;
}

constexpr int kMaxCols = 14;  // metadata message: rows, ncols, type ids

}  // namespace

std::unique_ptr<cudf::table> distribute_table(cudf::table_view global_table, Communicator* communicator)
{
  const int rank = communicator->mpi_rank, world = communicator->mpi_size;
  // root's (rows, ncols, dtypes) to everybody
  std::vector<int64_t> mine(2 + kMaxCols, 0), all((size_t)world * (2 + kMaxCols));
  if (rank == 0) {
    if (global_table.num_columns() > kMaxCols) throw std::runtime_error("distribute_table: too many columns");
    mine[0] = global_table.num_rows();
    mine[1] = global_table.num_columns();
    for (int c = 0; c < global_table.num_columns(); c++) mine[2 + c] = (int64_t)global_table.column(c).type().id();
  }
  communicator->allgather_i64(mine.data(), 2 + kMaxCols, all.data());
  const int64_t global_rows = all[0];
  const int ncols           = (int)all[1];
  const int64_t local_rows  = slice_rows(global_rows, rank, world);

  std::vector<std::unique_ptr<cudf::column>> local;
  for (int c = 0; c < ncols; c++)
    local.push_back(cudf::make_fixed_width_column(cudf::data_type((cudf::type_id)all[2 + c]), (cudf::size_type)local_rows));
  CUDA_RT_CALL(cudaStreamSynchronize(nullptr));

  for (int c = 0; c < ncols; c++) {
    const int es = (int)cudf::size_of(local[c]->type());
    communicator->start();
    if (rank == 0) {
      const char* src = global_table.column(c).head<char>();
      for (int r = 1; r < world; r++)
        communicator->send(src + slice_begin(global_rows, r, world) * es, slice_rows(global_rows, r, world), es, r);
      CUDA_RT_CALL(cudaMemcpy(local[c]->mutable_view().head<char>(), src, (size_t)local_rows * es,
                              cudaMemcpyDeviceToDevice));
    } else {
      communicator->recv(local[c]->mutable_view().head<char>(), local_rows, es, 0);
    }
    communicator->stop();
  }
  return std::make_unique<cudf::table>(std::move(local));
}

std::unique_ptr<cudf::table> collect_tables(cudf::table_view table, Communicator* communicator)
{
  const int rank = communicator->mpi_rank, world = communicator->mpi_size;
  const int ncols = table.num_columns();
  int64_t nrows   = table.num_rows();
  std::vector<int64_t> rows(world);
  communicator->allgather_i64(&nrows, 1, rows.data());
  std::vector<int64_t> scan(world + 1, 0);
  for (int r = 0; r < world; r++) scan[r + 1] = scan[r] + rows[r];
  if (scan[world] > INT32_MAX) throw std::runtime_error("collect_tables: result exceeds cudf::size_type rows");

  std::vector<std::unique_ptr<cudf::column>> merged;
  if (rank == 0) {
    for (int c = 0; c < ncols; c++)
      merged.push_back(cudf::make_fixed_width_column(table.column(c).type(), (cudf::size_type)scan[world]));
    CUDA_RT_CALL(cudaStreamSynchronize(nullptr));
  }
  for (int c = 0; c < ncols; c++) {
    const int es = (int)cudf::size_of(table.column(c).type());
    communicator->start();
    if (rank == 0) {
      char* dst = merged[c]->mutable_view().head<char>();
      for (int r = 1; r < world; r++) communicator->recv(dst + scan[r] * es, rows[r], es, r);
      CUDA_RT_CALL(cudaMemcpy(dst, table.column(c).head<char>(), (size_t)rows[0] * es, cudaMemcpyDeviceToDevice));
    } else {
      communicator->send(table.column(c).head<char>(), nrows, es, 0);
    }
    communicator->stop();
  }
  if (rank != 0) return std::unique_ptr<cudf::table>(nullptr);
  return std::make_unique<cudf::table>(std::move(merged));
}
