// test_shuffle_on.cu -- the reference's shuffle invariant (test/test_shuffle_on.cpp:78-83): after
// shuffle_on with HASH_IDENTITY every key on a rank has the same residue modulo the number of
// ranks (co-location; not "residue == rank").
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "../host/bootstrap.hpp"
#include "../host/compression.hpp"
#include "../host/error.hpp"
#include "../host/setup.hpp"
#include "../host/shuffle_on.hpp"

int main(int argc, char* argv[])
{
  dj_bootstrap::init(&argc, &argv);
  set_cuda_device();
  Communicator* communicator{nullptr};
  registered_memory_resource* registered_mr{nullptr};
  rmm::mr::pool_memory_resource<rmm::mr::device_memory_resource>* pool_mr{nullptr};
  setup_memory_pool_and_communicator(communicator, registered_mr, pool_mr, "NCCL", "none", 0);
  const int rank = communicator->mpi_rank, world = communicator->mpi_size;
  const int64_t size = 1'000'000;

  std::vector<int64_t> h_keys(size), h_pay(size);
  srand(1234 + rank);
  for (int64_t i = 0; i < size; i++) {
    h_keys[i] = rand() % (10 * size);
    h_pay[i]  = i;
  }
  std::vector<std::unique_ptr<cudf::column>> cols;
  for (int c = 0; c < 2; c++) cols.push_back(cudf::make_numeric_column(cudf::data_type(cudf::type_id::INT64), (cudf::size_type)size));
  CUDA_RT_CALL(cudaMemcpy(cols[0]->mutable_view().head(), h_keys.data(), size * 8, cudaMemcpyHostToDevice));
  CUDA_RT_CALL(cudaMemcpy(cols[1]->mutable_view().head(), h_pay.data(), size * 8, cudaMemcpyHostToDevice));
  cudf::table input(std::move(cols));

  auto shuffled = shuffle_on(input.view(), {0}, communicator, generate_none_compression_options(input.view()),
                             cudf::hash_id::HASH_IDENTITY);
  std::vector<int64_t> got(shuffled->num_rows());
  CUDA_RT_CALL(cudaMemcpy(got.data(), shuffled->view().column(0).head(), got.size() * 8, cudaMemcpyDeviceToHost));
  int64_t bad = 0;
  for (size_t i = 1; i < got.size(); i++)
    if (got[i] % world != got[0] % world) bad++;
  const int64_t total_bad  = dj_bootstrap::allreduce_sum(bad);
  const int64_t total_rows = dj_bootstrap::allreduce_sum(shuffled->num_rows());
  destroy_memory_pool_and_communicator(communicator, registered_mr, pool_mr, "NCCL", "none");
  dj_bootstrap::finalize();
  if (total_bad != 0 || total_rows != size * world) {
    if (rank == 0) std::cerr << "test_shuffle_on FAILED: " << total_bad << " misplaced keys, " << total_rows << " rows" << std::endl;
    return 1;
  }
  if (rank == 0) std::cerr << "Test case \"test_shuffle_on\" passes successfully." << std::endl;
  return 0;
}
