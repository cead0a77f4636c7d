#include "distributed_join.hpp"

#include <chrono>
#include <cmath>
#include <iostream>
#include <stdexcept>

#include "all_to_all_comm.hpp"
#include "bootstrap.hpp"
#include "error.hpp"
#include "shuffle_on.hpp"

using cudf::table;
using std::vector;

// Observable behaviour of the reference's helper (src/distributed_join.cpp:60-69), including its
// quirk of starting the divisor search at ceil(sqrt(N)): tests pin N<=NVL and
// NVL in {1,2} only.
static int get_nvl_partition_size(int mpi_size, int nvlink_domain_size)
{
  if (nvlink_domain_size >= mpi_size) return mpi_size;
  for (int size = (int)std::ceil(std::sqrt((double)mpi_size)); size > 0; size--)
    if (mpi_size % size == 0 && size <= nvlink_domain_size) return size;
  return 1;
}

static bool is_key_payload(cudf::table_view t, vector<cudf::size_type> const& on)
{
  return t.num_columns() == 2 && on.size() == 1 && on[0] == 0 && cudf::is_fixed_width(t.column(0).type()) &&
         cudf::is_fixed_width(t.column(1).type());
}

static std::unique_ptr<table> local_join_helper(cudf::table_view left, cudf::table_view right,
                                                vector<cudf::size_type> const& left_on,
                                                vector<cudf::size_type> const& right_on)
{
  // either side empty -> empty result, as the reference guards (src/distributed_join.cpp:76-82)
  if (left.num_rows() && right.num_rows()) return cudf::inner_join(left, right, left_on, right_on);
  return std::make_unique<table>();
}

// Whole hot path in one library call: hash partition -> NCCL all-to-all -> local join, batches
// overlapped on two streams, results appended into one output (dj_distributed_inner_join_i64).
static std::unique_ptr<table> fused_join(cudf::table_view left, cudf::table_view right, NCCLCommunicator* nccl,
                                         int over_decom_factor, bool report_timing)
{
  const int64_t nl = left.num_rows(), nr = right.num_rows();
  const int world  = nccl->mpi_size;
  const size_t ws_bytes = dj_distributed_inner_join_workspace_bytes(nl, nr, world, over_decom_factor);
  // The library pushes buckets into peers' workspaces through CUDA IPC (copy engines over NVLink),
  // which needs plain cudaMalloc memory: keep one grow-only workspace per process instead of
  // taking it from the stream-ordered pool (with pool memory the library falls back to NCCL).
  static struct Workspace {
    void* p = nullptr;
    size_t n = 0;
    void* data() { return p; }
    ~Workspace()
    {
      if (p) cudaFree(p);
    }
  } ws;
  // Growing is collective: peers hold CUDA IPC mappings of the old allocation, which must be closed
  // before it is freed (dj_comm_release_workspace).  Every rank evaluates the same condition.
  auto grow = [&](size_t want) {
    vector<int64_t> wants(world);
    int64_t mine = ws.n < want ? (int64_t)want : 0;
    nccl->allgather_i64(&mine, 1, wants.data());
    bool any = false;
    for (int64_t w : wants) any = any || w > 0;
    if (!any) return;
    DJ_CALL(dj_comm_release_workspace(nccl->comm));
    if (ws.n < want) {
      if (ws.p) CUDA_RT_CALL(cudaFree(ws.p));
      CUDA_RT_CALL(cudaMalloc(&ws.p, want));
      ws.n = want;
    }
  };
  grow(ws_bytes);
  // first guess for the output: as many rows as the larger input; retried collectively if short
  int64_t capacity = std::max<int64_t>(std::max(nl, nr), 1);
  for (;;) {
    vector<std::unique_ptr<cudf::column>> cols;
    for (int c = 0; c < 4; c++)
      cols.push_back(cudf::make_fixed_width_column(cudf::data_type(cudf::type_id::INT64), (cudf::size_type)capacity));
    int64_t n_out = 0;
    dj_join_options opts{};
    opts.over_decom_factor = over_decom_factor;
    opts.report_timing     = report_timing ? 1 : 0;
    int rc = dj_distributed_inner_join_i64(
      nccl->comm, left.column(0).head<int64_t>(), left.column(1).head<int64_t>(), nl,
      right.column(0).head<int64_t>(), right.column(1).head<int64_t>(), nr,
      cols[0]->mutable_view().head<int64_t>(), cols[1]->mutable_view().head<int64_t>(),
      cols[2]->mutable_view().head<int64_t>(), cols[3]->mutable_view().head<int64_t>(), capacity, &n_out, &opts,
      ws.data(), ws.n, nullptr);
    if (rc == DJ_ERR_WORKSPACE && world > 1) {
      // skewed sizes: a rank receives more than the balanced estimate.  Every rank returned this
      // code together; each grows to what it was told it needs (plus headroom) and all retry.
      grow(opts.workspace_needed > 0 ? (size_t)(opts.workspace_needed + opts.workspace_needed / 8) : ws.n);
      continue;
    }
    if (rc == DJ_OK) {
      for (auto& c : cols) c->set_size((cudf::size_type)n_out);
      return std::make_unique<table>(std::move(cols));
    }
    if (rc != DJ_ERR_OVERFLOW) DJ_CALL(rc);
    // every rank got DJ_ERR_OVERFLOW together: agree on the largest need and try again
    vector<int64_t> needs(world);
    nccl->allgather_i64(&n_out, 1, needs.data());
    for (int64_t n : needs) capacity = std::max(capacity, n);
    if (capacity > INT32_MAX) throw std::runtime_error("join result exceeds cudf::size_type rows per rank");
  }
}

std::unique_ptr<table> distributed_inner_join(cudf::table_view left, cudf::table_view right,
                                              vector<cudf::size_type> const& left_on,
                                              vector<cudf::size_type> const& right_on, Communicator* communicator,
                                              vector<ColumnCompressionOptions> left_compression_options,
                                              vector<ColumnCompressionOptions> right_compression_options,
                                              int over_decom_factor, bool report_timing,
                                              void* preallocated_pinned_buffer, int nvlink_domain_size)
{
  if (over_decom_factor < 1) throw std::runtime_error("over_decom_factor must be >= 1");
  const int mpi_rank = communicator->mpi_rank, mpi_size = communicator->mpi_size;
  const int group = get_nvl_partition_size(mpi_size, nvlink_domain_size);
  auto clock      = [] { return std::chrono::high_resolution_clock::now(); };
  auto ms_since   = [&](std::chrono::high_resolution_clock::time_point t0) {
    return std::chrono::duration_cast<std::chrono::milliseconds>(clock() - t0).count();
  };

  // ---- fast path: one NVLink domain covering all ranks, int64 key + int64 payload
  auto* nccl = dynamic_cast<NCCLCommunicator*>(communicator);
  if (nccl && group == mpi_size && is_key_payload(left, left_on) && is_key_payload(right, right_on)) {
    if (cudf::all_i64(left) && cudf::all_i64(right)) {
      CUDA_RT_CALL(cudaStreamSynchronize(nullptr));
      return fused_join(left, right, nccl, over_decom_factor, report_timing);
    }
    // 4-byte columns: widen into temporary INT64 tables, join, narrow back to the callers' types
    std::vector<cudf::data_type> types;
    for (auto const& c : left) types.push_back(c.type());
    for (auto const& c : right) types.push_back(c.type());
    auto wl = cudf::widen_to_i64(left), wr = cudf::widen_to_i64(right);
    CUDA_RT_CALL(cudaStreamSynchronize(nullptr));
    auto joined = fused_join(wl->view(), wr->view(), nccl, over_decom_factor, report_timing);
    return cudf::narrow_like(joined->view(), types);
  }

  // ---- general path, stage by stage like src/distributed_join.cpp:152-339
  std::unique_ptr<table> shuffled_left_ib, shuffled_right_ib;
  if (group != mpi_size) {
    constexpr uint32_t hash_partition_seed_ib = 87654321;  // src/distributed_join.cpp:160
    shuffled_left_ib  = shuffle_on(left, left_on, CommunicationGroup(mpi_size, group), communicator,
                                   left_compression_options, cudf::hash_id::HASH_MURMUR3, hash_partition_seed_ib,
                                   report_timing, preallocated_pinned_buffer);
    shuffled_right_ib = shuffle_on(right, right_on, CommunicationGroup(mpi_size, group), communicator,
                                   right_compression_options, cudf::hash_id::HASH_MURMUR3, hash_partition_seed_ib,
                                   report_timing, preallocated_pinned_buffer);
    left  = shuffled_left_ib->view();
    right = shuffled_right_ib->view();
  }
  if (group == 1) {
    auto t0     = clock();
    auto result = local_join_helper(left, right, left_on, right_on);
    CUDA_RT_CALL(cudaStreamSynchronize(nullptr));
    if (report_timing)  // the reference prints this label for the local join (src/distributed_join.cpp:194)
      std::cout << "Rank " << mpi_rank << ": Hash partition takes " << ms_since(t0) << "ms" << std::endl;
    return result;
  }

  constexpr uint32_t hash_partition_seed = 12345678;  // src/distributed_join.cpp:211
  auto t0      = clock();
  auto hashed_l = cudf::hash_partition(left, left_on, group * over_decom_factor, cudf::hash_id::HASH_MURMUR3,
                                       hash_partition_seed);
  auto hashed_r = cudf::hash_partition(right, right_on, group * over_decom_factor, cudf::hash_id::HASH_MURMUR3,
                                       hash_partition_seed);
  CUDA_RT_CALL(cudaStreamSynchronize(nullptr));
  shuffled_left_ib.reset();
  shuffled_right_ib.reset();
  hashed_l.second.push_back(hashed_l.first->num_rows());
  hashed_r.second.push_back(hashed_r.first->num_rows());
  if (report_timing)
    std::cout << "Rank " << mpi_rank << ": Hash partition takes " << ms_since(t0) << "ms" << std::endl;

  // Batch b exchanges buckets [b*G, (b+1)*G]; the join of a batch is enqueued right after its
  // exchange returns.  Kernels are asynchronous, so the next batch's exchange (on the
  // communicator's stream) overlaps this join without the reference's spinning helper thread.
  vector<std::unique_ptr<table>> batch_results;
  vector<std::unique_ptr<table>> keep_left, keep_right;
  for (int b = 0; b < over_decom_factor; b++) {
    auto tb = clock();
    auto slice = [&](vector<cudf::size_type> const& off) {
      return vector<cudf::size_type>(off.begin() + b * group, off.begin() + (b + 1) * group + 1);
    };
    AllToAllCommunicator ex_l(hashed_l.first->view(), slice(hashed_l.second), CommunicationGroup(group, 1),
                              communicator, generate_none_compression_options(hashed_l.first->view()), true);
    AllToAllCommunicator ex_r(hashed_r.first->view(), slice(hashed_r.second), CommunicationGroup(group, 1),
                              communicator, generate_none_compression_options(hashed_r.first->view()), true);
    keep_left.push_back(ex_l.allocate_communicated_table());
    keep_right.push_back(ex_r.allocate_communicated_table());
    ex_l.launch_communication(keep_left.back()->mutable_view(), report_timing, preallocated_pinned_buffer);
    ex_r.launch_communication(keep_right.back()->mutable_view(), report_timing, preallocated_pinned_buffer);
    if (report_timing)
      std::cout << "Rank " << mpi_rank << ": All-to-all communication on batch " << b << " takes " << ms_since(tb)
                << "ms" << std::endl;
    auto tj = clock();
    batch_results.push_back(local_join_helper(keep_left.back()->view(), keep_right.back()->view(), left_on, right_on));
    if (report_timing) {
      CUDA_RT_CALL(cudaStreamSynchronize(nullptr));
      std::cout << "Rank " << mpi_rank << ": Local join on batch " << b << " takes " << ms_since(tj) << "ms"
                << std::endl;
    }
  }
  CUDA_RT_CALL(cudaStreamSynchronize(nullptr));
  if (batch_results.size() == 1) return std::move(batch_results[0]);
  vector<cudf::table_view> views;
  for (auto& t : batch_results) views.push_back(t->view());
  return cudf::concatenate(views);
}
