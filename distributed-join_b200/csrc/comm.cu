// comm.cu -- NCCL plumbing and the distributed inner join driver (include/dj_b200.h).
//
// Replaces NCCLCommunicator (src/communicator.cpp:799-875), communicate_sizes
// (src/all_to_all_comm.cpp:54-111), the table all-to-all (:126-189,307-356) and the
// orchestration of distributed_inner_join (src/distributed_join.cpp:134-340):
//   * no MPI: counts travel by ncclAllGather, the unique id comes from the launcher;
//   * no staging copies: buckets are sent from / received into their final buffers;
//   * one ncclGroup per batch covering every column of both tables;
//   * the reference's spinning join thread + std::atomic flags (:100-132,283-322) become two
//     CUDA streams and events: batch b+1's exchange overlaps batch b's local join;
//   * batch results are appended into one output (no cudf::concatenate, :333-339).
#include <cuda.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dj_device.cuh"
#include "dj_internal.h"

struct dj_comm {
  ncclComm_t nccl = nullptr;
  int rank = 0, size = 1, device = 0;
  ncclComm_t nccl_ctrl     = nullptr;  // duplicate communicator for counts / verdicts, so that tiny
                                       // control collectives never queue behind the bulk exchange
  cudaStream_t comm_stream = nullptr;
  cudaStream_t ctrl_stream = nullptr;
  cudaEvent_t ev_ready     = nullptr;
  cudaEvent_t ev_part[2]   = {nullptr, nullptr};
  cudaEvent_t ev_seg[2]    = {nullptr, nullptr};
  std::vector<cudaEvent_t> ev_batch;
  // Peer-memory exchange (copy engines over NVLink, no SMs): every rank maps every peer's flag
  // block and -- per call -- workspace through CUDA IPC, pushes its buckets with cudaMemcpyAsync
  // and signals with a stream write; receivers wait with a stream wait-value.
  bool peer_ok = false;
  uint32_t* d_flags = nullptr;               // [size][kFlagSlots], written by peers
  std::vector<uint32_t*> peer_flags;         // peers' d_flags mapped here
  std::vector<cudaStream_t> peer_stream;     // one push stream per peer
  struct IpcEntry { cudaIpcMemHandle_t h; char* base; };
  std::vector<std::vector<IpcEntry>> ipc_cache;  // per peer: opened workspace allocations
  uint32_t seq = 0;
  bool flag_by_memcpy = false;
  bool wait_flush = true;  // CU_STREAM_WAIT_VALUE_FLUSH is dropped if the driver refuses it  // fallback when stream write-value is refused on peer memory
  CUresult (*fn_wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
  CUresult (*fn_write32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
  CUresult (*fn_addr_range)(CUdeviceptr*, size_t*, CUdeviceptr) = nullptr;
  int64_t* h_pinned = nullptr;  // pinned scratch
  int64_t* d_small  = nullptr;  // device scratch for tiny collectives
  size_t small_elems = 0;
};

namespace dj {

#define DJ_NCCL_TRY(expr)                                                                     \
  do {                                                                                        \
    ncclResult_t _r = (expr);                                                                 \
    if (_r != ncclSuccess) {                                                                  \
      dj::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, ncclGetErrorString(_r));    \
      return DJ_ERR_NCCL;                                                                     \
    }                                                                                         \
  } while (0)

constexpr int kFlagSlots = 64;
constexpr size_t kSmallElems = 1 << 20;  // int64 entries of pinned + device scratch per communicator

static int ensure_events(dj_comm* c, int n)
{
  while ((int)c->ev_batch.size() < n) {
    cudaEvent_t e;
    DJ_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->ev_batch.push_back(e);
  }
  return DJ_OK;
}

}  // namespace dj

using namespace dj;

static int ctrl_allgather(dj_comm* c, const int64_t* h_mine, int n, int64_t* h_all);

// Maps every peer's flag block; decides (collectively) whether the copy-engine exchange is usable.
static int setup_peer_exchange(dj_comm* c)
{
  const char* mode = getenv("DJ_EXCHANGE");
  bool ok = !(mode && (mode[0] == 'n' || mode[0] == 'N'));  // DJ_EXCHANGE=nccl forces the NCCL path
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) ok = false;
  c->fn_wait32 = (decltype(c->fn_wait32))fn;
  fn = nullptr;
  if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) ok = false;
  c->fn_write32 = (decltype(c->fn_write32))fn;
  fn = nullptr;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) ok = false;
  c->fn_addr_range = (decltype(c->fn_addr_range))fn;
  cudaGetLastError();

  DJ_CUDA_TRY(cudaMalloc(&c->d_flags, (size_t)c->size * kFlagSlots * sizeof(uint32_t)));
  DJ_CUDA_TRY(cudaMemset(c->d_flags, 0, (size_t)c->size * kFlagSlots * sizeof(uint32_t)));
  cudaIpcMemHandle_t mine;
  if (cudaIpcGetMemHandle(&mine, c->d_flags) != cudaSuccess) {
    ok = false;
    memset(&mine, 0, sizeof(mine));
    cudaGetLastError();
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  std::vector<int64_t> send(9), all((size_t)c->size * 9);
  send[0] = ok ? 1 : 0;
  memcpy(&send[1], &mine, 64);
  int rc = ctrl_allgather(c, send.data(), 9, all.data());
  if (rc) return rc;
  for (int r = 0; r < c->size; r++) ok = ok && all[(size_t)r * 9] == 1;
  c->peer_flags.assign(c->size, nullptr);
  c->peer_stream.assign(c->size, nullptr);
  c->ipc_cache.assign(c->size, {});
  int64_t opened = 1;
  if (ok) {
    for (int r = 0; r < c->size; r++) {
      if (r == c->rank) {
        c->peer_flags[r] = c->d_flags;
        continue;
      }
      cudaIpcMemHandle_t h;
      memcpy(&h, &all[(size_t)r * 9 + 1], 64);
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        opened = 0;
        cudaGetLastError();
        break;
      }
      c->peer_flags[r] = (uint32_t*)p;
      DJ_CUDA_TRY(cudaStreamCreateWithFlags(&c->peer_stream[r], cudaStreamNonBlocking));
    }
  }
  std::vector<int64_t> oks(c->size);
  rc = ctrl_allgather(c, &opened, 1, oks.data());
  if (rc) return rc;
  for (int r = 0; r < c->size; r++) ok = ok && oks[r] == 1;
  c->peer_ok = ok;
  return DJ_OK;
}

// Peer view of rank `peer`'s workspace allocation described by (handle, offset); opened once.
static char* map_peer_workspace(dj_comm* c, int peer, const cudaIpcMemHandle_t& h, int64_t offset)
{
  for (auto& e : c->ipc_cache[peer])
    if (memcmp(&e.h, &h, sizeof(h)) == 0) return e.base + offset;
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  c->ipc_cache[peer].push_back({h, (char*)p});
  return (char*)p + offset;
}

extern "C" int dj_comm_unique_id(void* h_id128)
{
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  DJ_NCCL_TRY(ncclGetUniqueId(&id));
  memcpy(h_id128, &id, sizeof(id));
  return DJ_OK;
}

extern "C" int dj_comm_create(int rank, int size, const void* h_id128, dj_comm_t** out)
{
  DJ_REQUIRE(out && size >= 1 && rank >= 0 && rank < size, "comm_create: bad rank/size");
  dj_comm* c = new dj_comm();
  c->rank    = rank;
  c->size    = size;
  DJ_CUDA_TRY(cudaGetDevice(&c->device));
  if (size > 1) {
    DJ_REQUIRE(h_id128, "comm_create: unique id missing");
    ncclUniqueId id;
    memcpy(&id, h_id128, sizeof(id));
    DJ_NCCL_TRY(ncclCommInitRank(&c->nccl, size, id, rank));
    DJ_NCCL_TRY(ncclCommSplit(c->nccl, 0, rank, &c->nccl_ctrl, nullptr));
  }
  DJ_CUDA_TRY(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
  DJ_CUDA_TRY(cudaStreamCreateWithFlags(&c->ctrl_stream, cudaStreamNonBlocking));
  DJ_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
  for (int i = 0; i < 2; i++) {
    DJ_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_part[i], cudaEventDisableTiming));
    DJ_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_seg[i], cudaEventDisableTiming));
  }
  DJ_CUDA_TRY(cudaMallocHost(&c->h_pinned, kSmallElems * sizeof(int64_t)));
  DJ_CUDA_TRY(cudaMalloc(&c->d_small, kSmallElems * sizeof(int64_t)));
  c->small_elems = kSmallElems;
  if (size > 1) {
    int rc = setup_peer_exchange(c);
    if (rc) return rc;
  }
  *out           = c;
  return DJ_OK;
}

extern "C" int dj_comm_destroy(dj_comm_t* c)
{
  if (!c) return DJ_OK;
  cudaDeviceSynchronize();
  if (c->nccl_ctrl) ncclCommDestroy(c->nccl_ctrl);
  if (c->nccl) ncclCommDestroy(c->nccl);
  for (int i = 0; i < 2; i++) {
    if (c->ev_part[i]) cudaEventDestroy(c->ev_part[i]);
    if (c->ev_seg[i]) cudaEventDestroy(c->ev_seg[i]);
  }
  if (c->ctrl_stream) cudaStreamDestroy(c->ctrl_stream);
  for (auto ps : c->peer_stream)
    if (ps) cudaStreamDestroy(ps);
  for (int i = 0; i < (int)c->peer_flags.size(); i++)
    if (c->peer_flags[i] && i != c->rank) cudaIpcCloseMemHandle(c->peer_flags[i]);
  for (auto& v : c->ipc_cache)
    for (auto& e : v) cudaIpcCloseMemHandle(e.base);
  if (c->d_flags) cudaFree(c->d_flags);
  for (auto e : c->ev_batch) cudaEventDestroy(e);
  if (c->ev_ready) cudaEventDestroy(c->ev_ready);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->d_small) cudaFree(c->d_small);
  delete c;
  return DJ_OK;
}

extern "C" int dj_comm_rank(const dj_comm_t* c) { return c ? c->rank : 0; }
extern "C" int dj_comm_size(const dj_comm_t* c) { return c ? c->size : 1; }

extern "C" int dj_comm_allgather_i64(dj_comm_t* c, const int64_t* h_mine, int n, int64_t* h_all,
                                     void* stream)
{
  DJ_REQUIRE(c && n >= 0, "allgather: bad argument");
  if (c->size == 1) {
    memcpy(h_all, h_mine, (size_t)n * 8);
    return DJ_OK;
  }
  DJ_REQUIRE((size_t)n * (c->size + 1) <= c->small_elems, "allgather: %d values per rank is too many", n);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t* d_send = c->d_small;
  int64_t* d_recv = c->d_small + n;
  memcpy(c->h_pinned, h_mine, (size_t)n * 8);
  DJ_CUDA_TRY(cudaMemcpyAsync(d_send, c->h_pinned, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  DJ_NCCL_TRY(ncclAllGather(d_send, d_recv, (size_t)n, ncclInt64, c->nccl, st));
  DJ_CUDA_TRY(cudaMemcpyAsync(c->h_pinned + n, d_recv, (size_t)n * c->size * 8,
                              cudaMemcpyDeviceToHost, st));
  DJ_CUDA_TRY(cudaStreamSynchronize(st));
  memcpy(h_all, c->h_pinned + n, (size_t)n * c->size * 8);
  return DJ_OK;
}

// Control-plane all-gather on the duplicate communicator and its own stream (blocking, tiny).
static int ctrl_allgather(dj_comm* c, const int64_t* h_mine, int n, int64_t* h_all)
{
  if (c->size == 1) {
    memcpy(h_all, h_mine, (size_t)n * 8);
    return DJ_OK;
  }
  const size_t half = c->small_elems / 2;  // second half of both scratch areas
  DJ_REQUIRE((size_t)n * (c->size + 1) <= half / 4, "allgather: %d values per rank is too many", n);
  cudaStream_t st = c->ctrl_stream;
  int64_t* hs = c->h_pinned + half + half / 2;
  int64_t* ds = c->d_small + half;
  memcpy(hs, h_mine, (size_t)n * 8);
  DJ_CUDA_TRY(cudaMemcpyAsync(ds, hs, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  DJ_NCCL_TRY(ncclAllGather(ds, ds + n, (size_t)n, ncclInt64, c->nccl_ctrl, st));
  DJ_CUDA_TRY(cudaMemcpyAsync(hs + n, ds + n, (size_t)n * c->size * 8, cudaMemcpyDeviceToHost, st));
  DJ_CUDA_TRY(cudaStreamSynchronize(st));
  memcpy(h_all, hs + n, (size_t)n * c->size * 8);
  return DJ_OK;
}

extern "C" int dj_comm_barrier(dj_comm_t* c, void* stream)
{
  int64_t mine = 1;
  std::vector<int64_t> all(c ? c->size : 1);
  if (!c || c->size == 1) return cudaStreamSynchronize((cudaStream_t)stream) == cudaSuccess ? DJ_OK : DJ_ERR_CUDA;
  return dj_comm_allgather_i64(c, &mine, 1, all.data(), stream);
}

extern "C" int dj_comm_group_start(dj_comm_t*)
{
  DJ_NCCL_TRY(ncclGroupStart());
  return DJ_OK;
}
extern "C" int dj_comm_group_end(dj_comm_t*)
{
  DJ_NCCL_TRY(ncclGroupEnd());
  return DJ_OK;
}
extern "C" int dj_comm_send(dj_comm_t* c, const void* d_buf, int64_t nbytes, int dest, void* stream)
{
  DJ_REQUIRE(c && c->nccl, "send: communicator has no NCCL (size 1)");
  DJ_NCCL_TRY(ncclSend(d_buf, (size_t)nbytes, ncclInt8, dest, c->nccl, (cudaStream_t)stream));
  return DJ_OK;
}
extern "C" int dj_comm_recv(dj_comm_t* c, void* d_buf, int64_t nbytes, int source, void* stream)
{
  DJ_REQUIRE(c && c->nccl, "recv: communicator has no NCCL (size 1)");
  DJ_NCCL_TRY(ncclRecv(d_buf, (size_t)nbytes, ncclInt8, source, c->nccl, (cudaStream_t)stream));
  return DJ_OK;
}

extern "C" int dj_all_to_all(dj_comm_t* c, int group_size, const int* h_group_ranks, int self_idx,
                             const void* const* h_send_cols, void* const* h_recv_cols,
                             const int64_t* h_send_offsets, const int64_t* h_recv_offsets,
                             const int* h_elem_sizes, int ncols, int include_self, void* stream)
{
  DJ_REQUIRE(c && group_size >= 1 && self_idx >= 0 && self_idx < group_size && ncols >= 0,
             "all_to_all: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (include_self) {
    const int64_t n = h_send_offsets[self_idx + 1] - h_send_offsets[self_idx];
    DJ_REQUIRE(n == h_recv_offsets[self_idx + 1] - h_recv_offsets[self_idx],
               "all_to_all: self send/recv sizes differ");
    for (int col = 0; col < ncols && n > 0; col++) {
      const size_t es = (size_t)h_elem_sizes[col];
      DJ_CUDA_TRY(cudaMemcpyAsync((char*)h_recv_cols[col] + h_recv_offsets[self_idx] * es,
                                  (const char*)h_send_cols[col] + h_send_offsets[self_idx] * es,
                                  (size_t)n * es, cudaMemcpyDeviceToDevice, st));
    }
  }
  if (group_size == 1) return DJ_OK;
  DJ_REQUIRE(c->nccl, "all_to_all: communicator has no NCCL (size 1)");
  DJ_NCCL_TRY(ncclGroupStart());
  for (int col = 0; col < ncols; col++) {
    const size_t es = (size_t)h_elem_sizes[col];
    for (int i = 0; i < group_size; i++) {
      if (i == self_idx) continue;
      const int64_t ns = h_send_offsets[i + 1] - h_send_offsets[i];
      const int64_t nr = h_recv_offsets[i + 1] - h_recv_offsets[i];
      if (ns > 0)
        DJ_NCCL_TRY(ncclSend((const char*)h_send_cols[col] + h_send_offsets[i] * es, (size_t)ns * es,
                             ncclInt8, h_group_ranks[i], c->nccl, st));
      if (nr > 0)
        DJ_NCCL_TRY(ncclRecv((char*)h_recv_cols[col] + h_recv_offsets[i] * es, (size_t)nr * es,
                             ncclInt8, h_group_ranks[i], c->nccl, st));
    }
  }
  DJ_NCCL_TRY(ncclGroupEnd());
  return DJ_OK;
}

// ------------------------------------------------------------------------- distributed join

static const uint32_t kNvlinkSeed = 12345678u;  // src/distributed_join.cpp:211
// Buckets handed to NCCL start on 32-row (256-byte) boundaries on both the send and the receive
// side: NCCL's peer copies drop to narrow accesses on pointers that are not 16-byte aligned (the
// reference works around the same effect with two staging copies, src/communicator.cpp:820-869).
constexpr int kAlignRows = 32;

static inline int64_t pad_rows(int64_t n) { return (n + kAlignRows - 1) / kAlignRows * kAlignRows; }

static size_t dist_ws_bytes(int64_t nl, int64_t nr, int world, int odf, double slack)
{
  if (world <= 1) return local_join_workspace(nl < nr ? nl : nr, nl < nr ? nr : nl) + 8192;
  const int nparts = world * odf;
  size_t total     = 1 << 16;
  // partitioned (padded) copies of both tables
  total += align_up((size_t)(nl + (int64_t)nparts * kAlignRows) * sizeof(Row), 256) +
           align_up((size_t)(nr + (int64_t)nparts * kAlignRows) * sizeof(Row), 256);
  total += 2 * pass_workspace_bytes(1, kMaxFanout) + 4 * align_up(((size_t)kMaxFanout + 1) * 8, 256);
  // receive buffers (balanced estimate with slack) + per-(source, sub-bucket) segment tables
  const size_t rl = (size_t)((double)nl * slack) + (size_t)nparts * kAlignRows + 4096;
  const size_t rr = (size_t)((double)nr * slack) + (size_t)nparts * kAlignRows + 4096;
  total += align_up(rl * sizeof(Row), 256) + align_up(rr * sizeof(Row), 256) + (size_t)odf * 8 * 256 +
           (size_t)odf * 2 * 3 * align_up((size_t)kMaxFanout * 8, 256);
  // join scratch for the largest batch (both sides stay alive until the join kernel has run)
  const int64_t bl = (int64_t)(rl / odf) + 4096, br = (int64_t)(rr / odf) + 4096;
  const RadixPlan plan = plan_for(bl < br ? bl : br, true);
  total += side_ws_bytes(bl, plan, kMaxFanout) + side_ws_bytes(br, plan, kMaxFanout);
  return total + 8192;
}

extern "C" size_t dj_distributed_inner_join_workspace_bytes(int64_t nleft, int64_t nright, int world,
                                                            int over_decom_factor)
{
  return dist_ws_bytes(nleft, nright, world, over_decom_factor < 1 ? 1 : over_decom_factor, 1.15);
}

// DJ_TRACE=1: device-side timeline of one call (CUDA event timestamps relative to its start)
struct Trace {
  bool on = false;
  cudaEvent_t base = nullptr;
  std::vector<std::pair<const char*, cudaEvent_t>> marks;
  std::vector<std::pair<const char*, double>> host_marks;  // host wall clock, ms since init
  std::chrono::high_resolution_clock::time_point t0;
  void host(const char* name)
  {
    if (!on) return;
    host_marks.push_back({name, std::chrono::duration<double, std::milli>(
                                  std::chrono::high_resolution_clock::now() - t0).count()});
  }
  void init(cudaStream_t st)
  {
    const char* e = getenv("DJ_TRACE");
    on            = e && e[0] == '1';
    if (!on) return;
    cudaEventCreate(&base);
    cudaEventRecord(base, st);
    t0 = std::chrono::high_resolution_clock::now();
  }
  void mark(const char* name, cudaStream_t st)
  {
    if (!on) return;
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    cudaEventRecord(ev, st);
    marks.push_back({name, ev});
  }
  void dump(int rank)
  {
    if (!on) return;
    cudaDeviceSynchronize();
    for (auto& m : marks) {
      float ms = 0;
      cudaEventElapsedTime(&ms, base, m.second);
      printf("[trace rank %d] %8.3f ms  %s\n", rank, ms, m.first);
      cudaEventDestroy(m.second);
    }
    cudaEventDestroy(base);
    for (auto& m : host_marks) printf("[trace rank %d] host %8.3f ms  %s\n", rank, m.second, m.first);
    fflush(stdout);
  }
};

static double ms_since(std::chrono::high_resolution_clock::time_point t0)
{
  return std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
}

extern "C" int dj_distributed_inner_join_i64(dj_comm_t* comm, const int64_t* d_left_key,
                                             const int64_t* d_left_payload, int64_t nleft,
                                             const int64_t* d_right_key,
                                             const int64_t* d_right_payload, int64_t nright,
                                             int64_t* d_out_lk, int64_t* d_out_lp,
                                             int64_t* d_out_rk, int64_t* d_out_rp,
                                             int64_t out_capacity, int64_t* h_out_count,
                                             dj_join_options* opts, void* d_workspace,
                                             size_t workspace_bytes, void* stream)
{
  DJ_REQUIRE(nleft >= 0 && nright >= 0 && h_out_count && d_workspace, "distributed_inner_join: bad argument");
  cudaStream_t st   = (cudaStream_t)stream;
  const int world   = comm ? comm->size : 1;
  const int rank    = comm ? comm->rank : 0;
  const int odf     = (opts && opts->over_decom_factor > 1) ? opts->over_decom_factor : 1;
  const bool timing = opts && opts->report_timing;
  if (opts) {
    opts->t_partition_ms = opts->t_comm_ms = opts->t_join_ms = 0;
    opts->bytes_sent = 0;
  }
  Arena arena(d_workspace, workspace_bytes);
  int64_t* d_count = arena.take<int64_t>(32);
  DJ_REQUIRE(d_count, "distributed_inner_join: workspace too small");
  DJ_CUDA_TRY(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
  int64_t* out[4] = {d_out_lk, d_out_lp, d_out_rk, d_out_rp};
  auto t0         = std::chrono::high_resolution_clock::now();

  if (world == 1) {
    // src/distributed_join.cpp:186-199 -- one rank: the local join is the whole job.
    const bool swap = nright < nleft;  // build on the smaller table
    int rc = swap ? local_join(d_right_key, d_right_payload, nright, d_left_key, d_left_payload,
                               nleft, out, out_capacity, d_count, true, arena, st)
                  : local_join(d_left_key, d_left_payload, nleft, d_right_key, d_right_payload,
                               nright, out, out_capacity, d_count, false, arena, st);
    if (rc) return rc;
    DJ_CUDA_TRY(cudaMemcpyAsync(comm ? comm->h_pinned : h_out_count, d_count, 8,
                                cudaMemcpyDeviceToHost, st));
    DJ_CUDA_TRY(cudaStreamSynchronize(st));
    if (comm) *h_out_count = comm->h_pinned[0];
    if (timing) {
      opts->t_join_ms = ms_since(t0);
      // the reference labels this branch's time "Hash partition" (src/distributed_join.cpp:194)
      printf("Rank %d: Hash partition takes %.0fms\n", rank, opts->t_join_ms);
    }
    return *h_out_count > out_capacity ? (set_error("join output needs %lld rows, capacity %lld",
                                                    (long long)*h_out_count, (long long)out_capacity),
                                          DJ_ERR_OVERFLOW)
                                       : DJ_OK;
  }

  const int G      = world;  // one NVSwitch box: the NVLink group is every rank
  const int nparts = G * odf;
  DJ_REQUIRE(nparts <= kMaxFanout, "distributed_inner_join: %d partitions exceed %d", nparts, kMaxFanout);
  DJ_REQUIRE(nleft < ((int64_t)1 << 31) && nright < ((int64_t)1 << 31),
             "distributed_inner_join: per-rank tables are limited to 2^31 rows");
  int rc = ensure_events(comm, 2 * odf);
  if (rc) return rc;
  Trace trace;
  trace.init(st);
  // The persistent partition / join kernels leave a couple of SMs idle for the whole call: the
  // control plane's tiny NCCL all-gathers are kernels too, and a GPU saturated by persistent CTAs
  // would make each of them wait for a kernel boundary (measured: up to 4 ms per collective).
  struct ReserveGuard {
    ReserveGuard()
    {
      const char* e = getenv("DJ_SM_RESERVE");
      set_sm_reserve(e ? atoi(e) : 2);
    }
    ~ReserveGuard() { set_sm_reserve(0); }
  } reserve_guard;

  // ---- 0. agree on the join's radix plan from the global table sizes.  When the plan has two
  //         levels, the first one is FUSED into the rank partition on the sender: bucket =
  //         (destination, sub-bucket), so the receiver only runs the second level.
  int sub_bits = 0;
  RadixPlan plan{0, 0, 1};
  {
    int64_t sizes[2] = {nleft, nright};
    std::vector<int64_t> alls((size_t)world * 2);
    rc = ctrl_allgather(comm, sizes, 2, alls.data());
    if (rc) return rc;
    int64_t tot[2] = {0, 0};
    for (int r = 0; r < world; r++) {
      tot[0] += alls[(size_t)r * 2];
      tot[1] += alls[(size_t)r * 2 + 1];
    }
    const int64_t est_build = std::min(tot[0], tot[1]) / nparts + 1;  // rows per rank and batch
    plan                    = plan_for(est_build, true);
    const int bits          = plan.bits1 + plan.bits2;
    int fit                 = 0;  // largest sub_bits with nparts << sub_bits <= kMaxFanout
    while ((nparts << (fit + 1)) <= kMaxFanout) fit++;
    const char* nofuse = getenv("DJ_NO_FUSE");
    if (plan.bits2 > 0 && fit > 0 && !(nofuse && nofuse[0] == '1')) {
      const int b1 = std::min(plan.bits1, fit);
      if (bits - b1 <= 10) {
        sub_bits   = b1;
        plan.bits1 = b1;
        plan.bits2 = bits - b1;
      }
    }
  }
  trace.host("plan agreed");
  // ---- 0b. copy-engine exchange: map every peer's workspace for this call (cached per allocation)
  bool use_peer = comm->peer_ok && 2 * odf <= kFlagSlots;
  std::vector<char*> peer_ws(world, nullptr);
  if (use_peer) {
    CUdeviceptr base = 0;
    size_t alloc_sz  = 0;
    cudaIpcMemHandle_t wh;
    memset(&wh, 0, sizeof(wh));
    bool ok = comm->fn_addr_range(&base, &alloc_sz, (CUdeviceptr)d_workspace) == CUDA_SUCCESS &&
              cudaIpcGetMemHandle(&wh, (void*)base) == cudaSuccess;
    cudaGetLastError();
    std::vector<int64_t> send(10), allh((size_t)world * 10);
    send[0] = ok ? 1 : 0;
    send[1] = ok ? (int64_t)((CUdeviceptr)d_workspace - base) : 0;
    memcpy(&send[2], &wh, 64);
    rc = ctrl_allgather(comm, send.data(), 10, allh.data());
    if (rc) return rc;
    for (int r = 0; r < world; r++) ok = ok && allh[(size_t)r * 10] == 1;
    int64_t mapped = 1;
    if (ok)
      for (int r = 0; r < world && mapped; r++) {
        if (r == rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, &allh[(size_t)r * 10 + 2], 64);
        peer_ws[r] = map_peer_workspace(comm, r, h, allh[(size_t)r * 10 + 1]);
        if (!peer_ws[r]) mapped = 0;
      }
    std::vector<int64_t> oks(world);
    rc = ctrl_allgather(comm, &mapped, 1, oks.data());
    if (rc) return rc;
    for (int r = 0; r < world; r++) ok = ok && oks[r] == 1;
    use_peer = ok;  // identical on every rank; otherwise fall back to the NCCL exchange
  }
  trace.host("peer workspaces mapped");
  const uint32_t seq = use_peer ? ++comm->seq : 0;
  std::vector<int64_t> peer_piece_off;  // [table][rank][batch][key|pay] byte offsets in the peer's workspace
  peer_piece_off.assign((size_t)2 * world * odf * 2, 0);

  const int F1s = 1 << sub_bits;       // sub-buckets per destination in the sender's partition
  const int nbk = nparts << sub_bits;  // buckets of the sender's partition
  const int nseg = G * F1s;            // (source, sub-bucket) segments of a received piece

  // ---- 1. hash partition both tables (src/distributed_join.cpp:213-225) on the caller's stream;
  //         every destination's run of buckets starts on kAlignRows so NCCL sends straight from it
  const int64_t n_in[2]    = {nleft, nright};
  const int64_t* in_key[2] = {d_left_key, d_right_key};
  const int64_t* in_pay[2] = {d_left_payload, d_right_payload};
  const size_t pw          = pass_workspace_bytes(1, nbk);
  Row* prow[2];
  int64_t *d_off[2], *d_cnt[2];
  for (int t = 0; t < 2; t++) {
    char* pws         = arena.take<char>(pw);
    const size_t rows = (size_t)(n_in[t] + (int64_t)nparts * kAlignRows);
    prow[t]           = arena.take<Row>(rows);
    d_off[t]          = arena.take<int64_t>((size_t)nbk + 1);
    d_cnt[t]          = arena.take<int64_t>((size_t)nbk + 1);
    if (!pws || !prow[t] || !d_off[t] || !d_cnt[t]) {
      set_error("distributed_inner_join: workspace too small for the partitioned tables");
      return DJ_ERR_WORKSPACE;
    }
    PassDesc desc{sub_bits ? 2 : 0, kNvlinkSeed, DJ_HASH_MURMUR3, 0, nbk, 1, 1, kAlignRows, nparts, sub_bits};
    PassBuffers pb{};
    pb.in_key = in_key[t]; pb.in_pay[0] = in_pay[t]; pb.out_rows = prow[t];
    pb.nrows = n_in[t]; pb.d_child_off = d_off[t]; pb.d_child_cnt = d_cnt[t];
    rc = run_partition_pass(desc, pb, pws, pw, st);
    if (rc) return rc;
    DJ_CUDA_TRY(cudaEventRecord(comm->ev_part[t], st));
    trace.mark(t ? "partition(R) done" : "partition(L) done", st);
    trace.host(t ? "partition(R) launched" : "partition(L) launched");
  }

  // ---- 2-4. table by table: sizes (communicate_sizes, on the control communicator), receive
  //           layout (allocate_communicated_table), exchange.  The left table's exchange starts
  //           while the right table is still being partitioned; an event per (batch, table) hands
  //           each received piece to the compute stream, so radix passes overlap later exchanges.
  struct Piece {
    std::vector<int64_t> begin, count;  // per source
    int64_t span = 0, rows = 0;
    Row* data = nullptr;
    int64_t *d_seg_begin = nullptr, *d_seg_end = nullptr;
    int* d_seg_parent = nullptr;
  };
  std::vector<Piece> pieces((size_t)odf * 2);
  std::vector<int64_t> off[2], cntv[2], allc[2];
  int64_t* hp   = comm->h_pinned + (256 << 10);  // D2H landing zone for offsets / counts
  int64_t* hseg = comm->h_pinned + (320 << 10);  // pinned staging for the segment tables
  DJ_REQUIRE(pieces.size() * 3 * (size_t)nseg <= (192u << 10), "distributed_inner_join: too many segments");
  int64_t max_span[2] = {0, 0};
  bool exchange_in_flight = false;
  auto tcomm = std::chrono::high_resolution_clock::now();

  // rows of source `src`'s table t in sub-bucket `sub` of destination bucket q
  auto cnt = [&](int src, int t, int q, int sub) {
    return allc[t][(size_t)src * nbk + ((size_t)q << sub_bits) + sub];
  };
  auto issue_exchange = [&](int b, int t) -> int {
    Piece& pc = pieces[(size_t)b * 2 + t];
    auto send_begin = [&](int dest) { return off[t][((size_t)b * G + dest) << sub_bits]; };
    auto send_count = [&](int dest) {
      int64_t c = 0;
      for (int sub = 0; sub < F1s; sub++) c += cntv[t][(((size_t)b * G + dest) << sub_bits) + sub];
      return c;
    };
    trace.mark(t ? "exchange(R) begin" : "exchange(L) begin", comm->comm_stream);
    // own bucket: device copy (src/all_to_all_comm.cpp:610-653); the rest over NVLink
    if (pc.count[rank] > 0)
      DJ_CUDA_TRY(cudaMemcpyAsync(pc.data + pc.begin[rank], prow[t] + send_begin(rank),
                                  (size_t)pc.count[rank] * sizeof(Row), cudaMemcpyDeviceToDevice, comm->comm_stream));
    if (use_peer) {
      // push every peer's bucket into ITS receive piece with the copy engines (no SMs, so the
      // radix passes running meanwhile keep the whole GPU), then raise that peer's flag
      const int slot = (b * 2 + t) % kFlagSlots;
      for (int i = 0; i < G; i++) {
        if (i == rank) continue;
        cudaStream_t ps = comm->peer_stream[i];
        DJ_CUDA_TRY(cudaStreamWaitEvent(ps, comm->ev_part[t], 0));
        const int64_t ns = send_count(i);
        if (ns > 0) {
          // where my rows start inside peer i's piece: padded counts of the sources before me
          int64_t dst_begin = 0;
          for (int s2 = 0; s2 < rank; s2++) {
            int64_t c = 0;
            for (int sub = 0; sub < F1s; sub++) c += cnt(s2, t, b * G + i, sub);
            dst_begin += pad_rows(c);
          }
          const int64_t* po = &peer_piece_off[(((size_t)t * world + i) * odf + b) * 2];
          DJ_CUDA_TRY(cudaMemcpyAsync(peer_ws[i] + po[0] + dst_begin * sizeof(Row), prow[t] + send_begin(i),
                                      (size_t)ns * sizeof(Row), cudaMemcpyDefault, ps));
          if (opts) opts->bytes_sent += 16 * ns;
        }
        uint32_t* flag = comm->peer_flags[i] + (size_t)rank * kFlagSlots + slot;
        if (!comm->flag_by_memcpy &&
            comm->fn_write32((CUstream)ps, (CUdeviceptr)flag, seq, 0) != CUDA_SUCCESS)
          comm->flag_by_memcpy = true;
        if (comm->flag_by_memcpy) {
          // 4-byte copy from a pinned word (ordered behind the data copies on the same stream)
          uint32_t* src = reinterpret_cast<uint32_t*>(comm->h_pinned + (900 << 10)) + (seq % 4096);
          *src          = seq;
          DJ_CUDA_TRY(cudaMemcpyAsync(flag, src, 4, cudaMemcpyDefault, ps));
        }
      }
      DJ_CUDA_TRY(cudaEventRecord(comm->ev_batch[(size_t)b * 2 + t], comm->comm_stream));
      exchange_in_flight = true;
      return DJ_OK;
    }
    DJ_NCCL_TRY(ncclGroupStart());
    for (int i = 0; i < G; i++) {
      if (i == rank) continue;
      const int64_t ns = send_count(i), nr = pc.count[i];
      if (ns > 0) {
        DJ_NCCL_TRY(ncclSend(prow[t] + send_begin(i), (size_t)ns * sizeof(Row), ncclInt8, i, comm->nccl,
                             comm->comm_stream));
        if (opts) opts->bytes_sent += 16 * ns;
      }
      if (nr > 0) {
        DJ_NCCL_TRY(ncclRecv(pc.data + pc.begin[i], (size_t)nr * sizeof(Row), ncclInt8, i, comm->nccl,
                             comm->comm_stream));
      }
    }
    DJ_NCCL_TRY(ncclGroupEnd());
    DJ_CUDA_TRY(cudaEventRecord(comm->ev_batch[(size_t)b * 2 + t], comm->comm_stream));
    trace.mark(t ? "exchange(R) end" : "exchange(L) end", comm->comm_stream);
    exchange_in_flight = true;
    return DJ_OK;
  };
  auto drain_exchange = [&]() {
    if (!exchange_in_flight) return;
    cudaStreamSynchronize(comm->comm_stream);
    if (use_peer)
      for (int i = 0; i < G; i++)
        if (i != rank) cudaStreamSynchronize(comm->peer_stream[i]);
  };
  for (int t = 0; t < 2; t++) {
    // sizes of table t: wait only for ITS partition pass
    DJ_CUDA_TRY(cudaStreamWaitEvent(comm->ctrl_stream, comm->ev_part[t], 0));
    DJ_CUDA_TRY(cudaMemcpyAsync(hp, d_off[t], (size_t)(nbk + 1) * 8, cudaMemcpyDeviceToHost, comm->ctrl_stream));
    DJ_CUDA_TRY(cudaMemcpyAsync(hp + nbk + 1, d_cnt[t], (size_t)nbk * 8, cudaMemcpyDeviceToHost, comm->ctrl_stream));
    DJ_CUDA_TRY(cudaStreamSynchronize(comm->ctrl_stream));
    off[t].assign(hp, hp + nbk + 1);
    cntv[t].assign(hp + nbk + 1, hp + nbk + 1 + nbk);
    if (timing && t == 1) {
      opts->t_partition_ms = ms_since(t0);
      printf("Rank %d: Hash partition takes %.0fms\n", rank, opts->t_partition_ms);
      tcomm = std::chrono::high_resolution_clock::now();
    }
    trace.host(t ? "offsets(R) on host" : "offsets(L) on host");
    allc[t].resize((size_t)world * nbk);
    rc = ctrl_allgather(comm, cntv[t].data(), nbk, allc[t].data());
    if (rc) return rc;
    trace.host(t ? "counts(R) gathered" : "counts(L) gathered");

    // receive layout: per (batch) one padded piece per source, holding that source's F1s
    // sub-buckets back to back
    size_t need = arena.used;
    for (int b = 0; b < odf; b++) {
      Piece& pc = pieces[(size_t)b * 2 + t];
      pc.begin.resize(G);
      pc.count.resize(G);
      for (int s = 0; s < G; s++) {
        int64_t c = 0;
        for (int sub = 0; sub < F1s; sub++) c += cnt(s, t, b * G + rank, sub);
        pc.begin[s] = pc.span;
        pc.count[s] = c;
        pc.span += pad_rows(c);
        pc.rows += c;
      }
      need += align_up((size_t)pc.span * sizeof(Row) + 128, 256) + 3 * align_up((size_t)nseg * 8, 256) + 1024;
      max_span[t] = std::max(max_span[t], pc.span);
    }
    if (t == 1) need += side_ws_bytes(max_span[0], plan, nseg) + side_ws_bytes(max_span[1], plan, nseg) + 4096;
    // pieces are laid out first (pure arithmetic), then ONE collective carries both the
    // "it fits" verdict and the piece offsets the peers need for their pushes
    const bool fits = need <= workspace_bytes;
    for (int b = 0; b < odf && fits; b++) {
      const size_t i  = (size_t)b * 2 + t;
      Piece& pc       = pieces[i];
      pc.data         = arena.take<Row>((size_t)pc.span + 8);
      pc.d_seg_begin  = arena.take<int64_t>((size_t)nseg);
      pc.d_seg_end    = arena.take<int64_t>((size_t)nseg);
      pc.d_seg_parent = arena.take<int>((size_t)nseg);
    }
    {
      std::vector<int64_t> mine_off((size_t)odf * 2 + 1), all_off((size_t)world * (odf * 2 + 1));
      bool ok_local = fits;
      for (int b = 0; b < odf && fits; b++) {
        Piece& pc = pieces[(size_t)b * 2 + t];
        if (!pc.data || !pc.d_seg_begin || !pc.d_seg_end || !pc.d_seg_parent) ok_local = false;
      }
      mine_off[0] = ok_local ? 1 : 0;
      for (int b = 0; b < odf && ok_local; b++) {
        mine_off[1 + (size_t)b * 2]     = (char*)pieces[(size_t)b * 2 + t].data - (char*)d_workspace;
        mine_off[1 + (size_t)b * 2 + 1] = 0;
      }
      rc = ctrl_allgather(comm, mine_off.data(), odf * 2 + 1, all_off.data());
      if (rc) return rc;
      for (int r = 0; r < world; r++) {
        if (!all_off[(size_t)r * (odf * 2 + 1)]) {
          drain_exchange();
          set_error("distributed_inner_join: workspace too small on rank %d for its received partitions "
                    "(this rank needs %zu of %zu bytes)", r, need, workspace_bytes);
          return DJ_ERR_WORKSPACE;
        }
        for (int k = 0; k < odf * 2; k++)
          peer_piece_off[((size_t)t * world + r) * odf * 2 + k] = all_off[(size_t)r * (odf * 2 + 1) + 1 + k];
      }
    }
    for (int b = 0; b < odf; b++) {
      const size_t i = (size_t)b * 2 + t;
      Piece& pc      = pieces[i];
      int64_t* hb = hseg + i * 3 * nseg;
      int* hpar   = reinterpret_cast<int*>(hb + 2 * (size_t)nseg);
      for (int s = 0; s < G; s++) {
        int64_t at = pc.begin[s];
        for (int sub = 0; sub < F1s; sub++) {
          const int64_t c          = cnt(s, t, b * G + rank, sub);
          hb[s * F1s + sub]        = at;
          hb[nseg + s * F1s + sub] = at + c;
          hpar[s * F1s + sub]      = sub;
          at += c;
        }
      }
      DJ_CUDA_TRY(cudaMemcpyAsync(pc.d_seg_begin, hb, (size_t)nseg * 8, cudaMemcpyHostToDevice, comm->ctrl_stream));
      DJ_CUDA_TRY(cudaMemcpyAsync(pc.d_seg_end, hb + nseg, (size_t)nseg * 8, cudaMemcpyHostToDevice, comm->ctrl_stream));
      DJ_CUDA_TRY(cudaMemcpyAsync(pc.d_seg_parent, hpar, (size_t)nseg * 4, cudaMemcpyHostToDevice, comm->ctrl_stream));
    }
    DJ_CUDA_TRY(cudaEventRecord(comm->ev_seg[t], comm->ctrl_stream));

    // exchanges in batch order (b,L),(b,R); what can start now: (0,L) after the left table's
    // sizes, everything else once the right table's sizes are known
    trace.host(t ? "pieces(R) laid out" : "pieces(L) laid out");
    DJ_CUDA_TRY(cudaStreamWaitEvent(comm->comm_stream, comm->ev_part[t], 0));
    if (t == 0) {
      rc = issue_exchange(0, 0);
      if (rc) return rc;
    } else {
      rc = issue_exchange(0, 1);
      if (rc) return rc;
      for (int b = 1; b < odf; b++)
        for (int tt = 0; tt < 2; tt++) {
          rc = issue_exchange(b, tt);
          if (rc) return rc;
        }
    }
  }
  if (timing) {
    DJ_CUDA_TRY(cudaStreamSynchronize(comm->comm_stream));
    opts->t_comm_ms = ms_since(tcomm);
    for (int b = 0; b < odf; b++)
      printf("Rank %d: All-to-all communication on batch %d takes %.0fms\n", rank, b, opts->t_comm_ms / odf);
  }
  trace.host("exchanges issued");
  DJ_CUDA_TRY(cudaStreamWaitEvent(st, comm->ev_seg[0], 0));
  DJ_CUDA_TRY(cudaStreamWaitEvent(st, comm->ev_seg[1], 0));

  const size_t join_mark = arena.used;
  for (int b = 0; b < odf; b++) {
    auto tj    = std::chrono::high_resolution_clock::now();
    Piece& L   = pieces[(size_t)b * 2];
    Piece& R   = pieces[(size_t)b * 2 + 1];
    arena.used = join_mark;  // join scratch is reused batch after batch (same stream)
    if (L.rows == 0 || R.rows == 0) continue;  // src/distributed_join.cpp:76-82
    const bool swap = R.rows < L.rows;         // build on the smaller side
    PreparedSide side[2];
    for (int t = 0; t < 2; t++) {
      Piece& pc = t ? R : L;
      DJ_CUDA_TRY(cudaStreamWaitEvent(st, comm->ev_batch[(size_t)b * 2 + t], 0));
      if (use_peer) {
        const int slot = (b * 2 + t) % kFlagSlots;
        for (int src = 0; src < G; src++) {
          if (src == rank) continue;
          const CUdeviceptr fa = (CUdeviceptr)(comm->d_flags + (size_t)src * kFlagSlots + slot);
          CUresult wr = CUDA_ERROR_NOT_SUPPORTED;
          if (comm->wait_flush) {
            wr = comm->fn_wait32((CUstream)st, fa, seq, CU_STREAM_WAIT_VALUE_GEQ | CU_STREAM_WAIT_VALUE_FLUSH);
            if (wr != CUDA_SUCCESS) comm->wait_flush = false;
          }
          if (wr != CUDA_SUCCESS) wr = comm->fn_wait32((CUstream)st, fa, seq, CU_STREAM_WAIT_VALUE_GEQ);
          if (wr != CUDA_SUCCESS) {
            drain_exchange();
            set_error("distributed_inner_join: cuStreamWaitValue32 failed with CUresult %d", (int)wr);
            return DJ_ERR_CUDA;
          }
        }
        trace.mark(t ? "arrived(R)" : "arrived(L)", st);
      }
      TableInput in{nullptr, nullptr, pc.data, pc.span, pc.d_seg_begin, pc.d_seg_end, nseg, pc.d_seg_parent,
                    sub_bits > 0};
      trace.mark(t ? "radix(R) begin" : "radix(L) begin", st);
      rc = prepare_side(in, plan, &side[t], arena, st);
      if (rc) return rc;
      trace.mark(t ? "radix(R) end" : "radix(L) end", st);
    }
    rc = join_prepared(side[swap ? 1 : 0], side[swap ? 0 : 1], plan, out, out_capacity, d_count, swap, st);
    if (rc) return rc;
    trace.mark("join end", st);
    if (timing) {
      DJ_CUDA_TRY(cudaStreamSynchronize(st));
      double ms = ms_since(tj);
      opts->t_join_ms += ms;
      printf("Rank %d: Local join on batch %d takes %.0fms\n", rank, b, ms);
    }
  }
  trace.host("join launched");
  DJ_CUDA_TRY(cudaMemcpyAsync(comm->h_pinned, d_count, 8, cudaMemcpyDeviceToHost, st));
  DJ_CUDA_TRY(cudaStreamSynchronize(st));
  DJ_CUDA_TRY(cudaStreamSynchronize(comm->comm_stream));
  trace.host("streams drained");
  if (use_peer)
    for (int i = 0; i < G; i++)
      if (i != rank) DJ_CUDA_TRY(cudaStreamSynchronize(comm->peer_stream[i]));  // my buckets may be reused now
  *h_out_count = comm->h_pinned[0];
  trace.dump(rank);
  // the overflow verdict is collective: every rank returns DJ_ERR_OVERFLOW if any rank's
  // output did not fit, so that callers can retry together
  {
    int64_t over = *h_out_count > out_capacity ? 1 : 0;
    std::vector<int64_t> overs(world);
    rc = ctrl_allgather(comm, &over, 1, overs.data());
    if (rc) return rc;
    for (int r = 0; r < world; r++)
      if (overs[r]) {
        set_error("join output does not fit on rank %d (this rank: %lld rows, capacity %lld)", r,
                  (long long)*h_out_count, (long long)out_capacity);
        return DJ_ERR_OVERFLOW;
      }
  }
  return DJ_OK;
}

extern "C" size_t dj_distributed_inner_join_host_workspace_bytes(int64_t nleft, int64_t nright,
                                                                 int64_t out_capacity, int world,
                                                                 int over_decom_factor)
{
  return dj_distributed_inner_join_workspace_bytes(nleft, nright, world, over_decom_factor) +
         2 * (align_up((size_t)nleft * 8, 256) + align_up((size_t)nright * 8, 256)) +
         4 * align_up((size_t)out_capacity * 8, 256) + 8192;
}

extern "C" int dj_distributed_inner_join_i64_host(
  dj_comm_t* comm, const int64_t* h_left_key, const int64_t* h_left_payload, int64_t nleft,
  const int64_t* h_right_key, const int64_t* h_right_payload, int64_t nright, int64_t* h_out_lk,
  int64_t* h_out_lp, int64_t* h_out_rk, int64_t* h_out_rp, int64_t out_capacity,
  int64_t* h_out_count, dj_join_options* opts, void* d_workspace, size_t workspace_bytes,
  void* stream)
{
  DJ_REQUIRE(d_workspace && h_out_count, "distributed_inner_join_host: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  Arena arena(d_workspace, workspace_bytes);
  int64_t* dlk = arena.take<int64_t>((size_t)nleft);
  int64_t* dlp = arena.take<int64_t>((size_t)nleft);
  int64_t* drk = arena.take<int64_t>((size_t)nright);
  int64_t* drp = arena.take<int64_t>((size_t)nright);
  int64_t* o[4];
  for (int c = 0; c < 4; c++) o[c] = arena.take<int64_t>((size_t)out_capacity);
  if (!dlk || !dlp || !drk || !drp || !o[0] || !o[1] || !o[2] || !o[3]) {
    set_error("distributed_inner_join_host: workspace too small");
    return DJ_ERR_WORKSPACE;
  }
  DJ_CUDA_TRY(cudaMemcpyAsync(dlk, h_left_key, (size_t)nleft * 8, cudaMemcpyHostToDevice, st));
  DJ_CUDA_TRY(cudaMemcpyAsync(dlp, h_left_payload, (size_t)nleft * 8, cudaMemcpyHostToDevice, st));
  DJ_CUDA_TRY(cudaMemcpyAsync(drk, h_right_key, (size_t)nright * 8, cudaMemcpyHostToDevice, st));
  DJ_CUDA_TRY(cudaMemcpyAsync(drp, h_right_payload, (size_t)nright * 8, cudaMemcpyHostToDevice, st));
  const size_t off = align_up(arena.used, 256);
  int rc = dj_distributed_inner_join_i64(comm, dlk, dlp, nleft, drk, drp, nright, o[0], o[1], o[2],
                                         o[3], out_capacity, h_out_count, opts,
                                         (char*)d_workspace + off, workspace_bytes - off, stream);
  if (rc && rc != DJ_ERR_OVERFLOW) return rc;
  const int64_t n = *h_out_count < out_capacity ? *h_out_count : out_capacity;
  int64_t* h[4]   = {h_out_lk, h_out_lp, h_out_rk, h_out_rp};
  for (int c = 0; c < 4; c++)
    DJ_CUDA_TRY(cudaMemcpyAsync(h[c], o[c], (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  DJ_CUDA_TRY(cudaStreamSynchronize(st));
  return rc;
}
