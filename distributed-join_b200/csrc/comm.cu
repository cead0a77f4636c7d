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
  cudaEvent_t ev_hist      = nullptr;
  cudaEvent_t ev_part[2]   = {nullptr, nullptr};
  cudaEvent_t ev_seg[2]    = {nullptr, nullptr};
  std::vector<cudaEvent_t> ev_batch;
  // Peer-memory exchange (copy engines over NVLink, no SMs): every rank maps every peer's flag
  // block and -- per call -- workspace through CUDA IPC, pushes its buckets with cudaMemcpyAsync
  // and signals with a stream write; receivers wait with a stream wait-value.
  bool peer_ok = false;
  uint32_t* d_flags = nullptr;               // [size][kFlagSlots], written by peers
  std::vector<uint32_t*> peer_flags;         // peers' d_flags mapped here
  int64_t* d_inbox = nullptr;                // [size][kInbox] control messages deposited by peers (same allocation)
  std::vector<int64_t*> peer_inbox;          // peers' d_inbox mapped here
  uint32_t cseq = 0;                         // control-message sequence number
  std::vector<cudaStream_t> peer_stream;     // one push stream per peer
  struct IpcEntry { cudaIpcMemHandle_t h; char* base; };
  std::vector<std::vector<IpcEntry>> ipc_cache;  // per peer: opened workspace allocations (most recent last)
  std::vector<int64_t> last_handle;              // [size][8] workspace handles seen in the previous call
  uint32_t** d_peer_flags = nullptr;             // device copy of peer_flags (verdict_kernel)
  std::vector<cudaEvent_t> ev_xbeg, ev_xend;     // [2][size] timing events around the pushes (measure_exchange)
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

constexpr int kFlagSlots = 64;      // per source rank: data flags 0..61, control inbox flag 62, verdict 63
constexpr int kCtrlSlot  = kFlagSlots - 2;
constexpr int kDataSlots = kFlagSlots - 2;
constexpr int kInbox     = 4096;  // int64 words every source rank may deposit per control message
// Messages of one join call (hello, mapping ack, bucket counts) land in separate banks of the inbox: a
// fast rank may already send its NEXT message while a slow peer has not yet read the previous one.
enum { kBankHello = 0, kBankAck = 1, kBankCounts = 2, kBankMisc = 3, kInboxBanks = 4 };
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

static int ensure_xevents(dj_comm* c, int n)
{
  while ((int)c->ev_xbeg.size() < n) {
    cudaEvent_t a, b;
    DJ_CUDA_TRY(cudaEventCreate(&a));
    DJ_CUDA_TRY(cudaEventCreate(&b));
    c->ev_xbeg.push_back(a);
    c->ev_xend.push_back(b);
  }
  return DJ_OK;
}

// The stream waits until *d_flag >= value (a peer raises it over NVLink).
static int stream_wait_flag(dj_comm* c, cudaStream_t st, const uint32_t* d_flag, uint32_t value)
{
  const CUdeviceptr fa = (CUdeviceptr)d_flag;
  CUresult wr          = CUDA_ERROR_NOT_SUPPORTED;
  if (c->wait_flush) {
    wr = c->fn_wait32((CUstream)st, fa, value, CU_STREAM_WAIT_VALUE_GEQ | CU_STREAM_WAIT_VALUE_FLUSH);
    if (wr != CUDA_SUCCESS) c->wait_flush = false;
  }
  if (wr != CUDA_SUCCESS) wr = c->fn_wait32((CUstream)st, fa, value, CU_STREAM_WAIT_VALUE_GEQ);
  if (wr != CUDA_SUCCESS) {
    set_error("distributed_inner_join: cuStreamWaitValue32 failed with CUresult %d", (int)wr);
    return DJ_ERR_CUDA;
  }
  return DJ_OK;
}

// Control-plane all-gather WITHOUT kernels: every rank deposits `n` int64 words (device or pinned host
// memory) into every peer's inbox with copy-engine copies over NVLink, raises the peer's control
// flag with a stream memory operation, waits for the peers' flags and reads its inbox back.  Unlike
// an NCCL collective it needs no SM, so it is never held up by the persistent partition / join
// kernels that fill the GPU (an NCCL all-gather issued next to them waited for a kernel boundary:
// 4.1 ms measured at N=2).  Every message kind of a call has its own inbox bank (a fast rank's
// counts must not overwrite a hello that a slow peer has not read yet); a bank is safe to reuse in
// the next call because a rank leaves a join only after every peer has published its verdict,
// i.e. long after all banks of that call were read.
static int peer_allgather(dj_comm* c, int bank, const void* src, int n, int64_t* h_all)
{
  DJ_REQUIRE(n >= 1 && n <= kInbox, "control message of %d words exceeds the inbox", n);
  cudaStream_t st    = c->ctrl_stream;
  const uint32_t seq = ++c->cseq;
  const size_t bank_off = (size_t)bank * c->size * kInbox;
  for (int k = 0; k < c->size; k++) {
    const int i = (c->rank + k) % c->size;  // staggered: no two ranks start with the same destination
    DJ_CUDA_TRY(cudaMemcpyAsync(c->peer_inbox[i] + bank_off + (size_t)c->rank * kInbox, src, (size_t)n * 8,
                                cudaMemcpyDefault, st));
    if (i == c->rank) continue;
    uint32_t* flag = c->peer_flags[i] + (size_t)c->rank * kFlagSlots + kCtrlSlot;
    if (!c->flag_by_memcpy && c->fn_write32((CUstream)st, (CUdeviceptr)flag, seq, 0) != CUDA_SUCCESS)
      c->flag_by_memcpy = true;
    if (c->flag_by_memcpy) {
      uint32_t* w = reinterpret_cast<uint32_t*>(c->h_pinned + (910 << 10)) + (seq % 4096);
      *w          = seq;
      DJ_CUDA_TRY(cudaMemcpyAsync(flag, w, 4, cudaMemcpyDefault, st));
    }
  }
  for (int srcr = 0; srcr < c->size; srcr++) {
    if (srcr == c->rank) continue;
    int rc = stream_wait_flag(c, st, c->d_flags + (size_t)srcr * kFlagSlots + kCtrlSlot, seq);
    if (rc) return rc;
  }
  int64_t* land = c->h_pinned + (512 << 10);  // [size][n]
  DJ_REQUIRE((size_t)n * c->size <= (256u << 10), "control message too large for the landing zone");
  DJ_CUDA_TRY(cudaMemcpy2DAsync(land, (size_t)n * 8, c->d_inbox + bank_off, (size_t)kInbox * 8, (size_t)n * 8, c->size,
                                cudaMemcpyDeviceToHost, st));
  DJ_CUDA_TRY(cudaStreamSynchronize(st));
  memcpy(h_all, land, (size_t)n * c->size * 8);
  return DJ_OK;
}

// One thread: this rank's overflow verdict goes into every peer's flag block (slot `slot` of row
// `rank`) as (seq << 1) | overflowed, with system-scope release stores over NVLink.
__global__ void verdict_kernel(const unsigned long long* d_count, unsigned long long capacity,
                               uint32_t* const* peer_flags, int world, int rank, int slots, int slot, uint32_t seq)
{
  if (threadIdx.x != 0) return;
  const uint32_t v = (seq << 1) | (*d_count > capacity ? 1u : 0u);
  for (int p = 0; p < world; p++) {
    uint32_t* f = peer_flags[p] + (size_t)rank * slots + slot;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(v) : "memory");
  }
}

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

  const size_t flag_bytes = align_up((size_t)c->size * kFlagSlots * sizeof(uint32_t), 256);
  const size_t ctl_bytes  = flag_bytes + (size_t)kInboxBanks * c->size * kInbox * sizeof(int64_t);
  DJ_CUDA_TRY(cudaMalloc(&c->d_flags, ctl_bytes));
  DJ_CUDA_TRY(cudaMemset(c->d_flags, 0, ctl_bytes));
  c->d_inbox = reinterpret_cast<int64_t*>(reinterpret_cast<char*>(c->d_flags) + flag_bytes);
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
  c->peer_inbox.assign(c->size, nullptr);
  c->peer_stream.assign(c->size, nullptr);
  c->ipc_cache.assign(c->size, {});
  int64_t opened = 1;
  if (ok) {
    for (int r = 0; r < c->size; r++) {
      if (r == c->rank) {
        c->peer_flags[r] = c->d_flags;
        c->peer_inbox[r] = c->d_inbox;
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
      c->peer_inbox[r] = reinterpret_cast<int64_t*>(reinterpret_cast<char*>(p) + flag_bytes);
      DJ_CUDA_TRY(cudaStreamCreateWithFlags(&c->peer_stream[r], cudaStreamNonBlocking));
    }
  }
  std::vector<int64_t> oks(c->size);
  rc = ctrl_allgather(c, &opened, 1, oks.data());
  if (rc) return rc;
  for (int r = 0; r < c->size; r++) ok = ok && oks[r] == 1;
  if (ok) {
    DJ_CUDA_TRY(cudaMalloc(&c->d_peer_flags, (size_t)c->size * sizeof(uint32_t*)));
    DJ_CUDA_TRY(cudaMemcpy(c->d_peer_flags, c->peer_flags.data(), (size_t)c->size * sizeof(uint32_t*),
                           cudaMemcpyHostToDevice));
  }
  c->peer_ok = ok;
  return DJ_OK;
}

// Peer view of rank `peer`'s workspace allocation described by (handle, offset); opened once.
static char* map_peer_workspace(dj_comm* c, int peer, const cudaIpcMemHandle_t& h, int64_t offset)
{
  auto& cache = c->ipc_cache[peer];
  for (size_t i = 0; i < cache.size(); i++)
    if (memcmp(&cache[i].h, &h, sizeof(h)) == 0) {
      const dj_comm::IpcEntry e = cache[i];
      cache.erase(cache.begin() + i);
      cache.push_back(e);  // most recently used last
      return e.base + offset;
    }
  // a peer rarely alternates between more than two live workspaces: older mappings are closed so
  // that the memory behind them can really be released by its owner
  while (cache.size() >= 2) {
    cudaIpcCloseMemHandle(cache.front().base);
    cache.erase(cache.begin());
  }
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
  DJ_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_hist, cudaEventDisableTiming));
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
  if (c->d_peer_flags) cudaFree(c->d_peer_flags);
  for (auto e : c->ev_xbeg) cudaEventDestroy(e);
  for (auto e : c->ev_xend) cudaEventDestroy(e);
  if (c->ev_hist) cudaEventDestroy(c->ev_hist);
  for (auto e : c->ev_batch) cudaEventDestroy(e);
  if (c->ev_ready) cudaEventDestroy(c->ev_ready);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->d_small) cudaFree(c->d_small);
  delete c;
  return DJ_OK;
}

// Collective: every rank closes its mappings of the peers' workspaces, then all ranks meet, so that
// a caller may cudaFree / shrink / regrow its workspace afterwards (freeing memory that an importer
// still has open is undefined behaviour in CUDA IPC).
extern "C" int dj_comm_release_workspace(dj_comm_t* c)
{
  if (!c || c->size == 1) return DJ_OK;
  cudaDeviceSynchronize();
  for (auto& v : c->ipc_cache) {
    for (auto& e : v) cudaIpcCloseMemHandle(e.base);
    v.clear();
  }
  c->last_handle.clear();
  cudaGetLastError();
  int64_t one = 1;
  std::vector<int64_t> all(c->size);
  return ctrl_allgather(c, &one, 1, all.data());
}

extern "C" void* dj_comm_nccl_handle(dj_comm_t* c) { return c ? (void*)c->nccl : nullptr; }
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

// Control all-gather of host words: kernel-free peer path when available, NCCL otherwise.
static int ctrl_gather_host(dj_comm* c, int bank, const int64_t* h_mine, int n, int64_t* h_all)
{
  if (c->size > 1 && c->peer_ok && n <= kInbox) {
    int64_t* stage = c->h_pinned + (129 << 10) + (size_t)bank * 64;  // pinned copy: outlives the async copies
    DJ_REQUIRE(n <= 64, "host control message too long");
    memcpy(stage, h_mine, (size_t)n * 8);
    return peer_allgather(c, bank, stage, n, h_all);
  }
  return ctrl_allgather(c, h_mine, n, h_all);
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

// Which side the hash tables are built on.  The caller's LEFT table is the build side (the reference's
// drivers pass the unique-key build table as `left`, benchmark/distributed_join.cu:266-283) unless the
// right table is clearly smaller: received slices of equal-sized tables differ by a few rows per rank,
// and letting that noise pick the side made some ranks build on the duplicate-laden probe table
// (measured at N=2: 11.2 ms instead of 7.9 ms for the same join).
static inline bool build_on_right(int64_t nleft, int64_t nright) { return nright + nright / 8 < nleft; }
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
    opts->workspace_needed = 0;
    opts->t_exchange_ms[0] = opts->t_exchange_ms[1] = 0;
    opts->t_exchange_total_ms = 0;
  }
  Arena arena(d_workspace, workspace_bytes);
  int64_t* d_count = arena.take<int64_t>(32);
  DJ_REQUIRE(d_count, "distributed_inner_join: workspace too small");
  DJ_CUDA_TRY(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
  int64_t* out[4] = {d_out_lk, d_out_lp, d_out_rk, d_out_rp};
  auto t0         = std::chrono::high_resolution_clock::now();

  if (world == 1) {
    // src/distributed_join.cpp:186-199 -- one rank: the local join is the whole job.
    const bool swap = build_on_right(nleft, nright);
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
  DJ_REQUIRE(((uintptr_t)d_workspace & 255) == 0, "distributed_inner_join: the workspace must be 256-byte aligned");
  int rc = ensure_events(comm, 2 * odf);
  if (rc) return rc;
  Trace trace;
  trace.init(st);
  // NCCL fallback only: the persistent partition / join kernels leave a couple of SMs idle for the
  // whole call, because the control plane's tiny NCCL all-gathers are kernels too and a GPU
  // saturated by persistent CTAs makes each of them wait for a kernel boundary (measured: up to
  // 4 ms per collective).  The default control plane moves its messages with copy engines.
  struct ReserveGuard {
    explicit ReserveGuard(int dflt)
    {
      const char* e = getenv("DJ_SM_RESERVE");
      set_sm_reserve(e ? atoi(e) : dflt);
    }
    ~ReserveGuard() { set_sm_reserve(0); }
  } reserve_guard(comm->peer_ok ? 0 : 2);  // the peer-memory control plane launches no kernels

  // ---- 0. ONE control all-gather opens the call: table sizes (-> the radix plan every rank derives
  //         identically), workspace size, and the CUDA IPC identity of the workspace.  Everything a
  //         rank later needs to know about a peer's memory layout follows from these numbers and
  //         from the bucket-count matrix (step 2) by pure arithmetic -- no further agreement rounds.
  constexpr int kHello = 13;  // nleft, nright, workspace bytes, ipc ok, offset in allocation, handle[8]
  std::vector<int64_t> hello((size_t)world * kHello);
  {
    int64_t mine[kHello] = {nleft, nright, (int64_t)workspace_bytes, 0, 0};
    if (comm->peer_ok) {
      CUdeviceptr base = 0;
      size_t alloc_sz  = 0;
      cudaIpcMemHandle_t wh;
      memset(&wh, 0, sizeof(wh));
      const bool ok = comm->fn_addr_range(&base, &alloc_sz, (CUdeviceptr)d_workspace) == CUDA_SUCCESS &&
                      cudaIpcGetMemHandle(&wh, (void*)base) == cudaSuccess;
      cudaGetLastError();
      mine[3] = ok ? 1 : 0;
      mine[4] = ok ? (int64_t)((CUdeviceptr)d_workspace - base) : 0;
      memcpy(&mine[5], &wh, 64);
    }
    rc = ctrl_gather_host(comm, kBankHello, mine, kHello, hello.data());
    if (rc) return rc;
  }
  auto H = [&](int r, int f) -> int64_t { return hello[(size_t)r * kHello + f]; };
  int sub_bits = 0;
  RadixPlan plan{0, 0, 1};
  {
    int64_t tot[2] = {0, 0};
    for (int r = 0; r < world; r++) {
      DJ_REQUIRE(H(r, 0) < ((int64_t)1 << 31) && H(r, 1) < ((int64_t)1 << 31),
                 "distributed_inner_join: per-rank tables are limited to 2^31 rows (rank %d)", r);
      tot[0] += H(r, 0);
      tot[1] += H(r, 1);
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
  trace.host("hello gathered, plan agreed");

  const int F1s  = 1 << sub_bits;       // sub-buckets per destination in the sender's partition
  const int nbk  = nparts << sub_bits;  // buckets of the sender's partition
  const int nseg = G * F1s;             // (source, sub-bucket) segments of a received piece
  const size_t pw = pass_workspace_bytes(1, nbk);

  // Workspace layout of ANY rank, as offsets from its workspace base: a pure function of that
  // rank's table sizes and (for the receive pieces) of the rows it receives -- so every rank can
  // compute where its rows go inside every peer without asking.
  struct WsLayout {
    size_t count = 0, pws[2] = {0, 0}, prow[2] = {0, 0}, doff[2] = {0, 0}, dcnt[2] = {0, 0}, pbase[2] = {0, 0};
    std::vector<size_t> piece, seg_begin, seg_end, seg_parent;  // [odf*2]
    size_t join_mark = 0, need = 0;
  };
  auto off_of = [](const void* p) { return (size_t)reinterpret_cast<uintptr_t>(p); };
  auto lay_out = [&](int64_t nl, int64_t nr, const int64_t* spans /* [odf*2], nullptr: partition part only */) {
    WsLayout L;
    Arena va(nullptr, ~(size_t)0 >> 1);
    L.count = off_of(va.take<int64_t>(32));
    const int64_t n2[2] = {nl, nr};
    for (int t = 0; t < 2; t++) {
      L.pws[t]  = off_of(va.take<char>(pw));
      L.prow[t] = off_of(va.take<Row>((size_t)(n2[t] + (int64_t)nparts * kAlignRows)));
      L.doff[t] = off_of(va.take<int64_t>((size_t)nbk + 1));
      L.pbase[t] = off_of(va.take<Row*>((size_t)nparts));
    }
    L.dcnt[0] = off_of(va.take<int64_t>((size_t)2 * nbk + 2));  // both tables' counts, contiguous: one message
    L.dcnt[1] = L.dcnt[0] + (size_t)nbk * 8;
    L.need = va.used;
    if (!spans) return L;
    L.piece.resize((size_t)odf * 2);
    L.seg_begin.resize((size_t)odf * 2);
    L.seg_end.resize((size_t)odf * 2);
    L.seg_parent.resize((size_t)odf * 2);
    int64_t max_span[2] = {0, 0};
    for (int t = 0; t < 2; t++)
      for (int b = 0; b < odf; b++) {
        const size_t i  = (size_t)b * 2 + t;
        L.piece[i]      = off_of(va.take<Row>((size_t)spans[i] + 8));
        L.seg_begin[i]  = off_of(va.take<int64_t>((size_t)nseg));
        L.seg_end[i]    = off_of(va.take<int64_t>((size_t)nseg));
        L.seg_parent[i] = off_of(va.take<int>((size_t)nseg));
        max_span[t]     = std::max(max_span[t], spans[i]);
      }
    L.join_mark = va.used;
    L.need      = va.used + side_ws_bytes(max_span[0], plan, nseg) + side_ws_bytes(max_span[1], plan, nseg) + 4096;
    return L;
  };
  // every rank checks every rank's partition-stage fit: the verdict is identical everywhere
  for (int r = 0; r < world; r++) {
    const WsLayout L = lay_out(H(r, 0), H(r, 1), nullptr);
    if (L.need > (size_t)H(r, 2)) {
      if (opts) opts->workspace_needed = (int64_t)lay_out(nleft, nright, nullptr).need;
      set_error("distributed_inner_join: workspace too small on rank %d for the partitioned tables (%zu of %lld bytes)",
                r, L.need, (long long)H(r, 2));
      return DJ_ERR_WORKSPACE;
    }
  }

  // ---- 0b. copy-engine exchange: map every peer's workspace (cached per allocation; a peer whose
  //          allocation changed since the last call has its old mapping closed first)
  bool use_peer = comm->peer_ok && 2 * odf <= kDataSlots;
  for (int r = 0; r < world; r++) use_peer = use_peer && H(r, 3) == 1;
  std::vector<char*> peer_ws(world, nullptr);
  if (use_peer) {
    bool any_new = false;  // identical on every rank: everybody sees the same handles
    if (comm->last_handle.size() != (size_t)world * 8) {
      comm->last_handle.assign((size_t)world * 8, 0);
      any_new = true;
    }
    for (int r = 0; r < world; r++)
      if (memcmp(&comm->last_handle[(size_t)r * 8], &hello[(size_t)r * kHello + 5], 64) != 0) any_new = true;
    int64_t mapped = 1;
    for (int r = 0; r < world && mapped; r++) {
      if (r == rank) continue;
      cudaIpcMemHandle_t h;
      memcpy(&h, &hello[(size_t)r * kHello + 5], 64);
      peer_ws[r] = map_peer_workspace(comm, r, h, H(r, 4));
      if (!peer_ws[r]) mapped = 0;
    }
    if (any_new) {
      // a mapping was (re)opened somewhere: agree that it worked before anybody pushes
      std::vector<int64_t> oks(world);
      rc = ctrl_gather_host(comm, kBankAck, &mapped, 1, oks.data());
      if (rc) return rc;
      for (int r = 0; r < world; r++) use_peer = use_peer && oks[r] == 1;
      for (int r = 0; r < world; r++) memcpy(&comm->last_handle[(size_t)r * 8], &hello[(size_t)r * kHello + 5], 64);
    } else if (!mapped) {
      set_error("distributed_inner_join: a cached peer mapping disappeared");
      return DJ_ERR_CUDA;
    }
  }
  trace.host("peer workspaces mapped");
  const uint32_t seq = use_peer ? ++comm->seq : 0;
  // Exchange flavours over peer memory (identical decision on every rank: same environment):
  //   copy   (default) partition into a local table, then copy engines push each bucket to its peer;
  //   fused  the partition kernel's own cp.async.bulk stores write every run straight into the
  //          destination rank's receive piece -- partition and all-to-all are ONE kernel
  //          (src/distributed_join.cpp:211-225 + src/communicator.cpp:811-869 collapsed).
  bool fused = false;
  {
    const char* e = getenv("DJ_EXCHANGE");
    fused         = use_peer && e && (e[0] == 'f' || e[0] == 'F');
  }

  // ---- 1. hash partition (src/distributed_join.cpp:213-225) on the caller's stream, as histogram
  //         halves first: the counts of BOTH tables leave for the host while the scatter kernels run.
  //         Every destination's run of buckets starts on kAlignRows so pushes go straight from it.
  const WsLayout my = lay_out(nleft, nright, nullptr);
  char* wsb         = (char*)d_workspace;
  const int64_t n_in[2]    = {nleft, nright};
  const int64_t* in_key[2] = {d_left_key, d_right_key};
  const int64_t* in_pay[2] = {d_left_payload, d_right_payload};
  Row* prow[2];
  int64_t* d_cnt[2];
  PassState pstate[2];
  for (int t = 0; t < 2; t++) {
    prow[t]  = (Row*)(wsb + my.prow[t]);
    d_cnt[t] = (int64_t*)(wsb + my.dcnt[t]);
    PassDesc desc{sub_bits ? 2 : 0, kNvlinkSeed, DJ_HASH_MURMUR3, 0, nbk, 1, 1, kAlignRows, nparts, sub_bits};
    PassBuffers pb{};
    pb.in_key = in_key[t]; pb.in_pay[0] = in_pay[t]; pb.out_rows = prow[t];
    pb.nrows = n_in[t]; pb.d_child_off = (int64_t*)(wsb + my.doff[t]); pb.d_child_cnt = d_cnt[t];
    rc = pass_histogram(desc, pb, wsb + my.pws[t], pw, st, &pstate[t]);
    if (rc) return rc;
  }
  DJ_CUDA_TRY(cudaEventRecord(comm->ev_hist, st));
  trace.mark("histograms done", st);
  for (int t = 0; t < 2 && !fused; t++) {
    rc = pass_scatter(pstate[t], st);
    if (rc) return rc;
    DJ_CUDA_TRY(cudaEventRecord(comm->ev_part[t], st));
    trace.mark(t ? "partition(R) done" : "partition(L) done", st);
  }
  trace.host("partition launched");

  // ---- 2. sizes (communicate_sizes, src/all_to_all_comm.cpp:54-111): one all-gather of both
  //         tables' bucket counts, as soon as the histograms are done
  std::vector<int64_t> allc((size_t)world * 2 * nbk);
  DJ_CUDA_TRY(cudaStreamWaitEvent(comm->ctrl_stream, comm->ev_hist, 0));
  if (comm->peer_ok && 2 * nbk <= kInbox) {
    // device -> every peer's inbox, straight from the histogram's output: no SM, no host hop
    rc = peer_allgather(comm, kBankCounts, d_cnt[0], 2 * nbk, allc.data());
    if (rc) return rc;
  } else {
    int64_t* hp = comm->h_pinned + (256 << 10);  // D2H landing zone
    DJ_CUDA_TRY(cudaMemcpyAsync(hp, d_cnt[0], (size_t)2 * nbk * 8, cudaMemcpyDeviceToHost, comm->ctrl_stream));
    DJ_CUDA_TRY(cudaStreamSynchronize(comm->ctrl_stream));
    rc = ctrl_allgather(comm, hp, 2 * nbk, allc.data());
    if (rc) return rc;
  }
  trace.host("counts gathered");
  if (timing) {
    DJ_CUDA_TRY(cudaEventSynchronize(comm->ev_part[1]));
    opts->t_partition_ms = ms_since(t0);
    printf("Rank %d: Hash partition takes %.0fms\n", rank, opts->t_partition_ms);
  }
  auto tcomm = std::chrono::high_resolution_clock::now();

  // rows of source `src`'s table t in sub-bucket `sub` of destination bucket q
  auto cnt = [&](int src, int t, int q, int sub) {
    return allc[((size_t)src * 2 + t) * nbk + ((size_t)q << sub_bits) + sub];
  };
  auto sent = [&](int src, int t, int q) {  // rows source `src` sends for destination bucket q
    int64_t c = 0;
    for (int sub = 0; sub < F1s; sub++) c += cnt(src, t, q, sub);
    return c;
  };
  // where source s's rows start inside destination rank r's piece (batch b, table t), and its span
  auto piece_begin = [&](int r, int b, int t, int s) {
    int64_t at = 0;
    for (int s2 = 0; s2 < s; s2++) at += pad_rows(sent(s2, t, b * G + r));
    return at;
  };
  // my own send offsets: the aligned_offsets_kernel arithmetic restated on the host
  std::vector<int64_t> send_begin((size_t)2 * nparts);
  for (int t = 0; t < 2; t++) {
    int64_t at = 0;
    for (int q = 0; q < nparts; q++) {
      send_begin[(size_t)t * nparts + q] = at;
      at += pad_rows(sent(rank, t, q));
    }
  }

  // ---- 3. receive layout of every rank (allocate_communicated_table) + the fit verdict, locally
  std::vector<WsLayout> lay(world);
  std::vector<int64_t> spans((size_t)odf * 2);
  for (int r = 0; r < world; r++) {
    for (int b = 0; b < odf; b++)
      for (int t = 0; t < 2; t++) spans[(size_t)b * 2 + t] = piece_begin(r, b, t, G);
    lay[r] = lay_out(H(r, 0), H(r, 1), spans.data());
  }
  for (int r = 0; r < world; r++)
    if (lay[r].need > (size_t)H(r, 2)) {
      if (opts) opts->workspace_needed = (int64_t)lay[rank].need;
      set_error("distributed_inner_join: workspace too small on rank %d for its received partitions "
                "(needs %zu of %lld bytes; this rank needs %zu)", r, lay[r].need, (long long)H(r, 2), lay[rank].need);
      // nothing has been pushed yet and every rank takes this branch: just drain our own kernels
      cudaStreamSynchronize(st);
      return DJ_ERR_WORKSPACE;
    }
  struct Piece {
    std::vector<int64_t> begin, count;  // per source
    int64_t span = 0, rows = 0;
    Row* data = nullptr;
    int64_t *d_seg_begin = nullptr, *d_seg_end = nullptr;
    int* d_seg_parent = nullptr;
  };
  std::vector<Piece> pieces((size_t)odf * 2);
  int64_t* hseg = comm->h_pinned + (320 << 10);  // pinned staging for the segment tables
  DJ_REQUIRE(pieces.size() * 3 * (size_t)nseg <= (192u << 10), "distributed_inner_join: too many segments");
  for (int t = 0; t < 2; t++)
    for (int b = 0; b < odf; b++) {
      const size_t i = (size_t)b * 2 + t;
      Piece& pc      = pieces[i];
      pc.begin.resize(G);
      pc.count.resize(G);
      pc.data         = (Row*)(wsb + lay[rank].piece[i]);
      pc.d_seg_begin  = (int64_t*)(wsb + lay[rank].seg_begin[i]);
      pc.d_seg_end    = (int64_t*)(wsb + lay[rank].seg_end[i]);
      pc.d_seg_parent = (int*)(wsb + lay[rank].seg_parent[i]);
      int64_t* hb = hseg + i * 3 * nseg;
      int* hpar   = reinterpret_cast<int*>(hb + 2 * (size_t)nseg);
      for (int s = 0; s < G; s++) {
        pc.begin[s] = pc.span;
        pc.count[s] = sent(s, t, b * G + rank);
        int64_t at  = pc.span;
        for (int sub = 0; sub < F1s; sub++) {
          const int64_t c          = cnt(s, t, b * G + rank, sub);
          hb[s * F1s + sub]        = at;
          hb[nseg + s * F1s + sub] = at + c;
          hpar[s * F1s + sub]      = sub;
          at += c;
        }
        pc.span += pad_rows(pc.count[s]);
        pc.rows += pc.count[s];
      }
      DJ_CUDA_TRY(cudaMemcpyAsync(pc.d_seg_begin, hb, (size_t)nseg * 8, cudaMemcpyHostToDevice, comm->ctrl_stream));
      DJ_CUDA_TRY(cudaMemcpyAsync(pc.d_seg_end, hb + nseg, (size_t)nseg * 8, cudaMemcpyHostToDevice, comm->ctrl_stream));
      DJ_CUDA_TRY(cudaMemcpyAsync(pc.d_seg_parent, hpar, (size_t)nseg * 4, cudaMemcpyHostToDevice, comm->ctrl_stream));
    }
  DJ_CUDA_TRY(cudaEventRecord(comm->ev_seg[0], comm->ctrl_stream));
  trace.host("pieces laid out");

  // ---- 4. exchange (src/all_to_all_comm.cpp:126-189): every (batch, table) bucket run goes straight
  //         from the partitioned table into the destination's receive piece
  bool exchange_in_flight = false;
  const bool measure      = opts && opts->measure_exchange && use_peer;
  auto issue_exchange = [&](int b, int t) -> int {
    Piece& pc = pieces[(size_t)b * 2 + t];
    auto sbeg = [&](int dest) { return send_begin[(size_t)t * nparts + (size_t)b * G + dest]; };
    trace.mark(t ? "exchange(R) begin" : "exchange(L) begin", comm->comm_stream);
    // own bucket: device copy (src/all_to_all_comm.cpp:610-653); the rest over NVLink
    if (pc.count[rank] > 0)
      DJ_CUDA_TRY(cudaMemcpyAsync(pc.data + pc.begin[rank], prow[t] + sbeg(rank),
                                  (size_t)pc.count[rank] * sizeof(Row), cudaMemcpyDeviceToDevice, comm->comm_stream));
    if (use_peer) {
      // push every peer's bucket into ITS receive piece with the copy engines (no SMs, so the
      // radix passes running meanwhile keep the whole GPU), then raise that peer's flag
      const int slot = b * 2 + t;
      // destinations in rank+1, rank+2, ... order: at every moment the ranks push along a
      // permutation, so no receiver sees all senders at once (the copy engines work through the
      // peer streams roughly in issue order; starting everybody at rank 0 is an incast)
      for (int k = 1; k < G; k++) {
        const int i     = (rank + k) % G;
        cudaStream_t ps = comm->peer_stream[i];
        DJ_CUDA_TRY(cudaStreamWaitEvent(ps, comm->ev_part[t], 0));
        if (measure && b == 0) DJ_CUDA_TRY(cudaEventRecord(comm->ev_xbeg[(size_t)t * G + i], ps));
        const int64_t ns = sent(rank, t, b * G + i);
        if (ns > 0) {
          char* dst = peer_ws[i] + lay[i].piece[(size_t)b * 2 + t] + (size_t)piece_begin(i, b, t, rank) * sizeof(Row);
          DJ_CUDA_TRY(cudaMemcpyAsync(dst, prow[t] + sbeg(i), (size_t)ns * sizeof(Row), cudaMemcpyDefault, ps));
          if (opts) opts->bytes_sent += (int64_t)sizeof(Row) * ns;
        }
        if (measure && b == odf - 1) DJ_CUDA_TRY(cudaEventRecord(comm->ev_xend[(size_t)t * G + i], ps));
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
      const int64_t ns = sent(rank, t, b * G + i), nr = pc.count[i];
      if (ns > 0) {
        DJ_NCCL_TRY(ncclSend(prow[t] + sbeg(i), (size_t)ns * sizeof(Row), ncclInt8, i, comm->nccl, comm->comm_stream));
        if (opts) opts->bytes_sent += (int64_t)sizeof(Row) * ns;
      }
      if (nr > 0)
        DJ_NCCL_TRY(ncclRecv(pc.data + pc.begin[i], (size_t)nr * sizeof(Row), ncclInt8, i, comm->nccl,
                             comm->comm_stream));
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
  if (measure) {
    rc = ensure_xevents(comm, 2 * G);
    if (rc) return rc;
  }
  if (fused) {
    // ---- 4'. fused partition + exchange.  Per table: the bucket cursors restart inside every
    //          destination part (a part = this rank's slot in one peer's receive piece), the part
    //          bases point into the peers' workspaces, the scatter kernel runs, and a flag per
    //          (batch, table) tells every peer that this rank's rows have landed.
    int64_t* hcur = comm->h_pinned + (700 << 10);  // [2][nbk] cursors, then [2][nparts] bases
    int64_t* hbas = hcur + 2 * (size_t)nbk;
    for (int t = 0; t < 2; t++) {
      for (int q = 0; q < nparts; q++) {
        const int b = q / G, i = q % G;
        char* base_ws = i == rank ? wsb : peer_ws[i];
        hbas[(size_t)t * nparts + q] =
          (int64_t)(uintptr_t)(base_ws + lay[i].piece[(size_t)b * 2 + t] + (size_t)piece_begin(i, b, t, rank) * sizeof(Row));
        int64_t at = 0;
        for (int sub = 0; sub < F1s; sub++) {
          hcur[(size_t)t * nbk + ((size_t)q << sub_bits) + sub] = at;
          at += cnt(rank, t, q, sub);
        }
        if (i != rank && opts) opts->bytes_sent += (int64_t)sizeof(Row) * at;
      }
      Row** d_pbase = (Row**)(wsb + my.pbase[t]);
      DJ_CUDA_TRY(cudaMemcpyAsync(pstate[t].dev.cursor, hcur + (size_t)t * nbk, (size_t)nbk * 8, cudaMemcpyHostToDevice, st));
      DJ_CUDA_TRY(cudaMemcpyAsync(d_pbase, hbas + (size_t)t * nparts, (size_t)nparts * 8, cudaMemcpyHostToDevice, st));
      pstate[t].dev.part_base  = d_pbase;
      pstate[t].dev.part_shift = sub_bits;
      if (measure) DJ_CUDA_TRY(cudaEventRecord(comm->ev_xbeg[t], st));
      rc = pass_scatter(pstate[t], st);
      if (rc) return rc;
      if (measure) DJ_CUDA_TRY(cudaEventRecord(comm->ev_xend[t], st));
      DJ_CUDA_TRY(cudaEventRecord(comm->ev_part[t], st));
      trace.mark(t ? "partition+exchange(R) done" : "partition+exchange(L) done", st);
      // the flags leave on the communication stream so that the next scatter is not held up
      DJ_CUDA_TRY(cudaStreamWaitEvent(comm->comm_stream, comm->ev_part[t], 0));
      for (int b = 0; b < odf; b++) {
        for (int k = 1; k < G; k++) {
          const int i    = (rank + k) % G;
          uint32_t* flag = comm->peer_flags[i] + (size_t)rank * kFlagSlots + (b * 2 + t);
          if (!comm->flag_by_memcpy &&
              comm->fn_write32((CUstream)comm->comm_stream, (CUdeviceptr)flag, seq, 0) != CUDA_SUCCESS)
            comm->flag_by_memcpy = true;
          if (comm->flag_by_memcpy) {
            uint32_t* src = reinterpret_cast<uint32_t*>(comm->h_pinned + (900 << 10)) + (seq % 4096);
            *src          = seq;
            DJ_CUDA_TRY(cudaMemcpyAsync(flag, src, 4, cudaMemcpyDefault, comm->comm_stream));
          }
        }
        DJ_CUDA_TRY(cudaEventRecord(comm->ev_batch[(size_t)b * 2 + t], comm->comm_stream));
      }
    }
    exchange_in_flight = true;
  }
  // batch order (0,L),(0,R),(1,L),...: the left table's pushes start while the right table is
  // still being partitioned
  for (int b = 0; b < odf && !fused; b++)
    for (int t = 0; t < 2; t++) {
      if (b == 0) DJ_CUDA_TRY(cudaStreamWaitEvent(comm->comm_stream, comm->ev_part[t], 0));
      rc = issue_exchange(b, t);
      if (rc) return rc;
    }
  if (timing) {
    DJ_CUDA_TRY(cudaStreamSynchronize(comm->comm_stream));
    if (use_peer)
      for (int i = 0; i < G; i++)
        if (i != rank) DJ_CUDA_TRY(cudaStreamSynchronize(comm->peer_stream[i]));
    opts->t_comm_ms = ms_since(tcomm);
    for (int b = 0; b < odf; b++)
      printf("Rank %d: All-to-all communication on batch %d takes %.0fms\n", rank, b, opts->t_comm_ms / odf);
  }
  trace.host("exchanges issued");
  DJ_CUDA_TRY(cudaStreamWaitEvent(st, comm->ev_seg[0], 0));

  // ---- 5. local join per batch (src/distributed_join.cpp:283-322), appending into one output
  arena.used             = lay[rank].join_mark;
  const size_t join_mark = arena.used;
  for (int b = 0; b < odf; b++) {
    auto tj    = std::chrono::high_resolution_clock::now();
    Piece& L   = pieces[(size_t)b * 2];
    Piece& R   = pieces[(size_t)b * 2 + 1];
    arena.used = join_mark;  // join scratch is reused batch after batch (same stream)
    auto await_piece = [&](int t) -> int {
      DJ_CUDA_TRY(cudaStreamWaitEvent(st, comm->ev_batch[(size_t)b * 2 + t], 0));
      if (use_peer) {
        const int slot = b * 2 + t;
        for (int src = 0; src < G; src++) {
          if (src == rank) continue;
          int r2 = stream_wait_flag(comm, st, comm->d_flags + (size_t)src * kFlagSlots + slot, seq);
          if (r2) return r2;
        }
        trace.mark(t ? "arrived(R)" : "arrived(L)", st);
      }
      return DJ_OK;
    };
    if (L.rows == 0 || R.rows == 0) {  // src/distributed_join.cpp:76-82 (arrivals are still awaited)
      for (int t = 0; t < 2; t++)
        if ((rc = await_piece(t))) {
          drain_exchange();
          return rc;
        }
      continue;
    }
    const bool swap = build_on_right(L.rows, R.rows);
    PreparedSide side[2];
    for (int t = 0; t < 2; t++) {
      Piece& pc = t ? R : L;
      // each piece is awaited right before ITS radix pass: the left table's pass overlaps the
      // right table's exchange
      if ((rc = await_piece(t))) {
        drain_exchange();
        return rc;
      }
      TableInput in{nullptr, nullptr, pc.data, pc.span, pc.d_seg_begin, pc.d_seg_end, nseg, pc.d_seg_parent,
                    sub_bits > 0};
      trace.mark(t ? "radix(R) begin" : "radix(L) begin", st);
      rc = prepare_side(in, plan, &side[t], arena, st);
      if (rc) {
        drain_exchange();
        return rc;
      }
      trace.mark(t ? "radix(R) end" : "radix(L) end", st);
    }
    rc = join_prepared(side[swap ? 1 : 0], side[swap ? 0 : 1], plan, out, out_capacity, d_count, swap, st);
    if (rc) {
      drain_exchange();
      return rc;
    }
    trace.mark("join end", st);
    if (timing) {
      DJ_CUDA_TRY(cudaStreamSynchronize(st));
      double ms = ms_since(tj);
      opts->t_join_ms += ms;
      printf("Rank %d: Local join on batch %d takes %.0fms\n", rank, b, ms);
    }
  }
  trace.host("join launched");

  // ---- 6. the overflow verdict is collective (every rank returns DJ_ERR_OVERFLOW if any rank's
  //         output did not fit, so that callers can retry together) and costs no collective: a
  //         one-thread kernel stores this rank's verdict into every peer's flag block over NVLink,
  //         the stream waits for the peers' words, and ONE synchronisation returns count + verdicts.
  int64_t* h_res = comm->h_pinned + (128 << 10);
  if (use_peer) {
    trace.mark("verdict begin", st);
    verdict_kernel<<<1, 32, 0, st>>>((const unsigned long long*)d_count, (unsigned long long)out_capacity,
                                     comm->d_peer_flags, world, rank, kFlagSlots, kFlagSlots - 1, seq);
    DJ_LAUNCH_CHECK();
    for (int src = 0; src < G; src++) {
      if (src == rank) continue;
      rc = stream_wait_flag(comm, st, comm->d_flags + (size_t)src * kFlagSlots + (kFlagSlots - 1), seq << 1);
      if (rc) {
        drain_exchange();
        return rc;
      }
    }
    trace.mark("verdicts received", st);
    DJ_CUDA_TRY(cudaMemcpy2DAsync(h_res + 1, 4, comm->d_flags + (kFlagSlots - 1), (size_t)kFlagSlots * 4, 4, world,
                                  cudaMemcpyDeviceToHost, st));
  }
  DJ_CUDA_TRY(cudaMemcpyAsync(h_res, d_count, 8, cudaMemcpyDeviceToHost, st));
  DJ_CUDA_TRY(cudaStreamSynchronize(st));
  trace.host("main stream drained");
  DJ_CUDA_TRY(cudaStreamSynchronize(comm->comm_stream));
  trace.host("streams drained");
  if (use_peer)
    for (int i = 0; i < G; i++)
      if (i != rank) DJ_CUDA_TRY(cudaStreamSynchronize(comm->peer_stream[i]));  // my buckets may be reused now
  *h_out_count = h_res[0];
  if (measure) {
    // per-direction NVLink throughput of this rank's pushes.  The copy engines work through the
    // peer streams one copy after the other, and an event recorded on a waiting stream is
    // timestamped when its copy starts -- so a table's window is taken from the FIRST stream's
    // begin event to the LATEST end event over all streams, not per stream.
    const int first = (rank + 1) % G;
    float latest_all = 0;
    for (int t = 0; t < 2; t++) {
      float worst = 0;
      if (fused) {  // the scatter kernel IS the exchange
        if (cudaEventElapsedTime(&worst, comm->ev_xbeg[t], comm->ev_xend[t]) != cudaSuccess) worst = 0;
        opts->t_exchange_ms[t] = worst;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, comm->ev_xbeg[0], comm->ev_xend[t]) == cudaSuccess) latest_all = std::max(latest_all, ms);
        continue;
      }
      for (int i = 0; i < G; i++) {
        if (i == rank) continue;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, comm->ev_xbeg[(size_t)t * G + first], comm->ev_xend[(size_t)t * G + i]) == cudaSuccess)
          worst = std::max(worst, ms);
        if (cudaEventElapsedTime(&ms, comm->ev_xbeg[(size_t)first], comm->ev_xend[(size_t)t * G + i]) == cudaSuccess)
          latest_all = std::max(latest_all, ms);
      }
      opts->t_exchange_ms[t] = worst;
    }
    opts->t_exchange_total_ms = latest_all;  // first push of the left table -> last push of the right table
    cudaGetLastError();
  }
  trace.dump(rank);
  int over_rank = -1;
  if (use_peer) {
    const uint32_t* hv = reinterpret_cast<const uint32_t*>(h_res + 1);
    for (int r = 0; r < world; r++)
      if (r != rank && (hv[r] & 1u)) over_rank = r;
    if (*h_out_count > out_capacity) over_rank = rank;
  } else {
    int64_t over = *h_out_count > out_capacity ? 1 : 0;
    std::vector<int64_t> overs(world);
    rc = ctrl_allgather(comm, &over, 1, overs.data());
    if (rc) return rc;
    for (int r = 0; r < world; r++)
      if (overs[r]) over_rank = r;
  }
  if (over_rank >= 0) {
    set_error("join output does not fit on rank %d (this rank: %lld rows, capacity %lld)", over_rank,
              (long long)*h_out_count, (long long)out_capacity);
    return DJ_ERR_OVERFLOW;
  }
  return DJ_OK;
}

static size_t streamed_ws_bytes(int64_t nleft, int64_t nright, int64_t out_capacity);

extern "C" size_t dj_distributed_inner_join_host_workspace_bytes(int64_t nleft, int64_t nright,
                                                                 int64_t out_capacity, int world,
                                                                 int over_decom_factor)
{
  const size_t staged = dj_distributed_inner_join_workspace_bytes(nleft, nright, world, over_decom_factor) +
                        2 * (align_up((size_t)nleft * 8, 256) + align_up((size_t)nright * 8, 256)) +
                        4 * align_up((size_t)out_capacity * 8, 256) + 8192;
  if (world > 1) return staged;
  return std::max(staged, streamed_ws_bytes(nleft, nright, out_capacity));
}

// Single-GPU end-to-end join with HOST tables, streamed: the PCIe link is the bottleneck (25.6 GB in,
// 7.7 GB out at 800M x 800M against ~50 ms of GPU work), so the call is organised around keeping
// both directions of the link busy:
//   1. the build table goes up first and is radix-partitioned while the probe table's first chunks
//      are already on the wire;
//   2. the probe table goes up in chunks (double-buffered); each chunk is partitioned with the same
//      radix plan and joined against the resident build buckets as soon as it has landed -- the GPU
//      re-inserts the build rows once per chunk, which costs HBM bandwidth that is idle anyway;
//   3. each chunk's matches go down on their own stream while the next chunk comes up (full duplex).
// What is left after the last byte has arrived is one chunk's join and one chunk's matches.
struct StreamedShape {
  bool swap;
  int64_t nb, np, chunk;
  int nchunks;
  RadixPlan plan;
};
static StreamedShape streamed_shape(int64_t nleft, int64_t nright)
{
  StreamedShape s{};
  s.swap      = build_on_right(nleft, nright);
  s.nb        = s.swap ? nright : nleft;
  s.np        = s.swap ? nleft : nright;
  s.plan      = plan_for(s.nb > 0 ? s.nb : 1, false);
  int nchunks = 16;
  const char* e = getenv("DJ_HOST_CHUNKS");
  if (e && atoi(e) > 0) nchunks = atoi(e);
  int64_t chunk = (s.np + nchunks - 1) / nchunks;
  if (chunk < (1 << 20)) chunk = std::min<int64_t>(s.np, 1 << 20);  // small tables: few chunks
  chunk     = std::max<int64_t>((chunk + 1) / 2 * 2, 2);  // even row counts keep the host columns 16-byte aligned
  s.chunk   = chunk;
  s.nchunks = (int)std::max<int64_t>((s.np + chunk - 1) / chunk, 1);
  return s;
}
// device bytes host_join_streamed takes from the workspace (same arithmetic as its arena walk)
static size_t streamed_ws_bytes(int64_t nleft, int64_t nright, int64_t out_capacity)
{
  const StreamedShape s = streamed_shape(nleft, nright);
  size_t total = 256 + 2 * align_up((size_t)s.nb * 8, 256);
  total += (s.nchunks > 1 ? 4 : 2) * align_up((size_t)s.chunk * 8, 256);
  total += 4 * align_up((size_t)out_capacity * 8, 256);
  total += side_ws_bytes(s.nb, s.plan, 0) + side_ws_bytes(s.chunk, s.plan, 0);
  return total + (64 << 10);
}

static int host_join_streamed(const int64_t* h_left_key, const int64_t* h_left_payload, int64_t nleft,
                              const int64_t* h_right_key, const int64_t* h_right_payload, int64_t nright,
                              int64_t* const h_out[4], int64_t out_capacity, int64_t* h_out_count,
                              dj_join_options* opts, void* d_workspace, size_t workspace_bytes, cudaStream_t st)
{
  *h_out_count = 0;
  if (opts) {
    opts->t_partition_ms = opts->t_comm_ms = opts->t_join_ms = 0;
    opts->bytes_sent = opts->workspace_needed = 0;
  }
  if (nleft == 0 || nright == 0) return DJ_OK;  // src/distributed_join.cpp:76-82
  const StreamedShape shape = streamed_shape(nleft, nright);
  const bool swap      = shape.swap;
  const int64_t nb     = shape.nb, np = shape.np, chunk = shape.chunk;
  const int nchunks    = shape.nchunks;
  const int64_t* h_bk  = swap ? h_right_key : h_left_key;
  const int64_t* h_bp  = swap ? h_right_payload : h_left_payload;
  const int64_t* h_pk  = swap ? h_left_key : h_right_key;
  const int64_t* h_pp  = swap ? h_left_payload : h_right_payload;
  const RadixPlan plan = shape.plan;
  int64_t* h_cnt = nullptr;  // pinned: running match count after every chunk (allocated before any work is queued)
  DJ_CUDA_TRY(cudaMallocHost(&h_cnt, (size_t)(nchunks + 1) * 8));
  struct PinGuard {
    int64_t* p;
    ~PinGuard() { cudaFreeHost(p); }
  } pin_guard{h_cnt};

  Arena arena(d_workspace, workspace_bytes);
  int64_t* d_count = arena.take<int64_t>(32);
  int64_t* dbk     = arena.take<int64_t>((size_t)nb);
  int64_t* dbp     = arena.take<int64_t>((size_t)nb);
  int64_t* dck[2], *dcp[2];
  for (int i = 0; i < 2; i++) {  // one buffer is enough for a single chunk
    dck[i] = (i == 0 || nchunks > 1) ? arena.take<int64_t>((size_t)chunk) : dck[0];
    dcp[i] = (i == 0 || nchunks > 1) ? arena.take<int64_t>((size_t)chunk) : dcp[0];
  }
  int64_t* o[4];
  for (int c = 0; c < 4; c++) o[c] = arena.take<int64_t>((size_t)out_capacity);
  if (!d_count || !dbk || !dbp || !dck[1] || !dcp[1] || !o[3]) {
    set_error("distributed_inner_join_host: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
              streamed_ws_bytes(nleft, nright, out_capacity));
    if (opts) opts->workspace_needed = (int64_t)streamed_ws_bytes(nleft, nright, out_capacity);
    return DJ_ERR_WORKSPACE;
  }
  cudaStream_t up = nullptr, down = nullptr;
  std::vector<cudaEvent_t> ev;  // [0] build up, then per chunk: landed, partitioned, joined
  auto cleanup = [&]() {
    for (auto e : ev) cudaEventDestroy(e);
    if (up) cudaStreamDestroy(up);
    if (down) cudaStreamDestroy(down);
  };
  struct Guard {
    decltype(cleanup)& f;
    ~Guard() { f(); }
  } guard{cleanup};
  DJ_CUDA_TRY(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
  DJ_CUDA_TRY(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
  ev.resize(2 + (size_t)nchunks * 3, nullptr);
  for (auto& e : ev) DJ_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  cudaEvent_t ev_start = ev[1];
  auto ev_landed = [&](int c) { return ev[2 + (size_t)c * 3]; };
  auto ev_parted = [&](int c) { return ev[3 + (size_t)c * 3]; };
  auto ev_joined = [&](int c) { return ev[4 + (size_t)c * 3]; };

  // the caller's stream orders the call: uploads start after whatever it had queued
  DJ_CUDA_TRY(cudaEventRecord(ev_start, st));
  DJ_CUDA_TRY(cudaStreamWaitEvent(up, ev_start, 0));
  DJ_CUDA_TRY(cudaStreamWaitEvent(down, ev_start, 0));
  DJ_CUDA_TRY(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
  DJ_CUDA_TRY(cudaMemcpyAsync(dbk, h_bk, (size_t)nb * 8, cudaMemcpyHostToDevice, up));
  DJ_CUDA_TRY(cudaMemcpyAsync(dbp, h_bp, (size_t)nb * 8, cudaMemcpyHostToDevice, up));
  DJ_CUDA_TRY(cudaEventRecord(ev[0], up));

  // build side: partitioned once, resident for the whole call
  PreparedSide build{}, probe{};
  DJ_CUDA_TRY(cudaStreamWaitEvent(st, ev[0], 0));
  TableInput tb{dbk, dbp, nullptr, nb, nullptr, nullptr, 0};
  int rc = prepare_side(tb, plan, &build, arena, st);
  if (rc) return rc;
  const size_t chunk_mark = arena.used;
  int64_t* out4[4] = {o[0], o[1], o[2], o[3]};
  int64_t done_rows = 0;  // output rows already on their way to the host
  auto drain = [&](int c) -> int {  // chunk c's matches -> host, on the download stream
    DJ_CUDA_TRY(cudaEventSynchronize(ev_joined(c)));
    int64_t upto = h_cnt[c] < out_capacity ? h_cnt[c] : out_capacity;
    if (upto > done_rows) {
      DJ_CUDA_TRY(cudaStreamWaitEvent(down, ev_joined(c), 0));
      for (int col = 0; col < 4; col++)
        DJ_CUDA_TRY(cudaMemcpyAsync(h_out[col] + done_rows, o[col] + done_rows, (size_t)(upto - done_rows) * 8,
                                    cudaMemcpyDeviceToHost, down));
      done_rows = upto;
    }
    return DJ_OK;
  };
  for (int c = 0; c < nchunks; c++) {
    const int64_t r0 = (int64_t)c * chunk, n = std::min(chunk, np - r0);
    const int bi     = c & 1;
    if (c >= 2) DJ_CUDA_TRY(cudaStreamWaitEvent(up, ev_parted(c - 2), 0));  // the buffer's previous chunk is consumed
    DJ_CUDA_TRY(cudaMemcpyAsync(dck[bi], h_pk + r0, (size_t)n * 8, cudaMemcpyHostToDevice, up));
    DJ_CUDA_TRY(cudaMemcpyAsync(dcp[bi], h_pp + r0, (size_t)n * 8, cudaMemcpyHostToDevice, up));
    DJ_CUDA_TRY(cudaEventRecord(ev_landed(c), up));

    DJ_CUDA_TRY(cudaStreamWaitEvent(st, ev_landed(c), 0));
    arena.used = chunk_mark;  // probe scratch is reused chunk after chunk (same stream)
    TableInput tp{dck[bi], dcp[bi], nullptr, n, nullptr, nullptr, 0};
    rc = prepare_side(tp, plan, &probe, arena, st);
    if (rc) return rc;
    DJ_CUDA_TRY(cudaEventRecord(ev_parted(c), st));
    rc = join_prepared(build, probe, plan, out4, out_capacity, d_count, swap, st);
    if (rc) return rc;
    DJ_CUDA_TRY(cudaMemcpyAsync(h_cnt + c, d_count, 8, cudaMemcpyDeviceToHost, st));
    DJ_CUDA_TRY(cudaEventRecord(ev_joined(c), st));
    // the previous chunk's matches go down while this chunk is being joined and the next comes up
    if (c >= 1 && (rc = drain(c - 1))) return rc;
  }
  if ((rc = drain(nchunks - 1))) return rc;
  DJ_CUDA_TRY(cudaStreamSynchronize(down));
  DJ_CUDA_TRY(cudaStreamSynchronize(st));
  *h_out_count = h_cnt[nchunks - 1];
  if (*h_out_count > out_capacity) {
    set_error("join output needs %lld rows, capacity %lld", (long long)*h_out_count, (long long)out_capacity);
    return DJ_ERR_OVERFLOW;
  }
  return DJ_OK;
}

extern "C" int dj_distributed_inner_join_i64_host(
  dj_comm_t* comm, const int64_t* h_left_key, const int64_t* h_left_payload, int64_t nleft,
  const int64_t* h_right_key, const int64_t* h_right_payload, int64_t nright, int64_t* h_out_lk,
  int64_t* h_out_lp, int64_t* h_out_rk, int64_t* h_out_rp, int64_t out_capacity,
  int64_t* h_out_count, dj_join_options* opts, void* d_workspace, size_t workspace_bytes,
  void* stream)
{
  DJ_REQUIRE(d_workspace && h_out_count && nleft >= 0 && nright >= 0, "distributed_inner_join_host: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t* h[4]   = {h_out_lk, h_out_lp, h_out_rk, h_out_rp};
  if (!comm || comm->size == 1)
    return host_join_streamed(h_left_key, h_left_payload, nleft, h_right_key, h_right_payload, nright, h,
                              out_capacity, h_out_count, opts, d_workspace, workspace_bytes, st);
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
  for (int c = 0; c < 4; c++)
    DJ_CUDA_TRY(cudaMemcpyAsync(h[c], o[c], (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  DJ_CUDA_TRY(cudaStreamSynchronize(st));
  return rc;
}
