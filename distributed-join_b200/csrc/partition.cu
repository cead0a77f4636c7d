// partition.cu -- radix / hash partition passes for sm_100a.
//
// One "pass" splits every parent bucket (a contiguous row range of SoA int64 columns) into
// F child buckets, in three stream-ordered launches and no host synchronisation:
//
//   plan_kernel    tile tables: which (parent, chunk) every tile index maps to
//   hist_kernel    reads keys only (8 B/row): exact child-bucket sizes -> exclusive scan
//   scatter_kernel reads every column once, writes it once (algorithmic 16 B * (1+npay)/2
//                  per row in each direction): each CTA stages a tile of rows in shared
//                  memory sorted by child bucket, reserves the tile's slice of every child
//                  bucket with one global atomicAdd per non-empty bucket, and streams the
//                  sorted tile out so that a warp's stores land in contiguous runs.
//
// It replaces cudf::hash_partition (reference call sites src/distributed_join.cpp:213-225,
// src/shuffle_on.cpp:59-60) in mode 0, and is the join's private sub-partitioner in mode 1.
#include <cub/device/device_scan.cuh>

#include "dj_device.cuh"
#include "dj_internal.h"

namespace dj {

namespace {

constexpr int kHistThreads    = 512;
constexpr int kHistTileRows   = 32768;
constexpr int kScatterThreads = 512;
constexpr int kRowsPerThread  = 8;
constexpr int kScatterTile    = kScatterThreads * kRowsPerThread;  // 4096 rows

struct PassDev {
  const int64_t* in_key;
  const int64_t* in_pay[kMaxPayload];
  int64_t* out_key;
  int64_t* out_pay[kMaxPayload];
  const int64_t* parent_off;  // [P+1]
  unsigned long long* counts;  // [P*F+1]
  unsigned long long* cursor;  // [P*F]
  const int* hist_tiles;       // [P+1] prefix of hist tiles per parent
  const int* scat_tiles;       // [P+1] prefix of scatter tiles per parent
  int P, F;
  uint32_t seed;
  int hash_id, shift, pow2;
};

template <int MODE>
__device__ __forceinline__ int bucket_of(int64_t key, const PassDev& d)
{
  if (MODE == 0) {
    uint32_t h = row_hash_i64(key, d.seed, d.hash_id);
    return d.pow2 ? (int)(h & (uint32_t)(d.F - 1)) : (int)(h % (uint32_t)d.F);
  } else {
    return (int)((local_hash_i64(key) >> d.shift) & (uint32_t)(d.F - 1));
  }
}

// largest p in [0, P) with prefix[p] <= t
__device__ __forceinline__ int find_parent(const int* prefix, int P, int t)
{
  int lo = 0, hi = P;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (prefix[mid] <= t) lo = mid; else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------- plan
// One CTA: parent offsets (or {0, n}) -> tile prefix tables for the two tile sizes.
__global__ void plan_kernel(const int64_t* parent_off_in, int64_t nrows, int P,
                            int64_t* parent_off_out, int* hist_tiles, int* scat_tiles)
{
  __shared__ int warp_sums[33];
  const int tid = threadIdx.x;
  int64_t lo = 0, hi = 0;
  if (tid < P) {
    lo = parent_off_in ? parent_off_in[tid] : 0;
    hi = parent_off_in ? parent_off_in[tid + 1] : nrows;
    if (!parent_off_in) {
      parent_off_out[0] = 0;
      parent_off_out[1] = nrows;
    }
  }
  int64_t n = hi - lo;
  int ht    = (int)((n + kHistTileRows - 1) / kHistTileRows);
  int st    = (int)((n + kScatterTile - 1) / kScatterTile);
  int he = block_exclusive_scan<1024>(ht, warp_sums);
  if (tid < P) hist_tiles[tid] = he;
  if (tid == 0) hist_tiles[P] = warp_sums[32];
  __syncthreads();
  int se = block_exclusive_scan<1024>(st, warp_sums);
  if (tid < P) scat_tiles[tid] = se;
  if (tid == 0) scat_tiles[P] = warp_sums[32];
}

// ---------------------------------------------------------------- histogram
template <int MODE>
__global__ void __launch_bounds__(kHistThreads) hist_kernel(PassDev d)
{
  extern __shared__ int s_hist[];
  __shared__ int s_parent;
  const int tid   = threadIdx.x;
  const int total = d.hist_tiles[d.P];
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    for (int i = tid; i < d.F; i += kHistThreads) s_hist[i] = 0;
    if (tid == 0) s_parent = find_parent(d.hist_tiles, d.P, t);
    __syncthreads();
    const int p       = s_parent;
    const int64_t beg = d.parent_off[p] + (int64_t)(t - d.hist_tiles[p]) * kHistTileRows;
    int64_t end       = beg + kHistTileRows;
    if (end > d.parent_off[p + 1]) end = d.parent_off[p + 1];
#pragma unroll 8
    for (int64_t i = beg + tid; i < end; i += kHistThreads)
      atomicAdd(&s_hist[bucket_of<MODE>(d.in_key[i], d)], 1);
    __syncthreads();
    for (int i = tid; i < d.F; i += kHistThreads) {
      int c = s_hist[i];
      if (c) atomicAdd(&d.counts[(size_t)p * d.F + i], (unsigned long long)c);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- scatter
// Dynamic shared memory layout (T = kScatterTile):
//   int64 skey[T]; int64 spay[NPAY][T]; int64 s_delta[F]; int s_start[F]; uint16 sbkt[T]
template <int MODE, int NPAY, bool WARP_AGG>
__global__ void __launch_bounds__(kScatterThreads, 2) scatter_kernel(PassDev d)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int T = kScatterTile;
  int64_t* skey   = reinterpret_cast<int64_t*>(smem_raw);
  int64_t* spay   = skey + T;
  int64_t* s_delta = spay + (size_t)NPAY * T;
  int* s_start     = reinterpret_cast<int*>(s_delta + d.F);
  uint16_t* sbkt   = reinterpret_cast<uint16_t*>(s_start + d.F);
  __shared__ int warp_sums[33];
  __shared__ int s_parent;

  const int tid   = threadIdx.x;
  const int lane  = tid & 31;
  const int F     = d.F;
  const int total = d.scat_tiles[d.P];
  const int bpt   = (F + kScatterThreads - 1) / kScatterThreads;

  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    for (int i = tid; i < F; i += kScatterThreads) s_start[i] = 0;
    if (tid == 0) s_parent = find_parent(d.scat_tiles, d.P, t);
    __syncthreads();
    const int p       = s_parent;
    const int64_t beg = d.parent_off[p] + (int64_t)(t - d.scat_tiles[p]) * T;
    int64_t end       = beg + T;
    if (end > d.parent_off[p + 1]) end = d.parent_off[p + 1];
    const int tile_n = (int)(end - beg);

    // phase 1: load keys, bucket, rank inside the tile's bucket
    int64_t key[kRowsPerThread];
    uint32_t brank[kRowsPerThread];
#pragma unroll
    for (int j = 0; j < kRowsPerThread; j++) {
      const int r = j * kScatterThreads + tid;
      key[j]      = r < tile_n ? d.in_key[beg + r] : 0;
    }
#pragma unroll
    for (int j = 0; j < kRowsPerThread; j++) {
      const int r      = j * kScatterThreads + tid;
      const bool valid = r < tile_n;
      const int b      = valid ? bucket_of<MODE>(key[j], d) : 0;
      int rank;
      if (WARP_AGG) {
        // few buckets: one shared-memory atomic per (warp, bucket) instead of per row
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        rank                 = 0;
        if (valid) {
          const unsigned peers = __match_any_sync(vmask, b);
          const int leader     = __ffs(peers) - 1;
          int base             = 0;
          if (lane == leader) base = atomicAdd(&s_start[b], __popc(peers));
          base = __shfl_sync(peers, base, leader);
          rank = base + __popc(peers & lanemask_lt());
        }
      } else {
        rank = valid ? atomicAdd(&s_start[b], 1) : 0;
      }
      brank[j] = ((uint32_t)b << 16) | (uint32_t)rank;
    }
    __syncthreads();

    // phase 2: exclusive scan of the tile histogram; reserve the tile's slice of each bucket
    {
      const int b0 = tid * bpt;
      int sum      = 0;
      for (int k = 0; k < bpt; k++)
        if (b0 + k < F) sum += s_start[b0 + k];
      int run = block_exclusive_scan<kScatterThreads>(sum, warp_sums);
      for (int k = 0; k < bpt; k++) {
        if (b0 + k < F) {
          const int c     = s_start[b0 + k];
          s_start[b0 + k] = run;
          if (c) {
            unsigned long long g =
              atomicAdd(&d.cursor[(size_t)p * F + b0 + k], (unsigned long long)c);
            s_delta[b0 + k] = (int64_t)g - run;
          }
          run += c;
        }
      }
    }
    __syncthreads();

    // phase 3: rows into bucket-sorted order in shared memory (payloads straight from HBM)
#pragma unroll
    for (int j = 0; j < kRowsPerThread; j++) {
      const int r = j * kScatterThreads + tid;
      if (r < tile_n) {
        const int b   = brank[j] >> 16;
        const int pos = s_start[b] + (int)(brank[j] & 0xffffu);
        skey[pos]     = key[j];
        sbkt[pos]     = (uint16_t)b;
#pragma unroll
        for (int c = 0; c < NPAY; c++) spay[(size_t)c * T + pos] = d.in_pay[c][beg + r];
      }
    }
    __syncthreads();

    // phase 4: stream the sorted tile out; consecutive threads hit consecutive addresses
    for (int i = tid; i < tile_n; i += kScatterThreads) {
      const int64_t dst = s_delta[sbkt[i]] + i;
      d.out_key[dst]    = skey[i];
#pragma unroll
      for (int c = 0; c < NPAY; c++) d.out_pay[c][dst] = spay[(size_t)c * T + i];
    }
    __syncthreads();
  }
}

size_t scatter_smem_bytes(int npay, int F)
{
  return (size_t)kScatterTile * 8 * (1 + npay) + (size_t)F * (8 + 4) + (size_t)kScatterTile * 2;
}

template <int MODE, int NPAY, bool AGG>
int launch_scatter(const PassDev& dev, int F, cudaStream_t stream)
{
  const size_t smem = scatter_smem_bytes(NPAY, F);
  auto kern         = scatter_kernel<MODE, NPAY, AGG>;
  DJ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  DJ_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kScatterThreads, smem));
  if (per_sm < 1) per_sm = 1;
  {
    ProfScope prof(DJ_PROF_SCATTER, stream);
    kern<<<sm_count() * per_sm, kScatterThreads, smem, stream>>>(dev);
  }
  DJ_LAUNCH_CHECK();
  return DJ_OK;
}

template <int MODE, int NPAY>
int launch_scatter_agg(const PassDev& dev, int F, cudaStream_t stream)
{
  return F <= 32 ? launch_scatter<MODE, NPAY, true>(dev, F, stream)
                 : launch_scatter<MODE, NPAY, false>(dev, F, stream);
}

template <int MODE>
int launch_scatter_npay(const PassDev& dev, int npay, int F, cudaStream_t stream)
{
  switch (npay) {
    case 1: return launch_scatter_agg<MODE, 1>(dev, F, stream);
    case 2: return launch_scatter_agg<MODE, 2>(dev, F, stream);
    case 3: return launch_scatter_agg<MODE, 3>(dev, F, stream);
  }
  set_error("partition: unsupported payload column count %d", npay);
  return DJ_ERR_ARG;
}

}  // namespace

// workspace: counts[P*F+1] | cursor[P*F] | parent_off[2] | hist_tiles[P+1] | scat_tiles[P+1] | cub temp
static size_t cub_scan_temp_bytes(size_t n)
{
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (unsigned long long*)nullptr,
                                (unsigned long long*)nullptr, (int)n);
  return bytes;
}

size_t pass_workspace_bytes(int P, int F)
{
  const size_t nb = (size_t)P * F;
  size_t total    = 0;
  total += align_up((nb + 1) * 8, 256);
  total += align_up(nb * 8, 256);
  total += 256;
  total += 2 * align_up(((size_t)P + 1) * 4, 256);
  total += align_up(cub_scan_temp_bytes(nb + 1), 256);
  return total + 1024;
}

int run_partition_pass(const PassDesc& desc, const PassBuffers& buf, void* d_ws, size_t ws_bytes,
                       cudaStream_t stream)
{
  DJ_REQUIRE(desc.F >= 1 && desc.F <= kMaxFanout, "partition: fan-out %d out of range", desc.F);
  DJ_REQUIRE(desc.P >= 1 && desc.P <= kMaxFanout, "partition: parent count %d out of range", desc.P);
  DJ_REQUIRE(desc.mode == 0 || (desc.F & (desc.F - 1)) == 0, "radix fan-out must be a power of 2");
  DJ_REQUIRE(desc.P == 1 || buf.d_parent_off != nullptr, "partition: parent offsets missing");
  const size_t nb = (size_t)desc.P * desc.F;
  Arena arena(d_ws, ws_bytes);
  auto* counts     = arena.take<unsigned long long>(nb + 1);
  auto* cursor     = arena.take<unsigned long long>(nb);
  auto* parent_off = arena.take<int64_t>(2);
  auto* hist_tiles = arena.take<int>(desc.P + 1);
  auto* scat_tiles = arena.take<int>(desc.P + 1);
  size_t cub_bytes = cub_scan_temp_bytes(nb + 1);
  auto* cub_temp   = arena.take<char>(cub_bytes);
  if (!counts || !cursor || !parent_off || !hist_tiles || !scat_tiles || !cub_temp) {
    set_error("partition pass: workspace too small (%zu bytes given)", ws_bytes);
    return DJ_ERR_WORKSPACE;
  }

  DJ_CUDA_TRY(cudaMemsetAsync(counts, 0, (nb + 1) * 8, stream));
  plan_kernel<<<1, 1024, 0, stream>>>(buf.d_parent_off, buf.nrows, desc.P, parent_off, hist_tiles,
                                      scat_tiles);
  DJ_LAUNCH_CHECK();

  PassDev dev{};
  dev.in_key  = buf.in_key;
  dev.out_key = buf.out_key;
  for (int c = 0; c < desc.npay; c++) {
    dev.in_pay[c]  = buf.in_pay[c];
    dev.out_pay[c] = buf.out_pay[c];
  }
  dev.parent_off = buf.d_parent_off ? buf.d_parent_off : parent_off;
  dev.counts     = counts;
  dev.cursor     = cursor;
  dev.hist_tiles = hist_tiles;
  dev.scat_tiles = scat_tiles;
  dev.P          = desc.P;
  dev.F          = desc.F;
  dev.seed       = desc.seed;
  dev.hash_id    = desc.hash_id;
  dev.shift      = desc.shift;
  dev.pow2       = (desc.F & (desc.F - 1)) == 0;

  const int hist_grid = sm_count() * 4;
  const size_t hsmem  = (size_t)desc.F * sizeof(int);
  {
    ProfScope prof(DJ_PROF_HIST, stream);
    if (desc.mode == 0)
      hist_kernel<0><<<hist_grid, kHistThreads, hsmem, stream>>>(dev);
    else
      hist_kernel<1><<<hist_grid, kHistThreads, hsmem, stream>>>(dev);
  }
  DJ_LAUNCH_CHECK();

  {
    ProfScope prof(DJ_PROF_OTHER, stream);
    DJ_CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_temp, cub_bytes, counts,
                                              (unsigned long long*)buf.d_child_off, (int)(nb + 1),
                                              stream));
    count_launch(2);
    DJ_CUDA_TRY(cudaMemcpyAsync(cursor, buf.d_child_off, nb * 8, cudaMemcpyDeviceToDevice, stream));
  }

  return desc.mode == 0 ? launch_scatter_npay<0>(dev, desc.npay, desc.F, stream)
                        : launch_scatter_npay<1>(dev, desc.npay, desc.F, stream);
}

}  // namespace dj
