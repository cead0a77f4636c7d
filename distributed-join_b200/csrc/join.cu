// join.cu -- shared-memory open-addressing hash build + probe for sm_100a.
//
// Replaces cudf::inner_join as called by local_join_helper (src/distributed_join.cpp:71-83).
// Both tables arrive radix-partitioned (partition.cu, mode 1) into buckets whose build side
// fits one CTA's shared-memory table.  Per bucket a CTA
//   1. inserts the build rows into a linear-probing table (64-bit key + 64-bit payload per
//      slot, occupancy bitmap claimed with atomicOr so no key value is reserved as "empty"),
//   2. streams the probe rows through it, one row per lane; matches are compacted with
//      __ballot_sync / popc into a CTA-wide staging tile in shared memory,
//   3. flushes the staging tile with ONE global atomicAdd per flush and coalesced stores to
//      the four output columns (so the output needs no cudf::concatenate afterwards).
// Multimap semantics: probing continues past a hit until an empty slot.  Build buckets larger
// than the table (skew / duplicates) are processed in chunks, re-streaming the probe side.
#include "dj_device.cuh"
#include "dj_internal.h"

namespace dj {

namespace {

constexpr int kJoinThreads = 512;
constexpr int kSlots       = 4096;             // table slots per CTA (power of 2)
constexpr int kChunkRows   = kSlots * 3 / 4;   // max build rows inserted at once
constexpr int kTargetRows  = kSlots * 3 / 8;   // planned average build rows per bucket
constexpr int kOutCap      = 1024;             // staged output rows per CTA
constexpr int kFlushAt     = kOutCap - kJoinThreads;

struct JoinDev {
  const int64_t* bk;
  const int64_t* bp;
  const int64_t* boff;
  const int64_t* pk;
  const int64_t* pp;
  const int64_t* poff;
  int nbuckets;
  int64_t* out[4];
  int64_t out_capacity;
  unsigned long long* out_count;
  int* work_counter;
};

struct __align__(16) JoinSmem {
  int64_t skey[kSlots];
  int64_t spay[kSlots];
  int64_t sout[4][kOutCap];
  unsigned occ[kSlots / 32];
  int scnt;
  int sbucket;
  unsigned long long sbase;
};

__device__ __forceinline__ void flush_staging(JoinSmem& s, const JoinDev& d, int n)
{
  // all threads call; n is uniform
  if (n > kOutCap) n = kOutCap;
  if (threadIdx.x == 0) s.sbase = atomicAdd(d.out_count, (unsigned long long)n);
  __syncthreads();
  const int64_t base = (int64_t)s.sbase;
#pragma unroll
  for (int c = 0; c < 4; c++)
    for (int i = threadIdx.x; i < n; i += kJoinThreads)
      if (base + i < d.out_capacity) d.out[c][base + i] = s.sout[c][i];
  __syncthreads();
  if (threadIdx.x == 0) s.scnt = 0;
  __syncthreads();
}

__global__ void __launch_bounds__(kJoinThreads, 2) bucket_join_kernel(JoinDev d)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  JoinSmem& s    = *reinterpret_cast<JoinSmem*>(smem_raw);
  const int tid  = threadIdx.x;
  const int lane = tid & 31;
  if (tid == 0) s.scnt = 0;

  while (true) {
    __syncthreads();
    if (tid == 0) s.sbucket = atomicAdd(d.work_counter, 1);
    __syncthreads();
    const int b = s.sbucket;
    if (b >= d.nbuckets) break;
    const int64_t b0 = d.boff[b], b1 = d.boff[b + 1];
    const int64_t p0 = d.poff[b], p1 = d.poff[b + 1];
    if (b1 == b0 || p1 == p0) continue;

    for (int64_t c0 = b0; c0 < b1; c0 += kChunkRows) {
      const int64_t c1 = (c0 + kChunkRows < b1) ? c0 + kChunkRows : b1;
      if (tid < kSlots / 32) s.occ[tid] = 0;
      __syncthreads();

      // ---- build: claim a slot bit, then fill the slot
      for (int64_t i = c0 + tid; i < c1; i += kJoinThreads) {
        const int64_t k = d.bk[i];
        const int64_t v = d.bp[i];
        uint32_t slot   = slot_hash_i64(k) & (kSlots - 1);
        while (true) {
          const unsigned bit = 1u << (slot & 31);
          const unsigned old = atomicOr(&s.occ[slot >> 5], bit);
          if (!(old & bit)) break;
          slot = (slot + 1) & (kSlots - 1);
        }
        s.skey[slot] = k;
        s.spay[slot] = v;
      }
      __syncthreads();

      // ---- probe: one row per lane and round, warp-synchronous chain walk
      for (int64_t r0 = p0; r0 < p1; r0 += kJoinThreads) {
        const int64_t i = r0 + tid;
        bool active     = i < p1;
        int64_t k = 0, v = 0;
        uint32_t slot = 0;
        if (active) {
          k    = d.pk[i];
          v    = d.pp[i];
          slot = slot_hash_i64(k) & (kSlots - 1);
        }
        while (__any_sync(0xffffffffu, active)) {
          bool match = false;
          if (active) {
            if (!((s.occ[slot >> 5] >> (slot & 31)) & 1u))
              active = false;
            else
              match = (s.skey[slot] == k);
          }
          const unsigned m = __ballot_sync(0xffffffffu, match);
          if (m) {
            const int leader = __ffs(m) - 1;
            int base         = 0;
            if (lane == leader) base = atomicAdd(&s.scnt, __popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (match) {
              const int pos = base + __popc(m & lanemask_lt());
              const int64_t bv = s.spay[slot];
              if (pos < kOutCap) {
                s.sout[0][pos] = k;
                s.sout[1][pos] = bv;
                s.sout[2][pos] = k;
                s.sout[3][pos] = v;
              } else {
                // staging full (many matches per probe row): direct, uncoalesced emit
                const int64_t g = (int64_t)atomicAdd(d.out_count, 1ull);
                if (g < d.out_capacity) {
                  d.out[0][g] = k;
                  d.out[1][g] = bv;
                  d.out[2][g] = k;
                  d.out[3][g] = v;
                }
              }
            }
          }
          if (active) slot = (slot + 1) & (kSlots - 1);
        }
        // barrier + uniform decision in one: the last emitting warp sees the final count
        const int need = __syncthreads_or(s.scnt > kFlushAt);
        if (need) flush_staging(s, d, s.scnt);
      }
    }
  }
  __syncthreads();
  const int n = s.scnt;
  __syncthreads();
  if (n > 0) flush_staging(s, d, n);
}

// Swap (build, probe) output halves when the caller's left table was used as the probe side.
}  // namespace

RadixPlan make_radix_plan(int64_t nbuild)
{
  RadixPlan p{0, 0, 1};
  int bits = 0;
  while (bits < 20 && (nbuild >> bits) > kTargetRows) bits++;
  if (bits <= 10) {
    p.bits1 = bits;
    p.bits2 = 0;
  } else {
    p.bits1 = bits / 2;
    p.bits2 = bits - p.bits1;
  }
  p.nbuckets = 1 << bits;
  return p;
}

int run_bucket_join(const JoinBuffers& jb, bool swap_output_sides, cudaStream_t stream)
{
  JoinDev d{};
  d.bk       = jb.bk;
  d.bp       = jb.bp;
  d.boff     = jb.d_build_off;
  d.pk       = jb.pk;
  d.pp       = jb.pp;
  d.poff     = jb.d_probe_off;
  d.nbuckets = jb.nbuckets;
  for (int c = 0; c < 4; c++) d.out[c] = jb.out[swap_output_sides ? (c + 2) % 4 : c];
  d.out_capacity = jb.out_capacity;
  d.out_count    = (unsigned long long*)jb.d_out_count;
  d.work_counter = jb.d_work_counter;

  const size_t smem = sizeof(JoinSmem);
  DJ_CUDA_TRY(cudaFuncSetAttribute(bucket_join_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  int per_sm = 1;
  DJ_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bucket_join_kernel,
                                                            kJoinThreads, smem));
  if (per_sm < 1) per_sm = 1;
  int grid = sm_count() * per_sm;
  if (grid > jb.nbuckets) grid = jb.nbuckets;
  {
    ProfScope prof(DJ_PROF_JOIN, stream);
    bucket_join_kernel<<<grid, kJoinThreads, smem, stream>>>(d);
  }
  DJ_LAUNCH_CHECK();
  return DJ_OK;
}

}  // namespace dj
