// mb_write_pattern.cu -- memory-system floor of one radix-partition pass as a function of the
// OUTPUT layout, with no in-SM sorting work: every CTA reads a 4096-row tile linearly and writes it
// as per-bucket runs reserved with one global atomicAdd per (tile, bucket), exactly the store
// stream scatter_tma_kernel produces.  Answers: what do SoA 8-byte columns vs AoS 16-byte rows,
// run length and run alignment cost at the L2/HBM level?
//
//   layout 0: SoA in (key[], pay[]) -> SoA out, 2 x STG.64 per row
//   layout 1: SoA in -> AoS out (16-byte rows), 1 x STG.128 per row
//   layout 2: AoS in -> AoS out
//   runlen  : average rows per (tile, bucket) run; F = 4096 / runlen buckets
//   jitter 0: every run is exactly `runlen` rows (runs stay aligned to runlen rows)
//   jitter 1: run lengths vary pseudo-randomly in [runlen/2, 3*runlen/2] (unaligned runs, the real case)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int THREADS = 1024, RPT = 4, T = THREADS * RPT;

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}

template <int LAYOUT>
__global__ void __launch_bounds__(THREADS, 2)
kernel(const int64_t* __restrict__ ik, const int64_t* __restrict__ ip, const int4* __restrict__ irow,
       int64_t* ok, int64_t* op, int4* orow, unsigned long long* cursor, const unsigned long long* base,
       int64_t nrows, int F, int runlen, int jitter, int transpose)
{
  __shared__ uint32_t s_dst[T / 2];      // destination of every run (rows)
  const int tid = threadIdx.x;
  const int64_t ntiles = nrows / T;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    // run lengths: buckets are paired, run 2m has runlen + d(m) rows and run 2m+1 runlen - d(m)
    auto dj = [&](int m) -> int {
      return jitter ? (int)(mix((uint32_t)t * 977u + (uint32_t)m) % (uint32_t)runlen) - runlen / 2 : 0;
    };
    for (int b = tid; b < F; b += THREADS) {
      const int d   = dj(b >> 1);
      const int len = (b & 1) ? runlen - d : runlen + d;
      // bucket permuted per tile so that consecutive runs go to unrelated buckets
      const int bb = (int)(((uint32_t)b * 2654435761u + (uint32_t)t * 40503u) % (uint32_t)F);
      s_dst[b] = (uint32_t)(base[bb] + atomicAdd(&cursor[bb], (unsigned long long)len));
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RPT; j++) {
      const int i0     = j * THREADS + tid;
      // transpose: consecutive lanes hold rows of different runs (direct scatter without sorting)
      const int i      = transpose ? (i0 % F) * runlen + i0 / F : i0;
      const int m      = i / (2 * runlen);
      const int within = i - m * 2 * runlen;
      const int first  = runlen + dj(m);
      const int run    = 2 * m + (within >= first);
      const uint32_t dst = s_dst[run] + (uint32_t)(within >= first ? within - first : within);
      const int64_t src  = t * T + i0;
      if (LAYOUT == 0) {
        ok[dst] = ik[src];
        op[dst] = ip[src];
      } else if (LAYOUT == 1) {
        const int64_t k = ik[src], p = ip[src];
        orow[dst] = make_int4((int)k, (int)(k >> 32), (int)p, (int)(p >> 32));
      } else {
        orow[dst] = irow[src];
      }
    }
    __syncthreads();
  }
}


// layout 3: AoS in -> AoS out, tile staged by one TMA bulk load, every run written by ONE
// cp.async.bulk shared->global issued by the thread that owns the bucket (no LDS/STG at all).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(THREADS, 1)
bulk_kernel(const int4* __restrict__ irow, int4* orow, unsigned long long* cursor, const unsigned long long* base,
            int64_t nrows, int F, int runlen, int jitter)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  int4* stage               = reinterpret_cast<int4*>(smem_raw);  // [2][T]
  unsigned long long* full  = reinterpret_cast<unsigned long long*>(smem_raw + 2 * T * 16);
  const int tid = threadIdx.x;
  const int64_t ntiles = nrows / T;
  if (tid == 0) {
    for (int i = 0; i < 2; i++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int k) {
    const int64_t t = blockIdx.x + (int64_t)k * gridDim.x;
    if (t >= ntiles) return;
    const int st = k & 1;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[st])), "r"(T * 16) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(stage + (size_t)st * T)), "l"(irow + t * T), "r"(T * 16), "r"(smem_u32(&full[st])) : "memory");
  };
  if (tid == 0) { issue(0); issue(1); }
  for (int k = 0; blockIdx.x + (int64_t)k * gridDim.x < ntiles; k++) {
    const int64_t t = blockIdx.x + (int64_t)k * gridDim.x;
    const int st = k & 1;
    {
      const uint32_t parity = (k >> 1) & 1;
      asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(
                     smem_u32(&full[st])), "r"(parity) : "memory");
    }
    auto dj = [&](int m) -> int {
      return jitter ? (int)(mix((uint32_t)t * 977u + (uint32_t)m) % (uint32_t)runlen) - runlen / 2 : 0;
    };
    for (int b = tid; b < F; b += THREADS) {
      const int d     = dj(b >> 1);
      const int len   = (b & 1) ? runlen - d : runlen + d;
      const int start = (b >> 1) * 2 * runlen + ((b & 1) ? runlen + d : 0);
      const int bb    = (int)(((uint32_t)b * 2654435761u + (uint32_t)t * 40503u) % (uint32_t)F);
      const unsigned long long dst = base[bb] + atomicAdd(&cursor[bb], (unsigned long long)len);
      if (len > 0)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(orow + dst),
                     "r"(smem_u32(stage + (size_t)st * T + start)), "r"(len * 16) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();  // every run of this stage has been read out of shared memory
    if (tid == 0) issue(k + 2);
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void copy_kernel(const int4* __restrict__ in, int4* out, int64_t n)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = in[i];
}

int main(int argc, char** argv)
{
  const int64_t nrows = (argc > 1 ? atoll(argv[1]) : 400000000LL) / T * T;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int64_t *ik, *ip, *ok, *op;
  cudaMalloc(&ik, nrows * 16);  // SoA: two column halves; AoS: nrows 16-byte rows
  ip = ik + nrows;
  const int64_t out_rows = nrows + nrows / 4 + (1 << 22);
  cudaMalloc(&ok, out_rows * 16);  // SoA: two column halves; AoS: out_rows 16-byte rows
  op = ok + out_rows;
  cudaMemset(ik, 1, nrows * 16);
  unsigned long long *cursor, *base;
  cudaMalloc(&cursor, 4096 * 8); cudaMalloc(&base, 4096 * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  // baseline: plain 16-byte copy of the same bytes
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    copy_kernel<<<sms * 8, 512>>>((const int4*)ik, (int4*)ok, nrows);  // nrows*16 bytes = both columns' worth
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  printf("copy 16 B/row in + 16 B/row out            : %7.3f ms  %7.1f GB/s\n", ms, 32.0 * nrows / ms * 1e-6);
  for (int layout = 0; layout < 3; layout++)
    for (int runlen : {4, 8, 16, 32})
      for (int jitter = 0; jitter < 2; jitter++) {
        const int F = T / runlen;
        const int64_t rows_in = nrows;
        // bucket capacities: every bucket gets rows/F * 1.2 rows
        unsigned long long hb[4096];
        const unsigned long long cap = (unsigned long long)(rows_in / F + rows_in / F / 5 + 64) / 8 * 8;
        for (int b = 0; b < F; b++) hb[b] = cap * b;
        if (cap * F > (unsigned long long)out_rows) { printf("skip\n"); continue; }
        cudaMemcpy(base, hb, F * 8, cudaMemcpyHostToDevice);
        float best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
          cudaMemset(cursor, 0, 4096 * 8);
          cudaEventRecord(e0);
          if (layout == 0)
            kernel<0><<<2 * sms, THREADS>>>(ik, ip, nullptr, ok, op, nullptr, cursor, base, nrows, F, runlen, jitter, 0);
          else if (layout == 1)
            kernel<1><<<2 * sms, THREADS>>>(ik, ip, nullptr, nullptr, nullptr, (int4*)ok, cursor, base, nrows, F, runlen, jitter, 0);
          else
            kernel<2><<<2 * sms, THREADS>>>(nullptr, nullptr, (const int4*)ik, nullptr, nullptr, (int4*)ok, cursor, base,
                                        nrows, F, runlen, jitter, 0);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) best = ms;
        }
        const double rows = nrows;
        printf("layout %d (%s) runlen %2d F %4d jitter %d : %7.3f ms  %7.1f GB/s  [%s]\n", layout,
               layout == 0 ? "SoA->SoA" : layout == 1 ? "SoA->AoS" : "AoS->AoS", runlen, F, jitter, best,
               32.0 * rows / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
      }
  {
    const size_t smem = 2 * T * 16 + 64;
    cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int runlen : {4, 8, 16, 32})
      for (int jitter = 0; jitter < 2; jitter++) {
        const int F = T / runlen;
        unsigned long long hb[4096];
        const unsigned long long cap = (unsigned long long)(nrows / F + nrows / F / 5 + 64) / 8 * 8;
        for (int b = 0; b < F; b++) hb[b] = cap * b;
        cudaMemcpy(base, hb, F * 8, cudaMemcpyHostToDevice);
        float best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
          cudaMemset(cursor, 0, 4096 * 8);
          cudaEventRecord(e0);
          bulk_kernel<<<sms, THREADS, smem>>>((const int4*)ik, (int4*)ok, cursor, base, nrows, F, runlen, jitter);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) best = ms;
        }
        printf("layout 3 (AoS->AoS, TMA load + cp.async.bulk store per run) runlen %2d F %4d jitter %d : %7.3f ms  %7.1f GB/s  [%s]\n",
               runlen, F, jitter, best, 32.0 * nrows / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
      }
    // direct scatter: every lane of a warp stores to a different bucket (no in-SM sort needed)
    for (int layout = 1; layout < 3; layout++) {
      const int runlen = 4, F = T / runlen;
      unsigned long long hb[4096];
      const unsigned long long cap = (unsigned long long)(nrows / F + nrows / F / 5 + 64) / 8 * 8;
      for (int b = 0; b < F; b++) hb[b] = cap * b;
      cudaMemcpy(base, hb, F * 8, cudaMemcpyHostToDevice);
      float best = 1e9;
      for (int rep = 0; rep < 3; rep++) {
        cudaMemset(cursor, 0, 4096 * 8);
        cudaEventRecord(e0);
        if (layout == 1)
          kernel<1><<<2 * sms, THREADS>>>(ik, ip, nullptr, nullptr, nullptr, (int4*)ok, cursor, base, nrows, F, runlen, 0, 1);
        else
          kernel<2><<<2 * sms, THREADS>>>(nullptr, nullptr, (const int4*)ik, nullptr, nullptr, (int4*)ok, cursor, base, nrows, F, runlen, 0, 1);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      printf("direct scatter layout %d (lanes -> 32 different buckets, F %d) : %7.3f ms  %7.1f GB/s  [%s]\n", layout, F, best,
             32.0 * nrows / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
