// mb_smem_rank.cu -- per-SM cost (cycles per 32 rows) of the scatter kernel's building blocks.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int THREADS = 1024, RPT = 4, T = THREADS * RPT, ITERS = 200;

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}

// mode 0: atomicAdd ranking; 1: ballot-match ranking with warp-private uint16 counters;
// 2: random STS.128 scatter + linear LDS.128; 3: random 2 x STS.64 scatter + linear 2 x LDS.64
template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) kernel(int F, int fbits, unsigned long long* cycles, uint32_t* sink)
{
  extern __shared__ __align__(16) unsigned char smem[];
  int* hist       = reinterpret_cast<int*>(smem);                       // F ints
  uint16_t* whist = reinterpret_cast<uint16_t*>(smem + 4096);           // 32 warps x F
  int4* rows      = reinterpret_cast<int4*>(smem + 4096 + 65536);       // T x 16 B
  int64_t* ka     = reinterpret_cast<int64_t*>(rows);
  int64_t* pa     = ka + T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t acc = 0;
  for (int i = tid; i < F; i += THREADS) hist[i] = 0;
  for (int i = tid; i < 32 * F; i += THREADS) whist[i] = 0;
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int j = 0; j < RPT; j++) {
      const uint32_t h = mix((blockIdx.x * ITERS + it) * T + j * THREADS + tid);
      const int b      = h & (F - 1);
      if (MODE == 0) {
        acc += atomicAdd(&hist[b], 1);
      } else if (MODE == 1) {
        unsigned peers = 0xffffffffu;
        for (int bit = 0; bit < fbits; bit++) {
          const unsigned m = __ballot_sync(0xffffffffu, (b >> bit) & 1);
          peers &= ((b >> bit) & 1) ? m : ~m;
        }
        const int leader = __ffs(peers) - 1;
        uint16_t* c      = &whist[warp * F + b];
        int base         = 0;
        if (lane == leader) {
          base = *c;
          *c   = (uint16_t)(base + __popc(peers));
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        acc += base + __popc(peers & ((1u << lane) - 1));
      } else if (MODE == 2) {
        const int pos = (h >> 8) & (T - 1);
        rows[pos]     = make_int4(h, tid, j, it);
      } else {
        const int pos = (h >> 8) & (T - 1);
        ka[pos]       = h;
        pa[pos]       = tid;
      }
    }
    if (MODE >= 2) {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < RPT; j++) {
        const int i = j * THREADS + tid;
        if (MODE == 2) acc += rows[i].x + rows[i].z;
        else acc += (uint32_t)ka[i] + (uint32_t)pa[i];
      }
      __syncthreads();
    }
  }
  const unsigned long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * THREADS + tid] = acc;
}

template <int MODE>
void run(const char* name, int F, int fbits)
{
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  unsigned long long* d_cycles;
  uint32_t* d_sink;
  cudaMalloc(&d_cycles, sms * 8);
  cudaMalloc(&d_sink, (size_t)sms * THREADS * 4);
  const size_t smem = 4096 + 65536 + (size_t)T * 16;
  cudaFuncSetAttribute(kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kernel<MODE><<<sms, THREADS, smem>>>(F, fbits, d_cycles, d_sink);
  kernel<MODE><<<sms, THREADS, smem>>>(F, fbits, d_cycles, d_sink);
  unsigned long long h[256] = {0};
  cudaMemcpy(h, d_cycles, sms * 8, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms; i++) avg += (double)h[i];
  avg /= sms;
  printf("%-44s F=%4d : %7.1f cycles per 32 rows (SM-wide)  [%s]\n", name, F, avg / (ITERS * (T / 32.0)),
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(d_cycles);
  cudaFree(d_sink);
}

int main()
{
  for (int fbits : {3, 6, 9, 10}) {
    run<0>("rank: shared atomicAdd with return", 1 << fbits, fbits);
    run<1>("rank: ballot match + warp-private counters", 1 << fbits, fbits);
  }
  run<2>("sort: random STS.128 + linear LDS.128", 1024, 10);
  run<3>("sort: random 2xSTS.64 + linear 2xLDS.64", 1024, 10);
  return 0;
}
