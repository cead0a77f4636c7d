// mb_l2_join.cu -- build / probe rate of a global-memory fingerprint table (one table shared by the
// whole GPU) as a function of its size: is an L2-resident join of ~1M-row buckets competitive
// with a second radix pass + shared-memory join?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint64_t k)
{
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33;
  return (uint32_t)(k >> 16);
}

__global__ void build(const int64_t* keys, int n, uint32_t* slots, uint32_t mask)
{
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t h = mix((uint64_t)keys[i]);
    uint32_t s = h & mask;
    const uint32_t e = 0x80000000u | ((h >> 11) << 20 & 0x7ff00000u) | (uint32_t)(i & 0xfffff);
    while (atomicCAS(&slots[s], 0u, e) != 0u) s = (s + 1) & mask;
  }
}

__global__ void probe(const int64_t* bkeys, const int64_t* pkeys, int n, const uint32_t* slots, uint32_t mask,
                      unsigned long long* matches)
{
  unsigned long long m = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int64_t k  = pkeys[i];
    const uint32_t h = mix((uint64_t)k);
    uint32_t s = h & mask;
    const uint32_t want = ((h >> 11) << 20) & 0x7ff00000u;
    for (uint32_t e = slots[s]; e != 0u; s = (s + 1) & mask, e = slots[s])
      if ((e & 0x7ff00000u) == want && bkeys[e & 0xfffff] == k) m++;
  }
  atomicAdd(matches, m);
}

int main()
{
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int logn = 16; logn <= 20; logn++) {
    const int n = 1 << logn;  // build rows == probe rows; idx field holds 20 bits
    const uint32_t nslots = 4u << logn;  // load 0.25
    int64_t *bk, *pk;
    uint32_t* slots;
    unsigned long long* matches;
    cudaMalloc(&bk, (size_t)n * 8); cudaMalloc(&pk, (size_t)n * 8);
    cudaMalloc(&slots, (size_t)nslots * 4); cudaMalloc(&matches, 8);
    int64_t* h = new int64_t[n];
    for (int i = 0; i < n; i++) h[i] = (int64_t)i * 2654435761LL + 12345;
    cudaMemcpy(bk, h, (size_t)n * 8, cudaMemcpyHostToDevice);
    for (int i = 0; i < n; i++) h[i] = (i % 3 == 0) ? h[i] : ~h[i];  // selectivity 1/3
    cudaMemcpy(pk, h, (size_t)n * 8, cudaMemcpyHostToDevice);
    delete[] h;
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    float tb = 0, tp = 0;
    const int reps = 20;
    for (int r = 0; r < reps + 2; r++) {
      cudaMemset(slots, 0, (size_t)nslots * 4);
      cudaMemset(matches, 0, 8);
      cudaEventRecord(e0);
      build<<<sms * 4, 512>>>(bk, n, slots, nslots - 1);
      cudaEventRecord(e1);
      probe<<<sms * 4, 512>>>(bk, pk, n, slots, nslots - 1, matches);
      cudaEventRecord(e2);
      cudaEventSynchronize(e2);
      float a, b;
      cudaEventElapsedTime(&a, e0, e1);
      cudaEventElapsedTime(&b, e1, e2);
      if (r >= 2) { tb += a; tp += b; }
    }
    unsigned long long hm = 0;
    cudaMemcpy(&hm, matches, 8, cudaMemcpyDeviceToHost);
    printf("rows 2^%d, table %6.1f MB: build %7.2f Grows/s, probe %7.2f Grows/s, matches %llu [%s]\n", logn,
           nslots * 4 / 1e6, n / (tb / reps * 1e-3) / 1e9, n / (tp / reps * 1e-3) / 1e9, hm,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(bk); cudaFree(pk); cudaFree(slots); cudaFree(matches);
  }
  return 0;
}
