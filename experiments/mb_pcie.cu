// mb_pcie.cu -- host<->device copy bandwidth from pinned memory: one direction, both directions at
// once (two streams), and chunked copies; decides the chunk size and overlap plan of the
// end-to-end (host-buffer) join entry.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

static float timed(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b); return ms; }

int main(int argc, char** argv)
{
  const size_t bytes = (size_t)(argc > 1 ? atoll(argv[1]) : 4096) << 20;
  char *h_a, *h_b, *d_a, *d_b;
  cudaHostAlloc(&h_a, bytes, cudaHostAllocDefault);
  cudaHostAlloc(&h_b, bytes, cudaHostAllocDefault);
  memset(h_a, 1, bytes); memset(h_b, 2, bytes);
  cudaMalloc(&d_a, bytes); cudaMalloc(&d_b, bytes);
  cudaStream_t s0, s1;
  cudaStreamCreate(&s0); cudaStreamCreate(&s1);
  cudaEvent_t e0, e1, f0, f1;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&f0); cudaEventCreate(&f1);
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0, s0); cudaMemcpyAsync(d_a, h_a, bytes, cudaMemcpyHostToDevice, s0); cudaEventRecord(e1, s0);
    float ms = timed(e0, e1);
    printf("H2D alone        %6zu MB: %8.2f ms %6.1f GB/s\n", bytes >> 20, ms, bytes / ms * 1e-6);
    cudaEventRecord(e0, s0); cudaMemcpyAsync(h_b, d_b, bytes, cudaMemcpyDeviceToHost, s0); cudaEventRecord(e1, s0);
    ms = timed(e0, e1);
    printf("D2H alone        %6zu MB: %8.2f ms %6.1f GB/s\n", bytes >> 20, ms, bytes / ms * 1e-6);
    cudaDeviceSynchronize();
    cudaEventRecord(e0, s0); cudaEventRecord(f0, s1);
    cudaMemcpyAsync(d_a, h_a, bytes, cudaMemcpyHostToDevice, s0);
    cudaMemcpyAsync(h_b, d_b, bytes, cudaMemcpyDeviceToHost, s1);
    cudaEventRecord(e1, s0); cudaEventRecord(f1, s1);
    float m0 = timed(e0, e1), m1 = timed(f0, f1);
    printf("duplex           %6zu MB: H2D %8.2f ms %6.1f GB/s | D2H %8.2f ms %6.1f GB/s\n", bytes >> 20, m0,
           bytes / m0 * 1e-6, m1, bytes / m1 * 1e-6);
    for (size_t chunk_mb : {16, 64, 256}) {
      const size_t chunk = chunk_mb << 20;
      cudaDeviceSynchronize();
      cudaEventRecord(e0, s0);
      for (size_t off = 0; off < bytes; off += chunk)
        cudaMemcpyAsync(d_a + off, h_a + off, chunk < bytes - off ? chunk : bytes - off, cudaMemcpyHostToDevice, s0);
      cudaEventRecord(e1, s0);
      ms = timed(e0, e1);
      printf("H2D chunks %4zu MB        : %8.2f ms %6.1f GB/s\n", chunk_mb, ms, bytes / ms * 1e-6);
    }
  }
  // pageable source for comparison (what a caller without pinned buffers gets)
  char* h_p = (char*)malloc(bytes / 4);
  memset(h_p, 3, bytes / 4);
  cudaEventRecord(e0, s0); cudaMemcpyAsync(d_a, h_p, bytes / 4, cudaMemcpyHostToDevice, s0); cudaEventRecord(e1, s0);
  float ms = timed(e0, e1);
  printf("H2D pageable     %6zu MB: %8.2f ms %6.1f GB/s\n", (bytes / 4) >> 20, ms, bytes / 4 / ms * 1e-6);
  return 0;
}
