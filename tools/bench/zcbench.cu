// Zero-copy scatter microbenchmark: can the correction step write its (at most t) changed 64-bit words per
// codeword straight into the caller's page-locked host array (UVA-mapped), instead of copying every 1 KiB
// word back?  Per "codeword" (1 KiB region of a pinned host buffer) E random 8-byte words are written by a
// kernel; compared with the plain D2H DMA of the same buffer, alone and with an H2D DMA running alongside.
//   nvcc -O2 -o zcbench zcbench.cu ; ./zcbench [codewords = 1000000] [E = 8]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void k_scatter(uint64_t *__restrict__ host_words, const uint64_t *__restrict__ dev_words, uint32_t wpc,
                          uint32_t batch, uint32_t E, uint32_t seed)
{
  const uint32_t cw = blockIdx.x * blockDim.x + threadIdx.x;
  if (cw >= batch) return;
  uint32_t s = cw * 2654435761u + seed;
  for (uint32_t e = 0; e < E; e++) {
    s = s * 1664525u + 1013904223u;
    const uint32_t w = (s >> 8) % wpc;
    host_words[(size_t)cw * wpc + w] = dev_words[(size_t)cw * wpc + w] ^ (1ull << (s & 63u));
  }
}

int main(int argc, char **argv)
{
  const uint32_t B = argc > 1 ? (uint32_t)atol(argv[1]) : 1000000u, E = argc > 2 ? (uint32_t)atoi(argv[2]) : 8u;
  const uint32_t wpc = 128;
  const size_t bytes = (size_t)B * wpc * 8;
  uint64_t *h, *h2, *d, *d2;
  CK(cudaMallocHost((void **)&h, bytes));
  CK(cudaMallocHost((void **)&h2, bytes));
  memset(h, 0, bytes);
  memset(h2, 1, bytes);
  CK(cudaMalloc((void **)&d, bytes));
  CK(cudaMalloc((void **)&d2, bytes));
  CK(cudaMemset(d, 0, bytes));
  cudaStream_t s0, s1;
  CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  float ms;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(a, s0));
    k_scatter<<<(B + 255) / 256, 256, 0, s0>>>(h, d, wpc, B, E, rep);
    CK(cudaEventRecord(b, s0));
    CK(cudaStreamSynchronize(s0));
    CK(cudaEventElapsedTime(&ms, a, b));
    printf("{\"mode\": \"zero-copy scatter\", \"codewords\": %u, \"writes_per_codeword\": %u, \"ms\": %.3f, \"Mwrites_per_s\": %.1f}\n",
           B, E, ms, (double)B * E / ms / 1e3);
  }
  // check one value landed
  {
    uint32_t s = 0 * 2654435761u + 2u;
    s = s * 1664525u + 1013904223u;
    const uint32_t w = (s >> 8) % wpc;
    if (h[w] == 0) { fprintf(stderr, "scatter did not land\n"); return 1; }
  }
  CK(cudaEventRecord(a, s0));
  CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s0));
  CK(cudaEventRecord(b, s0));
  CK(cudaStreamSynchronize(s0));
  CK(cudaEventElapsedTime(&ms, a, b));
  printf("{\"mode\": \"d2h dma of every word\", \"ms\": %.3f, \"GB_per_s\": %.1f}\n", ms, bytes / ms / 1e6);
  CK(cudaEventRecord(a, s1));
  CK(cudaMemcpyAsync(d2, h2, bytes, cudaMemcpyHostToDevice, s1));
  CK(cudaEventRecord(b, s1));
  CK(cudaStreamSynchronize(s1));
  CK(cudaEventElapsedTime(&ms, a, b));
  printf("{\"mode\": \"h2d dma alone\", \"ms\": %.3f, \"GB_per_s\": %.1f}\n", ms, bytes / ms / 1e6);
  // scatter while an H2D DMA of the same size runs
  for (int rep = 0; rep < 2; rep++) {
    cudaEvent_t c, e;
    CK(cudaEventCreate(&c));
    CK(cudaEventCreate(&e));
    CK(cudaEventRecord(c, s1));
    CK(cudaMemcpyAsync(d2, h2, bytes, cudaMemcpyHostToDevice, s1));
    CK(cudaEventRecord(e, s1));
    CK(cudaEventRecord(a, s0));
    for (int k = 0; k < 4; k++) k_scatter<<<(B + 255) / 256, 256, 0, s0>>>(h, d, wpc, B, E, 10 + k);
    CK(cudaEventRecord(b, s0));
    CK(cudaDeviceSynchronize());
    float ms_h2d;
    CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaEventElapsedTime(&ms_h2d, c, e));
    printf("{\"mode\": \"4 x scatter alongside h2d dma\", \"scatter_ms_each\": %.3f, \"h2d_ms\": %.3f, \"h2d_GB_per_s\": %.1f}\n",
           ms / 4, ms_h2d, bytes / ms_h2d / 1e6);
  }
  return 0;
}
