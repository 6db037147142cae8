// Micro-benchmark (development aid): throughput of the bitsliced constant multiply on one GPU,
// for several instruction mixes.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I bch_fft_b200/csrc -DBCHFFT_ALU_ROWS=.. \
//        tools/bench/mulbench.cu -o /tmp/mulbench && /tmp/mulbench
#include <cstdio>
#include <cstdint>
#include "bs_common.cuh"
using namespace bchfft;

#ifndef MB_M
#define MB_M 13
#endif
#ifndef MB_VARIANT
#define MB_VARIANT 0
#endif

template <int M>
__device__ __forceinline__ void mad_variant(uint32_t (&acc)[M], const uint32_t (&a)[M], uint32_t c)
{
#if MB_VARIANT == 0
  bs_mad_scalar<M>(acc, a, c);
#elif MB_VARIANT == 1
  // Horner over the bits of c: acc += c_j * (a * x^j), a*x^j kept reduced (no 2M-1 accumulator)
  uint32_t t[M];
#pragma unroll
  for (int i = 0; i < M; ++i) t[i] = a[i];
  constexpr uint32_t taps = field_poly(M) & ((1u << M) - 1u);
#pragma unroll
  for (int j = 0; j < M; ++j) {
    const uint32_t mask = 0u - ((c >> j) & 1u);
#pragma unroll
    for (int i = 0; i < M; ++i) acc[i] ^= t[i] & mask;
    // t <- t * x
    const uint32_t top = t[M - 1];
#pragma unroll
    for (int i = M - 1; i >= 1; --i) t[i] = t[i - 1] ^ (((taps >> i) & 1u) ? top : 0u);
    t[0] = top;   // taps bit 0 is always set
  }
#elif MB_VARIANT == 2
  // predicate-select formulation: p ^= (bit ? a : 0) written with a ternary
  uint32_t p[2 * M - 1];
#pragma unroll
  for (int k = 0; k < 2 * M - 1; ++k) p[k] = (k < M) ? acc[k] : 0u;
#pragma unroll
  for (int j = 0; j < M; ++j) {
    const bool on = (c >> j) & 1u;
#pragma unroll
    for (int i = 0; i < M; ++i) p[i + j] ^= on ? a[i] : 0u;
  }
  bs_reduce<M>(p);
#pragma unroll
  for (int i = 0; i < M; ++i) acc[i] = p[i];
#elif MB_VARIANT == 3
  // warp-uniform constant: branch over the set bits of c, XOR only (no masks)
  uint32_t p[2 * M - 1];
#pragma unroll
  for (int k = 0; k < 2 * M - 1; ++k) p[k] = (k < M) ? acc[k] : 0u;
#pragma unroll
  for (int j = 0; j < M; ++j) {
    if ((c >> j) & 1u) {
#pragma unroll
      for (int i = 0; i < M; ++i) p[i + j] ^= a[i];
    }
  }
  bs_reduce<M>(p);
#pragma unroll
  for (int i = 0; i < M; ++i) acc[i] = p[i];
#elif MB_VARIANT == 4
  // warp-uniform constant, two bits of c per branch: both set -> one three-input XOR per plane
  uint32_t p[2 * M];
#pragma unroll
  for (int k = 0; k < 2 * M; ++k) p[k] = (k < M) ? acc[k] : 0u;
#pragma unroll
  for (int j = 0; j < M; j += 2) {
    const uint32_t two = (c >> j) & 3u;
    if (j + 1 == M) {
      if (two & 1u) {
#pragma unroll
        for (int i = 0; i < M; ++i) p[i + j] ^= a[i];
      }
    } else if (two == 1u) {
#pragma unroll
      for (int i = 0; i < M; ++i) p[i + j] ^= a[i];
    } else if (two == 2u) {
#pragma unroll
      for (int i = 0; i < M; ++i) p[i + j + 1] ^= a[i];
    } else if (two == 3u) {
      p[j] ^= a[0];
#pragma unroll
      for (int i = 1; i < M; ++i) p[i + j] = xor3(p[i + j], a[i], a[i - 1]);
      p[j + M] ^= a[M - 1];
    }
  }
  uint32_t q[2 * M - 1];
#pragma unroll
  for (int k = 0; k < 2 * M - 1; ++k) q[k] = p[k];
  bs_reduce<M>(q);
#pragma unroll
  for (int i = 0; i < M; ++i) acc[i] = q[i];
#elif MB_VARIANT == 6
  // warp-uniform constant, three bits of c per branch: any two set bits fold into one three-input
  // XOR per plane, three set bits into two
  uint32_t p[2 * M + 1];
#pragma unroll
  for (int k = 0; k < 2 * M + 1; ++k) p[k] = (k < M) ? acc[k] : 0u;
#pragma unroll
  for (int j = 0; j < M; j += 3) {
    const uint32_t g = (c >> j) & 7u & ((1u << (M - j < 3 ? M - j : 3)) - 1u);
    if (g == 0u) continue;
    switch (g) {
      case 1u:
#pragma unroll
        for (int i = 0; i < M; ++i) p[i + j] ^= a[i];
        break;
      case 2u:
#pragma unroll
        for (int i = 0; i < M; ++i) p[i + j + 1] ^= a[i];
        break;
      case 4u:
#pragma unroll
        for (int i = 0; i < M; ++i) p[i + j + 2] ^= a[i];
        break;
      case 3u:
#pragma unroll
        for (int k = 0; k <= M; ++k) p[k + j] = xor3(p[k + j], k < M ? a[k] : 0u, k >= 1 ? a[k - 1] : 0u);
        break;
      case 6u:
#pragma unroll
        for (int k = 0; k <= M; ++k) p[k + j + 1] = xor3(p[k + j + 1], k < M ? a[k] : 0u, k >= 1 ? a[k - 1] : 0u);
        break;
      case 5u:
#pragma unroll
        for (int k = 0; k <= M + 1; ++k) p[k + j] = xor3(p[k + j], k < M ? a[k] : 0u, k >= 2 ? a[k - 2] : 0u);
        break;
      default:   // 7
#pragma unroll
        for (int k = 0; k <= M + 1; ++k)
          p[k + j] = xor3(p[k + j], k < M ? a[k] : 0u, (k >= 1 && k - 1 < M) ? a[k - 1] : 0u) ^ (k >= 2 ? a[k - 2] : 0u);
        break;
    }
  }
  uint32_t q[2 * M - 1];
#pragma unroll
  for (int k = 0; k < 2 * M - 1; ++k) q[k] = p[k];
  bs_reduce<M>(q);
#pragma unroll
  for (int i = 0; i < M; ++i) acc[i] = q[i];
#elif MB_VARIANT == 7
  // warp-uniform constant, two bits of c per branch, running multiple t = a x^j kept reduced:
  // no double-width accumulator, no final reduction, no register moves (the names rotate statically)
  constexpr uint32_t taps = field_poly(M) & ((1u << M) - 1u);
  uint32_t t[M];
#pragma unroll
  for (int i = 0; i < M; ++i) t[i] = a[i];
#pragma unroll
  for (int j = 0; j < M; j += 2) {
    const uint32_t two = (c >> j) & 3u;
    // tx = t x
    uint32_t tx[M];
    {
      const uint32_t top = t[M - 1];
#pragma unroll
      for (int i = M - 1; i >= 1; --i) tx[i] = ((taps >> i) & 1u) ? (t[i - 1] ^ top) : t[i - 1];
      tx[0] = top;
    }
    if (j + 1 == M) {
      if (two & 1u) {
#pragma unroll
        for (int i = 0; i < M; ++i) acc[i] ^= t[i];
      }
    } else if (two == 1u) {
#pragma unroll
      for (int i = 0; i < M; ++i) acc[i] ^= t[i];
    } else if (two == 2u) {
#pragma unroll
      for (int i = 0; i < M; ++i) acc[i] ^= tx[i];
    } else if (two == 3u) {
#pragma unroll
      for (int i = 0; i < M; ++i) acc[i] = xor3(acc[i], t[i], tx[i]);
    }
    if (j + 2 < M) {
      const uint32_t top = tx[M - 1];
#pragma unroll
      for (int i = M - 1; i >= 1; --i) t[i] = ((taps >> i) & 1u) ? (tx[i - 1] ^ top) : tx[i - 1];
      t[0] = top;
    }
  }
#elif MB_VARIANT == 5
  // warp-uniform constant, predication instead of branches: one predicate per bit of c guards the
  // M XORs of that bit (one asm block each, so that ptxas keeps them predicated)
  static_assert(M == 13, "variant 5 is written for M = 13");
  uint32_t p[2 * M - 1];
#pragma unroll
  for (int k = 0; k < 2 * M - 1; ++k) p[k] = (k < M) ? acc[k] : 0u;
#pragma unroll
  for (int j = 0; j < M; ++j) {
    const uint32_t bit = (c >> j) & 1u;
    asm("{ .reg .pred q; setp.ne.u32 q, %26, 0;\n"
        "@q xor.b32 %0, %0, %13; @q xor.b32 %1, %1, %14; @q xor.b32 %2, %2, %15; @q xor.b32 %3, %3, %16;\n"
        "@q xor.b32 %4, %4, %17; @q xor.b32 %5, %5, %18; @q xor.b32 %6, %6, %19; @q xor.b32 %7, %7, %20;\n"
        "@q xor.b32 %8, %8, %21; @q xor.b32 %9, %9, %22; @q xor.b32 %10, %10, %23; @q xor.b32 %11, %11, %24;\n"
        "@q xor.b32 %12, %12, %25; }"
        : "+r"(p[j + 0]), "+r"(p[j + 1]), "+r"(p[j + 2]), "+r"(p[j + 3]), "+r"(p[j + 4]), "+r"(p[j + 5]),
          "+r"(p[j + 6]), "+r"(p[j + 7]), "+r"(p[j + 8]), "+r"(p[j + 9]), "+r"(p[j + 10]), "+r"(p[j + 11]),
          "+r"(p[j + 12])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]),
          "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(bit));
  }
  bs_reduce<M>(p);
#pragma unroll
  for (int i = 0; i < M; ++i) acc[i] = p[i];
#endif
}

#if MB_VARIANT >= 3
#define MB_CONST_INDEX ((blockIdx.x * 37u + (threadIdx.x >> 5)) + it)
#else
#define MB_CONST_INDEX (tid + it)
#endif

__global__ void __launch_bounds__(512, 1) k(uint32_t *out, const uint32_t *consts, int iters)
{
  uint32_t acc[MB_M], a[MB_M];
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int i = 0; i < MB_M; ++i) {
    acc[i] = tid * 2654435761u + i;
    a[i] = tid * 40503u + 7 * i + 1;
  }
  for (int it = 0; it < iters; ++it) {
    const uint32_t c = consts[MB_CONST_INDEX & 1023];
    mad_variant<MB_M>(acc, a, c);
#pragma unroll
    for (int i = 0; i < MB_M; ++i) a[i] ^= acc[(i + 1) % MB_M];
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < MB_M; ++i) r ^= acc[i];
  out[tid] = r;
}

int main()
{
  uint32_t *out, *consts;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&consts, 1024 * 4);
  uint32_t h[1024];
  for (int i = 0; i < 1024; i++) h[i] = (i * 2654435761u >> 7) & ((1u << MB_M) - 1u);
  cudaMemcpy(consts, h, sizeof h, cudaMemcpyHostToDevice);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int iters = 4000;
  k<<<148, 512>>>(out, consts, 100);
  cudaEventRecord(a);
  k<<<148, 512>>>(out, consts, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  static uint32_t hout[148 * 512];
  cudaMemcpy(hout, out, sizeof hout, cudaMemcpyDeviceToHost);
  uint32_t sum = 0;
  for (int i = 0; i < 148 * 512; i++) sum = sum * 31u + hout[i];
  printf("checksum %08x  ", sum);
  const double mults = 148.0 * 512 * iters;
  printf("M=%d variant=%d alu_rows=%d: %.3f ms, %.1f G mult/s, %.1f clk per mult per SMSP-warp (at 1.9 GHz: %.0f)\n",
         MB_M, MB_VARIANT, BCHFFT_ALU_ROWS, ms, mults / ms / 1e6, 0.0, ms * 1e-3 * 1.9e9 / (iters * 4.0));
  return 0;
}
