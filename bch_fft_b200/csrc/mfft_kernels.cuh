// Batched multiplicative (mixed-radix Cooley-Tukey) FFT over the cyclic group of GF(2^m):
// out[i] = P(x^i), i < N = 2^m - 1 (replaces libfft/multiplicative_fft.c:106-324 of the
// reference).  Secondary kernel: N * sum(prime factors) constant multiplications against
// ~1.5 N m for the additive path.
//
// Decimation in time over the factor list f_1 .. f_k (descending), S_l = f_1 ... f_l:
//   D_0[off][0]  = a_off                                            off < N
//   D_l[off][i]  = sum_{r < f_l} D_{l-1}[off + (N/S_l) r][i mod S_{l-1}] * x^((N/S_l) i r)
//                                                                    off < N/S_l, i < S_l
//   out[i]       = D_k[0][i]
// D_l is stored at position off * S_l + i of a slab buffer (same bitsliced layout as the
// additive FFT).  For a polynomial with ncoef <= N/S_l coefficients level l is a pure
// broadcast (only r = 0 contributes) and is skipped -- the counterpart of the reference's
// truncated evaluation (multiplicative_fft.c:290-295), largest factors first.
#pragma once
#include "bs_common.cuh"

namespace bchfft {

// one level: grid = (ceil(N / blockDim), slabs).  `first`: src holds the coefficients
// themselves (every skipped level is a broadcast), indexed by offset only.
template <int M>
__global__ void __launch_bounds__(256)
k_mfft_stage(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, size_t slab_words,
             uint32_t N, uint32_t S_prev, uint32_t p, uint32_t step, uint32_t first, uint32_t ncoef,
             const uint32_t *__restrict__ gexp)
{
  constexpr int PS = plane_stride(M);
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N) return;
  const uint32_t S = S_prev * p;
  const uint32_t off = idx / S, i = idx - off * S, ip = i % S_prev;
  const uint32_t *sbase = src + (size_t)blockIdx.y * slab_words;
  const uint32_t stepi = (uint32_t)(((uint64_t)step * i) % N);
  uint32_t acc[M];
#pragma unroll
  for (int b = 0; b < M; ++b) acc[b] = 0u;
  uint32_t e = 0;
  for (uint32_t r = 0; r < p; ++r) {
    const uint32_t off2 = off + step * r;
    size_t at;
    if (first) {
      if (off2 >= ncoef) break;   // coefficients beyond ncoef are zero, off2 grows with r
      at = off2;
    } else {
      at = (size_t)off2 * S_prev + ip;
    }
    const uint4 *in = reinterpret_cast<const uint4 *>(sbase + at * PS);
    uint32_t a[M];
#pragma unroll
    for (int v = 0; v < PS / 4; ++v) {
      const uint4 q = in[v];
      if (4 * v + 0 < M) a[4 * v + 0] = q.x;
      if (4 * v + 1 < M) a[4 * v + 1] = q.y;
      if (4 * v + 2 < M) a[4 * v + 2] = q.z;
      if (4 * v + 3 < M) a[4 * v + 3] = q.w;
    }
    if (e == 0u) {
#pragma unroll
      for (int b = 0; b < M; ++b) acc[b] ^= a[b];
    } else {
      bs_mad_scalar<M>(acc, a, gexp[e]);
    }
    e += stepi;
    if (e >= N) e -= N;
  }
  uint4 *out = reinterpret_cast<uint4 *>(dst + (size_t)blockIdx.y * slab_words + (size_t)idx * PS);
#pragma unroll
  for (int v = 0; v < PS / 4; ++v) {
    uint4 q;
    q.x = (4 * v + 0 < M) ? acc[4 * v + 0] : 0u;
    q.y = (4 * v + 1 < M) ? acc[4 * v + 1] : 0u;
    q.z = (4 * v + 2 < M) ? acc[4 * v + 2] : 0u;
    q.w = (4 * v + 3 < M) ? acc[4 * v + 3] : 0u;
    out[v] = q;
  }
}

}  // namespace bchfft
