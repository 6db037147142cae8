// Multiplicative FFT over GF(2^16), N = 2^16 - 1 = 255 x 257 (replaces libfft/multiplicative_fft.c:106-257
// for the BASELINE field; tools/mfft_model.py is the numpy model of every step below).
//
// out[i] = sum_k a_k x^(ik), i < N.  The reference runs a mixed-radix decimation over 3, 5, 17, 257 with a
// direct p-point DFT per stage: 282 table multiplications per point, 257 of them in the last stage.  Here:
//
//  * Good-Thomas (prime-factor) index maps, so that no stage needs twiddle factors:
//        k = (257 k1 + 255 k2) mod N,   i = (32896 i1 + 32640 i2) mod N   (i = i1 mod 255, i = i2 mod 257)
//        out[i] = sum_k1 w1^(k1 i1) sum_k2 w2^(k2 i2) a_k,   w1 = x^257 (order 255), w2 = x^255 (order 257)
//  * pass 1 -- 255 transforms of length 257 per polynomial by the CYCLOTOMIC method: the exponents 1..256
//    split into 16 cosets {k_s 2^i mod 257} of 16 (2 has order 16 mod 257), so
//        A(y) = a_0 + sum_s L_s(y^(k_s)),   L_s(z) = sum_i a_(k_s 2^i) z^(2^i)   GF(2)-linear in z.
//    With a normal basis g^(2^b), b < 16, and z = sum_b c_b(z) g^(2^b):
//        L_s(z) = sum_b c_b(z) V_(s,b),   V_(s,b) = L_s(g^(2^b)) = sum_i a_(k_s 2^i) g^(2^((b+i) mod 16)).
//    256 values V, each 16 multiplications by one of only 16 COMPILE-TIME constants (a constant multiply is a
//    fixed 16 x 16 bit matrix: ~70 LOP3 on bitsliced data, no masks, no branches), then every output is a
//    masked XOR of the V: 16 multiplications + 256 masked additions per point instead of 256 multiplications.
//  * pass 2 -- 257 transforms of length 255 per polynomial as a 3 x 5 x 17 prime-factor transform
//    (k1 = 85 a + 51 b + 15 c, i1 = 85 a' + 51 b' + 120 c'; kernels (w1^85)^(a a'), (w1^51)^(b b'),
//    (w1^15)^(c c')): 2 + 4 + 16 constant multiplications per point, all compile-time constants.
//
// Data: bitsliced slabs, element-major [position][16 planes] (64 bytes per element of 32 polynomials), so
// the strided gathers of both passes move whole sectors.
#pragma once
#include <utility>
#include "bs_common.cuh"

namespace bchfft {
namespace mf16 {

constexpr uint32_t kPoly = 0x1002Du;          // field_poly(16)
constexpr int kN = 65535, kN1 = 255, kN2 = 257;
constexpr int kNormalLog = 11;                // g = x^11: its 16 conjugates are a basis (checked on the host)

__host__ __device__ constexpr uint32_t cmul(uint32_t a, uint32_t b)
{
  uint32_t r = 0;
  for (int i = 15; i >= 0; --i) {
    r <<= 1;
    if (r & 0x10000u) r ^= kPoly;
    if ((b >> i) & 1u) r ^= a;
  }
  return r;
}
__host__ __device__ constexpr uint32_t cpow_x(uint32_t e)   // x^e
{
  uint32_t r = 1, base = 2;
  while (e) {
    if (e & 1u) r = cmul(r, base);
    base = cmul(base, base);
    e >>= 1;
  }
  return r;
}
// mask of the input planes b that feed output plane r of the map a -> a * C
__host__ __device__ constexpr uint32_t row_mask(uint32_t C, int r)
{
  uint32_t m = 0;
  for (int b = 0; b < 16; ++b)
    if ((cmul(C, 1u << b) >> r) & 1u) m |= 1u << b;
  return m;
}

template <uint32_t C, int R>
__device__ __forceinline__ uint32_t const_row(const uint32_t (&in)[16])
{
  constexpr uint32_t mask = row_mask(C, R);
  uint32_t acc = 0u;
#pragma unroll
  for (int b = 0; b < 16; ++b)
    if ((mask >> b) & 1u) acc ^= in[b];
  return acc;
}
template <uint32_t C, int... R>
__device__ __forceinline__ void mad_const_rows(uint32_t (&acc)[16], const uint32_t (&in)[16],
                                               std::integer_sequence<int, R...>)
{
  ((acc[R] ^= const_row<C, R>(in)), ...);
}
// acc += in * C for a compile-time field constant C
template <uint32_t C>
__device__ __forceinline__ void mad_const(uint32_t (&acc)[16], const uint32_t (&in)[16])
{
  mad_const_rows<C>(acc, in, std::make_integer_sequence<int, 16>());
}

constexpr int kAS = 264;   // smem row stride (words) of the 257 / 256 / 255-entry plane rows

// ---- pass 1: one 257-point transform per CTA ------------------------------------------------------
// grid = (255, slabs), 256 threads.  cos16[s * 16 + i] = k_s 2^i mod 257; ctab[e] = coordinates of w2^e in
// the normal basis (bit b <-> g^(2^b)).  src / dst: [slab][position][16]; dst position = k1 * 257 + i2.
template <int C>
__device__ __forceinline__ void p1_term(uint32_t (&acc)[16], const uint32_t *A, const uint32_t *cos16, uint32_t s,
                                        uint32_t b)
{
  constexpr uint32_t G = cpow_x(((uint32_t)kNormalLog << C) % (uint32_t)kN);   // g^(2^C)
  const uint32_t k = cos16[s * 16u + ((C - b) & 15u)];
  uint32_t a[16];
#pragma unroll
  for (int p = 0; p < 16; ++p) a[p] = A[p * kAS + k];
  mad_const<G>(acc, a);
}
template <int... C>
__device__ __forceinline__ void p1_all_terms(uint32_t (&acc)[16], const uint32_t *A, const uint32_t *cos16,
                                             uint32_t s, uint32_t b, std::integer_sequence<int, C...>)
{
  (p1_term<C>(acc, A, cos16, s, b), ...);
}

__global__ void __launch_bounds__(256, 3)
k_mfft16_p257(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, size_t slab_words,
              const uint32_t *__restrict__ g_cos16, const uint32_t *__restrict__ g_ctab)
{
  BCHFFT_DYN_SMEM(uint32_t, sm);
  uint32_t *A = sm, *V = sm + 16 * kAS, *cos16 = V + 16 * kAS, *ctab = cos16 + 256;
  const uint32_t t = threadIdx.x, k1 = blockIdx.x;
  const uint32_t *in = src + (size_t)blockIdx.y * slab_words;
  uint32_t *out = dst + (size_t)blockIdx.y * slab_words;
  for (uint32_t i = t; i < 256u; i += blockDim.x) cos16[i] = g_cos16[i];
  for (uint32_t i = t; i < 257u; i += blockDim.x) ctab[i] = g_ctab[i];
  for (uint32_t k2 = t; k2 < 257u; k2 += blockDim.x) {
    const uint32_t pos = (257u * k1 + 255u * k2) % (uint32_t)kN;
    const uint4 *e = reinterpret_cast<const uint4 *>(in + (size_t)pos * 16u);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const uint4 q = e[v];
      A[(4 * v + 0) * kAS + k2] = q.x;
      A[(4 * v + 1) * kAS + k2] = q.y;
      A[(4 * v + 2) * kAS + k2] = q.z;
      A[(4 * v + 3) * kAS + k2] = q.w;
    }
  }
  __syncthreads();
  // V_(s,b) = sum_c a_(k_s 2^((c - b) mod 16)) g^(2^c): every thread uses the same constant in every term
  for (uint32_t w = t; w < 256u; w += blockDim.x) {
    const uint32_t s = w >> 4, b = w & 15u;
    uint32_t acc[16];
#pragma unroll
    for (int p = 0; p < 16; ++p) acc[p] = 0u;
    p1_all_terms(acc, A, cos16, s, b, std::make_integer_sequence<int, 16>());
#pragma unroll
    for (int p = 0; p < 16; ++p) V[p * kAS + w] = acc[p];
  }
  __syncthreads();
  // A(w2^j) = a_0 + sum_(s,b) c_b(w2^(j k_s)) V_(s,b); two outputs per thread share every load of V
  for (uint32_t j0 = t; j0 < 129u; j0 += blockDim.x) {
    const uint32_t j1 = j0 + 129u;          // 129 .. 257: the last one does not exist
    uint32_t o0[16], o1[16];
#pragma unroll
    for (int p = 0; p < 16; ++p) o0[p] = o1[p] = A[p * kAS];
    for (uint32_t s = 0; s < 16u; ++s) {
      const uint32_t ks = cos16[s * 16u];
      const uint32_t c0 = ctab[(j0 * ks) % 257u], c1 = ctab[(j1 * ks) % 257u];
#pragma unroll 4
      for (uint32_t b = 0; b < 16u; ++b) {
        const uint32_t m0 = 0u - ((c0 >> b) & 1u), m1 = 0u - ((c1 >> b) & 1u);
        const uint32_t *v = V + s * 16u + b;
#pragma unroll
        for (int p = 0; p < 16; ++p) {
          const uint32_t x = v[p * kAS];
          o0[p] ^= x & m0;
          o1[p] ^= x & m1;
        }
      }
    }
    uint4 *e0 = reinterpret_cast<uint4 *>(out + ((size_t)k1 * 257u + j0) * 16u);
#pragma unroll
    for (int v = 0; v < 4; ++v) e0[v] = make_uint4(o0[4 * v], o0[4 * v + 1], o0[4 * v + 2], o0[4 * v + 3]);
    if (j1 < 257u) {
      uint4 *e1 = reinterpret_cast<uint4 *>(out + ((size_t)k1 * 257u + j1) * 16u);
#pragma unroll
      for (int v = 0; v < 4; ++v) e1[v] = make_uint4(o1[4 * v], o1[4 * v + 1], o1[4 * v + 2], o1[4 * v + 3]);
    }
  }
}

// ---- pass 2: one 255-point transform (3 x 5 x 17 prime-factor) per CTA --------------------------------
// P-point stage along one axis of the [3][5][17] cube (linear index 85 a + 17 b + c): STRIDE = index step of
// the axis, output (j, rest).  out_j = x_0 + sum_(c = 1 .. P-1) W^c x_(c / j mod P)  (j != 0);  out_0 = sum x_r.
template <int P, uint32_t WLOG, int C>
__device__ __forceinline__ void p2_term(uint32_t (&acc)[16], const uint32_t *X, uint32_t base, uint32_t stride,
                                        uint32_t jinv)
{
  constexpr uint32_t W = cpow_x((uint32_t)(((uint64_t)WLOG * (uint32_t)(C + 1)) % (uint32_t)kN));   // root^(C+1)
  const uint32_t r = ((uint32_t)(C + 1) * jinv) % (uint32_t)P;
  uint32_t a[16];
#pragma unroll
  for (int p = 0; p < 16; ++p) a[p] = X[p * kAS + base + r * stride];
  mad_const<W>(acc, a);
}
template <int P, uint32_t WLOG, int... C>
__device__ __forceinline__ void p2_all_terms(uint32_t (&acc)[16], const uint32_t *X, uint32_t base, uint32_t stride,
                                             uint32_t jinv, std::integer_sequence<int, C...>)
{
  (p2_term<P, WLOG, C>(acc, X, base, stride, jinv), ...);
}

__device__ __forceinline__ uint32_t inv_mod_small(uint32_t j, uint32_t p)   // j^-1 mod p, p <= 17 prime, j != 0
{
  uint32_t r = 1;
  for (uint32_t e = 0; e + 2u < p; ++e) r = (r * j) % p;   // j^(p-2)
  return r;
}

// one stage; threads are ordered j-major so that the (multiplication-free) j = 0 outputs fill whole warps.
// idx(j, rest): linear cube index of output j at the other two coordinates `rest`.
template <int P, uint32_t WLOG>
__device__ __forceinline__ void p2_stage(const uint32_t *X, uint32_t *Y, uint32_t stride)
{
  constexpr uint32_t others = 255u / (uint32_t)P;
  for (uint32_t t = threadIdx.x; t < 255u; t += blockDim.x) {
    const uint32_t j = t / others, rest = t - j * others;
    // base index with the axis coordinate = 0
    uint32_t base;
    if (P == 17) base = rest * 17u;                               // rest = 5 a + b, axis c (stride 1)
    else if (P == 5) base = (rest / 17u) * 85u + rest % 17u;      // rest = 17 a + c, axis b (stride 17)
    else base = rest;                                             // rest = 17 b + c, axis a (stride 85)
    uint32_t acc[16];
    if (j == 0u) {
#pragma unroll
      for (int p = 0; p < 16; ++p) acc[p] = 0u;
      for (uint32_t r = 0; r < (uint32_t)P; ++r)
#pragma unroll
        for (int p = 0; p < 16; ++p) acc[p] ^= X[p * kAS + base + r * stride];
    } else {
#pragma unroll
      for (int p = 0; p < 16; ++p) acc[p] = X[p * kAS + base];
      p2_all_terms<P, WLOG>(acc, X, base, stride, inv_mod_small(j, (uint32_t)P), std::make_integer_sequence<int, P - 1>());
    }
#pragma unroll
    for (int p = 0; p < 16; ++p) Y[p * kAS + base + j * stride] = acc[p];
  }
}

// grid = (257, slabs), 256 threads.  src: pass-1 layout (position k1 * 257 + i2); dst position = final index i.
__global__ void __launch_bounds__(256, 4)
k_mfft16_p255(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, size_t slab_words)
{
  BCHFFT_DYN_SMEM(uint32_t, sm);
  uint32_t *X = sm, *Y = sm + 16 * kAS;
  const uint32_t i2 = blockIdx.x;
  const uint32_t *in = src + (size_t)blockIdx.y * slab_words;
  uint32_t *out = dst + (size_t)blockIdx.y * slab_words;
  for (uint32_t k1 = threadIdx.x; k1 < 255u; k1 += blockDim.x) {
    // k1 = (85 a + 51 b + 15 c) mod 255  <=>  a = k1 mod 3, b = k1 mod 5, c = 8 k1 mod 17
    const uint32_t idx = (k1 % 3u) * 85u + (k1 % 5u) * 17u + (8u * k1) % 17u;
    const uint4 *e = reinterpret_cast<const uint4 *>(in + ((size_t)k1 * 257u + i2) * 16u);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const uint4 q = e[v];
      X[(4 * v + 0) * kAS + idx] = q.x;
      X[(4 * v + 1) * kAS + idx] = q.y;
      X[(4 * v + 2) * kAS + idx] = q.z;
      X[(4 * v + 3) * kAS + idx] = q.w;
    }
  }
  __syncthreads();
  // roots: w1 = x^257;  axis c: w1^15,  axis b: w1^51,  axis a: w1^85
  p2_stage<17, 257u * 15u>(X, Y, 1u);
  __syncthreads();
  p2_stage<5, 257u * 51u>(Y, X, 17u);
  __syncthreads();
  p2_stage<3, 257u * 85u>(X, Y, 85u);
  __syncthreads();
  for (uint32_t t = threadIdx.x; t < 255u; t += blockDim.x) {
    const uint32_t a = t / 85u, b = (t % 85u) / 17u, c = t % 17u;
    const uint32_t i1 = (85u * a + 51u * b + 120u * c) % 255u;
    const uint32_t pos = (32896u * i1 + 32640u * i2) % (uint32_t)kN;
    uint4 *e = reinterpret_cast<uint4 *>(out + (size_t)pos * 16u);
#pragma unroll
    for (int v = 0; v < 4; ++v)
      e[v] = make_uint4(Y[(4 * v + 0) * kAS + t], Y[(4 * v + 1) * kAS + t], Y[(4 * v + 2) * kAS + t],
                        Y[(4 * v + 3) * kAS + t]);
  }
}

}  // namespace mf16
}  // namespace bchfft
