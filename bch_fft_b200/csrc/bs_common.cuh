// Bitsliced GF(2^m) primitives shared by every kernel.
//
// Representation ("slab"): 32 independent items (polynomials or codewords) are processed
// together.  One field element per item is held as M 32-bit plane words; bit l of plane b is
// bit b (coefficient of x^b, polynomial basis of libbase/finite_field.c:19-48 in the
// reference) of item l's element.  Field addition is a plane-wise XOR, multiplication is
// AND/XOR logic only (LOP3 on sm_100a): no log/exp tables, no bank conflicts, no branches.
#pragma once
#include <stdint.h>

#ifdef BCHFFT_EMU
// tools/cudaemu/cudaemu.h is force-included by the emulator build
#else
#include <cuda_runtime.h>
#define BCHFFT_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define BCHFFT_DYN_SMEM(type, name) \
  extern __shared__ __align__(16) unsigned char bchfft_dyn_smem_raw[]; \
  type *name = reinterpret_cast<type *>(bchfft_dyn_smem_raw)
#endif

// Launch shapes of the shared-memory tile kernels (threads per CTA, minimum resident CTAs per SM
// the kernel is compiled for -- the latter bounds registers per thread).  Measured on B200:
//  * k_gm_pass keeps one 128 KB tile per SM, so it wants as many threads as registers allow;
//  * k_locate / k_syndrome use half-size tiles, two CTAs of 256 threads per SM: while one CTA
//    sits in a load phase or at a barrier the other one keeps the integer pipe busy.
#ifndef BCHFFT_NT_PASS
#define BCHFFT_NT_PASS 512
#endif
#ifndef BCHFFT_MINB_PASS
#define BCHFFT_MINB_PASS 1
#endif
// Taylor-only passes: resident CTAs per SM and Taylor levels fused per sweep (2^(levels+1)
// registers of data per thread)
#ifndef BCHFFT_MINB_TY
#define BCHFFT_MINB_TY 3
#endif
#ifndef BCHFFT_TY_LEVELS
#define BCHFFT_TY_LEVELS 4
#endif
// one-plane Taylor passes of the plane-major path: tile bits and resident CTAs per SM
#ifndef BCHFFT_PLANE_T
#define BCHFFT_PLANE_T 14
#endif
#ifndef BCHFFT_MINB_PLANE
#define BCHFFT_MINB_PLANE 2
#endif
#ifndef BCHFFT_NT_LOC
#define BCHFFT_NT_LOC 256
#endif
#ifndef BCHFFT_MINB_LOC
#define BCHFFT_MINB_LOC 2
#endif
// butterfly sweeps of the FFT pass kernels use the branch-over-set-bits multiply where the
// multiplier is warp-uniform (tile_butterfly*, UNI)
// radix-4 butterfly sweeps (two levels per shared-memory round trip, four elements in registers)
#ifndef BCHFFT_BF_R4
#define BCHFFT_BF_R4 1
#endif
#ifndef BCHFFT_UNI_BF
#define BCHFFT_UNI_BF 1
#endif
// radix-4 sweeps of the FFT pass kernels: the two butterflies that share a multiplier in one branch
// sequence (bs_mad_uniform2; more registers live at once)
// smallest base 2^h whose runs of generalised Taylor levels take the line sweep (up to six levels per
// shared-memory round trip) instead of the two-level register programs
#ifndef BCHFFT_TX_LINES_MIN_H
#define BCHFFT_TX_LINES_MIN_H 5
#endif
#ifndef BCHFFT_R4_MAD2
#define BCHFFT_R4_MAD2 0
#endif
#ifndef BCHFFT_MINB_SYNP
#define BCHFFT_MINB_SYNP 3
#endif
#ifndef BCHFFT_MINB_LOCU
#define BCHFFT_MINB_LOCU 3
#endif
// k_elp_fwd_u / k_elp_fwd wait on chains of dependent global loads (list -> codeword -> degree / coefficients):
// resident CTAs per SM the kernels are compiled for.  Measured, k_elp_fwd_u per 10^6 words: 1 CTA (254 registers
// per thread, 8 warps per SM) 0.32 ms, 2: 0.18, 3: 0.15, 4 (64 registers): 0.13 ms.
#ifndef BCHFFT_MINB_ELPU
#define BCHFFT_MINB_ELPU 4
#endif
#ifndef BCHFFT_NT_SYN
#define BCHFFT_NT_SYN 256
#endif
#ifndef BCHFFT_MINB_SYN
#define BCHFFT_MINB_SYN 2
#endif

namespace bchfft {

// Reduction polynomials of the reference's field table (libbase/finite_field.c:9-18,32-35),
// x^m term included; index = m.  m >= 26 are not primitive there and are not supported here.
__host__ __device__ constexpr uint32_t field_poly(int m)
{
  constexpr uint32_t t[26] = {0,        0,        0x7,      0xB,      0x13,     0x25,    0x43,
                              0x83,     0x171,    0x211,    0x409,    0x805,    0x1099,  0x201B,
                              0x5803,   0x8003,   0x1002D,  0x20009,  0x40081,  0x80063, 0x100009,
                              0x200005, 0x400003, 0x800021, 0x100001B, 0x2000009};
  return (m >= 2 && m <= 25) ? t[m] : 0;
}

// plane words per element in HBM slabs: M rounded up to a multiple of 4 so that one element is
// a whole number of 16-byte vectors.  Build option: a multiple of 8 above GF(2^BCHFFT_PS_ALIGN8_ABOVE), so
// that an element is a whole number of 32-byte sectors (an 80-byte element of GF(2^20) straddles 3.5 sectors
// on average) and the Taylor-only passes take 8-plane groups (k_ty_pass).  Measured at m = 20 with the
// option at 16: 22.3 ms per 256 polynomials against 19.2 ms -- 20 % more bytes per pass cost more than the
// aligned sectors give back; off by default.
#ifndef BCHFFT_PS_ALIGN8_ABOVE
#define BCHFFT_PS_ALIGN8_ABOVE 99
#endif
__host__ __device__ constexpr int plane_stride(int m)
{
  return m > BCHFFT_PS_ALIGN8_ABOVE ? ((m + 7) & ~7) : ((m + 3) & ~3);
}

// three-input XOR: one LOP3
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c)
{
#ifdef BCHFFT_EMU
  return a ^ b ^ c;
#else
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#endif
}

// p[0 .. 2M-2] -> p[0 .. M-1] modulo the field polynomial.  Gather form: plane j receives plane
// j + M - e for every tap e of the polynomial (x^M = sum_e x^e); going from the top plane down,
// every source is already final when it is used.  The (at most five) inputs of a plane are
// folded with three-input XORs: 2 LOP3 per plane for the pentanomials, 1 for the trinomials.
template <int M>
__device__ __forceinline__ void bs_reduce(uint32_t (&p)[2 * M - 1])
{
  constexpr uint32_t taps = field_poly(M) & ((1u << M) - 1u);
#pragma unroll
  for (int j = 2 * M - 3; j >= 0; --j) {
    uint32_t acc = p[j], pend = 0u;
    bool have = false;
#pragma unroll
    for (int e = 0; e < M; ++e) {
      const int k = j + M - e;
      if (((taps >> e) & 1u) && k >= M && k <= 2 * M - 2) {
        if (have) {
          acc = xor3(acc, pend, p[k]);
          have = false;
        } else {
          pend = p[k];
          have = true;
        }
      }
    }
    if (have) acc ^= pend;
    p[j] = acc;
  }
}

// Products a[i] * c_j accumulate into p[i+j].  Two instruction mixes per partial product:
//   "alu": p ^= a & mask_j        one LOP3 on the alu pipe (mask_j = 0 / ~0 from bit j of c)
//   "fma": p ^= a * bit_j         one IMAD on the fma pipe + XOR; pairs of such XORs fuse into one
//                                 3-input LOP3, so the alu pipe sees half an op per product
// The alu and fma pipes issue independently (each one warp instruction per two cycles per
// SMSP on sm_100), so splitting the M*M partial products between them raises the bound from
// alu-only throughput.  BCHFFT_ALU_ROWS rows (values of j) use the alu mix, the rest the fma mix.
#ifndef BCHFFT_ALU_ROWS
#define BCHFFT_ALU_ROWS 32
#endif

template <int M>
__device__ __forceinline__ void bs_accumulate_products(uint32_t (&p)[2 * M - 1], const uint32_t (&a)[M],
                                                       uint32_t c)
{
#pragma unroll
  for (int j = 0; j < M; ++j) {
    const uint32_t bit = (c >> j) & 1u;
    if (j < BCHFFT_ALU_ROWS) {
      // 0 / ~0 from bit j: shift it to the sign position, arithmetic shift back
      const uint32_t mask = (uint32_t)((int32_t)(c << (31 - j)) >> 31);
#pragma unroll
      for (int i = 0; i < M; ++i) p[i + j] ^= a[i] & mask;
    } else {
#pragma unroll
      for (int i = 0; i < M; ++i) p[i + j] ^= a[i] * bit;
    }
  }
}

// a <- a * c, c one field element shared by the 32 items of the slab (every additive-FFT
// multiplier is such a data-independent constant): M*M partial products + reduction.
template <int M>
__device__ __forceinline__ void bs_mul_scalar(uint32_t (&a)[M], uint32_t c)
{
  uint32_t p[2 * M - 1];
#pragma unroll
  for (int k = 0; k < 2 * M - 1; ++k) p[k] = 0u;
  bs_accumulate_products<M>(p, a, c);
  bs_reduce<M>(p);
#pragma unroll
  for (int i = 0; i < M; ++i) a[i] = p[i];
}

// acc <- acc + a * c
template <int M>
__device__ __forceinline__ void bs_mad_scalar(uint32_t (&acc)[M], const uint32_t (&a)[M], uint32_t c)
{
  uint32_t p[2 * M - 1];
#pragma unroll
  for (int k = 0; k < 2 * M - 1; ++k) p[k] = (k < M) ? acc[k] : 0u;
  bs_accumulate_products<M>(p, a, c);
  bs_reduce<M>(p);
#pragma unroll
  for (int i = 0; i < M; ++i) acc[i] = p[i];
}

// acc <- acc + a * c for a constant c that is the same in every lane of the warp (kernels whose
// lanes are 32 slabs at one position).  The bits of c are then warp-uniform branch conditions:
// only the set bits cost anything (no masks), two bits are taken per branch and a pair of set
// bits is folded with one three-input XOR per plane.  Measured (tools/bench/mulbench.cu, B200):
// 278 clk per warp multiply at m = 13 against 437 for the masked form, 381 against 636 at m = 16.
template <int M>
__device__ __forceinline__ void bs_mad_uniform(uint32_t (&acc)[M], const uint32_t (&a)[M], uint32_t c)
{
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
}

// two multiply-adds with the same warp-uniform constant (the two butterflies of a radix-4 block
// that share a multiplier): one branch sequence, twice the independent work per branch
template <int M>
__device__ __forceinline__ void bs_mad_uniform2(uint32_t (&acc0)[M], const uint32_t (&a0)[M],
                                                uint32_t (&acc1)[M], const uint32_t (&a1)[M], uint32_t c)
{
  uint32_t p[2 * M], r[2 * M];
#pragma unroll
  for (int k = 0; k < 2 * M; ++k) {
    p[k] = (k < M) ? acc0[k] : 0u;
    r[k] = (k < M) ? acc1[k] : 0u;
  }
#pragma unroll
  for (int j = 0; j < M; j += 2) {
    const uint32_t two = (c >> j) & 3u;
    if (j + 1 == M) {
      if (two & 1u) {
#pragma unroll
        for (int i = 0; i < M; ++i) {
          p[i + j] ^= a0[i];
          r[i + j] ^= a1[i];
        }
      }
    } else if (two == 1u) {
#pragma unroll
      for (int i = 0; i < M; ++i) {
        p[i + j] ^= a0[i];
        r[i + j] ^= a1[i];
      }
    } else if (two == 2u) {
#pragma unroll
      for (int i = 0; i < M; ++i) {
        p[i + j + 1] ^= a0[i];
        r[i + j + 1] ^= a1[i];
      }
    } else if (two == 3u) {
      p[j] ^= a0[0];
      r[j] ^= a1[0];
#pragma unroll
      for (int i = 1; i < M; ++i) {
        p[i + j] = xor3(p[i + j], a0[i], a0[i - 1]);
        r[i + j] = xor3(r[i + j], a1[i], a1[i - 1]);
      }
      p[j + M] ^= a0[M - 1];
      r[j + M] ^= a1[M - 1];
    }
  }
  uint32_t q[2 * M - 1];
#pragma unroll
  for (int k = 0; k < 2 * M - 1; ++k) q[k] = p[k];
  bs_reduce<M>(q);
#pragma unroll
  for (int i = 0; i < M; ++i) acc0[i] = q[i];
#pragma unroll
  for (int k = 0; k < 2 * M - 1; ++k) q[k] = r[k];
  bs_reduce<M>(q);
#pragma unroll
  for (int i = 0; i < M; ++i) acc1[i] = q[i];
}

// In-place transpose of a 32x32 bit matrix held as 32 words, LSB-first: afterwards bit j of
// x[i] is the old bit i of x[j].  Converts 32 per-item values <-> 32 plane words.
__device__ __forceinline__ void transpose32(uint32_t (&x)[32])
{
#pragma unroll
  for (int s = 4; s >= 0; --s) {
    const int j = 1 << s;
    const uint32_t m = (s == 4) ? 0x0000FFFFu : (s == 3) ? 0x00FF00FFu : (s == 2) ? 0x0F0F0F0Fu
                     : (s == 1) ? 0x33333333u : 0x55555555u;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if (k & j) continue;
      const uint32_t t = ((x[k] >> j) ^ x[k + j]) & m;
      x[k + j] ^= t;
      x[k] ^= t << j;
    }
  }
}

// scalar (per-thread, table-free) GF(2^m) multiply used where an operand is data, e.g. the
// Berlekamp-Massey recurrence: shift-and-add with interleaved reduction.
template <int M>
__host__ __device__ __forceinline__ uint32_t gf_mul(uint32_t a, uint32_t b)
{
  constexpr uint32_t poly = field_poly(M);
  uint32_t r = 0;
#pragma unroll
  for (int i = M - 1; i >= 0; --i) {
    r <<= 1;
    if (r & (1u << M)) r ^= poly;
    if ((b >> i) & 1u) r ^= a;
  }
  return r;
}

}  // namespace bchfft
