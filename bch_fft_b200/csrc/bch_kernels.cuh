// Batched binary-BCH decode kernels (replace libbch/bch.c:228-562 of the reference run once per
// codeword).  A slab is 32 codewords; bit l of a plane word belongs to codeword l of the slab.
//
//   k_syndrome_p   packed codewords -> bitsliced odd syndromes S_1, S_3, .. (bch.c:228-326) through
//                  the universal field-polynomial LFSR on permuted positions (all j coprime to the order)
//   k_syndrome     same result, general path: one masked LFSR per class
//   k_syn_finish   un-bitslice, S_2j = S_j^2, zero test (stage-level API; the decode pipeline lets
//                  k_bm do it)
//   k_bm           Berlekamp-Massey, one thread or a lane group per codeword (bch.c:328-404), shared-memory
//                  resident for small fields; k_bm_wide: one CTA per codeword
//   k_bucket_count / k_bucket_scatter   counting sort by locator degree; degrees 1 and 2 solved in closed form
//   k_elp_fwd[_u]  forward phase of the truncated FFT on the locator coefficients
//   k_locate_u     error-locator evaluation at every field point, K <= 7 levels, warp lanes = slabs,
//                  warp-uniform multipliers (bch.c:406-490)
//   k_locate       same result, general path (tile engine, lanes = positions)
//   k_correct      ok / corrected bookkeeping and the bit flips (bch.c:531-561)
//   k_encode / k_encode_wide   systematic encoder (bch.c:492-506)
#pragma once
#include <type_traits>
#include "bs_common.cuh"
#include "fft_kernels.cuh"

namespace bchfft {

constexpr int kSynSub = 256;  // positions per LFSR work item

// ---- syndromes ------------------------------------------------------------------------------
// S_j = c(x^j) = (c mod Q_j)(x^j) for any multiple Q_j of the minimal polynomial of x^j.  Every
// work item (odd j, sub-chunk q) runs a bitsliced LFSR of degree M over kSynSub positions
// (M LOP3 per position for 32 codewords), then folds its remainder r into the syndrome with
// the constants x^(j (b + q kSynSub)), b < M: S_j += sum_b r_b x^(j(b + off)).
//
// grid = (position chunks, slabs); chunk = `chunk_pos` positions (multiple of kSynSub, <= the
// shared-memory plane buffer); acc: [slab][nodd][M] plane words, zero-initialised, XOR-merged;
// parity: [slab] plane word = syndrome[0] of the reference: XOR of the codeword bits selected by
// `vbits` -- all ones on the reference's FFT path (c(1)), the bits (X^i mod g)(1) on its direct
// path, where it evaluates the remainder r = c mod g at 1 (bch.c:258-262, 301-314).
template <int M>
__global__ void __launch_bounds__(BCHFFT_NT_SYN, BCHFFT_MINB_SYN)
k_syndrome(const uint32_t *__restrict__ words32, uint32_t words32_per_cw, uint32_t length_bits,
           uint32_t batch, uint32_t chunk_pos, uint32_t nodd, const uint32_t *__restrict__ taps,
           const uint32_t *__restrict__ fold, uint32_t nsub_total, const uint32_t *__restrict__ vbits,
           uint32_t *__restrict__ acc, uint32_t *__restrict__ parity)
{
  BCHFFT_DYN_SMEM(uint32_t, smem);
  const uint32_t slab = blockIdx.y;
  const uint32_t pos0 = blockIdx.x * chunk_pos;
  const uint32_t nsub = chunk_pos / kSynSub;
  // planes[p + p / 32]: one word of padding per 32 positions keeps both the transposing stores
  // of phase A (lanes 32 positions apart) and the loads of phase B (lanes of different
  // sub-chunks) on distinct banks
  uint32_t *planes = smem;
  uint32_t *sacc = smem + chunk_pos + chunk_pos / 32;  // [nodd][M]
  __shared__ uint32_t s_par;
  if (threadIdx.x == 0) s_par = 0u;
  for (uint32_t i = threadIdx.x; i < nodd * M; i += blockDim.x) sacc[i] = 0u;
  __syncthreads();

  // phase A: transpose 32 codewords x 32 positions blocks into plane words
  uint32_t par = 0u;
  for (uint32_t g = threadIdx.x; g < chunk_pos / 32; g += blockDim.x) {
    const uint32_t w = pos0 / 32 + g;
    uint32_t x[32];
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      const uint32_t item = slab * 32u + l;
      x[l] = (item < batch && w < words32_per_cw) ? words32[(size_t)item * words32_per_cw + w] : 0u;
    }
    // bits at positions >= length do not belong to the codeword (bch.c:258-262, 278)
    const uint32_t first = w * 32u;
    if (first + 32u > length_bits) {
      const uint32_t keep = first < length_bits ? length_bits - first : 0u;
      const uint32_t mask = keep ? (0xFFFFFFFFu >> (32u - keep)) : 0u;
#pragma unroll
      for (int l = 0; l < 32; ++l) x[l] &= mask;
    }
    transpose32(x);
    const uint32_t vw = (w < words32_per_cw) ? vbits[w] : 0u;
#pragma unroll
    for (int b = 0; b < 32; ++b) {
      planes[g * 33u + b] = x[b];
      par ^= x[b] & (0u - ((vw >> b) & 1u));
    }
  }
  if (par) atomicXor(&s_par, par);
  __syncthreads();

  // phase B: LFSR per (odd j, sub-chunk)
  for (uint32_t item = threadIdx.x; item < nodd * nsub; item += blockDim.x) {
    const uint32_t q = item / nodd, ji = item - q * nodd;
    const uint32_t tp = taps[ji];
    uint32_t tm[M], s[M];
#pragma unroll
    for (int b = 0; b < M; ++b) {
      tm[b] = 0u - ((tp >> b) & 1u);
      s[b] = 0u;
    }
    const uint32_t *src = planes + q * (kSynSub + kSynSub / 32);
    for (int p = kSynSub - 1; p >= 0; --p) {
      const uint32_t in = src[p + (p >> 5)];
      const uint32_t fb = s[M - 1];
#pragma unroll
      for (int b = M - 1; b >= 1; --b) s[b] = s[b - 1] ^ (fb & tm[b]);
      s[0] = in ^ (fb & tm[0]);
    }
    // fold: out = sum_b s[b] * fold[ji][q_global][b]
    const uint32_t qg = blockIdx.x * nsub + q;
    const uint32_t *fc = fold + ((size_t)ji * nsub_total + qg) * M;
    uint32_t out[M];
#pragma unroll
    for (int k = 0; k < M; ++k) out[k] = 0u;
#pragma unroll
    for (int b = 0; b < M; ++b) {
      const uint32_t c = fc[b];
#pragma unroll
      for (int k = 0; k < M; ++k) out[k] ^= s[b] & (0u - ((c >> k) & 1u));
    }
#pragma unroll
    for (int k = 0; k < M; ++k)
      if (out[k]) atomicXor(&sacc[ji * M + k], out[k]);
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < nodd * M; i += blockDim.x)
    if (sacc[i]) atomicXor(&acc[(size_t)slab * nodd * M + i], sacc[i]);
  if (threadIdx.x == 0 && s_par) atomicXor(&parity[slab], s_par);
}


// ---- syndromes through the universal LFSR (orders coprime to every j) -------------------------
// When gcd(j, 2^M - 1) = 1 the map p -> e = j p mod (2^M - 1) permutes the positions, and
//   S_j = sum_p c_p x^(j p) = ( sum_e c_(p(e)) X^e  mod  f(X) )(x),      p(e) = e j^-1 mod (2^M - 1),
// f the field polynomial.  So every class runs the SAME LFSR -- division by f, whose taps are
// compile-time constants (4 LOP3 per position for a pentanomial instead of M masked ones) -- on
// the codeword read in permuted order, and the remainder IS the syndrome in the field's
// polynomial basis.  The whole codeword of a slab sits in shared memory as plane words
// (padded: word p + p/32); work item = (class, chunk of lc = ceil(order/32) consecutive e), the 32
// lanes of a warp are the 32 chunks of one class; a chunk's remainder is shifted by x^(q lc) with
// the fold constants.  Bank conflicts: the lanes of a warp read positions that differ by multiples
// of D = lc j^-1 and all move by the same step, so the conflict pattern is a property of (j, lc).
// The host picks, per class, the chunk length lc and the conjugate j' = j 2^i to work with
// (S_j = S_j'^(2^(M-i)); the Frobenius power is GF(2)-linear, so it is folded into the fold
// constants at no cost) that minimise the simulated number of shared-memory wavefronts.
template <int M>
__device__ __forceinline__ void lfsr_field_step(uint32_t (&s)[M], uint32_t in)
{
  constexpr uint32_t taps = field_poly(M) & ((1u << M) - 1u);
  const uint32_t fb = s[M - 1];
#pragma unroll
  for (int b = M - 1; b >= 1; --b) s[b] = ((taps >> b) & 1u) ? (s[b - 1] ^ fb) : s[b - 1];
  s[0] = in ^ fb;   // bit 0 of the field polynomial is always set
}

// A CTA takes TWO slabs: their plane words are interleaved (uint2 per position), so that one
// index update and one 64-bit shared load feed two LFSRs.  The position index is kept as a byte
// offset (8 p), so a step costs three integer operations and one LDS.64 besides the two LFSRs.
// grid = ceil(slabs / 2); dynamic smem = 2 * (32 * ceil(order / 32) + nodd * M) words.
//
// Phase A (transpose into plane words): thread = (slab, group of 32 positions); the 32 lanes of a
// warp own 32 consecutive groups, i.e. addresses 256 bytes apart, so every lane stores its plane
// words in an order rotated by its lane number (a five-stage barrel rotation of the register array).
// Phase B: the 32 chunks of one class are the 32 lanes of one warp, so the class's remainder is one
// warp XOR-reduction per plane; nothing else ever writes that accumulator.
__device__ __forceinline__ uint32_t warp_xor(uint32_t v)
{
#ifdef BCHFFT_EMU
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v ^= __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
#else
  return __reduce_xor_sync(0xFFFFFFFFu, v);
#endif
}

template <int M>
__global__ void __launch_bounds__(BCHFFT_NT_SYN, BCHFFT_MINB_SYNP)
k_syndrome_p(const uint32_t *__restrict__ words32, uint32_t words32_per_cw, uint32_t length_bits,
             uint32_t batch, uint32_t nodd, const uint32_t *__restrict__ lcs, const uint32_t *__restrict__ jinv,
             const uint32_t *__restrict__ start_idx, const uint32_t *__restrict__ foldp,
             const uint32_t *__restrict__ vbits, uint32_t *__restrict__ acc, uint32_t *__restrict__ parity)
{
  BCHFFT_DYN_SMEM(uint32_t, smem);
  constexpr uint32_t order = (1u << M) - 1u;
  const uint32_t slab0 = blockIdx.x * 2u;
  const uint32_t nslabs = (batch + 31u) >> 5;
  const uint32_t ngrp = (order + 31u) / 32u;          // 32-position groups of the codeword
  uint2 *planes = reinterpret_cast<uint2 *>(smem);
  uint32_t *sacc = smem + 2u * ngrp * 32u;             // [2][nodd][M]
  __shared__ uint32_t s_par[2];
  if (threadIdx.x < 2) s_par[threadIdx.x] = 0u;
  __syncthreads();

  // phase A: transpose 32 codewords x 32 positions blocks into plane words (both slabs)
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t w = threadIdx.x; w < 2u * ngrp; w += blockDim.x) {
    const uint32_t which = w >= ngrp ? 1u : 0u, g = w - which * ngrp;
    const uint32_t slab = slab0 + which;
    uint32_t x[32];
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      const uint32_t item = slab * 32u + l;
      x[l] = (item < batch && g < words32_per_cw) ? words32[(size_t)item * words32_per_cw + g] : 0u;
    }
    const uint32_t first = g * 32u;
    if (first + 32u > length_bits) {
      const uint32_t keep = first < length_bits ? length_bits - first : 0u;
      const uint32_t mask = keep ? (0xFFFFFFFFu >> (32u - keep)) : 0u;
#pragma unroll
      for (int l = 0; l < 32; ++l) x[l] &= mask;
    }
    transpose32(x);
    const uint32_t vw = (g < words32_per_cw) ? vbits[g] : 0u;
    uint32_t par = 0u;
#pragma unroll
    for (int b = 0; b < 32; ++b) par ^= x[b] & (0u - ((vw >> b) & 1u));
    if (par) atomicXor(&s_par[which], par);
    // x[k] <- x[(k + lane) & 31]
#pragma unroll
    for (int sft = 0; sft < 5; ++sft) {
      const bool on = (lane >> sft) & 1u;
      uint32_t y[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) y[k] = on ? x[(k + (1 << sft)) & 31] : x[k];
#pragma unroll
      for (int k = 0; k < 32; ++k) x[k] = y[k];
    }
    uint32_t *dst = reinterpret_cast<uint32_t *>(planes + g * 32u) + which;
#pragma unroll
    for (int k = 0; k < 32; ++k) dst[2u * ((k + lane) & 31u)] = x[k];
  }
  __syncthreads();

  // phase B: uniform trip count for every thread (a warp is active as a whole)
  const unsigned char *pbytes = reinterpret_cast<const unsigned char *>(planes);
  constexpr uint32_t order8 = order * 8u;
  for (uint32_t base = 0; base < nodd * 32u; base += blockDim.x) {
    const uint32_t item = base + threadIdx.x;
    const bool active = item < nodd * 32u;
    const uint32_t ji = item >> 5, q = item & 31u;
    uint32_t o0[M], o1[M];
#pragma unroll
    for (int k = 0; k < M; ++k) o0[k] = o1[k] = 0u;
    if (active) {
      const uint32_t lc = lcs[ji];
      const uint32_t e0 = q * lc, e1 = (e0 + lc < order) ? e0 + lc : order;   // chunk [e0, e1)
      const uint32_t nsteps = e1 > e0 ? e1 - e0 : 0u;
      const uint32_t step8 = jinv[ji] * 8u;
      uint32_t idx8 = start_idx[item] * 8u;        // 8 p(e1 - 1)
      uint32_t s0[M], s1[M];
#pragma unroll
      for (int b = 0; b < M; ++b) s0[b] = s1[b] = 0u;
      // descending e: s <- s X + c_(p(e));  p(e - 1) = p(e) - j^-1 mod order
      // (all chunks of a class but the last have the same length, so the trip counts are warp-uniform).
      // The main loop takes M steps per trip: after M steps the rotating register names of the LFSR
      // are back where they started, so the loop carries no register moves.
      const uint32_t pro = nsteps % (uint32_t)M, main_steps = nsteps - pro;
      for (uint32_t k = 0; k < pro; ++k) {
        const uint2 in = *reinterpret_cast<const uint2 *>(pbytes + idx8);
        lfsr_field_step<M>(s0, in.x);
        lfsr_field_step<M>(s1, in.y);
        const uint32_t t = idx8 - step8, t2 = t + order8;
        idx8 = t < t2 ? t : t2;   // unsigned min: t wrapped around 2^32 exactly when idx8 < step8
      }
      for (uint32_t k = 0; k < main_steps; k += (uint32_t)M) {
#pragma unroll
        for (int u = 0; u < M; ++u) {
          const uint2 in = *reinterpret_cast<const uint2 *>(pbytes + idx8);
          lfsr_field_step<M>(s0, in.x);
          lfsr_field_step<M>(s1, in.y);
          const uint32_t t = idx8 - step8, t2 = t + order8;
          idx8 = t < t2 ? t : t2;
        }
      }
      // shift by x^(q lc): out = sum_b s[b] * x^(q lc + b)
      const uint32_t *fc = foldp + (size_t)item * M;
#pragma unroll
      for (int b = 0; b < M; ++b) {
        const uint32_t c = fc[b];
#pragma unroll
        for (int k = 0; k < M; ++k) {
          const uint32_t mask = 0u - ((c >> k) & 1u);
          o0[k] ^= s0[b] & mask;
          o1[k] ^= s1[b] & mask;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < M; ++k) {
      const uint32_t r0 = warp_xor(o0[k]), r1 = warp_xor(o1[k]);
      if (active && q == 0u) {
        sacc[ji * M + k] = r0;
        sacc[(nodd + ji) * M + k] = r1;
      }
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < 2u * nodd * M; i += blockDim.x) {
    const uint32_t which = i >= nodd * M ? 1u : 0u;
    if (slab0 + which < nslabs) acc[(size_t)(slab0 + which) * nodd * M + (i - which * nodd * M)] = sacc[i];
  }
  if (threadIdx.x < 2 && slab0 + threadIdx.x < nslabs) parity[slab0 + threadIdx.x] = s_par[threadIdx.x];
}

// One thread per codeword: syndromes[cw][0..d-1] (u32), flag[cw] (bch.c:317-324).
// syndrome[0] (not a BCH syndrome; method dependent in the reference, see k_syndrome) comes from
// the parity word.
template <int M>
__global__ void k_syn_finish(const uint32_t *__restrict__ acc, const uint32_t *__restrict__ parity,
                             uint32_t nodd, uint32_t d, uint32_t batch,
                             uint32_t *__restrict__ syn, int32_t *__restrict__ flag)
{
  const uint32_t cw = blockIdx.x * blockDim.x + threadIdx.x;
  if (cw >= batch) return;
  const uint32_t slab = cw >> 5, l = cw & 31u;
  uint32_t *s = syn + (size_t)cw * d;
  uint32_t any = 0u;
  s[0] = (parity[slab] >> l) & 1u;
  any |= s[0];
  for (uint32_t i = 1; i < d; ++i) {
    uint32_t v;
    if (i & 1u) {
      const uint32_t *a = acc + ((size_t)slab * nodd + (i >> 1)) * M;
      v = 0u;
#pragma unroll
      for (int b = 0; b < M; ++b) v |= ((a[b] >> l) & 1u) << b;
    } else {
      const uint32_t h = s[i >> 1];
      v = gf_mul<M>(h, h);
    }
    s[i] = v;
    any |= v;
  }
  flag[cw] = any ? 1 : 0;
}

// ---- syndromes of few long codewords through the additive FFT ---------------------------------
// The reference's own choice for large d (bch.c:278-292): unpack the word to one coefficient per
// bit, evaluate it at every x^i with one additive FFT, the syndromes are entries 0 .. d-1.  The
// LFSR kernels above cost (positions x classes) per slab even when the slab holds one codeword;
// for a handful of codewords with many classes the transform is cheaper (host rule: syn_fft()).
// k_unpack_bits: grid = (ceil(n / 256), batch); coef[cw][p] = bit p of the word, 0 for p >= length.
__global__ void k_unpack_bits(const uint32_t *__restrict__ words32, uint32_t words32_per_cw, uint32_t length,
                              uint32_t n, uint32_t *__restrict__ coef)
{
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x, cw = blockIdx.y;
  if (p >= n) return;
  uint32_t v = 0u;
  if (p < length) v = (words32[(size_t)cw * words32_per_cw + (p >> 5)] >> (p & 31u)) & 1u;
  coef[(size_t)cw * n + p] = v;
}

// One CTA per codeword: syn[cw][j] = P(x^j) from the transform for j = 1 .. d-1, syn[cw][0] = the
// reference's method-dependent entry (XOR of the word bits selected by vbits, see k_syndrome),
// flag[cw] = any of them non-zero (bch.c:317-324).
__global__ void k_syn_from_fft(const uint32_t *__restrict__ res, uint32_t res_stride,
                               const uint32_t *__restrict__ words32, uint32_t words32_per_cw,
                               const uint32_t *__restrict__ vbits, uint32_t d,
                               uint32_t *__restrict__ syn, int32_t *__restrict__ flag)
{
  __shared__ uint32_t s_par, s_any;
  const uint32_t cw = blockIdx.x;
  if (threadIdx.x == 0) {
    s_par = 0u;
    s_any = 0u;
  }
  __syncthreads();
  uint32_t par = 0u;
  for (uint32_t w = threadIdx.x; w < words32_per_cw; w += blockDim.x)
    par ^= words32[(size_t)cw * words32_per_cw + w] & vbits[w];
  par = (uint32_t)__popc(par) & 1u;
  if (par) atomicXor(&s_par, 1u);
  uint32_t any = 0u;
  for (uint32_t j = 1u + threadIdx.x; j < d; j += blockDim.x) {
    const uint32_t v = res[(size_t)cw * res_stride + j];
    syn[(size_t)cw * d + j] = v;
    any |= v;
  }
  if (any) atomicOr(&s_any, 1u);
  __syncthreads();
  if (threadIdx.x == 0) {
    syn[(size_t)cw * d] = s_par;
    flag[cw] = (s_any | s_par) ? 1 : 0;
  }
}

// ---- Berlekamp-Massey -----------------------------------------------------------------------
// One thread per codeword, faithful to bch.c:328-404 including the rotation of the three
// coefficient buffers (so that C[deg C + 1 .. L], which the reference leaves stale, is
// reproduced from zero-initialised buffers).  bm: [3][d+1][batch] scratch, zero on entry
// (coefficient-major: threads of a warp touch consecutive words).  Outputs: L[cw] and
// elp[cw][0..d] = the final C buffer.
// Both operands of the multiplies are data here, so this is the one place that uses the
// reference's log/exp formulation (finite_field.c:59-70).
//   SM = true  (fields up to GF(2^14), d small enough): everything a thread touches lives in
//        shared memory -- 16-bit log / exp tables staged by the CTA, the three coefficient buffers
//        and the syndromes as [index][thread] columns (conflict-free) -- so the recurrence runs at
//        shared-memory latency.  dynamic smem = 2 * (2^(M+1) + (4 d + 3) * threads) bytes.
//   SM = false: coefficient buffers in the zero-initialised global scratch bm: [3][d+1][batch]
//        (coefficient-major), table-free shift-and-add multiply.
//        With acc != nullptr (decode pipeline) the kernel also does k_syn_finish's job: it builds
//        its syndrome column straight from the bitsliced accumulators (S_2j = S_j^2 through the
//        tables), writes flag[], and the global syndrome array is never touched.
// LPC = lanes per codeword (SM only; a power of two up to 32): the LPC lanes of a group split the
// coefficient index j modulo LPC -- 1/LPC of the sequential work per codeword and 1/LPC of the shared
// memory per thread, so that the (4 d + 3) 16-bit entries per codeword still fit next to the tables when t
// is large (t = 64: LPC = 8, t = 200: LPC = 16; without that those codes fall back to global scratch and
// run at L2 latency).  One XOR reduction (discrepancy) and at most one max reduction (degree after
// cancellation) per iteration go through warp shuffles.  `cs`, the stride between coefficient rows, is
// padded by the host so that the LPC rows one warp touches fall on disjoint banks.
template <int M, bool SM, int LPC>
__global__ void k_bm(const uint32_t *__restrict__ syn, int32_t *__restrict__ flag, uint32_t d,
                     uint32_t batch, const uint32_t *__restrict__ gexp, const uint32_t *__restrict__ glog,
                     uint32_t *__restrict__ bm, uint32_t *__restrict__ elp, uint32_t *__restrict__ Lout,
                     const uint32_t *__restrict__ acc, const uint32_t *__restrict__ parity, uint32_t nodd,
                     uint32_t elp_limit, uint32_t cs_sm, uint32_t binary_word)
{
  static_assert(LPC == 1 || (SM && LPC <= 32 && (LPC & (LPC - 1)) == 0), "lane groups need the shared-memory variant");
  typedef typename std::conditional<SM, uint16_t, uint32_t>::type coef_t;
  const uint32_t order = (1u << M) - 1u;
  BCHFFT_DYN_SMEM(uint16_t, tabs);
  uint16_t *slog = tabs, *sexp = tabs + (1u << M);
  const uint32_t per_cta = blockDim.x / LPC;
  const uint32_t cwl = threadIdx.x / LPC, sub = threadIdx.x % LPC;
  const uint32_t cw = blockIdx.x * per_cta + cwl;
  // the LPC lanes of this codeword within the warp
  const uint32_t pairmask = LPC == 32 ? 0xFFFFFFFFu
                          : LPC >= 2 ? (((1u << (LPC & 31)) - 1u) << (threadIdx.x & 31u & ~(uint32_t)(LPC - 1))) : 0u;
  auto pair_sync = [&]() {
    if (LPC >= 2) __syncwarp(pairmask);
  };
  auto pair_xor = [&](uint32_t v) -> uint32_t {
#pragma unroll
    for (int o = LPC / 2; o >= 1; o >>= 1) v ^= __shfl_xor_sync(pairmask, v, o);
    return v;
  };
  auto pair_or = [&](uint32_t v) -> uint32_t {
#pragma unroll
    for (int o = LPC / 2; o >= 1; o >>= 1) v |= __shfl_xor_sync(pairmask, v, o);
    return v;
  };
  auto pair_max = [&](int v) -> int {
#pragma unroll
    for (int o = LPC / 2; o >= 1; o >>= 1) {
      const int w = __shfl_xor_sync(pairmask, v, o);
      v = w > v ? w : v;
    }
    return v;
  };
  // shared-memory indices fit 32 bits (no 64-bit address arithmetic in the recurrence)
  typedef typename std::conditional<SM, uint32_t, size_t>::type stride_t;
  const stride_t cs = SM ? (stride_t)cs_sm : (stride_t)batch;   // stride between coefficients
  const stride_t bs = (stride_t)(d + 1) * cs;                   // stride between buffers
  coef_t *A, *B, *C;
  const coef_t *s;
  uint32_t flagged = 0u;
  if (SM) {
    for (uint32_t i = threadIdx.x; i <= order; i += blockDim.x) {
      slog[i] = (uint16_t)glog[i];
      sexp[i] = (uint16_t)gexp[i];
    }
    coef_t *poly = (coef_t *)(tabs + (2u << M)) + cwl;
    A = poly;
    B = A + bs;
    C = B + bs;
    coef_t *sc = C + bs;
    for (uint32_t j = sub; j < 3u * (d + 1u); j += LPC) poly[j * cs] = 0;
    s = sc;
    __syncthreads();   // tables complete
    if (acc) {
      // syndromes from the bitsliced accumulators [slab][nodd][M] (cf. k_syn_finish): the lanes of the
      // group split the odd classes; S_(o 2^e) = S_o^(2^e) follows from the odd one's logarithm
      uint32_t any = 0u;
      if (cw < batch) {
        const uint32_t slab = cw >> 5, l = cw & 31u;
        if (sub == 0u) {
          const uint32_t v = (parity[slab] >> l) & 1u;
          sc[0] = (coef_t)v;
          any = v;
        }
        // (two loops: the global loads of several classes are in flight together)
#pragma unroll 4
        for (uint32_t o = 1u + 2u * sub; o < d; o += 2u * LPC) {
          const uint32_t *a = acc + ((size_t)slab * nodd + (o >> 1)) * M;
          uint32_t v = 0u;
#pragma unroll
          for (int b = 0; b < M; ++b) v |= ((a[b] >> l) & 1u) << b;
          sc[o * cs] = (coef_t)v;
          any |= v;
        }
        for (uint32_t o = 1u + 2u * sub; 2u * o < d; o += 2u * LPC) {
          const uint32_t v = sc[o * cs];
          uint32_t lg = v ? (uint32_t)slog[v] : 0u;
          for (uint32_t i = 2u * o; i < d; i *= 2u) {
            lg *= 2u;
            if (lg >= order) lg -= order;
            sc[i * cs] = (coef_t)(v ? (uint32_t)sexp[lg] : 0u);
          }
        }
      }
      any = pair_or(any);
      if (cw < batch && sub == 0u) flag[cw] = any ? 1 : 0;
      flagged = any ? 1u : 0u;
    } else {
      for (uint32_t j = sub; j < d; j += LPC) sc[j * cs] = (coef_t)(cw < batch ? syn[(size_t)cw * d + j] : 0u);
      flagged = cw < batch ? (uint32_t)flag[cw] : 0u;
    }
    pair_sync();
  } else {
    A = (coef_t *)bm + cw;
    B = A + bs;
    C = B + bs;
    s = (const coef_t *)syn + (size_t)cw * d;
    flagged = cw < batch ? (uint32_t)flag[cw] : 0u;
  }
  const stride_t ss = SM ? cs : 1;   // stride between syndromes
  auto mul = [&](uint32_t a, uint32_t b) -> uint32_t {
    if (!SM) return gf_mul<M>(a, b);
    if (!a || !b) return 0u;
    uint32_t lg = (uint32_t)slog[a] + (uint32_t)slog[b];
    if (lg >= order) lg -= order;
    return sexp[lg];
  };
  // a * x^lg (= a * elem)
  auto mul_log = [&](uint32_t a, uint32_t lg, uint32_t elem) -> uint32_t {
    if (!SM) return gf_mul<M>(a, elem);
    if (!a) return 0u;
    lg += slog[a];
    if (lg >= order) lg -= order;
    return sexp[lg];
  };
  if (cw >= batch) return;
  uint32_t L = 0;
  if (flagged) {
    uint32_t degA = 0, degB = 0, degC = 0, gap = 1, inv_b_log = 0;
    if (sub == 0u) {
      B[0] = 1u;
      C[0] = 1u;
    }
    pair_sync();
    for (uint32_t i = 0; i + 1 < d; ++i) {
      // Syndromes of a binary word are power sums in characteristic 2 (S_2j = S_j^2), for which every
      // second discrepancy -- the one that brings in S_(i+1) with i + 1 even -- is zero whatever the
      // word (Berlekamp's binary simplification).  The decode pipeline says so (binary_word); the
      // stage-level entry point takes arbitrary syndrome arrays and computes every discrepancy.
      if (binary_word && (i & 1u)) {
        gap++;
        continue;
      }
      // four terms in flight: the loads of a batch are issued before the first table lookup, so the
      // dependent shared-memory latencies (coefficient -> log -> exp) overlap instead of adding up
      uint32_t disc = sub == 0u ? (uint32_t)s[(i + 1) * ss] : 0u;
      for (uint32_t j0 = 1u + sub; j0 <= degC; j0 += 4u * LPC) {
        uint32_t cv[4], sv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t j = j0 + (uint32_t)u * LPC;
          const bool in = j <= degC;
          cv[u] = in ? (uint32_t)C[j * cs] : 0u;
          sv[u] = in ? (uint32_t)s[(i + 1 - j) * ss] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) disc ^= mul(cv[u], sv[u]);
      }
      disc = pair_xor(disc);
      if (!disc) {
        gap++;
        continue;
      }
      const uint32_t dlog = SM ? (uint32_t)slog[disc] : glog[disc];
      uint32_t sl = dlog + inv_b_log;
      if (sl >= order) sl -= order;
      const uint32_t scale = SM ? (uint32_t)sexp[sl] : gexp[sl];
      const uint32_t hiB = degB + gap;
      // A = C below `gap`, scale * B shifted by gap (+ C) from there on -- the reference's four copy
      // loops and its XOR loop (bch.c:352-385) in one sweep over j, split between the lanes by j
      const uint32_t top = hiB > degC ? hiB : degC;
      int last_nz = -1;
      for (uint32_t j0 = sub; j0 <= top; j0 += 4u * LPC) {
        uint32_t bv[4], cv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t j = j0 + (uint32_t)u * LPC;
          bv[u] = (j >= gap && j <= hiB) ? (uint32_t)B[(j - gap) * cs] : 0u;
          cv[u] = j <= degC ? (uint32_t)C[j * cs] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t j = j0 + (uint32_t)u * LPC;
          const uint32_t v = mul_log(bv[u], sl, scale) ^ cv[u];
          if (j >= gap && j <= degC && v) last_nz = (int)j;
          if (j <= top) A[j * cs] = (coef_t)v;
        }
      }
      if (degC != hiB) {
        degA = top;
      } else {
        last_nz = pair_max(last_nz);
        if (last_nz >= 0) degA = (uint32_t)last_nz;   // otherwise degA keeps its previous value (bch.c:379-384)
      }
      if (2u * L <= i) {
        L = i + 1u - L;
        degB = degC;
        coef_t *t = B; B = C; C = t;
        inv_b_log = order - dlog;
        gap = 1;
      } else {
        gap++;
      }
      { coef_t *t = C; C = A; A = t; }
      degC = degA;
      pair_sync();   // the partner's half of the new C is visible before the next discrepancy
    }
  } else if (sub == 0u) {
    C[0] = 1u;
  }
  pair_sync();
  if (sub == 0u) Lout[cw] = L;
  // elp_limit = d: the whole buffer (stage-level API, bch.c leaves it in m_C_elp); smaller: only what
  // the root search reads, C[0..L] of the codewords it will look at (L <= elp_limit)
  uint32_t *e = elp + (size_t)cw * (d + 1);
  const uint32_t upto = elp_limit >= d ? d : (L <= elp_limit ? L : 0u);
  if (elp_limit >= d || (L >= 1u && L <= elp_limit))
    for (uint32_t j = sub; j <= upto; j += LPC) e[j] = C[j * cs];
}

// ---- Berlekamp-Massey, one CTA per codeword ---------------------------------------------------
// Same recurrence and buffer rotation as k_bm (bch.c:328-404), for FEW codewords with LARGE t
// (the reference's single decode at n = 65535, t = 1000: 2t iterations of up to t multiply-adds are
// 2*10^6 dependent steps for one thread).  Here the blockDim.x threads of a CTA split the
// coefficient index j of both inner loops; per iteration one XOR reduction (discrepancy) and, when
// the degrees tie, one max reduction (degree of the new C) go through shared memory.  All control
// state (L, gap, degrees, buffer roles) is replicated in every thread and stays uniform because it
// only depends on reduced values read behind a barrier.  The three coefficient buffers and the
// syndromes live in shared memory: dynamic smem = (4 d + 3) words.  Reduction cells rotate over
// three slots: the cell of iteration i + 2 is cleared by thread 0 after the barrier of iteration i,
// and nobody adds to it before the barrier of iteration i + 1.
template <int M>
__global__ void k_bm_wide(const uint32_t *__restrict__ syn, const int32_t *__restrict__ flag, uint32_t d,
                          uint32_t batch, const uint32_t *__restrict__ gexp, const uint32_t *__restrict__ glog,
                          uint32_t *__restrict__ elp, uint32_t *__restrict__ Lout, uint32_t elp_limit,
                          uint32_t binary_word)
{
  const uint32_t order = (1u << M) - 1u;
  BCHFFT_DYN_SMEM(uint32_t, poly);
  __shared__ uint32_t sh_disc[3];
  __shared__ int sh_last[3];
  const uint32_t cw = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  if (cw >= batch) return;
  uint32_t *A = poly, *B = A + (d + 1u), *C = B + (d + 1u), *sc = C + (d + 1u);
  for (uint32_t j = tid; j < 3u * (d + 1u); j += nt) poly[j] = 0u;
  for (uint32_t j = tid; j < d; j += nt) sc[j] = syn[(size_t)cw * d + j];
  if (tid < 3u) {
    sh_disc[tid] = 0u;
    sh_last[tid] = -1;
  }
  __syncthreads();
  if (tid == 0u) {
    B[0] = 1u;
    C[0] = 1u;
  }
  __syncthreads();
  auto add_disc = [&](uint32_t slot, uint32_t v) {
#ifndef BCHFFT_EMU
    v = __reduce_xor_sync(0xFFFFFFFFu, v);
    if ((tid & 31u) == 0u && v) atomicXor(&sh_disc[slot], v);
#else
    if (v) atomicXor(&sh_disc[slot], v);
#endif
  };
  auto add_last = [&](uint32_t slot, int v) {
#ifndef BCHFFT_EMU
    v = __reduce_max_sync(0xFFFFFFFFu, v);
    if ((tid & 31u) == 0u && v >= 0) atomicMax(&sh_last[slot], v);
#else
    if (v >= 0) atomicMax(&sh_last[slot], v);
#endif
  };
  uint32_t L = 0;
  if (flag[cw]) {
    uint32_t degA = 0, degB = 0, degC = 0, gap = 1, inv_b_log = 0, slot = 0;
    // (the reduction cells rotate per EXECUTED iteration: a skipped one has no barrier)
    auto next_slot = [&]() { slot = slot == 2u ? 0u : slot + 1u; };
    for (uint32_t i = 0; i + 1 < d; ++i) {
      if (binary_word && (i & 1u)) {   // see k_bm: this discrepancy is zero for the syndromes of a binary word
        gap++;
        continue;
      }
      uint32_t part = 0u;
      for (uint32_t j = 1u + tid; j <= degC; j += nt) part ^= gf_mul<M>(C[j], sc[i + 1u - j]);
      add_disc(slot, part);
      __syncthreads();
      const uint32_t disc = sh_disc[slot] ^ sc[i + 1u];
      const uint32_t clear = slot == 0u ? 2u : slot - 1u;   // = (slot + 2) % 3: the cell of iteration i + 2
      if (tid == 0u) {
        sh_disc[clear] = 0u;
        sh_last[clear] = -1;
      }
      if (!disc) {
        gap++;
        next_slot();
        continue;
      }
      const uint32_t dlog = glog[disc];
      uint32_t sl = dlog + inv_b_log;
      if (sl >= order) sl -= order;
      const uint32_t scale = gexp[sl];
      const uint32_t hiB = degB + gap;
      const uint32_t top = hiB > degC ? hiB : degC;
      int last_nz = -1;
      for (uint32_t j = tid; j <= top; j += nt) {
        const uint32_t bv = (j >= gap && j <= hiB) ? B[j - gap] : 0u;
        const uint32_t cv = j <= degC ? C[j] : 0u;
        const uint32_t v = gf_mul<M>(bv, scale) ^ cv;
        if (j >= gap && j <= degC && v) last_nz = (int)j;
        A[j] = v;
      }
      if (degC != hiB) {
        degA = top;
        __syncthreads();   // the new coefficients are visible before the next discrepancy
      } else {
        add_last(slot, last_nz);
        __syncthreads();
        const int ln = sh_last[slot];
        if (ln >= 0) degA = (uint32_t)ln;   // otherwise degA keeps its previous value (bch.c:379-384)
      }
      if (2u * L <= i) {
        L = i + 1u - L;
        degB = degC;
        uint32_t *t = B; B = C; C = t;
        inv_b_log = order - dlog;
        gap = 1;
      } else {
        gap++;
      }
      { uint32_t *t = C; C = A; A = t; }
      degC = degA;
      next_slot();
    }
  }
  __syncthreads();
  if (tid == 0u) Lout[cw] = L;
  uint32_t *e = elp + (size_t)cw * (d + 1);
  const uint32_t upto = elp_limit >= d ? d : (L <= elp_limit ? L : 0u);
  if (elp_limit >= d || (L >= 1u && L <= elp_limit))
    for (uint32_t j = tid; j <= upto; j += nt) e[j] = C[j];
}

// ---- root search ----------------------------------------------------------------------------
// The error locator of every codeword is evaluated at all 2^M field points with the truncated
// additive FFT (the device counterpart of bch.c:406-490's FFT path, with the reference's
// low-degree shortcut of additive_fft.c:103-110 taken to its conclusion).  The cost of the
// truncated transform is K = ceil(log2(L+1)) butterfly levels over all points, so codewords are
// first sorted into buckets by K and packed into slabs per bucket: a slab of degree-3 locators
// runs 2 levels, not the 5 a degree-16 locator needs.
//   k_bucket_*  counting sort of the codewords that need a root search by locator degree (buckets K)
//   k_elp_fwd   per bucket: bitslice the locator coefficients of 2^sb slabs per CTA and run the
//               forward phase (twists + Taylor expansions) on the 2^K coefficients
//   k_locate    one launch over the slabs of all buckets: per tile of 2^t positions broadcast
//               the 2^K forward values over the cosets, K butterfly levels, zero test fused
//               into the last level
// A zero at the tile position that holds the field point k = x^i != 0 is an error at bit `order - i`
// (bch.c:479-488, including the bit-0 quirk: i = 0 reports position `order`).

constexpr int kMaxBuckets = 12;   // K = 1 .. 12

__device__ __forceinline__ uint32_t locator_bucket(uint32_t L)   // ceil(log2(L + 1))
{
  return 32u - (uint32_t)__clz((int)L);
}

// Closed-form roots of a locator of degree <= 2 (C[0] = 1): 1 + c1 x has the root 1 / c1;
// 1 + c1 x + c2 x^2 with c1, c2 != 0 becomes y^2 + y = u = c2 / c1^2 under x = (c1 / c2) y, which has
// the solutions y0, y0 + 1 exactly when Tr(u) = 0, y0 = sum over the set bits k of u of h_k with
// h_k^2 + h_k = x^k (+ a fixed trace-one element when Tr(x^k) = 1; those cancel in pairs);
// c1 = 0: the double root sqrt(1 / c2) is one point.  qs = {trace mask, h_0 .. h_(m-1)}.  Reports what
// the exhaustive search would: a root x^i counts as position order - i (bch.c:479-488).
// Returns the number of roots.
__device__ __forceinline__ uint32_t locator_roots_direct(uint32_t L, uint32_t c1, uint32_t c2, uint32_t order,
                                                         const uint32_t *__restrict__ glog,
                                                         const uint32_t *__restrict__ gexp,
                                                         const uint32_t *__restrict__ qs, uint32_t (&where)[2])
{
  if (L < 2u) c2 = 0u;
  if (!c2) {
    if (!c1) return 0u;
    where[0] = order - (order - glog[c1]) % order;          // log(1 / c1) = (order - log c1) mod order
    return 1u;
  }
  const uint32_t l2 = glog[c2];
  if (!c1) {
    // x^2 = 1 / c2: halve the logarithm (2^-1 = (order + 1) / 2 mod order)
    const uint32_t lg = (uint32_t)(((uint64_t)((order - l2) % order) * ((order + 1u) / 2u)) % order);
    where[0] = order - lg;
    return 1u;
  }
  const uint32_t l1 = glog[c1];
  const uint32_t u = gexp[(l2 + 2u * (order - l1)) % order];
  if (__popc(u & qs[0]) & 1u) return 0u;                     // Tr(u) = 1: no root in the field
  uint32_t y0 = 0u;
  for (uint32_t k = 0, v = u; v; ++k, v >>= 1)
    if (v & 1u) y0 ^= qs[1u + k];
  const uint32_t lr = (l1 + order - l2) % order;             // log(c1 / c2)
  const uint32_t la = (lr + glog[y0]) % order, lb = (lr + glog[y0 ^ 1u]) % order;
  where[0] = order - la;
  where[1] = order - lb;
  return 2u;
}

// Counting sort of the codewords that need a root search, by locator degree L: bucket K =
// ceil(log2(L + 1)) holds its codewords in ascending L, so that the 1024 codewords of a slab group
// have (almost) the same degree and k_locate_u can skip the forward values that are zero for the whole
// group (coefficient i of the converted locator vanishes for i > L).
//   counts:  [kMaxBuckets + 1] codewords per bucket (written by k_bucket_scatter)
//   by_deg:  [2][tcap + 1] zero-initialised: codewords per degree, scatter cursors per degree
//   lists:   [kMaxBuckets + 1][batch] codeword indices
// Codewords with nothing to locate (not flagged, L = 0: constant locator, or L > tcap) stay out.
// direct_max (0, 1 or 2): locators of degree <= direct_max are solved in closed form by k_bucket_count
// (count[] and positions[] written) and stay out of the buckets as well.
// Both kernels: dynamic smem = 2 (tcap + 1) words, same grid.
__global__ void k_bucket_count(const int32_t *__restrict__ flag, const uint32_t *__restrict__ Lin,
                               uint32_t tcap, uint32_t batch, uint32_t *__restrict__ by_deg, uint32_t direct_max,
                               const uint32_t *__restrict__ elp, uint32_t d, uint32_t order,
                               const uint32_t *__restrict__ glog, const uint32_t *__restrict__ gexp,
                               const uint32_t *__restrict__ qs, uint32_t *__restrict__ count,
                               uint32_t *__restrict__ positions, uint32_t pos_stride)
{
  BCHFFT_DYN_SMEM(uint32_t, hist);
  for (uint32_t i = threadIdx.x; i <= tcap; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  const uint32_t cw = blockIdx.x * blockDim.x + threadIdx.x;
  if (cw < batch && flag[cw]) {
    const uint32_t L = Lin[cw];
    if (L >= 1u && L <= tcap) {
      if (L <= direct_max) {
        const uint32_t *e = elp + (size_t)cw * (d + 1);
        uint32_t where[2];
        const uint32_t nr = locator_roots_direct(L, e[1], L >= 2u ? e[2] : 0u, order, glog, gexp, qs, where);
        for (uint32_t k = 0; k < nr && k < pos_stride; ++k) positions[(size_t)cw * pos_stride + k] = where[k];
        count[cw] = nr;
      } else {
        atomicAdd(&hist[L], 1u);
      }
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i <= tcap; i += blockDim.x)
    if (hist[i]) atomicAdd(&by_deg[i], hist[i]);
}

__global__ void k_bucket_scatter(const int32_t *__restrict__ flag, const uint32_t *__restrict__ Lin,
                                 uint32_t tcap, uint32_t batch, uint32_t *__restrict__ counts,
                                 uint32_t *__restrict__ by_deg, uint32_t *__restrict__ lists, uint32_t direct_max)
{
  BCHFFT_DYN_SMEM(uint32_t, sm);
  uint32_t *hist = sm, *start = sm + (tcap + 1u);      // per degree: codewords of this CTA, first list slot
  const uint32_t *per_deg = by_deg;
  uint32_t *cursor = by_deg + (tcap + 1u);
  for (uint32_t i = threadIdx.x; i <= tcap; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  const uint32_t cw = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t L = 0u, rank = 0u;
  if (cw < batch && flag[cw]) {
    const uint32_t l = Lin[cw];
    if (l > direct_max && l >= 1u && l <= tcap) {
      L = l;
      rank = atomicAdd(&hist[L], 1u);
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i <= tcap; i += blockDim.x) {
    if (!i) continue;
    // degrees below i in the same bucket come first
    const uint32_t lowest = 1u << (locator_bucket(i) - 1u);
    uint32_t before = 0u;
    for (uint32_t j = lowest; j < i; ++j) before += per_deg[j];
    if (blockIdx.x == 0u && (i == tcap || locator_bucket(i + 1u) != locator_bucket(i)))
      counts[locator_bucket(i)] = before + per_deg[i];
    start[i] = hist[i] ? before + atomicAdd(&cursor[i], hist[i]) : 0u;
  }
  __syncthreads();
  if (L) lists[(size_t)locator_bucket(L) * batch + start[L] + rank] = cw;
}

// Forward phase for bucket K = fwd.dom_bits.  grid = worst-case number of slab groups (CTAs
// beyond the bucket's population exit); dynamic smem = M * 2^(K+sb) words.
// Fout: [virtual slab][2^Kmax][PS] with virtual slab = vbase(bucket) + slab in bucket.
template <int M>
__global__ void __launch_bounds__(256, BCHFFT_MINB_ELPU)
k_elp_fwd(const uint32_t *__restrict__ elp, const uint32_t *__restrict__ Lin, uint32_t d,
          const uint32_t *__restrict__ counts, const uint32_t *__restrict__ lists, uint32_t batch,
          uint32_t sb, uint32_t kmax, FftTables tab, PassDesc fwd, uint32_t *__restrict__ Fout)
{
  constexpr int PS = plane_stride(M);
  BCHFFT_DYN_SMEM(uint32_t, S);
  const uint32_t K = fwd.dom_bits, TK = 1u << K, tt = K + sb, T = 1u << tt;
  const uint32_t pop = counts[K], nslabs = (pop + 31u) >> 5;
  const uint32_t slab0 = blockIdx.x << sb;
  if (slab0 >= nslabs) return;
  uint32_t vbase = 0u;
  for (uint32_t k = 1; k < K; ++k) vbase += (counts[k] + 31u) >> 5;
  const uint32_t *list = lists + (size_t)K * batch;
  // coefficient index i of the 32 codewords of a slab -> M plane words.  Coefficients above L
  // are not part of the polynomial (bch.c:444-445, 466).
  for (uint32_t idx = threadIdx.x; idx < T; idx += blockDim.x) {
    const uint32_t slab = slab0 + (idx >> K), i = idx & (TK - 1u);
    uint32_t x[32];
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      const uint32_t at = slab * 32u + l;
      uint32_t v = 0u;
      if (at < pop) {
        const uint32_t cw = list[at];
        if (i <= Lin[cw]) v = elp[(size_t)cw * (d + 1) + i];
      }
      x[l] = v;
    }
    transpose32(x);
    const uint32_t sl = swz(idx);
#pragma unroll
    for (int b = 0; b < M; ++b) S[b * T + sl] = x[b];
  }
  __syncthreads();
  TileGeom g;
  g.fixed = 0u;
  g.a0 = 0u;
  g.n0 = tt;
  g.a1 = tt;
  g.mask0 = T - 1u;
  g.pmask = TK - 1u;        // constants depend on the coefficient index only, not on the slab
  tile_run_ops<M>(S, T, tt, g, tab, fwd);
  for (uint32_t idx = threadIdx.x; idx < T; idx += blockDim.x) {
    const uint32_t slab = slab0 + (idx >> K), i = idx & (TK - 1u);
    if (slab >= nslabs) continue;
    uint4 *out = reinterpret_cast<uint4 *>(Fout + (((size_t)(vbase + slab) << kmax) + i) * PS);
    const uint32_t sl = swz(idx);
#pragma unroll
    for (int v = 0; v < PS / 4; ++v) {
      uint4 q;
      q.x = (4 * v + 0 < M) ? S[(4 * v + 0) * T + sl] : 0u;
      q.y = (4 * v + 1 < M) ? S[(4 * v + 1) * T + sl] : 0u;
      q.z = (4 * v + 2 < M) ? S[(4 * v + 2) * T + sl] : 0u;
      q.w = (4 * v + 3 < M) ? S[(4 * v + 3) * T + sl] : 0u;
      out[v] = q;
    }
  }
}

// zero test + root compaction on the final values of a tile
template <int M>
struct RootEpilogue {
  const TileGeom *g;
  const uint32_t *errloc;   // tile position -> reported error position (0xFFFFFFFF: the point 0)
  const uint32_t *list;     // codeword index of every lane of the slab
  uint32_t *count, *positions;
  uint32_t active, pos_stride;
  __device__ __forceinline__ void operator()(uint32_t l, const uint32_t *v) const
  {
    uint32_t nz = 0u;
#pragma unroll
    for (int b = 0; b < M; ++b) nz |= v[b];
    uint32_t roots = ~nz & active;
    if (!roots) return;
    const uint32_t where = errloc[g->pos(l)];
    if (where == 0xFFFFFFFFu) return;   // P(0) = C[0] = 1 for every active codeword; x^i never reaches 0
    while (roots) {
      const uint32_t lane = (uint32_t)__ffs((int)roots) - 1u;
      roots &= roots - 1u;
      const uint32_t cw = list[lane];
      const uint32_t idx = atomicAdd(&count[cw], 1u);
      if (idx < pos_stride) positions[(size_t)cw * pos_stride + idx] = where;
    }
  }
};

// grid = (tiles per slab, upper bound of the virtual slab count); dynamic smem = M * 2^t words.
// bwd_descs[K]: the K butterfly levels on a tile of 2^t positions.
template <int M>
__global__ void __launch_bounds__(BCHFFT_NT_LOC, BCHFFT_MINB_LOC)
k_locate(const uint32_t *__restrict__ F, const uint32_t *__restrict__ counts,
         const uint32_t *__restrict__ lists, uint32_t batch, uint32_t kmax, FftTables tab,
         const PassDesc *__restrict__ bwd_descs, const uint32_t *__restrict__ errloc,
         uint32_t *__restrict__ count, uint32_t *__restrict__ positions, uint32_t pos_stride)
{
  constexpr int PS = plane_stride(M);
  BCHFFT_DYN_SMEM(uint32_t, S);
  // virtual slab -> (bucket K, slab within the bucket)
  uint32_t v = blockIdx.y, K = 1u;
  for (; K <= kmax; ++K) {
    const uint32_t ns = (counts[K] + 31u) >> 5;
    if (v < ns) break;
    v -= ns;
  }
  if (K > kmax) return;   // beyond the last populated slab (uniform across the CTA)
  const uint32_t pop = counts[K];
  const uint32_t lanes = pop - v * 32u;
  const uint32_t active = lanes >= 32u ? 0xFFFFFFFFu : ((1u << lanes) - 1u);
  const PassDesc &bwd = bwd_descs[K];
  const uint32_t T = 1u << bwd.t;
  const TileGeom g = tile_geom(bwd, blockIdx.x);
  tile_load<M>(S, T, F + ((size_t)blockIdx.y << kmax) * PS, g, (1u << K) - 1u);
  __syncthreads();
  RootEpilogue<M> epi;
  epi.g = &g;
  epi.errloc = errloc;
  epi.list = lists + (size_t)K * batch + v * 32u;
  epi.count = count;
  epi.positions = positions;
  epi.active = active;
  epi.pos_stride = pos_stride;
  tile_run_ops<M, RootEpilogue<M>>(S, T, bwd.t, g, tab, bwd, &epi);
}


// ---- root search, low locator degrees (K <= 5): warp lanes = 32 slabs -------------------------
// The butterfly multipliers depend on the position only.  When the 32 lanes of a warp hold 32
// different slabs at the SAME positions, every multiplier is warp-uniform and the multiply
// branches over its set bits instead of masking all M*M partial products (bs_mad_uniform).  A
// "group" is 32 slabs of one bucket (1024 codewords); its forward values are stored
// [group][coefficient 2^kmax][plane M][lane 32].  A CTA takes a range of 32-position tiles of one
// group; per tile the K levels run as radix-4 / radix-2 steps, one block per warp, through a
// [32 positions][M planes][32 lanes] shared-memory buffer (first step reads the forward values,
// last step feeds the zero test).
constexpr int kLocUMaxK = 7;
// root queue entries per CTA (kernel parameter qcap; BCHFFT_LOCU_QUEUE overrides it so that tests can
// force the drain and overflow paths; the emulator build defaults to a tiny queue for the same reason).
// Shared memory the queue does not take is L1 for the forward-value loads of the first step, frequent drains
// cost barriers: measured per 10^6 words at (8191, 16), 4096 entries 4.46 ms, 2048: 4.30, 1024: 4.31, 512: 4.34,
// 256: 4.46, 128: 4.72 ms.
#ifdef BCHFFT_EMU
constexpr uint32_t kLocUQueue = 24;
#else
constexpr uint32_t kLocUQueue = 2048;
#endif

// F2 index of (group, coefficient i, plane b, lane)
__device__ __forceinline__ size_t f2_index(uint32_t group, uint32_t kmax, uint32_t i, uint32_t M_, uint32_t b,
                                           uint32_t lane)
{
  return ((((size_t)group << kmax) + i) * M_ + b) * 32u + lane;
}

// forward phase like k_elp_fwd, output in the group layout; all buckets in one launch:
// grid = (ceil(slabs / 8), kmax), bucket K = blockIdx.y + 1 (K <= 5 here), a CTA takes 256
// coefficient columns = 2^(8-K) slabs; fwd_descs[K] = forward pass of bucket K.
template <int M>
__global__ void __launch_bounds__(256, BCHFFT_MINB_ELPU)
k_elp_fwd_u(const uint32_t *__restrict__ elp, const uint32_t *__restrict__ Lin, uint32_t d,
            const uint32_t *__restrict__ counts, const uint32_t *__restrict__ lists, uint32_t batch,
            uint32_t kmax, FftTables tab, const PassDesc *__restrict__ fwd_descs, uint32_t *__restrict__ F2)
{
  BCHFFT_DYN_SMEM(uint32_t, S);
  const uint32_t K = blockIdx.y + 1u, sb = 8u - K;
  const PassDesc &fwd = fwd_descs[K];
  const uint32_t TK = 1u << K, tt = K + sb, T = 1u << tt;
  const uint32_t pop = counts[K], nslabs = (pop + 31u) >> 5;
  const uint32_t slab0 = blockIdx.x << sb;
  if (slab0 >= nslabs) return;
  uint32_t gbase = 0u;
  for (uint32_t k = 1; k < K; ++k) gbase += (((counts[k] + 31u) >> 5) + 31u) >> 5;
  const uint32_t *list = lists + (size_t)K * batch;
  for (uint32_t idx = threadIdx.x; idx < T; idx += blockDim.x) {
    const uint32_t slab = slab0 + (idx >> K), i = idx & (TK - 1u);
    uint32_t x[32];
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      const uint32_t at = slab * 32u + l;
      uint32_t v = 0u;
      if (at < pop) {
        const uint32_t cw = list[at];
        if (i <= Lin[cw]) v = elp[(size_t)cw * (d + 1) + i];
      }
      x[l] = v;
    }
    transpose32(x);
    const uint32_t sl = swz(idx);
#pragma unroll
    for (int b = 0; b < M; ++b) S[b * T + sl] = x[b];
  }
  __syncthreads();
  TileGeom g;
  g.fixed = 0u;
  g.a0 = 0u;
  g.n0 = tt;
  g.a1 = tt;
  g.mask0 = T - 1u;
  g.pmask = TK - 1u;
  tile_run_ops<M>(S, T, tt, g, tab, fwd);
  for (uint32_t idx = threadIdx.x; idx < T; idx += blockDim.x) {
    const uint32_t slab = slab0 + (idx >> K), i = idx & (TK - 1u);
    if (slab >= nslabs) continue;
    const uint32_t sl = swz(idx);
#pragma unroll
    for (int b = 0; b < M; ++b) F2[f2_index(gbase + (slab >> 5), kmax, i, M, b, slab & 31u)] = S[b * T + sl];
  }
}

// Drain the CTA's root queue: entry -> (codeword, reported position); every thread takes
// independent entries, so the global atomics overlap.  Called by all threads after a barrier.
__device__ __forceinline__ void locu_flush(const uint32_t *queue, uint32_t *qn, uint32_t qcap,
                                           const uint32_t *__restrict__ glist,
                                           const uint32_t *__restrict__ errloc, uint32_t *__restrict__ count,
                                           uint32_t *__restrict__ positions, uint32_t pos_stride)
{
  const uint32_t n = *qn < qcap ? *qn : qcap;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t e = queue[i];
    const uint32_t where = errloc[e >> 10];
    if (where == 0xFFFFFFFFu) continue;   // the point 0: C(0) = 1, never a root of an active codeword
    const uint32_t cw = glist[e & 1023u];
    const uint32_t idx = atomicAdd(&count[cw], 1u);
    if (idx < pos_stride) positions[(size_t)cw * pos_stride + idx] = where;
  }
  __syncthreads();
  if (threadIdx.x == 0) *qn = 0u;
  __syncthreads();
}

template <int M>
struct LocU {
  // per-thread view of the job
  const uint32_t *F2g;        // forward values of this group (+ lane)
  uint32_t *W;                // shared buffer (+ lane)
  const uint32_t *gam;
  const uint32_t *gam_off;
  uint32_t *queue, *qn;       // shared root queue of the CTA: (tile position << 10) | codeword of the group
  const uint32_t *glist, *errloc;   // overflow path: the group's codeword list, position table
  uint32_t *count, *positions;
  uint32_t active, kmask, tile0, lane, pos_stride, qcap;
  uint32_t lmax;              // largest locator degree of the group: forward values above it are zero

  __device__ __forceinline__ void load(uint32_t (&e)[M], uint32_t p, bool from_f) const
  {
    if (from_f) {
      const uint32_t *src = F2g + (size_t)(p & kmask) * M * 32u;
#pragma unroll
      for (int b = 0; b < M; ++b) e[b] = src[b * 32];
    } else {
#pragma unroll
      for (int b = 0; b < M; ++b) e[b] = W[(p * M + b) * 32u];
    }
  }
  __device__ __forceinline__ void store(const uint32_t (&e)[M], uint32_t p) const
  {
#pragma unroll
    for (int b = 0; b < M; ++b) W[(p * M + b) * 32u] = e[b];
  }
  // zero test of the final value at tile position p.  Roots go to the CTA's shared queue (one
  // shared-memory atomic each); the global bookkeeping happens once per CTA in locu_flush.
  __device__ __forceinline__ void root_test(const uint32_t (&e)[M], uint32_t p) const
  {
    uint32_t nz = 0u;
#pragma unroll
    for (int b = 0; b < M; ++b) nz |= e[b];
    uint32_t roots = ~nz & active;
    while (roots) {
      const uint32_t bit = (uint32_t)__ffs((int)roots) - 1u;
      roots &= roots - 1u;
      const uint32_t slot = atomicAdd(qn, 1u);
      if (slot < qcap) {
        queue[slot] = ((tile0 + p) << 10) | (lane << 5) | bit;
      } else {
        // queue full (only with pathological root clustering): record directly
        const uint32_t where = errloc[tile0 + p];
        if (where != 0xFFFFFFFFu) {
          const uint32_t cw = glist[(lane << 5) | bit];
          const uint32_t idx = atomicAdd(&count[cw], 1u);
          if (idx < pos_stride) positions[(size_t)cw * pos_stride + idx] = where;
        }
      }
    }
  }
  // BF(d), BF(d-1) on block bw (0..7) of the tile
  __device__ __forceinline__ void radix4(uint32_t d, uint32_t bw, bool from_f, bool last) const
  {
    const uint32_t lb = d - 1u;
    const uint32_t p0 = ((bw >> lb) << (lb + 2u)) | (bw & ((1u << lb) - 1u));
    const uint32_t o = 1u << lb;
    uint32_t e0[M], e1[M], e2[M], e3[M];
    // first step: the upper half of the forward values is zero above the group's largest degree
    // (the codewords of a bucket are sorted by degree) -- nothing to load, nothing to multiply
    const bool z2 = from_f && ((p0 + 2u * o) & kmask) > lmax, z3 = from_f && ((p0 + 3u * o) & kmask) > lmax;
    load(e0, p0, from_f);
    load(e1, p0 + o, from_f);
    if (z2) {
#pragma unroll
      for (int b = 0; b < M; ++b) e2[b] = 0u;
    } else {
      load(e2, p0 + 2u * o, from_f);
    }
    if (z3) {
#pragma unroll
      for (int b = 0; b < M; ++b) e3[b] = 0u;
    } else {
      load(e3, p0 + 3u * o, from_f);
    }
    const uint32_t gp = tile0 + p0;
    const uint32_t G = gam[gam_off[d] + (gp >> (d + 1u))];
    if (G) {
      if (!z3) bs_mad_uniform2<M>(e0, e2, e1, e3, G);
      else if (!z2) bs_mad_uniform<M>(e0, e2, G);
    }
#pragma unroll
    for (int b = 0; b < M; ++b) {
      e2[b] ^= e0[b];
      e3[b] ^= e1[b];
    }
    const uint32_t Ga = gam[gam_off[d - 1u] + (gp >> d)], Gb = gam[gam_off[d - 1u] + (gp >> d) + 1u];
    if (Ga) bs_mad_uniform<M>(e0, e1, Ga);
    if (Gb) bs_mad_uniform<M>(e2, e3, Gb);
#pragma unroll
    for (int b = 0; b < M; ++b) {
      e1[b] ^= e0[b];
      e3[b] ^= e2[b];
    }
    if (last) {
      root_test(e0, p0);
      root_test(e1, p0 + o);
      root_test(e2, p0 + 2u * o);
      root_test(e3, p0 + 3u * o);
    } else {
      store(e0, p0);
      store(e1, p0 + o);
      store(e2, p0 + 2u * o);
      store(e3, p0 + 3u * o);
    }
  }
  // BF(d), BF(d-1) for d = 5, 6 (first step of the buckets K = d + 1 = 6, 7): the radix-4 group
  // {i, i+o, i+2o, i+3o}, o = 2^(d-1), spans 2 or 4 tiles.  Its inputs are forward values, which every
  // tile can read, so a tile recomputes the part of the group that ends in it: both products of level
  // d, then one butterfly of level d-1 -- 3 multiplies for 2 outputs (d = 5: local positions i, i+16)
  // or for 1 output (d = 6: local position i) instead of 4 multiplies for 4 outputs.
  __device__ __forceinline__ void radix4_wide(uint32_t d, uint32_t i) const
  {
    const uint32_t o = 1u << (d - 1u);
    const uint32_t r0 = tile0 & kmask;          // first position of this tile within its 2^(d+1) block
    const uint32_t hi = r0 >> d;                // the tile lies in the upper half of level d
    uint32_t e0[M], e1[M], e2[M], e3[M];
    const bool z2 = i + 2u * o > lmax, z3 = i + 3u * o > lmax;   // see radix4
    load(e0, i, true);
    load(e1, i + o, true);
    if (z2) {
#pragma unroll
      for (int b = 0; b < M; ++b) e2[b] = 0u;
    } else {
      load(e2, i + 2u * o, true);
    }
    if (z3) {
#pragma unroll
      for (int b = 0; b < M; ++b) e3[b] = 0u;
    } else {
      load(e3, i + 3u * o, true);
    }
    const uint32_t q = tile0 >> (d + 1u);
    const uint32_t G = gam[gam_off[d] + q];
    if (G) {
      if (!z3) bs_mad_uniform2<M>(e0, e2, e1, e3, G);
      else if (!z2) bs_mad_uniform<M>(e0, e2, G);
    }
    if (hi) {
#pragma unroll
      for (int b = 0; b < M; ++b) {
        e0[b] ^= e2[b];
        e1[b] ^= e3[b];
      }
    }
    const uint32_t Gx = gam[gam_off[d - 1u] + 2u * q + hi];
    if (Gx) bs_mad_uniform<M>(e0, e1, Gx);
#pragma unroll
    for (int b = 0; b < M; ++b) e1[b] ^= e0[b];
    if (d == 5u) {
      store(e0, i);
      store(e1, i + 16u);
    } else if ((r0 >> 5) & 1u) {
      store(e1, i);
    } else {
      store(e0, i);
    }
  }
  // BF(0) on pair pw (0..15) of the tile; always the last step
  __device__ __forceinline__ void radix2_last(uint32_t pw, bool from_f) const
  {
    const uint32_t p0 = pw << 1;
    uint32_t lo[M], hi[M];
    load(lo, p0, from_f);
    load(hi, p0 + 1u, from_f);
    const uint32_t G = gam[gam_off[0] + ((tile0 + p0) >> 1)];
    if (G) bs_mad_uniform<M>(lo, hi, G);
#pragma unroll
    for (int b = 0; b < M; ++b) hi[b] ^= lo[b];
    root_test(lo, p0);
    root_test(hi, p0 + 1u);
  }
};

// grid = (tile ranges per group, upper bound of the number of groups); dynamic smem =
// 32 * M * 32 words.  ngroups_max >= sum_K ceil(ceil(count_K / 32) / 32).
template <int M>
__global__ void __launch_bounds__(BCHFFT_NT_LOC, BCHFFT_MINB_LOCU)
k_locate_u(const uint32_t *__restrict__ F2, const uint32_t *__restrict__ counts,
           const uint32_t *__restrict__ lists, const uint32_t *__restrict__ Lin, uint32_t batch, uint32_t kmax,
           FftTables tab,
           const uint32_t *__restrict__ errloc, uint32_t *__restrict__ count,
           uint32_t *__restrict__ positions, uint32_t pos_stride, uint32_t qcap)
{
  BCHFFT_DYN_SMEM(uint32_t, Wbuf);
  uint32_t *queue = Wbuf + 32u * M * 32u;
  __shared__ uint32_t s_qn;
  if (threadIdx.x == 0) s_qn = 0u;
  __syncthreads();
  // group -> (bucket K, group within the bucket)
  uint32_t v = blockIdx.y, K = 1u;
  for (; K <= kmax; ++K) {
    const uint32_t ng = (((counts[K] + 31u) >> 5) + 31u) >> 5;
    if (v < ng) break;
    v -= ng;
  }
  if (K > kmax) return;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const uint32_t pop = counts[K];
  const uint32_t slab = v * 32u + lane;                 // slab of this lane within the bucket
  const uint32_t first = slab * 32u;
  LocU<M> job;
  job.F2g = F2 + f2_index(blockIdx.y, kmax, 0u, M, 0u, lane);
  job.W = Wbuf + lane;
  job.gam = tab.gam;
  job.gam_off = tab.gam_off;
  job.queue = queue;
  job.qn = &s_qn;
  job.lane = lane;
  job.kmask = (1u << K) - 1u;
  const uint32_t *glist = lists + (size_t)K * batch + (size_t)v * 1024u;   // the group's codewords
  job.glist = glist;
  job.errloc = errloc;
  job.count = count;
  job.positions = positions;
  job.pos_stride = pos_stride;
  job.qcap = qcap;
  job.active = first >= pop ? 0u : (pop - first >= 32u ? 0xFFFFFFFFu : ((1u << (pop - first)) - 1u));
  {
    // the bucket's list ascends in degree: the group's last codeword has its largest one
    const uint32_t gfirst = v * 1024u, gend = pop - gfirst < 1024u ? pop : gfirst + 1024u;
    job.lmax = Lin ? Lin[glist[gend - 1u - gfirst]] : 0xFFFFFFFFu;
  }
  const uint32_t nsteps = (K + 1u) >> 1;
  const uint32_t ntiles = (1u << M) >> 5;
  const uint32_t per = (ntiles + gridDim.x - 1u) / gridDim.x;
  const uint32_t t_begin = blockIdx.x * per, t_end = (t_begin + per < ntiles) ? t_begin + per : ntiles;
  for (uint32_t tile = t_begin; tile < t_end; ++tile) {
    job.tile0 = tile << 5;
    // every eighth tile: drain the queue when it is half full (uniform decision behind a barrier;
    // about 32 roots per tile are expected, an overflow in between takes the direct path)
    if ((tile & 7u) == 7u) {
      __syncthreads();
      const uint32_t fill = s_qn;
      __syncthreads();   // nobody pushes before everybody has read the fill level
      if (fill > qcap / 2u) locu_flush(queue, &s_qn, qcap, glist, errloc, count, positions, pos_stride);
    }
    // steps: level pairs (d, d-1) as radix-4 blocks, a trailing level 0 as radix-2 pairs.  One
    // rolled loop, so that the kernel holds a single copy of each multiply body (the code of all
    // buckets then shares the instruction cache).
#pragma unroll 1
    for (uint32_t s = 0; s < nsteps; ++s) {
      const uint32_t d = K - 1u - 2u * s;
      const bool from_f = s == 0u, last = s + 1u == nsteps;
      if (d >= 5u) {
        // K = 6, 7: groups that span several tiles (first step only)
        const uint32_t ng = d == 5u ? 16u : 32u;
#pragma unroll 1
        for (uint32_t b = warp; b < ng; b += nwarps) job.radix4_wide(d, b);
      } else if (d >= 1u) {
        // a warp's blocks alternate between the two ends of the block range (w, 2 nwarps - 1 - w, ..): with
        // fewer than eight warps every warp gets cheap and expensive blocks of a zero-skipping first step
#pragma unroll 1
        for (uint32_t r = 0; r * nwarps < 8u; ++r) {
          const uint32_t b = r * nwarps + ((r & 1u) ? nwarps - 1u - warp : warp);
          if (b < 8u) job.radix4(d, b, from_f, last);      // (more than eight warps: the others idle)
        }
      } else {
#pragma unroll 1
        for (uint32_t r = 0; r * nwarps < 16u; ++r) {
          const uint32_t b = r * nwarps + ((r & 1u) ? nwarps - 1u - warp : warp);
          if (b < 16u) job.radix2_last(b, from_f);
        }
      }
      if (nsteps > 1u) __syncthreads();   // also keeps the next tile's first step off a buffer still being read
    }
  }
  __syncthreads();
  locu_flush(queue, &s_qn, qcap, glist, errloc, count, positions, pos_stride);
}


// ---- systematic encoder (bch.c:492-506) -------------------------------------
// codeword = data * X^dg + (data * X^dg mod g): data bit i -> codeword bit i + dg, parity in bits
// 0 .. dg-1.  One thread per codeword divides by g byte-wise: R <- (R << 8) ^ T[top byte of R ^
// data byte], T[v] = v(X) X^dg mod g (256 rows of NW 64-bit words in shared memory).  Requires
// 8 <= dg <= 64 NW.  data: [batch][dwords] packed data bits (bit i of item = bit i % 64 of word
// i / 64), only the first data_bits count.
// Memory side: a warp owns 32 codewords.  Their data travel through a [32][33]-word shared tile in
// chunks of 32 words per codeword, loaded with the lanes along a codeword (128-byte rows) and read
// back with one codeword per lane (conflict-free), so that neither the loads of the division nor the
// stores of the shifted copy data << dg (done by the warp together, lanes along the codeword) touch a
// sector for 8 bytes at a time -- round 1's one-thread-per-row version ran at 3 % of the copy bandwidth.
// The owner of a codeword finally writes the words that hold its parity bits.
// grid = ceil(batch / 128), 128 threads; dynamic smem = 256 * NW * 8 + 4 * 32 * 33 * 4 bytes.
template <int NW>
__global__ void __launch_bounds__(128)
k_encode(const uint64_t *__restrict__ data, uint32_t dwords, uint32_t data_bits, uint32_t dg,
         const uint64_t *__restrict__ table, uint64_t *__restrict__ cw_out, uint32_t wpc, uint32_t batch)
{
  BCHFFT_DYN_SMEM(uint64_t, T);
  uint32_t *tiles = reinterpret_cast<uint32_t *>(T + 256u * NW);
  for (uint32_t i = threadIdx.x; i < 256u * NW; i += blockDim.x) T[i] = table[i];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t *S = tiles + warp * (32u * 33u);
  const uint32_t cw0 = (blockIdx.x * (blockDim.x >> 5) + warp) * 32u;     // first codeword of the warp
  const uint32_t cw = cw0 + lane;                                           // this lane's codeword
  const uint32_t *d32 = reinterpret_cast<const uint32_t *>(data);
  const size_t dstride = (size_t)dwords * 2u;                               // 32-bit words per data row
  const uint32_t nw32 = (data_bits + 31u) >> 5;                             // data words in use
  const uint32_t tail = data_bits & 31u;                                    // bits of the last one (0 = all)
  // data word w of codeword c, bits beyond data_bits cleared, zero outside the row
  auto dword = [&](uint32_t c, uint32_t w) -> uint32_t {
    if (w >= nw32 || c >= batch) return 0u;
    uint32_t x = d32[(size_t)c * dstride + w];
    if (w == nw32 - 1u && tail) x &= (1u << tail) - 1u;
    return x;
  };

  // division: chunks of 32 data words from the top of the polynomial down
  uint64_t R[NW];
#pragma unroll
  for (int k = 0; k < NW; ++k) R[k] = 0ull;
  const uint32_t tw = (dg - 8u) >> 6, tsh = (dg - 8u) & 63u;   // top byte of R = bits dg-8 .. dg-1
  const uint32_t hw = (dg - 1u) >> 6;                           // last word of R in use
  const uint64_t hmask = (dg & 63u) ? ((1ull << (dg & 63u)) - 1ull) : ~0ull;
  const uint32_t nchunks = (nw32 + 31u) >> 5;
  for (uint32_t ch = nchunks; ch-- > 0;) {
    const uint32_t base = ch * 32u;
    for (uint32_t c = 0; c < 32u; ++c) S[c * 33u + lane] = dword(cw0 + c, base + lane);
    __syncwarp();
    for (uint32_t k = 32u; k-- > 0;) {
      if (base + k >= nw32) continue;                           // (uniform: nw32 is a kernel argument)
      const uint32_t w = S[lane * 33u + k];
#pragma unroll
      for (int by = 3; by >= 0; --by) {
        if ((base + k) * 32u + (uint32_t)by * 8u >= data_bits) continue;   // whole byte beyond the data
        const uint32_t dbyte = (w >> (by * 8)) & 255u;
        uint32_t top = 0u;
#pragma unroll
        for (int q = 0; q < NW; ++q) {
          if ((uint32_t)q == tw) top |= (uint32_t)(R[q] >> tsh);
          if ((uint32_t)q == tw + 1u && tsh > 56u) top |= (uint32_t)(R[q] << (64u - tsh));
        }
        const uint64_t *row = T + (size_t)((top ^ dbyte) & 255u) * NW;
#pragma unroll
        for (int q = NW - 1; q >= 0; --q) {
          const uint64_t lo = q > 0 ? (R[q - 1] >> 56) : 0ull;
          R[q] = ((R[q] << 8) | lo) ^ row[q];
        }
#pragma unroll
        for (int q = 0; q < NW; ++q) {
          if ((uint32_t)q == hw) R[q] &= hmask;
          if ((uint32_t)q > hw) R[q] = 0ull;
        }
      }
    }
    __syncwarp();                                               // the tile is reloaded next
  }

  // shifted copy: output word j (32 bits) = data bits 32 j - dg .. 32 j - dg + 31
  uint32_t *o32 = reinterpret_cast<uint32_t *>(cw_out);
  const size_t ostride = (size_t)wpc * 2u;
  const uint32_t sw = dg >> 5, sh = dg & 31u;
  const uint32_t jpar = (dg - 1u) >> 5;                         // last output word that holds parity bits
  auto shifted = [&](uint32_t c, uint32_t j) -> uint32_t {
    if (j < sw) return 0u;
    const uint32_t i = j - sw;
    uint32_t v = dword(c, i) << sh;
    if (sh && i >= 1u) v |= dword(c, i - 1u) >> (32u - sh);
    return v;
  };
  for (uint32_t c = 0; c < 32u; ++c) {
    if (cw0 + c >= batch) break;                                // (uniform)
    for (uint32_t j = jpar + 1u + lane; j < 2u * wpc; j += 32u) o32[(size_t)(cw0 + c) * ostride + j] = shifted(cw0 + c, j);
  }
  if (cw < batch) {
    for (uint32_t j = 0; j <= jpar && j < 2u * wpc; ++j) {
      uint32_t par = 0u;
#pragma unroll
      for (int q = 0; q < NW; ++q)
        if ((uint32_t)q == (j >> 1)) par = (uint32_t)(R[q] >> (32u * (j & 1u)));
      o32[(size_t)cw * ostride + j] = par | shifted(cw, j);
    }
  }
}


// Any generator degree >= 64 (the (8191, 200) and (65535, 1000) codes: deg g = 2444, 15360): one
// WARP per codeword.  The remainder S (deg g bits) lives in shared memory as NS = deg g / 64 + 1
// words; the data polynomial is consumed 64 bits at a time from the top:
//     S <- (S mod X^(dg-64)) X^64 + (top64(S) + chunk) X^dg  mod g
//        = (S << 64, truncated to dg bits)  ^  sum_{k : bit k of u set} rows[k],  u = top64(S) ^ chunk,
// rows[k] = X^(dg+k) mod g (64 rows of NS words, read through L1/L2).  u is warp-uniform, so the row
// loop runs over its set bits only; lane l owns words l, l + 32, ...  Two copies of S alternate, one
// __syncwarp per chunk.  grid = ceil(batch / warps per CTA), dynamic smem = warps * 2 * NS * 8.
__global__ void __launch_bounds__(256)
k_encode_wide(const uint64_t *__restrict__ data, uint32_t dwords, uint32_t data_bits, uint32_t dg,
              const uint64_t *__restrict__ rows, uint64_t *__restrict__ cw_out, uint32_t wpc, uint32_t batch)
{
  BCHFFT_DYN_SMEM(uint64_t, sm);
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const uint32_t cw = blockIdx.x * wpb + warp;
  if (cw >= batch) return;   // whole warps leave; the kernel uses no CTA-wide barrier
  const uint32_t NS = dg / 64u + 1u;
  uint64_t *cur = sm + (size_t)warp * 2u * NS, *nxt = cur + NS;
  const uint64_t *dsrc = data + (size_t)cw * dwords;
  for (uint32_t w = lane; w < 2u * NS; w += 32u) cur[w] = 0ull;
  __syncwarp();
  const uint32_t hw = dg >> 6, r = dg & 63u;
  const uint64_t hmask = r ? ((1ull << r) - 1ull) : 0ull;       // bits of word hw below dg
  const uint32_t lastd = data_bits ? (data_bits - 1u) >> 6 : 0u;
  auto dword = [&](uint32_t w) -> uint64_t {
    if (!data_bits || w > lastd) return 0ull;
    uint64_t x = dsrc[w];
    if (w == lastd && ((data_bits - 1u) & 63u) != 63u) x &= (2ull << ((data_bits - 1u) & 63u)) - 1ull;
    return x;
  };
  for (uint32_t c = data_bits ? lastd + 1u : 0u; c-- > 0;) {
    // bits [dg-64, dg) of S
    const uint64_t top = r ? ((cur[hw] << (64u - r)) | (cur[hw - 1u] >> r)) : cur[hw - 1u];
    const uint64_t u = top ^ dword(c);
    for (uint32_t w = lane; w < NS; w += 32u) {
      uint64_t v = w ? cur[w - 1u] : 0ull;                       // S << 64
      if (w == hw) v &= hmask;                                   // drop what moved to degree >= dg
      uint64_t bits = u;
      while (bits) {
        const uint32_t k = (uint32_t)__ffsll((long long)bits) - 1u;
        bits &= bits - 1ull;
        v ^= rows[(size_t)k * NS + w];
      }
      nxt[w] = v;
    }
    __syncwarp();
    uint64_t *t = cur; cur = nxt; nxt = t;
  }
  // codeword words: parity | data shifted up by dg
  uint64_t *out = cw_out + (size_t)cw * wpc;
  for (uint32_t j = lane; j < wpc; j += 32u) {
    uint64_t v = j < NS ? cur[j] : 0ull;
    if (data_bits && j >= hw) {
      const uint32_t i = j - hw;
      v |= dword(i) << r;
      if (r && i >= 1u) v |= dword(i - 1u) >> (64u - r);
    }
    out[j] = v;
  }
}

// ---- verdict + correction (bch.c:531-561) ---------------------------------------------------
// ok = 1, corrected = 0 for a clean word; L > t -> ok = 0, corrected = 0; otherwise corrected =
// #roots and the bits are flipped iff #roots == L.  Flips outside the packed array (only
// possible for shortened codes, where the reference writes out of bounds) are skipped.
// flip_on_device = 0 skips the flips (the caller applies them from `positions`).
// patch_dst (optional): the caller's page-locked HOST array for these codewords, addressed by the
// device (unified addressing).  Every 64-bit word a flip touched is then also written there, so the
// host array ends up corrected without the 1 KiB-per-codeword device-to-host copy: at most t
// eight-byte writes per codeword cross the bus instead (words nobody touched are already right --
// the host array is where the received words came from).
__global__ void k_correct(uint64_t *__restrict__ words, uint32_t words_per_cw, uint32_t batch,
                          const int32_t *__restrict__ flag, const uint32_t *__restrict__ Lin,
                          uint32_t tcap, const uint32_t *__restrict__ count,
                          const uint32_t *__restrict__ positions, uint32_t pos_stride,
                          uint8_t *__restrict__ ok, uint32_t *__restrict__ corrected,
                          uint64_t *__restrict__ patch_dst, uint32_t flip_on_device)
{
  const uint32_t cw = blockIdx.x * blockDim.x + threadIdx.x;
  if (cw >= batch) return;
  uint32_t good = 1u, found = 0u;
  if (flag[cw]) {
    const uint32_t L = Lin[cw];
    good = 0u;
    if (L <= tcap) {
      found = count[cw];
      if (found == L) good = 1u;
      // flip_on_device = 0: the host applies the flips itself from the position lists ("positions"
      // result mode of bchfft_bch_decode_host); only the verdict is needed here
      if (found == L && flip_on_device) {
        uint64_t *w = words + (size_t)cw * words_per_cw;
        if (!patch_dst) {
          // XOR reductions that return nothing: the thread issues all of its flips without waiting for any
          // (a read-modify-write per flip made this kernel a chain of found dependent DRAM round trips)
          const uint32_t *pp = positions + (size_t)cw * pos_stride;
          for (uint32_t i = 0; i < found; i += 4u) {
            uint32_t p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) p[u] = i + (uint32_t)u < found ? pp[i + (uint32_t)u] : 0xFFFFFFFFu;
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if ((p[u] >> 6) < words_per_cw)
                atomicXor(reinterpret_cast<unsigned long long *>(w + (p[u] >> 6)), 1ull << (p[u] & 63u));
          }
        } else {
          for (uint32_t i = 0; i < found; ++i) {
            const uint32_t p = positions[(size_t)cw * pos_stride + i];
            if ((p >> 6) < words_per_cw) w[p >> 6] ^= 1ull << (p & 63u);
          }
        }
        if (patch_dst) {
          // final values (two flips may share a word: both writes then carry the same value)
          uint64_t *hw = patch_dst + (size_t)cw * words_per_cw;
          for (uint32_t i = 0; i < found; ++i) {
            const uint32_t p = positions[(size_t)cw * pos_stride + i];
            if ((p >> 6) < words_per_cw) hw[p >> 6] = w[p >> 6];
          }
        }
      }
    }
  }
  if (ok) ok[cw] = (uint8_t)good;
  if (corrected) corrected[cw] = found;
}

// descending sort of each codeword's position list (bch.c:479-488 emits i ascending, i.e.
// positions descending); only the stage-level locate entry point needs the order.
__global__ void k_sort_positions(uint32_t *__restrict__ positions, uint32_t pos_stride,
                                 const uint32_t *__restrict__ count, uint32_t batch)
{
  const uint32_t cw = blockIdx.x * blockDim.x + threadIdx.x;
  if (cw >= batch) return;
  uint32_t *p = positions + (size_t)cw * pos_stride;
  const uint32_t n = count[cw] < pos_stride ? count[cw] : pos_stride;
  for (uint32_t i = 1; i < n; ++i) {
    const uint32_t v = p[i];
    uint32_t j = i;
    while (j > 0 && p[j - 1] < v) {
      p[j] = p[j - 1];
      --j;
    }
    p[j] = v;
  }
}

}  // namespace bchfft
