// Batched additive FFT over GF(2^m) on bitsliced slabs (replaces the stage loop of the
// reference's libfft/additive_fft.c:85-173 and the exp-order gather of :20).
//
// HBM layout of a slab buffer: [slab][position][PS] u32, PS = plane_stride(M); the PS words of
// one position are contiguous (one or more 16-byte vectors), so a tile made of any subset of
// position bits is read/written in whole 32-byte sectors.
// Shared-memory layout of a tile: S[plane][local position], lanes = consecutive positions.
#pragma once
#include "bs_common.cuh"
#include "gm_plan.h"

namespace bchfft {

constexpr int kMaxM = 25;

struct FftTables {
  const uint32_t *tw;
  const uint32_t *gam;
  uint32_t tw_off[kMaxM];
  uint32_t gam_off[kMaxM];
};

// ---- AoS u32 -> bitsliced slabs -----------------------------------------------------------
// src: [batch][item_stride] u32 (the reference ABI's coefficient arrays); coefficient index
// >= ncoef is read as zero (degree masking of additive_fft.c:99: coefficients 0..p_num_terms).
template <int M>
__global__ void k_bitslice_in(const uint32_t *__restrict__ src, size_t item_stride, uint32_t batch,
                              uint32_t ncoef, uint32_t npos_bits, uint32_t *__restrict__ dst)
{
  constexpr int PS = plane_stride(M);
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t slab = blockIdx.y;
  if (p >= (1u << npos_bits)) return;
  uint32_t x[32];
#pragma unroll
  for (int l = 0; l < 32; ++l) {
    const uint32_t item = slab * 32u + l;
    x[l] = (item < batch && p < ncoef) ? src[(size_t)item * item_stride + p] : 0u;
  }
  transpose32(x);
  uint4 *out = reinterpret_cast<uint4 *>(dst + ((((size_t)slab) << npos_bits) + p) * PS);
#pragma unroll
  for (int v = 0; v < PS / 4; ++v) {
    uint4 q;
    q.x = (4 * v + 0 < M) ? x[4 * v + 0] : 0u;
    q.y = (4 * v + 1 < M) ? x[4 * v + 1] : 0u;
    q.z = (4 * v + 2 < M) ? x[4 * v + 2] : 0u;
    q.w = (4 * v + 3 < M) ? x[4 * v + 3] : 0u;
    out[v] = q;
  }
}

// ---- tile helpers ---------------------------------------------------------------------------
struct TileGeom {
  uint32_t fixed, a0, n0, a1, mask0;
  __device__ __forceinline__ uint32_t pos(uint32_t l) const
  {
    return fixed | ((l & mask0) << a0) | ((l >> n0) << a1);
  }
  // local bit index of global position bit b (b must be inside the window)
  __device__ __forceinline__ uint32_t local_bit(uint32_t b) const
  {
    return (b >= a0 && b < a0 + n0) ? b - a0 : b - a1 + n0;
  }
};

__device__ __forceinline__ TileGeom tile_geom(const PassDesc &d, uint32_t tile)
{
  TileGeom g;
  g.a0 = d.a0;
  g.n0 = d.n0;
  g.a1 = d.a1;
  g.mask0 = (1u << d.n0) - 1u;
  const uint32_t g0 = d.a0, g1 = d.a1 - d.a0 - d.n0;
  g.fixed = (tile & ((1u << g0) - 1u)) | (((tile >> g0) & ((1u << g1) - 1u)) << (d.a0 + d.n0)) |
            ((tile >> (g0 + g1)) << (d.a1 + d.n1));
  return g;
}

template <int M>
__device__ __forceinline__ void tile_load(uint32_t *S, uint32_t T, const uint32_t *__restrict__ src,
                                          const TileGeom &g, uint32_t src_mask)
{
  constexpr int PS = plane_stride(M);
  for (uint32_t l = threadIdx.x; l < T; l += blockDim.x) {
    const uint32_t p = g.pos(l) & src_mask;
    const uint4 *in = reinterpret_cast<const uint4 *>(src + (size_t)p * PS);
#pragma unroll
    for (int v = 0; v < PS / 4; ++v) {
      const uint4 q = in[v];
      if (4 * v + 0 < M) S[(4 * v + 0) * T + l] = q.x;
      if (4 * v + 1 < M) S[(4 * v + 1) * T + l] = q.y;
      if (4 * v + 2 < M) S[(4 * v + 2) * T + l] = q.z;
      if (4 * v + 3 < M) S[(4 * v + 3) * T + l] = q.w;
    }
  }
}

template <int M>
__device__ __forceinline__ void tile_store(const uint32_t *S, uint32_t T, uint32_t *__restrict__ dst,
                                           const TileGeom &g)
{
  constexpr int PS = plane_stride(M);
  for (uint32_t l = threadIdx.x; l < T; l += blockDim.x) {
    uint4 *out = reinterpret_cast<uint4 *>(dst + (size_t)g.pos(l) * PS);
#pragma unroll
    for (int v = 0; v < PS / 4; ++v) {
      uint4 q;
      q.x = (4 * v + 0 < M) ? S[(4 * v + 0) * T + l] : 0u;
      q.y = (4 * v + 1 < M) ? S[(4 * v + 1) * T + l] : 0u;
      q.z = (4 * v + 2 < M) ? S[(4 * v + 2) * T + l] : 0u;
      q.w = (4 * v + 3 < M) ? S[(4 * v + 3) * T + l] : 0u;
      out[v] = q;
    }
  }
}

// TW(d): a[p] *= tw[d][p >> d]
template <int M>
__device__ __forceinline__ void tile_twist(uint32_t *S, uint32_t T, const TileGeom &g,
                                           const uint32_t *__restrict__ tw_d, uint32_t d)
{
  for (uint32_t l = threadIdx.x; l < T; l += blockDim.x) {
    const uint32_t c = tw_d[g.pos(l) >> d];
    if (c == 1u) continue;
    uint32_t a[M];
#pragma unroll
    for (int b = 0; b < M; ++b) a[b] = S[b * T + l];
    bs_mul_scalar<M>(a, c);
#pragma unroll
    for (int b = 0; b < M; ++b) S[b * T + l] = a[b];
  }
}

// TY(d, lo): one Taylor-expansion level at x^2+x on position bits (lo+1, lo): Q2 ^= Q3; Q1 ^= Q2
template <int M>
__device__ __forceinline__ void tile_taylor(uint32_t *S, uint32_t T, uint32_t lbl)
{
  const uint32_t groups = T >> 2, o1 = 1u << lbl, o2 = 2u << lbl;
  const uint32_t low_mask = o1 - 1u;
  for (uint32_t idx = threadIdx.x; idx < groups * M; idx += blockDim.x) {
    const uint32_t plane = idx / groups, gi = idx - plane * groups;
    const uint32_t base = ((gi & ~low_mask) << 2) | (gi & low_mask);
    uint32_t *row = S + plane * T + base;
    const uint32_t x3 = row[o2 + o1];
    const uint32_t x2 = row[o2] ^ x3;
    row[o2] = x2;
    row[o1] ^= x2;
  }
}

// BF(d): lo' = lo + gam[d][p >> (d+1)] * hi ; hi' = lo' + hi
template <int M>
__device__ __forceinline__ void tile_butterfly(uint32_t *S, uint32_t T, const TileGeom &g,
                                               const uint32_t *__restrict__ gam_d, uint32_t d, uint32_t lbd)
{
  const uint32_t o = 1u << lbd, low_mask = o - 1u;
  for (uint32_t j = threadIdx.x; j < (T >> 1); j += blockDim.x) {
    const uint32_t l = ((j & ~low_mask) << 1) | (j & low_mask);
    const uint32_t G = gam_d[g.pos(l) >> (d + 1)];
    uint32_t lo[M], hi[M];
#pragma unroll
    for (int b = 0; b < M; ++b) {
      lo[b] = S[b * T + l];
      hi[b] = S[b * T + l + o];
    }
    if (G) bs_mad_scalar<M>(lo, hi, G);
#pragma unroll
    for (int b = 0; b < M; ++b) {
      S[b * T + l] = lo[b];
      S[b * T + l + o] = lo[b] ^ hi[b];
    }
  }
}

template <int M>
__device__ __forceinline__ void tile_run_ops(uint32_t *S, uint32_t T, const TileGeom &g,
                                             const FftTables &tab, const PassDesc &desc)
{
  for (uint32_t i = 0; i < desc.nops; ++i) {
    const PassOp op = desc.ops[i];
    if (op.type == OP_TW) {
      tile_twist<M>(S, T, g, tab.tw + tab.tw_off[op.d], op.d);
    } else if (op.type == OP_TY) {
      tile_taylor<M>(S, T, g.local_bit(op.lo));
    } else {
      tile_butterfly<M>(S, T, g, tab.gam + tab.gam_off[op.d], op.d, g.local_bit(op.d));
    }
    __syncthreads();
  }
}

// ---- one pass over every tile of every slab -----------------------------------------------
// grid = (tiles per slab, slabs); dynamic smem = M * 2^t * 4 bytes.
template <int M>
__global__ void __launch_bounds__(512, 1)
k_gm_pass(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, size_t src_slab_words,
          size_t dst_slab_words, FftTables tab, PassDesc desc)
{
  BCHFFT_DYN_SMEM(uint32_t, S);
  const uint32_t T = 1u << desc.t;
  const TileGeom g = tile_geom(desc, blockIdx.x);
  tile_load<M>(S, T, src + (size_t)blockIdx.y * src_slab_words, g, desc.src_mask);
  __syncthreads();
  tile_run_ops<M>(S, T, g, tab, desc);
  tile_store<M>(S, T, dst + (size_t)blockIdx.y * dst_slab_words, g);
}

// ---- bitsliced slabs -> AoS u32, through a position permutation ---------------------------
// out[item][i] = value at position perm[i]; perm = bitrev_m(exp[i]) gives the reference's
// po_result (additive_fft.c:20), perm = bitrev_m(k) gives the clobbered p_poly (natural order).
template <int M>
__global__ void k_gather_out(const uint32_t *__restrict__ slabs, size_t slab_words,
                             const uint32_t *__restrict__ perm, uint32_t nout,
                             uint32_t *__restrict__ out, size_t item_stride, uint32_t batch)
{
  constexpr int PS = plane_stride(M);
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t slab = blockIdx.y;
  if (i >= nout) return;
  const uint4 *in = reinterpret_cast<const uint4 *>(slabs + (size_t)slab * slab_words + (size_t)perm[i] * PS);
  uint32_t x[32];
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (v < PS / 4) q = in[v];
    x[4 * v + 0] = (4 * v + 0 < M) ? q.x : 0u;
    x[4 * v + 1] = (4 * v + 1 < M) ? q.y : 0u;
    x[4 * v + 2] = (4 * v + 2 < M) ? q.z : 0u;
    x[4 * v + 3] = (4 * v + 3 < M) ? q.w : 0u;
  }
  transpose32(x);
#pragma unroll
  for (int l = 0; l < 32; ++l) {
    const uint32_t item = slab * 32u + l;
    if (item < batch) out[(size_t)item * item_stride + i] = x[l];
  }
}

}  // namespace bchfft
