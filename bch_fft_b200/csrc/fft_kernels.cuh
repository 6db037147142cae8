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

#ifdef BCHFFT_TRACE
// debug instrumentation: CTA (0,0) records (tag, clock64) pairs at sweep boundaries
__device__ unsigned long long g_trace[512];
__device__ unsigned int g_trace_n;
__device__ __forceinline__ void trace_mark(unsigned tag)
{
  if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
    const unsigned k = g_trace_n;
    if (k < 255) {
      g_trace[2 * k] = tag;
      g_trace[2 * k + 1] = clock64();
      g_trace_n = k + 1;
    }
  }
}
#define BCHFFT_MARK(tag) trace_mark(tag)
#else
#define BCHFFT_MARK(tag) ((void)0)
#endif

struct FftTables {
  const uint32_t *tw;
  const uint32_t *gam;
  uint32_t tw_off[kMaxM];
  uint32_t gam_off[kMaxM];
};

// ---- AoS u32 -> bitsliced slabs -----------------------------------------------------------
// src: [batch][item_stride] u32 (the reference ABI's coefficient arrays); coefficient index
// >= ncoef is read as zero (degree masking of additive_fft.c:99: coefficients 0..p_num_terms).
// plane_major = 1: dst is [slab][plane][position] (see k_ty_plane_pass) instead of
// [slab][position][PS].
template <int M>
__global__ void k_bitslice_in(const uint32_t *__restrict__ src, size_t item_stride, uint32_t batch,
                              uint32_t ncoef, uint32_t npos_bits, uint32_t *__restrict__ dst,
                              uint32_t plane_major = 0u)
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
  if (plane_major) {
    uint32_t *row = dst + ((((size_t)slab * PS) << npos_bits) + p);
#pragma unroll
    for (int b = 0; b < M; ++b) row[(size_t)b << npos_bits] = x[b];
    return;
  }
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
  uint32_t fixed, a0, n0, a1, mask0, pmask;
  __device__ __forceinline__ uint32_t pos(uint32_t l) const
  {
    return (fixed | ((l & mask0) << a0) | ((l >> n0) << a1)) & pmask;
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
  g.pmask = 0xFFFFFFFFu;
  const uint32_t g0 = d.a0, g1 = d.a1 - d.a0 - d.n0;
  g.fixed = (tile & ((1u << g0) - 1u)) | (((tile >> g0) & ((1u << g1) - 1u)) << (d.a0 + d.n0)) |
            ((tile >> (g0 + g1)) << (d.a1 + d.n1));
  return g;
}

// Shared-memory tiles are XOR-swizzled: local position l lives at word
// l ^ l[9:5] ^ l[14:10] of its plane, i.e. bank = l[4:0] ^ l[9:5] ^ l[14:10].  A warp whose lanes
// differ in five local bits with distinct residues mod 5 is conflict-free, whichever bits the
// threads themselves iterate over.  The map is XOR-linear: swz(a | b) = swz(a) ^ swz(b) for
// disjoint a, b, so a thread that owns a group of bits adds per-element constants to one base.
__device__ __forceinline__ uint32_t swz(uint32_t l) { return l ^ ((l >> 5) & 31u) ^ ((l >> 10) & 31u); }

// Enumeration of the 2^(t-r) work items of a sweep in which every thread owns the 2^r positions
// spanned by local bits [g0, g0+r).  Item bits 0..4 (the lanes of a warp) go to local bit i, or
// to i+5 when bit i belongs to the thread-owned group, so that the swizzle spreads them over all
// 32 banks; the remaining item bits fill the other free bits in ascending order.
struct SweepMap {
  uint32_t g5;            // lane bits that are shifted up by 5
  uint32_t start[4], len[4];
  uint32_t simple, g0, r;
  __device__ __forceinline__ uint32_t pos(uint32_t w) const
  {
    if (simple) return ((w >> g0) << (g0 + r)) | (w & ((1u << g0) - 1u));
    const uint32_t lane = w & 31u;
    uint32_t rest = w >> 5;
    uint32_t l = (lane & ~g5) | ((lane & g5) << 5);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      l |= (rest & ((1u << len[k]) - 1u)) << start[k];
      rest >>= len[k];
    }
    return l;
  }
};

__device__ __forceinline__ SweepMap make_sweep_map(uint32_t t, uint32_t g0, uint32_t r)
{
  SweepMap m;
  m.g0 = g0;
  m.r = r;
  m.g5 = 0u;
#pragma unroll
  for (int k = 0; k < 4; ++k) m.start[k] = m.len[k] = 0u;
  m.simple = (t < 10u || g0 >= 5u || r == 0u) ? 1u : 0u;
  if (m.simple) return m;
  const uint32_t group = ((1u << r) - 1u) << g0;
  m.g5 = group & 31u;
  uint32_t free_rest = ((1u << t) - 1u) & ~group & ~(31u & ~m.g5) & ~(m.g5 << 5);
  int k = 0;
  while (free_rest && k < 4) {
    const uint32_t s = (uint32_t)__ffs((int)free_rest) - 1u;
    uint32_t n = 0;
    while ((free_rest >> (s + n)) & 1u) n++;
    m.start[k] = s;
    m.len[k] = n;
    free_rest &= ~(((1u << n) - 1u) << s);
    k++;
  }
  return m;
}

// General form of SweepMap (any group position, tiles up to 2^15; used by the OP_TX register
// sweeps).  Item bits 0..4 (the lanes of a warp) go to five free local
// bits with distinct residues mod 5 (bit i, else i+5, else i+10), so that the swizzle spreads them
// over all 32 banks; the remaining item bits fill the other free bits in ascending order.
struct SweepMapX {
  uint32_t lane_bit[5];
  uint32_t start[8], len[8];
  uint32_t simple, g0, r;
  __device__ __forceinline__ uint32_t lane_part(uint32_t lane) const
  {
    uint32_t l = 0u;
#pragma unroll
    for (int j = 0; j < 5; ++j) l |= ((lane >> j) & 1u) << lane_bit[j];
    return l;
  }
  __device__ __forceinline__ uint32_t rest_part(uint32_t rest) const
  {
    uint32_t l = 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      l |= (rest & ((1u << len[k]) - 1u)) << start[k];
      rest >>= len[k];
    }
    return l;
  }
  __device__ __forceinline__ uint32_t pos(uint32_t w) const
  {
    if (simple) return ((w >> g0) << (g0 + r)) | (w & ((1u << g0) - 1u));
    return lane_part(w & 31u) | rest_part(w >> 5);
  }
};

__device__ __forceinline__ SweepMapX make_sweep_map_x(uint32_t t, uint32_t g0, uint32_t r)
{
  SweepMapX m;
  m.g0 = g0;
  m.r = r;
#pragma unroll
  for (int k = 0; k < 8; ++k) m.start[k] = m.len[k] = 0u;
#pragma unroll
  for (int j = 0; j < 5; ++j) m.lane_bit[j] = 0u;
  m.simple = (t < 10u || g0 >= 5u || r == 0u) ? 1u : 0u;
  if (m.simple) return m;
  const uint32_t group = ((1u << r) - 1u) << g0;
  uint32_t free_bits = ((1u << t) - 1u) & ~group;
  int nl = 0;
  for (uint32_t q = 0; q < 5u; ++q) {
    for (uint32_t c = q; c < t; c += 5u)
      if ((free_bits >> c) & 1u) {
        m.lane_bit[nl++] = c;
        free_bits &= ~(1u << c);
        break;
      }
  }
  while (nl < 5 && free_bits) {   // not enough residues: take what is left (conflicts accepted)
    const uint32_t c = (uint32_t)__ffs((int)free_bits) - 1u;
    m.lane_bit[nl++] = c;
    free_bits &= ~(1u << c);
  }
  int k = 0;
  while (free_bits && k < 8) {
    const uint32_t s = (uint32_t)__ffs((int)free_bits) - 1u;
    uint32_t n = 0;
    while ((free_bits >> (s + n)) & 1u) n++;
    m.start[k] = s;
    m.len[k] = n;
    free_bits &= ~(((1u << n) - 1u) << s);
    k++;
  }
  return m;
}

// Global <-> shared tile copies.  Four positions per thread are in flight at once (all vector
// loads issued before the first use) so that the copy is bandwidth- rather than latency-bound.
template <int M>
__device__ __forceinline__ void tile_load(uint32_t *S, uint32_t T, const uint32_t *__restrict__ src,
                                          const TileGeom &g, uint32_t src_mask)
{
  constexpr int PS = plane_stride(M);
  constexpr int U = 4;
  for (uint32_t base = threadIdx.x; base < T; base += U * blockDim.x) {
    uint4 q[U][PS / 4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t l = base + u * blockDim.x;
      if (l < T) {
        const uint4 *in = reinterpret_cast<const uint4 *>(src + (size_t)(g.pos(l) & src_mask) * PS);
#pragma unroll
        for (int v = 0; v < PS / 4; ++v) q[u][v] = in[v];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t l = base + u * blockDim.x;
      if (l < T) {
        const uint32_t sl = swz(l);
#pragma unroll
        for (int v = 0; v < PS / 4; ++v) {
          if (4 * v + 0 < M) S[(4 * v + 0) * T + sl] = q[u][v].x;
          if (4 * v + 1 < M) S[(4 * v + 1) * T + sl] = q[u][v].y;
          if (4 * v + 2 < M) S[(4 * v + 2) * T + sl] = q[u][v].z;
          if (4 * v + 3 < M) S[(4 * v + 3) * T + sl] = q[u][v].w;
        }
      }
    }
  }
}

template <int M>
__device__ __forceinline__ void tile_store(const uint32_t *S, uint32_t T, uint32_t *__restrict__ dst,
                                           const TileGeom &g)
{
  constexpr int PS = plane_stride(M);
  constexpr int U = 2;
  for (uint32_t base = threadIdx.x; base < T; base += U * blockDim.x) {
    uint4 q[U][PS / 4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t l = base + u * blockDim.x;
      if (l < T) {
        const uint32_t sl = swz(l);
#pragma unroll
        for (int v = 0; v < PS / 4; ++v) {
          q[u][v].x = (4 * v + 0 < M) ? S[(4 * v + 0) * T + sl] : 0u;
          q[u][v].y = (4 * v + 1 < M) ? S[(4 * v + 1) * T + sl] : 0u;
          q[u][v].z = (4 * v + 2 < M) ? S[(4 * v + 2) * T + sl] : 0u;
          q[u][v].w = (4 * v + 3 < M) ? S[(4 * v + 3) * T + sl] : 0u;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t l = base + u * blockDim.x;
      if (l < T) {
        uint4 *out = reinterpret_cast<uint4 *>(dst + (size_t)g.pos(l) * PS);
#pragma unroll
        for (int v = 0; v < PS / 4; ++v) out[v] = q[u][v];
      }
    }
  }
}

// TW(d): a[p] *= tw[d][p >> d].  UNI: when the five lowest local bits (the lanes of a warp) all map
// to global position bits below d, the multiplier is warp-uniform: branch over its set bits.
template <int M, bool UNI = false>
__device__ __forceinline__ void tile_twist(uint32_t *S, uint32_t T, const TileGeom &g,
                                           const uint32_t *__restrict__ tw_d, uint32_t d)
{
  // highest global bit driven by local bits 0..4
  const uint32_t top_lane_bit = g.n0 >= 5u ? g.a0 + 4u : g.a1 + (4u - g.n0);
  const bool uni = UNI && T >= 32u && top_lane_bit < d;
  constexpr int U = 4;   // constants of four positions are fetched before the first multiply
  for (uint32_t base = threadIdx.x; base < T; base += U * blockDim.x) {
    uint32_t c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t l = base + u * blockDim.x;
      c[u] = (l < T) ? tw_d[g.pos(l) >> d] : 1u;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (c[u] == 1u) continue;
      const uint32_t sl = swz(base + u * blockDim.x);
      uint32_t a[M];
#pragma unroll
      for (int b = 0; b < M; ++b) a[b] = S[b * T + sl];
      if (UNI && uni) {
        uint32_t r[M];
#pragma unroll
        for (int b = 0; b < M; ++b) r[b] = 0u;
        bs_mad_uniform<M>(r, a, c[u]);
#pragma unroll
        for (int b = 0; b < M; ++b) S[b * T + sl] = r[b];
      } else {
        bs_mul_scalar<M>(a, c[u]);
#pragma unroll
        for (int b = 0; b < M; ++b) S[b * T + sl] = a[b];
      }
    }
  }
}

// R-1 consecutive Taylor-expansion levels (at x^2+x) on local bits [g0, g0+R): the level on
// bit pair (lv+1, lv), lv = R-2 .. 0, does Q2 ^= Q3 then Q1 ^= Q2.  XOR only and plane-wise, so
// a work item is (plane, block of 2^R positions) held in registers.
template <int R>
__device__ __forceinline__ void tile_taylor_multi(uint32_t *S, uint32_t T, uint32_t t, uint32_t g0,
                                                  uint32_t nplanes)
{
  const SweepMap map = make_sweep_map(t, g0, R);
  const uint32_t blocks = T >> R;
  for (uint32_t idx = threadIdx.x; idx < blocks * nplanes; idx += blockDim.x) {
    const uint32_t plane = idx / blocks, w = idx - plane * blocks;
    const uint32_t sb = swz(map.pos(w));
    uint32_t *row = S + plane * T;
    uint32_t x[1 << R];
#pragma unroll
    for (int k = 0; k < (1 << R); ++k) x[k] = row[sb ^ swz((uint32_t)k << g0)];
#pragma unroll
    for (int lv = R - 2; lv >= 0; --lv) {
#pragma unroll
      for (int k = 0; k < (1 << R); ++k)
        if (((k >> lv) & 3) == 2) x[k] ^= x[k + (1 << lv)];
#pragma unroll
      for (int k = 0; k < (1 << R); ++k)
        if (((k >> lv) & 3) == 1) x[k] ^= x[k + (1 << lv)];
    }
#pragma unroll
    for (int k = 0; k < (1 << R); ++k) row[sb ^ swz((uint32_t)k << g0)] = x[k];
  }
}

// ---- generalised Taylor levels (OP_TX, see gm_plan.h) ---------------------------------------
// One level, any base: every thread takes targets B_i[off], i = 1 .. t-1, of a super-block; the
// thread that owns B_1[off] also does B_t[off] ^= B_(2t-1)[off] first (B_t is read by nobody
// else), so the level needs no barrier of its own.  L = local bit of the block size, h = log2 t.
static __device__ __noinline__ void tile_tx_generic(uint32_t *S, uint32_t T, uint32_t L, uint32_t h,
                                                uint32_t nplanes)
{
  const uint32_t half = T >> 1;                 // candidate targets per plane: blocks 0 .. t-1
  const uint32_t tm1 = (1u << h) - 1u;
  const uint32_t off = tm1 << L, lowmask = (1u << (L + h)) - 1u;
  for (uint32_t idx = threadIdx.x; idx < half * nplanes; idx += blockDim.x) {
    const uint32_t plane = idx / half, w = idx - plane * half;
    const uint32_t blk = (w >> L) & tm1;
    if (blk == 0u) continue;
    const uint32_t l = ((w & ~lowmask) << 1) | (w & lowmask);   // zero at the top bit of the block index
    uint32_t *row = S + plane * T;
    uint32_t src = row[swz(l + off)];
    if (blk == 1u) {
      // l + off lies in B_t; its own source B_(2t-1) is another t-1 blocks up
      src ^= row[swz(l + 2u * off)];
      row[swz(l + off)] = src;
    }
    row[swz(l)] ^= src;
  }
}

// compile-time level on a register block indexed by R local bits: REL = block-size bit relative to
// the group, H = log2 t.  Descending order makes B_t final before B_1 reads it.
template <int R, int REL, int H>
__device__ __forceinline__ void tx_level_regs(uint32_t (&x)[1 << R])
{
  constexpr int t = 1 << H;
  static_assert(REL + H + 1 <= R, "level does not fit the register block");
#pragma unroll
  for (int k = (1 << R) - 1; k >= 0; --k) {
    const int blk = (k >> REL) & (2 * t - 1);
    if (blk >= 1 && blk <= t) x[k] ^= x[k + ((t - 1) << REL)];
  }
}

// Register-resident sweeps over the level sequences the Cantor conversion produces:
//   TXP_C4   (R = 4)  TX(b+1,2) TX(b,2) TX(b,1) TX(b+2,1)            = whole conversion of 4 bits
//   TXP_X16  (R = 6)  TX(b+1,4) TX(b,4)                              two levels of base y^16 + y
//   TXP_X16C (R = 6)  TX(b+1,4) TX(b,4) followed by TXP_C4 at b      (last two levels + low conversion)
enum { TXP_C4 = 0, TXP_X16 = 1, TXP_X16C = 2 };

template <int PROG, int G0>
__device__ __noinline__ void tile_tx_regs_at(uint32_t *S, uint32_t T, uint32_t t, uint32_t nplanes)
{
  constexpr int R = (PROG == TXP_C4) ? 4 : 6;
  const SweepMapX map = make_sweep_map_x(t, (uint32_t)G0, R);
  const uint32_t blocks = T >> R;
  for (uint32_t idx = threadIdx.x; idx < blocks * nplanes; idx += blockDim.x) {
    const uint32_t plane = idx / blocks, w = idx - plane * blocks;
    const uint32_t sb = swz(map.pos(w));
    uint32_t *row = S + plane * T;
    uint32_t x[1 << R];
    // G0 is a compile-time constant: the per-element swizzle constants are immediates
#pragma unroll
    for (int k = 0; k < (1 << R); ++k) x[k] = row[sb ^ swz((uint32_t)k << G0)];
    if constexpr (PROG == TXP_X16 || PROG == TXP_X16C) {
      tx_level_regs<R, 1, 4>(x);
      tx_level_regs<R, 0, 4>(x);
    }
    if constexpr (PROG == TXP_C4 || PROG == TXP_X16C) {
      tx_level_regs<R, 1, 2>(x);
      tx_level_regs<R, 0, 2>(x);
      tx_level_regs<R, 0, 1>(x);
      tx_level_regs<R, 2, 1>(x);
    }
#pragma unroll
    for (int k = 0; k < (1 << R); ++k) row[sb ^ swz((uint32_t)k << G0)] = x[k];
  }
}

// group positions 0 .. 11 cover every window of tiles up to 2^15 positions (R <= 6 bits above g0)
template <int PROG>
__device__ __forceinline__ bool tile_tx_regs(uint32_t *S, uint32_t T, uint32_t t, uint32_t g0, uint32_t nplanes)
{
  switch (g0) {
    case 0: tile_tx_regs_at<PROG, 0>(S, T, t, nplanes); return true;
    case 1: tile_tx_regs_at<PROG, 1>(S, T, t, nplanes); return true;
    case 2: tile_tx_regs_at<PROG, 2>(S, T, t, nplanes); return true;
    case 3: tile_tx_regs_at<PROG, 3>(S, T, t, nplanes); return true;
    case 4: tile_tx_regs_at<PROG, 4>(S, T, t, nplanes); return true;
    case 5: tile_tx_regs_at<PROG, 5>(S, T, t, nplanes); return true;
    case 6: tile_tx_regs_at<PROG, 6>(S, T, t, nplanes); return true;
    case 7: tile_tx_regs_at<PROG, 7>(S, T, t, nplanes); return true;
    case 8: tile_tx_regs_at<PROG, 8>(S, T, t, nplanes); return true;
    case 9: tile_tx_regs_at<PROG, 9>(S, T, t, nplanes); return true;
    case 10: tile_tx_regs_at<PROG, 10>(S, T, t, nplanes); return true;
    case 11: tile_tx_regs_at<PROG, 11>(S, T, t, nplanes); return true;
    default: return false;
  }
}

// ---- line sweeps: several consecutive levels of one base in registers ------------------------
// Levels kappa = kmin+NLEV-1 .. kmin of base t = 2^h only ever combine positions that differ by
// multiples of u = (t-1) 2^kmin, and stay inside super-blocks of 2^(kmin+NLEV+h) positions.  A
// thread therefore owns the line r + u c, c = 0 .. 2^NLEV (r < u; the last element exists for
// r < 2^(NLEV+kmin) and is only ever a source), runs all NLEV levels on it in registers and writes
// it back: two memory accesses per element for the whole group of levels.  Level kappa pairs c
// with c + 2^(kappa-kmin) where the block index of the target is in [1, t]; descending c makes B_t
// final before B_1 reads it.  q0 = r >> kmin.
template <int NLEV>
__device__ __forceinline__ void tx_line_levels(uint32_t (&x)[(1 << NLEV) + 1], uint32_t q0, uint32_t h)
{
  constexpr int NE = (1 << NLEV) + 1;
  const uint32_t t = 1u << h, tm1 = t - 1u, m2 = 2u * t - 1u;
#pragma unroll
  for (int lv = NLEV - 1; lv >= 0; --lv) {
#pragma unroll
    for (int c = NE - 1 - (1 << lv); c >= 0; --c) {
      const uint32_t blk = ((q0 + tm1 * (uint32_t)c) >> lv) & m2;
      if (blk - 1u < t) x[c] ^= x[c + (1 << lv)];
    }
  }
}

// on a shared-memory tile: L = local bit of kmin
template <int NLEV>
__device__ __noinline__ void tile_tx_lines(uint32_t *S, uint32_t T, uint32_t L, uint32_t h, uint32_t nplanes)
{
  constexpr int NE = (1 << NLEV) + 1;
  const uint32_t u = ((1u << h) - 1u) << L;
  const uint32_t sbits = L + NLEV + h;
  const uint32_t nsuper = T >> sbits;
  const uint32_t items = nsuper * u;
  for (uint32_t idx = threadIdx.x; idx < items * nplanes; idx += blockDim.x) {
    const uint32_t plane = idx / items, w = idx - plane * items;
    const uint32_t sb = w / u, r = w - sb * u;
    uint32_t *row = S + plane * T;
    const uint32_t l0 = (sb << sbits) + r;
    const bool last = r < (1u << (NLEV + L));
    uint32_t x[NE];
#pragma unroll
    for (int c = 0; c < NE - 1; ++c) x[c] = row[swz(l0 + u * (uint32_t)c)];
    x[NE - 1] = last ? row[swz(l0 + u * (uint32_t)(NE - 1))] : 0u;
    tx_line_levels<NLEV>(x, r >> L, h);
#pragma unroll
    for (int c = 0; c < NE - 1; ++c) row[swz(l0 + u * (uint32_t)c)] = x[c];
  }
}

__device__ __forceinline__ void tile_tx_lines_dispatch(uint32_t *S, uint32_t T, uint32_t L, uint32_t h,
                                                       uint32_t nlev, uint32_t nplanes)
{
  if (nlev == 6) tile_tx_lines<6>(S, T, L, h, nplanes);
  else if (nlev == 5) tile_tx_lines<5>(S, T, L, h, nplanes);
  else if (nlev == 4) tile_tx_lines<4>(S, T, L, h, nplanes);
  else if (nlev == 3) tile_tx_lines<3>(S, T, L, h, nplanes);
  else tile_tx_lines<2>(S, T, L, h, nplanes);
}

// directly on a plane-major slab in HBM (no tile): one kernel pass for levels whose super-block
// is larger than a tile.  Consecutive threads own consecutive r, so every load / store of a warp is
// one contiguous row segment.  grid = (ceil(u / threads) * super-blocks per row, planes, slabs).
template <int NLEV>
__global__ void __launch_bounds__(256)
k_tx_line_pass(uint32_t *__restrict__ buf, uint32_t nbits, uint32_t ps, uint32_t kmin, uint32_t h)
{
  constexpr int NE = (1 << NLEV) + 1;
  const uint32_t u = ((1u << h) - 1u) << kmin;
  const uint32_t sbits = kmin + NLEV + h;
  const uint32_t per = (u + blockDim.x - 1u) / blockDim.x;
  const uint32_t sb = blockIdx.x / per, r = (blockIdx.x - sb * per) * blockDim.x + threadIdx.x;
  if (r >= u) return;
  uint32_t *row = buf + (((size_t)blockIdx.z * ps + blockIdx.y) << nbits) + ((size_t)sb << sbits) + r;
  const bool last = r < (1u << (NLEV + kmin));
  uint32_t x[NE];
#pragma unroll
  for (int c = 0; c < NE - 1; ++c) x[c] = row[(size_t)u * c];
  x[NE - 1] = last ? row[(size_t)u * (NE - 1)] : 0u;
  tx_line_levels<NLEV>(x, r >> kmin, h);
#pragma unroll
  for (int c = 0; c < NE - 1; ++c) row[(size_t)u * c] = x[c];
}

// Runs the OP_TX ops starting at op i that form one sweep; returns how many ops it consumed.
// Every sweep ends with the data in shared memory; the caller puts the barrier.
__device__ __forceinline__ uint32_t tile_tx_sweep(uint32_t *S, uint32_t T, uint32_t t, const TileGeom &g,
                                                  const PassDesc &desc, uint32_t i, uint32_t nplanes)
{
  const PassOp op = desc.ops[i];
  auto is_tx = [&](uint32_t k, uint32_t lo, uint32_t h) -> bool {
    return k < desc.nops && desc.ops[k].type == OP_TX && desc.ops[k].lo == lo && desc.ops[k].d == h;
  };
  auto is_c4 = [&](uint32_t k, uint32_t b) -> bool {
    return is_tx(k, b + 1u, 2u) && is_tx(k + 1u, b, 2u) && is_tx(k + 2u, b, 1u) && is_tx(k + 3u, b + 2u, 1u);
  };
  if (op.d >= BCHFFT_TX_LINES_MIN_H) {
    // run of levels of one large base: line sweep (at most six levels, super-block inside the tile)
    uint32_t n = 1u;
    while (n < 6u && op.lo >= n && is_tx(i + n, op.lo - n, op.d) &&
           g.local_bit(op.lo - n) + n == g.local_bit(op.lo))
      n++;
    if (n >= 2u && g.local_bit(op.lo) + op.d + 1u <= t) {
      tile_tx_lines_dispatch(S, T, g.local_bit(op.lo - (n - 1u)), op.d, n, nplanes);
      return n;
    }
  }
  if (op.d == 4u && op.lo >= 1u && t >= 6u && is_tx(i + 1u, op.lo - 1u, 4u)) {
    const uint32_t b = op.lo - 1u;
    if (is_c4(i + 2u, b)) {
      if (tile_tx_regs<TXP_X16C>(S, T, t, g.local_bit(b), nplanes)) return 6u;
    } else if (tile_tx_regs<TXP_X16>(S, T, t, g.local_bit(b), nplanes)) {
      return 2u;
    }
  } else if (op.d == 2u && op.lo >= 1u && t >= 4u && is_c4(i, op.lo - 1u)) {
    if (tile_tx_regs<TXP_C4>(S, T, t, g.local_bit(op.lo - 1u), nplanes)) return 4u;
  }
  tile_tx_generic(S, T, g.local_bit(op.lo), op.d, nplanes);
  return 1u;
}

// Epilogue hook of the last butterfly sweep of a pass: instead of writing the final values back
// to the tile, hand them to `epi(local position, planes)` (the root search tests them for zero).
struct NoEpilogue {
  __device__ __forceinline__ void operator()(uint32_t, const uint32_t *) const {}
};

__device__ __forceinline__ void tile_taylor_dispatch(uint32_t *S, uint32_t T, uint32_t t, uint32_t g0,
                                                     uint32_t levels, uint32_t nplanes)
{
  if (levels == 5) tile_taylor_multi<6>(S, T, t, g0, nplanes);
  else if (levels == 4) tile_taylor_multi<5>(S, T, t, g0, nplanes);
  else if (levels == 3) tile_taylor_multi<4>(S, T, t, g0, nplanes);
  else if (levels == 2) tile_taylor_multi<3>(S, T, t, g0, nplanes);
  else tile_taylor_multi<2>(S, T, t, g0, nplanes);
}

// number of consecutive Taylor ops starting at op i that one register-resident sweep can take:
// same depth, lo and local bit both descending by one, at most five levels (a six-bit group
// that reaches into the five lowest local bits cannot be made bank-conflict free by the
// swizzle: those stay at four)
__device__ __forceinline__ uint32_t taylor_run(const PassDesc &desc, const TileGeom &g, uint32_t i,
                                               uint32_t max_levels = 5u)
{
  const PassOp op = desc.ops[i];
  uint32_t c = 1;
  while (c < max_levels && i + c < desc.nops) {
    const PassOp nx = desc.ops[i + c];
    if (nx.type != OP_TY || nx.d != op.d || nx.lo + c != op.lo ||
        g.local_bit(nx.lo) + c != g.local_bit(op.lo))
      break;
    c++;
  }
  if (c == 5 && g.local_bit(op.lo) + 1u - c < 5u) c = 4;
  return c;
}

// BF(d): lo' = lo + gam[d][p >> (d+1)] * hi ; hi' = lo' + hi
// UNI: when the butterfly bit sits above local bit 4 the 32 lanes of a warp (local bits 0..4)
// belong to one block, i.e. share the multiplier: it is warp-uniform and the multiply branches
// over its set bits (bs_mad_uniform) instead of masking all M*M partial products.
template <int M, class Epi, bool UNI = false>
__device__ __forceinline__ void tile_butterfly(uint32_t *S, uint32_t T, uint32_t t, const TileGeom &g,
                                               const uint32_t *__restrict__ gam_d, uint32_t d, uint32_t lbd,
                                               const Epi *epi)
{
  const bool uni = UNI && lbd >= 5u;
  const SweepMap map = make_sweep_map(t, lbd, 1);
  const uint32_t o = 1u << lbd;
  for (uint32_t j = threadIdx.x; j < (T >> 1); j += blockDim.x) {
    const uint32_t l = map.pos(j);
    const uint32_t G = gam_d[g.pos(l) >> (d + 1)];
    const uint32_t s0 = swz(l), s1 = swz(l + o);
    uint32_t lo[M], hi[M];
#pragma unroll
    for (int b = 0; b < M; ++b) {
      lo[b] = S[b * T + s0];
      hi[b] = S[b * T + s1];
    }
    if (UNI && uni) {
      if (G) bs_mad_uniform<M>(lo, hi, G);
    } else if (G) {
      bs_mad_scalar<M>(lo, hi, G);
    }
#pragma unroll
    for (int b = 0; b < M; ++b) hi[b] ^= lo[b];
    if (epi) {
      (*epi)(l, lo);
      (*epi)(l + o, hi);
    } else {
#pragma unroll
      for (int b = 0; b < M; ++b) {
        S[b * T + s0] = lo[b];
        S[b * T + s1] = hi[b];
      }
    }
  }
}

// BF(d) then BF(d-1) on four positions held in registers (local bits lb, lb+1 = bits d-1, d)
template <int M, class Epi, bool UNI = false>
__device__ __forceinline__ void tile_butterfly_r4(uint32_t *S, uint32_t T, uint32_t t, const TileGeom &g,
                                                  const uint32_t *__restrict__ gam_d,
                                                  const uint32_t *__restrict__ gam_dm1, uint32_t d,
                                                  uint32_t lb, const Epi *epi)
{
  const bool uni = UNI && lb >= 5u;
  const SweepMap map = make_sweep_map(t, lb, 2);
  for (uint32_t j = threadIdx.x; j < (T >> 2); j += blockDim.x) {
    const uint32_t l = map.pos(j);
    const uint32_t p0 = g.pos(l);
    const uint32_t s0 = swz(l), s1 = swz(l + (1u << lb)), s2 = swz(l + (2u << lb)), s3 = swz(l + (3u << lb));
    uint32_t e0[M], e1[M], e2[M], e3[M];
#pragma unroll
    for (int b = 0; b < M; ++b) {
      e0[b] = S[b * T + s0];
      e1[b] = S[b * T + s1];
      e2[b] = S[b * T + s2];
      e3[b] = S[b * T + s3];
    }
    const uint32_t G = gam_d[p0 >> (d + 1)];
    if (UNI && uni) {
      if (G) {
#if BCHFFT_R4_MAD2
        bs_mad_uniform2<M>(e0, e2, e1, e3, G);
#else
        bs_mad_uniform<M>(e0, e2, G);
        bs_mad_uniform<M>(e1, e3, G);
#endif
      }
    } else if (G) {
      bs_mad_scalar<M>(e0, e2, G);
      bs_mad_scalar<M>(e1, e3, G);
    }
#pragma unroll
    for (int b = 0; b < M; ++b) {
      e2[b] ^= e0[b];
      e3[b] ^= e1[b];
    }
    const uint32_t Ga = gam_dm1[p0 >> d], Gb = gam_dm1[(p0 >> d) + 1u];
    if (UNI && uni) {
      if (Ga) bs_mad_uniform<M>(e0, e1, Ga);
      if (Gb) bs_mad_uniform<M>(e2, e3, Gb);
    } else {
      if (Ga) bs_mad_scalar<M>(e0, e1, Ga);
      if (Gb) bs_mad_scalar<M>(e2, e3, Gb);
    }
#pragma unroll
    for (int b = 0; b < M; ++b) {
      e1[b] ^= e0[b];
      e3[b] ^= e2[b];
    }
    if (epi) {
      (*epi)(l, e0);
      (*epi)(l + (1u << lb), e1);
      (*epi)(l + (2u << lb), e2);
      (*epi)(l + (3u << lb), e3);
    } else {
#pragma unroll
      for (int b = 0; b < M; ++b) {
        S[b * T + s0] = e0[b];
        S[b * T + s1] = e1[b];
        S[b * T + s2] = e2[b];
        S[b * T + s3] = e3[b];
      }
    }
  }
}

// Runs the ops of a pass on a tile.  Consecutive Taylor levels of one depth on adjacent bits are
// fused five at a time, consecutive butterfly levels two at a time (register-resident radix
// blocks): fewer shared-memory round trips and barriers than one sweep per level.
template <int M, class Epi = NoEpilogue, bool WITH_TX = false, bool UNI = false>
__device__ __forceinline__ void tile_run_ops(uint32_t *S, uint32_t T, uint32_t t, const TileGeom &g,
                                             const FftTables &tab, const PassDesc &desc,
                                             const Epi *epi = nullptr)
{
  uint32_t i = 0;
  while (i < desc.nops) {
    const PassOp op = desc.ops[i];
    BCHFFT_MARK(100u + op.type * 100u + op.d);
    if (op.type == OP_TW) {
      tile_twist<M, UNI>(S, T, g, tab.tw + tab.tw_off[op.d], op.d);
      i += 1;
    } else if (op.type == OP_TY) {
      const uint32_t c = taylor_run(desc, g, i);
      const uint32_t g0 = g.local_bit(op.lo) + 1u - c;   // local bit of the last level's lo
      tile_taylor_dispatch(S, T, t, g0, c, (uint32_t)M);
      i += c;
    } else if (WITH_TX && op.type == OP_TX) {
      if constexpr (WITH_TX) i += tile_tx_sweep(S, T, t, g, desc, i, (uint32_t)M);
    } else {
      bool pair = false;
      if (BCHFFT_BF_R4 && M <= 16 && i + 1 < desc.nops) {
        const PassOp nx = desc.ops[i + 1];
        pair = nx.type == OP_BF && nx.d + 1u == op.d && g.local_bit(nx.d) + 1u == g.local_bit(op.d);
      }
      // the epilogue, if any, replaces the write-back of the pass's final sweep
      if (pair) {
        tile_butterfly_r4<M, Epi, UNI>(S, T, t, g, tab.gam + tab.gam_off[op.d], tab.gam + tab.gam_off[op.d - 1u],
                                  op.d, g.local_bit(op.d) - 1u, (i + 2 == desc.nops) ? epi : nullptr);
        i += 2;
      } else {
        tile_butterfly<M, Epi, UNI>(S, T, t, g, tab.gam + tab.gam_off[op.d], op.d, g.local_bit(op.d),
                               (i + 1 == desc.nops) ? epi : nullptr);
        i += 1;
      }
    }
    __syncthreads();
  }
}

// ---- one pass over every tile of every slab -----------------------------------------------
// grid = (tiles per slab, slabs); dynamic smem = M * 2^t * 4 bytes.
template <int M>
__global__ void __launch_bounds__(BCHFFT_NT_PASS, BCHFFT_MINB_PASS)
k_gm_pass(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, size_t src_slab_words,
          size_t dst_slab_words, FftTables tab, PassDesc desc)
{
  BCHFFT_DYN_SMEM(uint32_t, S);
  const uint32_t T = 1u << desc.t;
  const TileGeom g = tile_geom(desc, blockIdx.x);
  BCHFFT_MARK(1);
  tile_load<M>(S, T, src + (size_t)blockIdx.y * src_slab_words, g, desc.src_mask);
  __syncthreads();
  tile_run_ops<M, NoEpilogue, true, (BCHFFT_UNI_BF != 0 && M >= 13)>(S, T, desc.t, g, tab, desc);
  BCHFFT_MARK(2);
  tile_store<M>(S, T, dst + (size_t)blockIdx.y * dst_slab_words, g);
  BCHFFT_MARK(3);
}

// ---- Taylor-only pass on a group of planes --------------------------------------------------
// A pass that contains only Taylor levels (XOR only, plane-wise) does not need whole elements:
// each CTA takes MP = 8 (or 4, when the plane stride is not a multiple of 8: GF(2^17) .. GF(2^20),
// GF(2^25)) of the PS plane words of every position of the tile, so its tile is a fraction of the
// size and three or more CTAs are resident per SM -- the load / sweep / store phases of neighbouring
// CTAs overlap.  grid = (tiles * PS/MP, slabs): the plane groups of one tile are adjacent CTAs, so they
// run together and share the sectors / 64-byte DRAM bursts.
template <int PSW, int MP>
__global__ void __launch_bounds__(256, BCHFFT_MINB_TY)
k_ty_pass(uint32_t *__restrict__ buf, size_t slab_words, PassDesc desc)
{
  static_assert(MP == 4 || MP == 8, "plane groups of one or two 16-byte vectors");
  static_assert(PSW % MP == 0, "plane stride must split into whole groups");
  constexpr int NG = PSW / MP, NV = MP / 4;
  BCHFFT_DYN_SMEM(uint32_t, S);
  const uint32_t T = 1u << desc.t;
  const TileGeom g = tile_geom(desc, blockIdx.x / NG);
  uint32_t *base = buf + (size_t)blockIdx.y * slab_words + (blockIdx.x % NG) * MP;
  constexpr int U = 4;
  for (uint32_t b0 = threadIdx.x; b0 < T; b0 += U * blockDim.x) {
    uint4 q[U][NV];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t l = b0 + u * blockDim.x;
      if (l < T) {
        const uint4 *in = reinterpret_cast<const uint4 *>(base + (size_t)g.pos(l) * PSW);
#pragma unroll
        for (int v = 0; v < NV; ++v) q[u][v] = in[v];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t l = b0 + u * blockDim.x;
      if (l < T) {
        const uint32_t sl = swz(l);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          S[(4 * v + 0) * T + sl] = q[u][v].x;
          S[(4 * v + 1) * T + sl] = q[u][v].y;
          S[(4 * v + 2) * T + sl] = q[u][v].z;
          S[(4 * v + 3) * T + sl] = q[u][v].w;
        }
      }
    }
  }
  __syncthreads();
  uint32_t i = 0;
  while (i < desc.nops) {
    const uint32_t c = taylor_run(desc, g, i, BCHFFT_TY_LEVELS);
    const uint32_t g0 = g.local_bit(desc.ops[i].lo) + 1u - c;
    tile_taylor_dispatch(S, T, desc.t, g0, c, (uint32_t)MP);
    i += c;
    __syncthreads();
  }
  for (uint32_t l = threadIdx.x; l < T; l += blockDim.x) {
    const uint32_t sl = swz(l);
    uint4 *out = reinterpret_cast<uint4 *>(base + (size_t)g.pos(l) * PSW);
#pragma unroll
    for (int v = 0; v < NV; ++v)
      out[v] = make_uint4(S[(4 * v + 0) * T + sl], S[(4 * v + 1) * T + sl], S[(4 * v + 2) * T + sl],
                          S[(4 * v + 3) * T + sl]);
  }
}

// ---- plane-major path (fully twist-free fields, e.g. GF(2^16); see gm_plan_pm) ----------------
// Slab layout [slab][plane][position].  k_ty_plane_pass: one plane of one slab per CTA, tile of
// 2^t1 positions, XOR-only levels (OP_TX / OP_TY).  grid = (tiles, M, slabs), dynamic smem = 4 * 2^t1 bytes.
template <int PLANES_PER_CTA>   // always 1 (a template so the header can be included everywhere)
__global__ void __launch_bounds__(256, BCHFFT_MINB_PLANE)
k_ty_plane_pass(uint32_t *__restrict__ buf, uint32_t nbits, uint32_t ps, PassDesc desc)
{
  BCHFFT_DYN_SMEM(uint32_t, S);
  const uint32_t T = 1u << desc.t;
  const TileGeom g = tile_geom(desc, blockIdx.x);
  uint32_t *row = buf + (((size_t)blockIdx.z * ps + blockIdx.y) << nbits);
  constexpr int U = 8;
  for (uint32_t b0 = threadIdx.x; b0 < T; b0 += U * blockDim.x) {
    uint32_t q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t l = b0 + u * blockDim.x;
      if (l < T) q[u] = row[g.pos(l)];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t l = b0 + u * blockDim.x;
      if (l < T) S[swz(l)] = q[u];
    }
  }
  __syncthreads();
  uint32_t i = 0;
  while (i < desc.nops) {
    if (desc.ops[i].type == OP_TX) {
      i += tile_tx_sweep(S, T, desc.t, g, desc, i, 1u);
    } else {
      const uint32_t c = taylor_run(desc, g, i, 5u);
      const uint32_t g0 = g.local_bit(desc.ops[i].lo) + 1u - c;
      tile_taylor_dispatch(S, T, desc.t, g0, c, 1u);
      i += c;
    }
    __syncthreads();
  }
  for (uint32_t l = threadIdx.x; l < T; l += blockDim.x) row[g.pos(l)] = S[swz(l)];
}

// whole-element tile <-> plane-major slab (lanes run along positions: 128-byte row segments when
// the window contains the low position bits)
template <int M>
__device__ __forceinline__ void tile_load_pm(uint32_t *S, uint32_t t, const uint32_t *__restrict__ src,
                                             uint32_t nbits, const TileGeom &g)
{
  const uint32_t T = 1u << t;
  constexpr int U = 8;
  for (uint32_t b0 = threadIdx.x; b0 < M * T; b0 += U * blockDim.x) {
    uint32_t q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t idx = b0 + u * blockDim.x;
      if (idx < M * T) q[u] = src[((size_t)(idx >> t) << nbits) + g.pos(idx & (T - 1u))];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t idx = b0 + u * blockDim.x;
      if (idx < M * T) S[(idx >> t) * T + swz(idx & (T - 1u))] = q[u];
    }
  }
}

template <int M>
__device__ __forceinline__ void tile_store_pm(const uint32_t *S, uint32_t t, uint32_t *__restrict__ dst,
                                              uint32_t nbits, const TileGeom &g)
{
  const uint32_t T = 1u << t;
  for (uint32_t idx = threadIdx.x; idx < M * T; idx += blockDim.x)
    dst[((size_t)(idx >> t) << nbits) + g.pos(idx & (T - 1u))] = S[(idx >> t) * T + swz(idx & (T - 1u))];
}

// butterfly pass reading the plane-major layout; writes it back in place, or (dst_em = 1) writes
// the element-major layout k_gather_out reads.  grid = (tiles, slabs).
template <int M>
__global__ void __launch_bounds__(BCHFFT_NT_PASS, BCHFFT_MINB_PASS)
k_gm_pass_pm(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, size_t slab_words,
             uint32_t nbits, uint32_t dst_em, FftTables tab, PassDesc desc)
{
  BCHFFT_DYN_SMEM(uint32_t, S);
  const uint32_t T = 1u << desc.t;
  const TileGeom g = tile_geom(desc, blockIdx.x);
  tile_load_pm<M>(S, desc.t, src + (size_t)blockIdx.y * slab_words, nbits, g);
  __syncthreads();
  tile_run_ops<M, NoEpilogue, true, (BCHFFT_UNI_BF != 0)>(S, T, desc.t, g, tab, desc);
  if (dst_em) tile_store<M>(S, T, dst + (size_t)blockIdx.y * slab_words, g);
  else tile_store_pm<M>(S, desc.t, dst + (size_t)blockIdx.y * slab_words, nbits, g);
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
