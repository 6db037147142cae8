// Host-side planning for the batched additive FFT: Gao-Mateer constants and the schedule of
// tile passes.  Pure C++ (no CUDA), so it is unit-testable on the CPU.
//
// The transform evaluated is the one of the reference's libfft/additive_fft.c:23-178
// (von zur Gathen-Gerhard, evaluation points = all field elements, natural order with basis
// 1, 2, 4, ...).  All arithmetic is exact, so any correct algorithm is bit-identical; we use
// the Gao-Mateer recursion (Taylor expansion at x^2+x) because every multiplier is a
// data-independent constant.  In-place iterative form (validated in tools/gm_model.py):
//
//   forward, depth d = 0..K-1 (domain: positions < 2^K, K = ceil(log2(#coefficients))):
//       TW(d)     a[p] *= tw[d][p >> d]
//       TY(d,lo)  lo = K-2 .. d:  on position bits (lo+1, lo):  Q2 ^= Q3 ; Q1 ^= Q2
//   broadcast the 2^K values over the 2^(m-K) cosets (positions p + c*2^K)
//   backward, depth d = K-1..0 (domain: all 2^m positions):
//       BF(d)     lo' = lo + gam[d][p >> (d+1)] * hi ; hi' = lo' + hi     (pair = bit d of p)
//   result: position p holds P(bitrev_m(p)).
#pragma once
#include <stdint.h>
#include <string.h>
#include <vector>

namespace bchfft {

// OP_TX(lo, h) (h stored in PassOp::d): one level of the generalised Taylor expansion at
// y^(2^h) + y (Gao-Mateer, base t = 2^h): blocks B_0 .. B_(2t-1) of 2^lo positions inside every
// super-block of 2^(lo+h+1) positions; B_t ^= B_(2t-1), then B_i ^= B_(i+t-1) for i = 1 .. t-1.
// OP_TX(lo, 1) is OP_TY on bits (lo+1, lo).
enum : uint8_t { OP_TW = 0, OP_TY = 1, OP_BF = 2, OP_TX = 3 };

struct PassOp {
  uint8_t type, d, lo, pad;
};

constexpr int kMaxOps = 120;

// One pass = one kernel launch; every CTA loads a tile of 2^t positions (window = up to two
// contiguous ranges of position bits), runs `nops` ops on it in shared memory, stores it.
struct PassDesc {
  uint32_t nops;
  uint32_t t;              // tile bits = n0 + n1
  uint32_t dom_bits;       // positions in the pass domain = 2^dom_bits
  uint32_t a0, n0, a1, n1; // window ranges [a0, a0+n0) and [a1, a1+n1); local bit i<n0 -> a0+i
  uint32_t src_mask;       // source position = p & src_mask (broadcast load when < 2^dom_bits-1)
  uint32_t backward;       // 0: forward-domain pass, 1: backward-domain pass
  PassOp ops[kMaxOps];
};

struct GmConstants {
  int m = 0;
  std::vector<uint32_t> tw;       // concatenated tw[d][0 .. 2^(m-d))
  std::vector<uint32_t> gam;      // concatenated gam[d][0 .. 2^(m-d-1))
  std::vector<uint32_t> tw_off;   // m entries
  std::vector<uint32_t> gam_off;  // m entries
  std::vector<uint32_t> basis;    // beta_1 .. beta_m: evaluation point J is sum_i J_i beta_(i+1)
  std::vector<uint32_t> pos_of_elem;  // field element k -> tile position holding P(k)
  std::vector<uint32_t> elem_of_pos;  // inverse map
  uint32_t twist_free = 0;        // bit d set: tau_d = 1, TW(d) is the identity
};

inline uint32_t bitrev_bits(uint32_t x, int bits)
{
  uint32_t r = 0;
  for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
}

// Solve y^2 + y = c over GF(2^m) (a GF(2)-linear system); returns false when Tr(c) = 1.
template <class Mul>
inline bool solve_artin_schreier(int m, Mul mul, uint32_t c, uint32_t *out)
{
  // echelon basis of the image of S(y) = y^2 + y, each vector with the y that produces it
  std::vector<uint32_t> img(m, 0u), pre(m, 0u);
  for (int b = 0; b < m; b++) {
    uint32_t y = 1u << b, v = mul(y, y) ^ y;
    for (int k = m - 1; k >= 0 && v; k--) {
      if (!((v >> k) & 1u)) continue;
      if (img[k]) {
        v ^= img[k];
        y ^= pre[k];
      } else {
        img[k] = v;
        pre[k] = y;
        v = 0;
      }
    }
  }
  uint32_t y = 0;
  for (int k = m - 1; k >= 0 && c; k--) {
    if (!((c >> k) & 1u)) continue;
    if (!img[k]) return false;
    c ^= img[k];
    y ^= pre[k];
  }
  *out = y;
  return c == 0;
}

// exp/log: the caller's TfiniteField tables (libbase/finite_field.c:19-48 layout:
// exp[i] = x^i, log[0] = order).
//
// Choice of the evaluation basis.  The transform evaluates at every field element, and the
// output is gathered through a permutation table anyway (exp order for po_result, natural order
// for the clobbered p_poly), so the order in which the kernels enumerate the points is free.
// We use a basis whose tail is a Cantor chain 1 = c_0, c_1, .., with c_k^2 + c_k = c_(k-1): every
// depth whose last basis element is 1 needs no twist at all (tw[d] == 1).  The chain has length
// 2^a for m = 2^a * odd: all 16 depths for GF(2^16) (multiplications drop from 1.5 n m to
// 0.5 n m), 4 for GF(2^20), 1 for odd m.  The remaining basis elements are powers of x.
inline GmConstants gm_constants(int m, const uint32_t *exp, const uint32_t *log)
{
  const uint32_t order = (1u << m) - 1u, n = 1u << m;
  auto mul = [&](uint32_t a, uint32_t b) -> uint32_t {
    if (!a || !b) return 0;
    uint32_t s = log[a] + log[b];
    if (s >= order) s -= order;
    return exp[s];
  };
  GmConstants c;
  c.m = m;
  std::vector<uint32_t> chain(1, 1u);
  while ((int)chain.size() < m) {
    uint32_t y;
    if (!solve_artin_schreier(m, mul, chain.back(), &y)) break;
    chain.push_back(y);
  }
  // basis = (fill .., chain reversed): beta_m = 1, beta_(m-1) = c_1, ...
  std::vector<uint32_t> basis(m, 0u);
  std::vector<uint32_t> ech(m, 0u);   // echelon form for the independence test
  auto independent_insert = [&](uint32_t v) -> bool {
    for (int k = m - 1; k >= 0 && v; k--) {
      if (!((v >> k) & 1u)) continue;
      if (ech[k]) v ^= ech[k];
      else { ech[k] = v; return true; }
    }
    return false;
  };
  const int cl = (int)chain.size();
  bool periodic = false;
  if (cl >= 2 && cl < m && m % cl == 0 && !(cl & (cl - 1))) {
    // Periodic Cantor basis (m = cl * r, the chain lives in the subfield GF(2^cl)): after the cl
    // twist-free depths of a group the remaining basis is the image of the unused elements under the
    // GF(2^cl)-linear map x -> S(x / tau), S(y) = y^(2^cl) + y = s_cl(y).  Its image is a GF(2^cl)-space, so it
    // contains tau' * (chain) for any nonzero tau' in it: choosing the next cl basis elements as preimages of
    // tau' * chain makes the next group start with ONE twist (by tau') followed by cl - 1 twist-free depths.
    // Twists drop from m - cl to m / cl - 1 (GF(2^20): 16 -> 4), and the Taylor levels of every group of
    // cl depths merge into a conversion of base x^(2^cl) + x (gm_plan).
    std::vector<uint32_t> taus;
    auto inv = [&](uint32_t v) -> uint32_t { return exp[(order - log[v]) % order]; };
    auto image = [&](uint32_t x) -> uint32_t {
      for (uint32_t tau : taus) {
        const uint32_t y = mul(x, inv(tau));
        uint32_t z = y;
        for (int i = 0; i < cl; i++) z = mul(z, z);
        x = z ^ y;
      }
      return x;
    };
    periodic = true;
    for (int j = 0; j * cl < m && periodic; j++) {
      const int k = m - j * cl;
      std::vector<uint32_t> img(m, 0u), pre(m, 0u);   // echelon over the images of the unit vectors
      for (int b = 0; b < m; b++) {
        uint32_t v = image(1u << b), pv = 1u << b;
        for (int kk = m - 1; kk >= 0 && v; kk--) {
          if (!((v >> kk) & 1u)) continue;
          if (img[kk]) {
            v ^= img[kk];
            pv ^= pre[kk];
          } else {
            img[kk] = v;
            pre[kk] = pv;
            v = 0;
          }
        }
      }
      uint32_t tau = 1u;
      if (j > 0) {
        tau = 0u;
        for (int kk = 0; kk < m && !tau; kk++) tau = img[kk];
      }
      if (!tau) {
        periodic = false;
        break;
      }
      for (int jj = 0; jj < cl; jj++) {
        uint32_t v = mul(tau, chain[jj]), pv = 0u;
        for (int kk = m - 1; kk >= 0 && v; kk--) {
          if (!((v >> kk) & 1u)) continue;
          if (!img[kk]) break;
          v ^= img[kk];
          pv ^= pre[kk];
        }
        if (v || !independent_insert(pv)) {
          periodic = false;
          break;
        }
        basis[k - 1 - jj] = pv;
      }
      taus.push_back(tau);
    }
    if (!periodic) {
      std::fill(basis.begin(), basis.end(), 0u);
      std::fill(ech.begin(), ech.end(), 0u);
    }
  }
  if (!periodic) {
    for (size_t k = 0; k < chain.size(); k++) {
      independent_insert(chain[k]);
      basis[m - 1 - k] = chain[k];
    }
    int slot = 0;
    for (int b = 0; b < m && slot < m - (int)chain.size(); b++)
      if (independent_insert(1u << b)) basis[slot++] = 1u << b;
  }
  c.basis = basis;
  // point tables: natural index J over the basis <-> tile position bitrev(J)
  c.pos_of_elem.assign(n, 0u);
  c.elem_of_pos.assign(n, 0u);
  {
    std::vector<uint32_t> elem(n, 0u);
    for (uint32_t J = 1; J < n; J++) {
      const int low = __builtin_ctz(J);
      elem[J] = elem[J & (J - 1u)] ^ basis[low];
    }
    for (uint32_t J = 0; J < n; J++) {
      const uint32_t p = bitrev_bits(J, m);
      c.elem_of_pos[p] = elem[J];
      c.pos_of_elem[elem[J]] = p;
    }
  }
  for (int d = 0; d < m; d++) {
    const int k = m - d;
    const uint32_t tau = basis[k - 1];
    if (tau == 1u) c.twist_free |= 1u << d;
    c.tw_off.push_back((uint32_t)c.tw.size());
    uint32_t pw = 1;
    for (uint32_t i = 0; i < (1u << k); i++) {
      c.tw.push_back(pw);
      pw = mul(pw, tau);
    }
    const uint32_t itau = exp[(order - log[tau]) % order];
    std::vector<uint32_t> g(k - 1);
    for (int i = 0; i < k - 1; i++) g[i] = mul(basis[i], itau);
    c.gam_off.push_back((uint32_t)c.gam.size());
    for (uint32_t q = 0; q < (1u << (k - 1)); q++) {
      const uint32_t j = bitrev_bits(q, k - 1);
      uint32_t v = 0;
      for (int b = 0; b < k - 1; b++)
        if ((j >> b) & 1u) v ^= g[b];
      c.gam.push_back(v);
    }
    std::vector<uint32_t> next(k - 1);
    for (int i = 0; i < k - 1; i++) next[i] = mul(g[i], g[i]) ^ g[i];
    basis = next;
  }
  return c;
}

// ---- pass planner -----------------------------------------------------------------------
// number of maximal runs of set bits
inline int bit_runs(uint32_t x)
{
  int r = 0;
  while (x) {
    x >>= __builtin_ctz(x);   // drop trailing zeros
    r++;
    while (x & 1u) x >>= 1;   // drop the run
  }
  return r;
}

// Grow `w` (a non-empty set of position bits within [0,dom), at most two runs) to
// min(t, dom) bits keeping at most two runs.  Low bits are preferred so that the rows a tile
// reads from HBM are contiguous: first extend the lowest run down to bit 0, then upward.
inline uint32_t pad_window(uint32_t w, int t, int dom)
{
  const int want = t < dom ? t : dom;
  while (__builtin_popcount(w) < want) {
    const int lo0 = __builtin_ctz(w);
    if (lo0 > 0) {
      w |= 1u << (lo0 - 1);
    } else {
      int n0 = 0;
      while ((w >> n0) & 1u) n0++;
      w |= 1u << n0;           // free by construction; merges with the upper run when adjacent
    }
  }
  return w;
}

inline void set_window(PassDesc &p, uint32_t w)
{
  // w has one or two runs
  const int lo0 = __builtin_ctz(w);
  int n0 = 0;
  while ((w >> (lo0 + n0)) & 1u) n0++;
  p.a0 = lo0;
  p.n0 = n0;
  const uint32_t rest = w & ~(((1u << n0) - 1u) << lo0);
  if (rest) {
    const int lo1 = __builtin_ctz(rest);
    int n1 = 0;
    while ((rest >> (lo1 + n1)) & 1u) n1++;
    p.a1 = lo1;
    p.n1 = n1;
  } else {
    p.a1 = lo0 + n0;
    p.n1 = 0;
  }
  p.t = p.n0 + p.n1;
}

// Position bits every window of the generic schedule must contain.  Default (0xFFFFFFFF = automatic): bit 0
// when an element ((m + 3) & ~3 plane words) is not a whole number of 32-byte sectors -- GF(2^17)..GF(2^20):
// 80 bytes, GF(2^25): 112 -- so that a tile row moves pairs of elements, i.e. whole sectors.  Measured at
// m = 20, ms per 256 polynomials: no required bit 11.5, bit 0 9.2, bits 0 and 1 (one pass more) 10.5.
#ifndef BCHFFT_GM_REQUIRED_BITS
#define BCHFFT_GM_REQUIRED_BITS 0xFFFFFFFFu
#endif
inline uint32_t gm_required_bits(int m)
{
  const uint32_t r = BCHFFT_GM_REQUIRED_BITS;
  if (r != 0xFFFFFFFFu) return r;
  return ((((m + 3) & ~3) * 4) % 32) ? 1u : 0u;
}

struct GmPlan {
  int m, K, t;
  std::vector<PassDesc> passes;  // forward passes first, then backward passes
  int first_backward;
};

struct PlanItem {
  PassOp op;
  uint32_t bits;
};

// Greedy packing of an op stream into passes: a pass takes ops while the union of their position
// bits (plus `required`, bits every window must contain) still fits t bits in at most two runs.
inline void build_passes(std::vector<PassDesc> &out, const std::vector<PlanItem> &stream, int dom,
                         int t_max, bool backward, uint32_t src_mask_first, uint32_t required = 0u)
{
  size_t i = 0;
  bool first = true;
  const int t = t_max < dom ? t_max : dom;
  do {
    PassDesc p;
    memset(&p, 0, sizeof p);
    uint32_t w = required;
    // (an op that does not fit next to the required bits gets a window without them)
    if (i < stream.size() && required) {
      const uint32_t cand = w | stream[i].bits;
      if (__builtin_popcount(cand) > t || bit_runs(cand) > 2) w = 0u;
    }
    while (i < stream.size() && p.nops < (uint32_t)kMaxOps) {
      const uint32_t cand = w | stream[i].bits;
      if (__builtin_popcount(cand) > t || bit_runs(cand) > 2) break;
      w = cand;
      p.ops[p.nops++] = stream[i].op;
      i++;
    }
    if (dom > 0) {
      if (w == 0) w = 1u;
      w = pad_window(w, t, dom);
      set_window(p, w);
    }
    p.dom_bits = dom;
    p.backward = backward ? 1u : 0u;
    p.src_mask = (first && backward) ? src_mask_first : 0xFFFFFFFFu;
    first = false;
    out.push_back(p);
  } while (i < stream.size());
}

inline void cantor_conversion_ops(std::vector<PlanItem> &out, int b0, int w);

// With `required` bits (see gm_required_bits) a pass is about a fifth cheaper (whole sectors per tile row)
// but the windows have fewer free bits: take them only when that costs at most one pass in eight.
inline void build_passes_sector_aware(std::vector<PassDesc> &out, const std::vector<PlanItem> &stream, int dom,
                                      int t_max, bool backward, uint32_t src_mask_first, uint32_t required)
{
  std::vector<PassDesc> plain, aligned;
  build_passes(plain, stream, dom, t_max, backward, src_mask_first, 0u);
  if (required && dom > t_max && t_max >= 8) {
    build_passes(aligned, stream, dom, t_max, backward, src_mask_first, required);
    if (aligned.size() * 8 <= plain.size() * 9) {
      out.insert(out.end(), aligned.begin(), aligned.end());
      return;
    }
  }
  out.insert(out.end(), plain.begin(), plain.end());
}

// t_max: largest tile (log2 positions) that fits the CTA's shared memory for this M.
// cantor_prefix: h consecutive depths d .. d + h - 1 without a twist between them (h a power of two; the
// periodic Cantor basis of gm_constants gives runs of the chain length: 4 for GF(2^20), 8 for GF(2^24), 2
// for GF(2^18)) have Taylor levels that are, as a formal identity on coefficient arrays, the expansion in
// base x^(2^h) + x on the index bits d .. K-1 followed by the Cantor conversion of bits d .. d+h-1:
// K - d - h levels OP_TX(k, h) and (h/2) log2(h) more instead of h (2 (K - d) - h - 1) / 2
// (tools/gm_model.py: tx_level / cantor_conversion; checked against the Taylor levels for K up to 14 and every
// offset d).  GF(2^20): 190 Taylor levels -> 60 generalised ones, 16 twists -> 4.
inline GmPlan gm_plan(int m, int K, int t_max, uint32_t twist_free = 0u, bool cantor_prefix = false)
{
  GmPlan plan;
  plan.m = m;
  plan.K = K;
  plan.t = t_max;
  std::vector<PlanItem> fwd, bwd;
  const int tcap = t_max < K ? t_max : K;
  for (int d = 0; d < K;) {
    if (!((twist_free >> d) & 1u)) fwd.push_back({{OP_TW, (uint8_t)d, 0, 0}, 0u});
    // depths d .. d + run - 1 have no twist between their Taylor levels
    int run = 1;
    while (d + run < K && ((twist_free >> (d + run)) & 1u)) run++;
    int h = 0;
    if (cantor_prefix) {
      for (h = 1; 2 * h <= run; h *= 2) {}
      while (h >= 2 && h + 1 > tcap) h /= 2;       // a level of base 2^h spans h + 1 position bits of one window
      if (h < 2 || K - d < h) h = 0;
    }
    if (h) {
      // Taylor levels of h consecutive depths = expansion in base x^(2^h) + x on the index bits d .. K-1
      // followed by the Cantor conversion of the h bits d .. d + h - 1
      for (int k = K - h - 1; k >= d; k--)
        fwd.push_back({{OP_TX, (uint8_t)h, (uint8_t)k, 0}, ((2u << h) - 1u) << k});
      cantor_conversion_ops(fwd, d, h);
      d += h;
      // the rest of the run continues without a twist: handled by the next iteration, which must not emit TW
      continue;
    }
    for (int lo = K - 2; lo >= d; lo--) fwd.push_back({{OP_TY, (uint8_t)d, (uint8_t)lo, 0}, 3u << lo});
    d++;
  }
  for (int d = K - 1; d >= 0; d--) bwd.push_back({{OP_BF, (uint8_t)d, 0, 0}, 1u << d});
  if (K == m) {
    // full-degree input: forward and backward phases work in place on the same buffer, so they
    // form one op stream (the last forward ops and the first butterflies usually share a pass;
    // a transform that fits one tile is a single pass)
    std::vector<PlanItem> all = fwd;
    all.insert(all.end(), bwd.begin(), bwd.end());
    plan.first_backward = 0;
    build_passes_sector_aware(plan.passes, all, m, t_max, true, 0xFFFFFFFFu, gm_required_bits(m));
    return plan;
  }
  if (!fwd.empty()) build_passes(plan.passes, fwd, K, t_max, false, 0xFFFFFFFFu);   // K = 0: constant polynomial
  plan.first_backward = (int)plan.passes.size();
  build_passes_sector_aware(plan.passes, bwd, m, t_max, true, K < m ? ((1u << K) - 1u) : 0xFFFFFFFFu,
                            gm_required_bits(m));
  return plan;
}

// Plane-major schedule for fully twist-free fields (m a power of two, e.g. GF(2^16)), full-degree
// input.
//
// Forward phase.  With every twist equal to one the forward phase is the change of basis from
// monomials to the products prod_d s_d(x)^(p_d) of the Cantor subspace polynomials
// s_0 = x, s_(d+1) = s_d^2 + s_d.  Because s_(2^k)(x) = x^(2^(2^k)) + x, that conversion splits
// recursively (Lin, Al-Naffouri, Han, Chung 2016, basis conversion for Cantor bases): for a
// window of w = 2h index bits starting at bit b0, expand in base y^(2^h) + y (h levels OP_TX(b0+k, h),
// k = h-1 .. 0), then convert the low and the high h bits independently.  m/2 * log2(m) levels
// instead of the m(m-1)/2 Taylor levels of the plain recursion: 32 instead of 120 at m = 16
// (tools/gm_model.py checks that both give the same coefficients).
//
// The phase is XOR-only and plane-wise, so it runs on a plane-major slab layout
// [slab][plane][position] with one-plane tiles of 2^t1 positions.  Every window contains position
// bits 0..2 (32-byte runs per plane row).  The butterfly passes read the same layout with
// whole-element tiles of 2^t2 positions; a trailing forward pass whose ops fit the first butterfly
// window is folded into it, and the last butterfly pass writes the element-major layout the output
// gather reads.
//
// Leading levels whose super-block (2^(lo+h+1) positions) is larger than a plane tile run as
// "line passes" straight on the slab in HBM (k_tx_line_pass: up to six consecutive levels of one
// base per pass, no tile).
struct GmPlanPM {
  std::vector<PassDesc> line_passes, plane_passes, bf_passes;
};

inline void cantor_conversion_ops(std::vector<PlanItem> &out, int b0, int w)
{
  if (w < 2) return;
  const int h = w / 2;
  for (int k = h - 1; k >= 0; k--)
    out.push_back({{OP_TX, (uint8_t)h, (uint8_t)(b0 + k), 0}, ((2u << h) - 1u) << (b0 + k)});
  cantor_conversion_ops(out, b0, h);
  cantor_conversion_ops(out, b0 + h, h);
}

inline bool gm_plan_pm(int m, int t1, int t2, uint32_t twist_free, GmPlanPM &out)
{
  if (m <= t2 || m > 24 || (m & (m - 1)) || t2 < 6 || t1 < 4 ||
      (twist_free & ((1u << m) - 1u)) != ((1u << m) - 1u))
    return false;
  std::vector<PlanItem> fwd, bwd;
  cantor_conversion_ops(fwd, 0, m);
  {
    const int tile = t1 < m ? t1 : m;
    size_t i = 0;
    while (i < fwd.size() && fwd[i].op.lo + fwd[i].op.d + 1 > tile) {
      PassDesc p;
      memset(&p, 0, sizeof p);
      p.dom_bits = p.t = m;
      p.n0 = m;
      p.a1 = m;
      p.backward = 1u;
      p.src_mask = 0xFFFFFFFFu;
      p.ops[p.nops++] = fwd[i].op;
      i++;
      while (i < fwd.size() && p.nops < 6u && fwd[i].op.d == p.ops[0].d &&
             fwd[i].op.lo + p.nops == p.ops[0].lo && fwd[i].op.lo + fwd[i].op.d + 1 > tile)
        p.ops[p.nops++] = fwd[i++].op;
      out.line_passes.push_back(p);
    }
    fwd.erase(fwd.begin(), fwd.begin() + i);
  }
  // every remaining level must fit a window next to the required low bits
  for (const PlanItem &it : fwd)
    if (__builtin_popcount(it.bits | 7u) > (t1 < m ? t1 : m) || bit_runs(it.bits | 7u) > 2) return false;
  if (!fwd.empty()) build_passes(out.plane_passes, fwd, m, t1, true, 0xFFFFFFFFu, 7u);
  // fold the last forward pass into the first butterfly pass when its bits fit that window
  if (!out.plane_passes.empty()) {
    const PassDesc &last = out.plane_passes.back();
    uint32_t bits = 7u | (1u << (m - 1));
    for (uint32_t i = 0; i < last.nops; i++) bits |= ((2u << last.ops[i].d) - 1u) << last.ops[i].lo;
    if (__builtin_popcount(bits) <= t2 && bit_runs(bits) <= 2) {
      for (uint32_t i = 0; i < last.nops; i++)
        bwd.push_back({last.ops[i], ((2u << last.ops[i].d) - 1u) << last.ops[i].lo});
      out.plane_passes.pop_back();
    }
  }
  for (int d = m - 1; d >= 0; d--) bwd.push_back({{OP_BF, (uint8_t)d, 0, 0}, 1u << d});
  build_passes(out.bf_passes, bwd, m, t2, true, 0xFFFFFFFFu, 7u);
  return true;
}

}  // namespace bchfft
