// Batched BCH decode driver (see include/bchfft_cuda.h, bchfft_code_* / bchfft_bch_*).
#include <stdlib.h>
#include <type_traits>
#include <string.h>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "cabi_common.h"
#include "bch_kernels.cuh"

struct bchfft_code {
  bchfft_field *f = nullptr;
  uint32_t d = 0, t = 0, length = 0, gen_degree = 0;
  uint32_t wpc = 0;          // 64-bit words per codeword
  uint32_t nodd = 0;         // odd syndromes computed directly: S_1, S_3, .., S_{2 nodd - 1}
  uint32_t syn_fft_rule = 0; // reference would take its FFT syndrome path (bch.c:240-254)
  uint32_t chunk_pos = 0, nchunks = 0, nsub_total = 0;
  uint32_t *d_taps = nullptr, *d_fold = nullptr, *d_vbits = nullptr;
  // universal-LFSR syndrome kernel (k_syndrome_p): usable when every odd j < d is coprime to the
  // field order; per class j^-1 mod order, per (class, chunk) the first permuted position, per
  // chunk the shift constants x^(q lc + b)
  uint32_t syn_p = 0;
  // systematic encoder (k_encode): T[v] = v(X) X^deg(g) mod g, 256 rows of enc_nw 64-bit words;
  // deg(g) > 2048 (enc_wide): rows X^(deg(g)+k) mod g, k < 64, of deg(g)/64+1 words (k_encode_wide)
  uint32_t enc_nw = 0, enc_wide = 0;
  uint64_t *d_enc_table = nullptr;
  uint32_t *d_jinv = nullptr, *d_start = nullptr, *d_foldp = nullptr, *d_lcs = nullptr;
  // work sets: [0] serves the device-pointer and stage-level entry points, [1] .. [3] are the
  // lanes of the pipelined host entry point (copy-in / decode / copy-out overlap)
  struct WorkSet {
    bchfft::DeviceBuffer scratch, fwd_buf, words, ok, corr;
    bchfft::DeviceBuffer fft_ws, fft_io;   // syndromes through the additive FFT (few long codewords)
    // root-search pass descriptors: a buffer of their own, allocated once, so that a growing
    // fwd_buf (DeviceBuffer::reserve frees and re-allocates) can never leave them stale
    bchfft::DeviceBuffer descs;
    // "positions" result mode of the host entry point: page-locked staging for ok / corrected / position lists
    bchfft::PinnedBuffer h_ok, h_corr, h_pos;
    cudaEvent_t results_ready = nullptr;
    cudaStream_t stream = nullptr;
    int descs_t = -1, descs_kmax = -1;   // (t, kmax) of the descriptors currently uploaded
  } work[4];
};

namespace bchfft {

static int ceil_log2u(uint32_t x)
{
  int k = 0;
  while ((1ull << k) < x) k++;
  return k;
}

// forward / backward descriptors of the truncated FFT used for root search: forward on the
// 2^K coefficients (window [0,K)), backward K butterfly levels on tiles [0,t)
static void locate_descs(int m, int K, int t, uint32_t twist_free, PassDesc &fwd, PassDesc &bwd)
{
  memset(&fwd, 0, sizeof fwd);
  memset(&bwd, 0, sizeof bwd);
  fwd.dom_bits = K;
  fwd.t = K;
  fwd.a0 = 0;
  fwd.n0 = K;
  fwd.a1 = K;
  fwd.n1 = 0;
  fwd.src_mask = 0xFFFFFFFFu;
  for (int d = 0; d < K; d++) {
    if (!((twist_free >> d) & 1u)) fwd.ops[fwd.nops++] = PassOp{OP_TW, (uint8_t)d, 0, 0};
    for (int lo = K - 2; lo >= d; lo--) fwd.ops[fwd.nops++] = PassOp{OP_TY, (uint8_t)d, (uint8_t)lo, 0};
  }
  bwd.dom_bits = m;
  bwd.t = t;
  bwd.a0 = 0;
  bwd.n0 = t;
  bwd.a1 = t;
  bwd.n1 = 0;
  bwd.backward = 1;
  bwd.src_mask = (1u << K) - 1u;
  for (int d = K - 1; d >= 0; d--) bwd.ops[bwd.nops++] = PassOp{OP_BF, (uint8_t)d, 0, 0};
}

struct Scratch {
  uint32_t *acc, *parity, *syn, *bm, *elp, *L, *count, *positions;
  int32_t *flag;
  uint32_t pos_stride;
  size_t bytes_per_cw;
};

static size_t scratch_per_cw(const bchfft_code *c, uint32_t pos_stride, int M)
{
  const size_t d = c->d;
  // acc+parity are per slab: ceil to one slab per 32 codewords
  return sizeof(uint32_t) * (d + 3 * (d + 1) + (d + 1) + 3 + pos_stride) +
         (sizeof(uint32_t) * ((size_t)c->nodd * M + 1) + 31) / 32 + 8;
}

static void carve(const bchfft_code *c, void *base, size_t cnt, uint32_t pos_stride, int M, Scratch &s)
{
  const size_t ns = (cnt + 31) / 32, d = c->d;
  uint32_t *p = (uint32_t *)base;
  s.acc = p;        p += ns * c->nodd * M;
  s.parity = p;     p += ns;
  s.count = p;      p += cnt;
  s.bm = p;         p += 3 * (d + 1) * cnt;   // acc..bm are zeroed with one memset
  s.syn = p;        p += cnt * d;
  s.elp = p;        p += cnt * (d + 1);
  s.L = p;          p += cnt;
  s.flag = (int32_t *)p; p += cnt;
  s.positions = p;  p += cnt * pos_stride;
  s.pos_stride = pos_stride;
}

static size_t scratch_bytes(const bchfft_code *c, size_t cnt, uint32_t pos_stride, int M)
{
  const size_t ns = (cnt + 31) / 32, d = c->d;
  return sizeof(uint32_t) * (ns * c->nodd * M + ns + cnt + 3 * (d + 1) * cnt + cnt * d +
                             cnt * (d + 1) + cnt + cnt + cnt * pos_stride) + 64;
}

// acc, parity, count and (with_bm) the global Berlekamp-Massey buffers; the shared-memory BM kernel
// does not touch the latter
static size_t zero_bytes(const bchfft_code *c, size_t cnt, int M, bool with_bm = true)
{
  const size_t ns = (cnt + 31) / 32, d = c->d;
  return sizeof(uint32_t) * (ns * c->nodd * M + ns + cnt + (with_bm ? 3 * (d + 1) * cnt : 0));
}

#ifdef BCHFFT_EMU
static const int kNtSyn = 32, kNtLoc = 32;
#else
static const int kNtSyn = BCHFFT_NT_SYN, kNtLoc = BCHFFT_NT_LOC;
#endif

// Few codewords, many syndrome classes: one additive FFT per word (the reference's FFT method,
// bch.c:278-292) instead of positions x classes LFSR steps on a slab that is nearly empty.
static bool syn_fft(const bchfft_code *c, size_t cnt)
{
  const char *env = getenv("BCHFFT_SYN_FFT");
  return cnt <= 64 && c->nodd >= 48u && c->f->m >= 8u && !(env && atoi(env) == 0);
}

template <int M>
static int launch_syndrome(bchfft_code *c, bchfft_code::WorkSet &w, const uint64_t *d_words, size_t cnt,
                           const Scratch &s, cudaStream_t st, bool finish = true)
{
  const unsigned ns = (unsigned)((cnt + 31) / 32);
  if (finish && syn_fft(c, cnt)) {
    const uint32_t n = c->f->n;
    int rc = w.fft_io.reserve((size_t)2 * cnt * n * sizeof(uint32_t));
    if (rc) return rc;
    uint32_t *coef = (uint32_t *)w.fft_io.p, *res = coef + cnt * (size_t)n;
    BCHFFT_RUN(k_unpack_bits, dim3((n + 255u) / 256u, (unsigned)cnt), dim3(256), 0, st,
               (const uint32_t *)d_words, c->wpc * 2u, c->length, n, coef);
    const uint32_t terms = c->length < n - 1u ? c->length : n - 1u;
    if ((rc = bchfft_additive_fft_dev_ws(c->f, w.fft_ws, coef, n, res, n, nullptr, 0, terms, cnt, (void *)st)))
      return rc;
    BCHFFT_RUN(k_syn_from_fft, dim3((unsigned)cnt), dim3(256), 0, st, (const uint32_t *)res, n,
               (const uint32_t *)d_words, c->wpc * 2u, (const uint32_t *)c->d_vbits, c->d, s.syn, s.flag);
    return BCHFFT_OK;
  }
  const char *penv = getenv("BCHFFT_SYN_P");
  if (c->syn_p && !(penv && atoi(penv) == 0)) {
    const size_t smem_p = 2u * sizeof(uint32_t) * ((size_t)((c->f->order + 31u) / 32u) * 32u + (size_t)c->nodd * M);
    BCHFFT_ENSURE_SMEM(k_syndrome_p<M>, c->f->device, smem_p);
    BCHFFT_RUN(k_syndrome_p<M>, dim3((ns + 1u) / 2u), dim3(kNtSyn), smem_p, st, (const uint32_t *)d_words, c->wpc * 2u,
               c->length, (uint32_t)cnt, c->nodd, (const uint32_t *)c->d_lcs, (const uint32_t *)c->d_jinv,
               (const uint32_t *)c->d_start, (const uint32_t *)c->d_foldp, (const uint32_t *)c->d_vbits,
               s.acc, s.parity);
    if (finish)
      BCHFFT_RUN(k_syn_finish<M>, dim3((unsigned)((cnt + 127) / 128)), dim3(128), 0, st,
                 (const uint32_t *)s.acc, (const uint32_t *)s.parity, c->nodd, c->d, (uint32_t)cnt, s.syn, s.flag);
    return BCHFFT_OK;
  }
  const size_t smem = sizeof(uint32_t) * ((size_t)c->chunk_pos + c->chunk_pos / 32 + (size_t)c->nodd * M);
  BCHFFT_ENSURE_SMEM(k_syndrome<M>, c->f->device, smem);
  BCHFFT_RUN(k_syndrome<M>, dim3(c->nchunks, ns), dim3(kNtSyn), smem, st,
                (const uint32_t *)d_words, c->wpc * 2u, c->length, (uint32_t)cnt, c->chunk_pos,
                c->nodd, (const uint32_t *)c->d_taps, (const uint32_t *)c->d_fold, c->nsub_total,
                (const uint32_t *)c->d_vbits, s.acc, s.parity);
  if (finish)
    BCHFFT_RUN(k_syn_finish<M>, dim3((unsigned)((cnt + 127) / 128)), dim3(128), 0, st,
                  (const uint32_t *)s.acc, (const uint32_t *)s.parity, c->nodd, c->d, (uint32_t)cnt,
                  s.syn, s.flag);
  return BCHFFT_OK;
}

#ifdef BCHFFT_EMU
static const unsigned kBmThreads = 32;
#else
static const unsigned kBmThreads = 256;
#endif

// Few codewords with a long recurrence (single decode at n = 65535, t = 1000): one CTA per codeword
// (k_bm_wide) instead of one thread.  Break-even is at about 16 t codewords.
static bool bm_wide(const bchfft_code *c, size_t cnt)
{
  const char *wenv = getenv("BCHFFT_BM_WIDE");
  const size_t smem_w = sizeof(uint32_t) * ((size_t)4 * c->d + 3u);
  return c->t >= 24u && cnt <= (size_t)16 * c->t && smem_w <= (size_t)200 * 1024 && !(wenv && atoi(wenv) == 0);
}

// Launch shape of the shared-memory Berlekamp-Massey kernel: lanes per codeword (lpc), threads, stride
// between coefficient rows (cs, 16-bit entries) and dynamic shared memory.
struct BmShape {
  unsigned lpc, nt, cs;
  size_t smem;
};

// One thread per codeword while (4 d + 3) 16-bit entries per thread fit next to the log / exp tables
// (measured at t = 16: lane pairs are slower, 1.12 ms against 1.02 ms); for larger d the smallest lane
// group (4, 8, 16, 32 lanes) that leaves room for three CTAs per SM.  Measured, k_bm per 262144 words at
// t = 64: 4 lanes 3.8 ms, 8 lanes 3.5 ms (the rule's choice), 16 lanes 4.0 ms, global scratch 6.8 ms; per
// 131072 words at t = 200: 8 lanes 28 ms, 16 lanes 17.7 ms, 32 lanes 15.9 ms (the rule's choice), global
// scratch 36 ms.  BCHFFT_BM_LPC forces a group size.
// The stride between coefficient rows is the smallest one for which the rows j, j+1, .. j+lpc-1 that
// one warp touches in one access fall on disjoint banks.  false: the code takes the global-scratch kernel.
static unsigned bm_row_stride(unsigned per_cta, unsigned lpc)
{
  unsigned cs = (per_cta + 1u) & ~1u;
  if (lpc == 1u) return per_cta;
  for (; cs < per_cta + 160u; cs += 2u) {
    uint32_t word_of_bank[32];
    bool used[32] = {false}, clash = false;
    for (unsigned lane = 0; lane < 32u && !clash; lane++) {
      const unsigned sub = lane % lpc, g = lane / lpc;
      const uint32_t word = (sub * cs + g) >> 1, bank = word & 31u;
      if (used[bank] && word_of_bank[bank] != word) clash = true;
      used[bank] = true;
      word_of_bank[bank] = word;
    }
    if (!clash) return cs;
  }
  return (per_cta + 1u) & ~1u;
}

template <int M>
static bool bm_in_smem(const bchfft_code *c, BmShape *out)
{
  const char *benv = getenv("BCHFFT_BM_SM");
  if (M > 14 || (benv && atoi(benv) == 0)) return false;
  const unsigned nt = kBmThreads;
  const char *lenv = getenv("BCHFFT_BM_LPC");
  const unsigned forced = lenv ? (unsigned)atoi(lenv) : 0u;
  // shared memory per CTA the shape may take: a lane group wants three CTAs per SM; when no group size
  // gets there, two CTAs or a single one still beat the global-scratch kernel
  const unsigned options[6] = {1u, 2u, 4u, 8u, 16u, 32u};
  const size_t limits[3] = {(size_t)75 * 1024, (size_t)110 * 1024, (size_t)200 * 1024};
  for (unsigned round = 0; round < 3; round++)
    for (unsigned k = 0; k < 6; k++) {
      const unsigned lpc = options[k];
      if (forced ? lpc != forced : lpc == 2u) continue;
      if (lpc > nt) break;
      const unsigned cs = bm_row_stride(nt / lpc, lpc);
      const size_t smem = sizeof(uint16_t) * (((size_t)2 << M) + (size_t)(4 * c->d + 3) * cs);
      const size_t limit = (lpc == 1u && round == 0) ? limits[1] : limits[round];
      if (smem <= limit) {
        out->lpc = lpc;
        out->nt = nt;
        out->cs = cs;
        out->smem = smem;
        return true;
      }
    }
  return false;
}

template <int M, int LPC>
static int launch_bm_sm(bchfft_code *c, size_t cnt, const Scratch &s, cudaStream_t st, bool fused, uint32_t elp_limit,
                        const BmShape &b, bool binary_word)
{
  const unsigned per_cta = b.nt / LPC;
  BCHFFT_ENSURE_SMEM((k_bm<M, true, LPC>), c->f->device, b.smem);
  BCHFFT_RUN((k_bm<M, true, LPC>), dim3((unsigned)((cnt + per_cta - 1) / per_cta)), dim3(b.nt), b.smem, st,
             (const uint32_t *)s.syn, s.flag, c->d, (uint32_t)cnt, (const uint32_t *)c->f->d_exp,
             (const uint32_t *)c->f->d_log, s.bm, s.elp, s.L,
             fused ? (const uint32_t *)s.acc : (const uint32_t *)nullptr, (const uint32_t *)s.parity, c->nodd,
             elp_limit, (uint32_t)b.cs, binary_word ? 1u : 0u);
  return BCHFFT_OK;
}

// fused: the syndromes are still bitsliced accumulators (launch_syndrome(..., finish = false));
// only valid when bm_in_smem()
template <int M>
// binary_word: the syndromes were computed from binary words by this library (decode pipeline), which lets the
// kernels skip the discrepancies that are zero for power sums in characteristic 2
static int launch_bm(bchfft_code *c, size_t cnt, const Scratch &s, cudaStream_t st, bool fused = false,
                     uint32_t elp_limit = 0xFFFFFFFFu, bool binary_word = false)
{
  if (getenv("BCHFFT_BM_ALL_DISC")) binary_word = false;
  if (!fused && bm_wide(c, cnt)) {
    const size_t smem_w = sizeof(uint32_t) * ((size_t)4 * c->d + 3u);
    unsigned ntw = ((c->t + 1u + 31u) / 32u) * 32u;
    ntw = ntw < 64u ? 64u : (ntw > 1024u ? 1024u : ntw);
    BCHFFT_ENSURE_SMEM(k_bm_wide<M>, c->f->device, smem_w);
    BCHFFT_RUN(k_bm_wide<M>, dim3((unsigned)cnt), dim3(ntw), smem_w, st, (const uint32_t *)s.syn,
               (const int32_t *)s.flag, c->d, (uint32_t)cnt, (const uint32_t *)c->f->d_exp,
               (const uint32_t *)c->f->d_log, s.elp, s.L, elp_limit, binary_word ? 1u : 0u);
    return BCHFFT_OK;
  }
  BmShape b;
  if constexpr (M <= 14) {
    if (bm_in_smem<M>(c, &b)) {
      switch (b.lpc) {
        case 1: return launch_bm_sm<M, 1>(c, cnt, s, st, fused, elp_limit, b, binary_word);
        case 2: return launch_bm_sm<M, 2>(c, cnt, s, st, fused, elp_limit, b, binary_word);
        case 4: return launch_bm_sm<M, 4>(c, cnt, s, st, fused, elp_limit, b, binary_word);
        case 8: return launch_bm_sm<M, 8>(c, cnt, s, st, fused, elp_limit, b, binary_word);
        case 16: return launch_bm_sm<M, 16>(c, cnt, s, st, fused, elp_limit, b, binary_word);
        default: return launch_bm_sm<M, 32>(c, cnt, s, st, fused, elp_limit, b, binary_word);
      }
    }
  }
  if (fused) {
    set_error("bch decode: internal error (fused Berlekamp-Massey without the shared-memory kernel)");
    return BCHFFT_ERR_ARG;
  }
  BCHFFT_RUN((k_bm<M, false, 1>), dim3((unsigned)((cnt + 127) / 128)), dim3(128), 0, st,
             (const uint32_t *)s.syn, s.flag, c->d, (uint32_t)cnt, (const uint32_t *)c->f->d_exp,
             (const uint32_t *)c->f->d_log, s.bm, s.elp, s.L, (const uint32_t *)nullptr,
             (const uint32_t *)nullptr, 0u, elp_limit, 0u, binary_word ? 1u : 0u);
  return BCHFFT_OK;
}

// tcap: largest L that is located (t for bch_decode, the largest requested degree for the
// stage-level entry point)
template <int M>
static int launch_locate(bchfft_code *c, bchfft_code::WorkSet &w, size_t cnt, const Scratch &s, uint32_t tcap,
                         cudaStream_t st)
{
  constexpr int PS = plane_stride(M);
  const int kmax = ceil_log2u(tcap + 1u) < 1 ? 1 : ceil_log2u(tcap + 1u);
  // half of the largest tile that fits an SM, so that two CTAs are resident (BCHFFT_MINB_LOC)
  int t = c->f->t_max - 1;
  if (t < kmax) t = kmax;
  if (t > M) t = M;
  const char *tenv = getenv("BCHFFT_LOCATE_T");
  if (tenv && atoi(tenv) >= kmax && atoi(tenv) <= c->f->t_max && atoi(tenv) <= M) t = atoi(tenv);
  if (kmax > t || kmax > M || kmax > kMaxBuckets || t > c->f->t_max) {
    set_error("bch locate: error-locator degree bound %u too large for one shared-memory tile", tcap);
    return BCHFFT_ERR_UNSUPPORTED;
  }
  const unsigned ns = (unsigned)((cnt + 31) / 32);
  const unsigned nvirt = ns + (unsigned)kmax;           // sum_K ceil(count_K / 32) <= ns + kmax
  // low locator degrees: lanes = slabs, warp-uniform multipliers (k_locate_u)
  const char *uenv = getenv("BCHFFT_LOCATE_U");
  // (buckets K = 6, 7 recompute part of their first step per tile: worth it for whole slab groups only --
  // a single bch_decode at n = 65535, t = 100 takes 0.44 ms through k_locate and 0.63 ms through k_locate_u)
  const bool use_u = (kmax <= 5 || (kmax <= kLocUMaxK && cnt >= 4096)) && M >= 5 && M <= 22 &&
                     !(uenv && atoi(uenv) == 0);
  const unsigned ngroups = (unsigned)(cnt / 1024) + (unsigned)kmax + 1u;   // >= sum_K ceil(ceil(count_K/32)/32)
  // descriptors: [0, kMaxBuckets] backward (butterfly) passes, [kMaxBuckets + 1, 2 kMaxBuckets + 1] forward passes
  const size_t desc_bytes = sizeof(PassDesc) * 2 * (kMaxBuckets + 1);
  // work buffer: bucket counts, bucket lists, forward values
  const size_t f_words = use_u ? (((size_t)ngroups << kmax) * M * 32u) : (((size_t)nvirt << kmax) * PS);
  const size_t head_words = (size_t)(kMaxBuckets + 1) + 2u * ((size_t)tcap + 1u);   // bucket counts, per-degree counts and cursors
  const size_t need = (head_words + (size_t)(kMaxBuckets + 1) * cnt + f_words) * 4 + 256;
  int rc = w.fwd_buf.reserve(need);
  if (rc) return rc;
  const bool fresh_descs = w.descs.p == nullptr;
  if ((rc = w.descs.reserve(desc_bytes))) return rc;
  PassDesc *d_descs = (PassDesc *)w.descs.p;
  uint32_t *counts = (uint32_t *)w.fwd_buf.p;
  uint32_t *by_deg = counts + (kMaxBuckets + 1);
  uint32_t *lists = counts + head_words;
  // forward values are read and written as 16-byte vectors
  uint32_t *F = (uint32_t *)(((uintptr_t)(lists + (size_t)(kMaxBuckets + 1) * cnt) + 15) & ~(uintptr_t)15);
  std::vector<PassDesc> fwd(kMaxBuckets + 1), bwd(kMaxBuckets + 1);
  for (int K = 1; K <= kmax; K++) locate_descs(M, K, t, c->f->twist_free, fwd[K], bwd[K]);
  if (fresh_descs || w.descs_t != t || w.descs_kmax != kmax) {
    // the descriptors only depend on (t, kmax): uploaded once per work set
    BCHFFT_CUDA_TRY(cudaMemcpyAsync(d_descs, bwd.data(), sizeof(PassDesc) * (kMaxBuckets + 1),
                                    cudaMemcpyHostToDevice, st));
    BCHFFT_CUDA_TRY(cudaMemcpyAsync(d_descs + (kMaxBuckets + 1), fwd.data(), sizeof(PassDesc) * (kMaxBuckets + 1),
                                    cudaMemcpyHostToDevice, st));
    BCHFFT_CUDA_TRY(cudaStreamSynchronize(st));          // bwd is a stack object
    w.descs_t = t;
    w.descs_kmax = kmax;
  }
  BCHFFT_CUDA_TRY(cudaMemsetAsync(counts, 0, head_words * 4, st));
  // locators of degree 1 and 2 are solved in closed form by k_bucket_count itself (BCHFFT_LOCATE_DIRECT=0:
  // all degrees go through the search)
  const char *denv = getenv("BCHFFT_LOCATE_DIRECT");
  const uint32_t direct_max = denv ? (uint32_t)(atoi(denv) < 0 ? 0 : (atoi(denv) > 2 ? 2 : atoi(denv))) : 2u;
  {
    const size_t smem_b = sizeof(uint32_t) * 2u * ((size_t)tcap + 1u);
    BCHFFT_ENSURE_SMEM(k_bucket_count, c->f->device, smem_b);
    BCHFFT_ENSURE_SMEM(k_bucket_scatter, c->f->device, smem_b);
    BCHFFT_RUN(k_bucket_count, dim3((unsigned)((cnt + 255) / 256)), dim3(256), smem_b, st, (const int32_t *)s.flag,
               (const uint32_t *)s.L, tcap, (uint32_t)cnt, by_deg, direct_max, (const uint32_t *)s.elp, c->d,
               c->f->order, (const uint32_t *)c->f->d_log, (const uint32_t *)c->f->d_exp,
               (const uint32_t *)c->f->d_qsolve, s.count, s.positions, s.pos_stride);
    BCHFFT_RUN(k_bucket_scatter, dim3((unsigned)((cnt + 255) / 256)), dim3(256), smem_b, st, (const int32_t *)s.flag,
               (const uint32_t *)s.L, tcap, (uint32_t)cnt, counts, by_deg, lists, direct_max);
  }
  if (use_u) {
    // forward phase of all buckets in one launch (blockIdx.y = K - 1): 256 coefficient columns per CTA,
    // i.e. 2^(8-K) slabs; CTAs beyond a bucket's population exit
    {
      const size_t smem_f = sizeof(uint32_t) * M * 256u;
      BCHFFT_ENSURE_SMEM(k_elp_fwd_u<M>, c->f->device, smem_f);
      const unsigned sb_min = kmax >= 5 ? 8u - (unsigned)kmax : 3u;       // slabs per CTA of the largest bucket: 2^sb_min
      BCHFFT_RUN(k_elp_fwd_u<M>, dim3((ns + (1u << sb_min) - 1u) >> sb_min, (unsigned)kmax), dim3(256), smem_f, st,
                 (const uint32_t *)s.elp, (const uint32_t *)s.L, c->d, (const uint32_t *)counts,
                 (const uint32_t *)lists, (uint32_t)cnt, (uint32_t)kmax, c->f->tab,
                 (const PassDesc *)(d_descs + (kMaxBuckets + 1)), F);
    }
    // enough CTAs per group to fill the machine and to even out the buckets' different costs
    const unsigned ntiles = (1u << M) >> 5;
    // (measured at 10^6 words of (8191, 16): 2400 CTAs 4.30 ms, 9000..40000 CTAs 4.00..4.06 ms)
    const char *cenv = getenv("BCHFFT_LOCU_CTAS");       // target number of CTAs of the launch
    const unsigned want_ctas = cenv && atoi(cenv) >= 1 ? (unsigned)atoi(cenv) : 16000u;
    unsigned gx = (want_ctas + ngroups - 1u) / ngroups;
    if (gx > ntiles) gx = ntiles;
    if (gx < 1u) gx = 1u;
    uint32_t qcap = kLocUQueue;
    const char *qenv = getenv("BCHFFT_LOCU_QUEUE");
    if (qenv && atoi(qenv) >= 1 && atoi(qenv) <= (int)kLocUQueue) qcap = (uint32_t)atoi(qenv);
    const size_t smem_u = sizeof(uint32_t) * (32u * M * 32u + qcap);
    BCHFFT_ENSURE_SMEM(k_locate_u<M>, c->f->device, smem_u);
    BCHFFT_RUN(k_locate_u<M>, dim3(gx, ngroups), dim3(kNtLoc), smem_u, st, (const uint32_t *)F,
               (const uint32_t *)counts, (const uint32_t *)lists,
               getenv("BCHFFT_LOCU_NOSKIP") ? (const uint32_t *)nullptr : (const uint32_t *)s.L, (uint32_t)cnt,
               (uint32_t)kmax, c->f->tab, (const uint32_t *)c->f->d_errloc, s.count, s.positions, s.pos_stride, qcap);
    return BCHFFT_OK;
  }
  for (int K = 1; K <= kmax; K++) {
    // forward phase: 2^sb slabs per CTA so that a CTA has 256 coefficient columns to work on
    uint32_t sb = 0;
    while ((1u << (K + sb)) < 256u) sb++;
    const size_t smem_f = sizeof(uint32_t) * M * ((size_t)1 << (K + sb));
    BCHFFT_ENSURE_SMEM(k_elp_fwd<M>, c->f->device, smem_f);
    BCHFFT_RUN(k_elp_fwd<M>, dim3((ns + (1u << sb) - 1u) >> sb), dim3(256), smem_f, st,
               (const uint32_t *)s.elp, (const uint32_t *)s.L, c->d, (const uint32_t *)counts,
               (const uint32_t *)lists, (uint32_t)cnt, sb, (uint32_t)kmax, c->f->tab, fwd[K], F);
  }
  const size_t smem = sizeof(uint32_t) * M * ((size_t)1 << t);
  BCHFFT_ENSURE_SMEM(k_locate<M>, c->f->device, smem);
  BCHFFT_RUN(k_locate<M>, dim3(1u << (M - t), nvirt), dim3(kNtLoc), smem, st, (const uint32_t *)F,
             (const uint32_t *)counts, (const uint32_t *)lists, (uint32_t)cnt, (uint32_t)kmax, c->f->tab,
             (const PassDesc *)d_descs, (const uint32_t *)c->f->d_errloc, s.count, s.positions,
             s.pos_stride);
  return BCHFFT_OK;
}

// patch_dst: see k_correct.  h_positions (pinned host, [batch][t + 1]): "positions" result mode -- the
// flips are left to the host, every chunk's position lists are copied there right after its verdict.
template <int M>
static int decode_run(bchfft_code *c, bchfft_code::WorkSet &w, uint64_t *d_words, size_t batch, uint8_t *d_ok,
                      uint32_t *d_corrected, cudaStream_t st, uint64_t *patch_dst = nullptr,
                      uint32_t *h_positions = nullptr)
{
  const uint32_t pos_stride = c->t + 1u;
  // Chunk the batch only to bound the scratch (about 5 d + t words per codeword).  The decode is
  // integer-pipe bound by a wide margin (intermediate traffic is ~1.5 KB per codeword), so L2
  // residency of the intermediates does not matter, whereas small chunks starve the
  // one-thread-per-codeword kernels of parallelism.
  const char *env = getenv("BCHFFT_DECODE_CHUNK");
  size_t chunk = env ? (size_t)atol(env) : ((size_t)1536 << 20) / scratch_per_cw(c, pos_stride, M);
  chunk = (chunk / 32) * 32;
  if (chunk < 32) chunk = 32;
  if (chunk > batch) chunk = batch;
  // gridDim.y limits: k_syndrome launches one row per slab, k_locate one per virtual slab (slabs + kmax)
  if (chunk > (size_t)(65535 - kMaxBuckets) * 32) chunk = (size_t)(65535 - kMaxBuckets) * 32;
  int rc = w.scratch.reserve(scratch_bytes(c, chunk, pos_stride, M));
  if (rc) return rc;
  for (size_t i0 = 0; i0 < batch; i0 += chunk) {
    const size_t cnt = batch - i0 < chunk ? batch - i0 : chunk;
    Scratch s;
    carve(c, w.scratch.p, cnt, pos_stride, M, s);
    uint64_t *wp = d_words + i0 * c->wpc;
    // the shared-memory Berlekamp-Massey kernel un-bitslices the syndromes itself
    BmShape bm_shape;
    const char *fenv = getenv("BCHFFT_BM_FUSED");
    const bool bm_sm = bm_in_smem<M>(c, &bm_shape);
    const bool fused = bm_sm && !bm_wide(c, cnt) && !(fenv && atoi(fenv) == 0);
    BCHFFT_CUDA_TRY(cudaMemsetAsync(w.scratch.p, 0, zero_bytes(c, cnt, M, !bm_sm), st));
    if ((rc = launch_syndrome<M>(c, w, wp, cnt, s, st, !fused))) return rc;
    if ((rc = launch_bm<M>(c, cnt, s, st, fused, c->t, true))) return rc;
    if ((rc = launch_locate<M>(c, w, cnt, s, c->t, st))) return rc;
    BCHFFT_RUN(k_correct, dim3((unsigned)((cnt + 127) / 128)), dim3(128), 0, st, wp, c->wpc,
                  (uint32_t)cnt, (const int32_t *)s.flag, (const uint32_t *)s.L, c->t,
                  (const uint32_t *)s.count, (const uint32_t *)s.positions, s.pos_stride,
                  d_ok ? d_ok + i0 : (uint8_t *)nullptr, d_corrected ? d_corrected + i0 : (uint32_t *)nullptr,
                  patch_dst ? patch_dst + i0 * c->wpc : (uint64_t *)nullptr, h_positions ? 0u : 1u);
    if (h_positions)
      BCHFFT_CUDA_TRY(cudaMemcpyAsync(h_positions + i0 * pos_stride, s.positions, cnt * pos_stride * sizeof(uint32_t),
                                      cudaMemcpyDeviceToHost, st));
  }
  return BCHFFT_OK;
}

// cudaMalloc + checked synchronous upload of a host vector (tables built once per code)
template <class T>
static int upload_table(const std::vector<T> &v, T **out)
{
  *out = nullptr;
  BCHFFT_CUDA_TRY(cudaMalloc((void **)out, v.size() * sizeof(T) + 16));
  BCHFFT_CUDA_TRY(cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return BCHFFT_OK;
}

// minimal polynomial of x^j over GF(2), as a bit mask (bit i = coefficient of X^i)
static uint64_t minimal_poly(const bchfft_field *f, uint32_t j, int *deg_out)
{
  const uint32_t order = f->order;
  const uint32_t *ex = f->h_exp.data(), *lg = f->h_log.data();
  auto mul = [&](uint32_t a, uint32_t b) -> uint32_t {
    if (!a || !b) return 0u;
    uint32_t s = lg[a] + lg[b];
    if (s >= order) s -= order;
    return ex[s];
  };
  std::vector<uint32_t> p(1, 1u);
  uint32_t e = j % order;
  const uint32_t start = e;
  do {
    const uint32_t root = ex[e];
    p.push_back(0u);
    for (size_t i = p.size() - 1; i > 0; i--) p[i] = p[i - 1] ^ mul(p[i], root);
    p[0] = mul(p[0], root);
    e = (uint32_t)(((uint64_t)e * 2u) % order);
  } while (e != start);
  uint64_t mask = 0;
  for (size_t i = 0; i < p.size(); i++)
    if (p[i] & 1u) mask |= 1ull << i;
  *deg_out = (int)p.size() - 1;
  return mask;
}

}  // namespace bchfft

using namespace bchfft;

extern "C" int bchfft_code_create(bchfft_field *f, uint32_t distance, uint32_t length,
                                  const uint64_t *packed_gen_poly, uint32_t gen_poly_degree,
                                  bchfft_code **out)
{
  if (!f || !out) {
    set_error("bchfft_code_create: null argument");
    return BCHFFT_ERR_ARG;
  }
  *out = nullptr;
  // any designed distance the reference accepts (bch.c:37: d < length): t = (d-1)/2 errors are
  // corrected, Berlekamp-Massey runs d-1 iterations on S_1 .. S_(d-1) (bch.c:340), so an even d
  // needs one more odd syndrome class than an odd one: nodd = d/2
  if (distance < 3 || length < 2 || length > f->n || distance >= length) {
    set_error("bchfft_code_create: need 3 <= distance < length <= 2^m");
    return BCHFFT_ERR_ARG;
  }
  BCHFFT_CUDA_TRY(cudaSetDevice(f->device));
  bchfft_code *c = new bchfft_code();
  c->f = f;
  c->d = distance;
  c->t = (distance - 1u) / 2u;
  c->length = length;
  c->gen_degree = gen_poly_degree;
  c->wpc = (length - 1u) / 64u + 1u;
  c->nodd = distance / 2u;
  {
    // method rule of bch_eval_syndrome (bch.c:240-254), float arithmetic as in the reference
    const float threshold = 0.5f * f->m * f->m;
    const float cost = ((float)distance) * ((float)gen_poly_degree) / ((float)f->n);
    c->syn_fft_rule = cost > threshold ? 1u : 0u;
  }
  const uint32_t M = f->m, order = f->order;
  const uint32_t npos = c->wpc * 64u;
  if ((size_t)c->nodd * M * 4 + (size_t)(kSynSub + 1) * 4 > (size_t)200 * 1024) {
    set_error("bchfft_code_create: t = %u too large for the shared-memory syndrome accumulators", c->t);
    delete c;
    return BCHFFT_ERR_UNSUPPORTED;
  }
  // position chunk per CTA: as much of the codeword as fits next to the accumulators
  size_t budget = ((size_t)200 * 1024) / 4 - (size_t)c->nodd * M;
  uint32_t chunk = 1u << 15;
  while (chunk > (uint32_t)kSynSub && (chunk + chunk / 32 > budget || chunk / 2 >= npos)) chunk >>= 1;
  c->chunk_pos = chunk;
  c->nchunks = (npos + chunk - 1u) / chunk;
  c->nsub_total = c->nchunks * (chunk / kSynSub);
  std::vector<uint32_t> taps(c->nodd), fold((size_t)c->nodd * c->nsub_total * M);
  for (uint32_t ji = 0; ji < c->nodd; ji++) {
    const uint32_t j = 2u * ji + 1u;
    int deg = 0;
    const uint64_t mp = minimal_poly(f, j, &deg);
    // Q_j = X^(M-deg) * M_j has degree exactly M; the LFSR taps are its low M coefficients
    const uint64_t q = mp << (M - (uint32_t)deg);
    taps[ji] = (uint32_t)(q & ((1ull << M) - 1ull));
    for (uint32_t qg = 0; qg < c->nsub_total; qg++)
      for (uint32_t b = 0; b < M; b++) {
        const uint64_t e = ((uint64_t)j % order) * (((uint64_t)b + (uint64_t)qg * kSynSub) % order) % order;
        fold[((size_t)ji * c->nsub_total + qg) * M + b] = f->h_exp[e];
      }
  }
  // syndrome[0] selector bits (see k_syndrome): all ones when the reference evaluates c at 1
  // (FFT path), (X^i mod g)(1) when it evaluates the remainder c mod g at 1 (direct path)
  std::vector<uint32_t> vbits(c->wpc * 2u, 0u);
  if (c->syn_fft_rule || !packed_gen_poly) {
    for (uint32_t i = 0; i < length; i++) vbits[i >> 5] |= 1u << (i & 31u);
  } else {
    const uint32_t gd = gen_poly_degree, gw = gd / 64u + 1u;
    std::vector<uint64_t> r(gw + 1, 0ull);   // r = X^i mod g, starts at 1
    r[0] = 1ull;
    uint32_t par = 1u;
    for (uint32_t i = 0; i < length; i++) {
      if (par) vbits[i >> 5] |= 1u << (i & 31u);
      // r <- r * X mod g
      uint64_t carry = 0;
      for (uint32_t k = 0; k < gw; k++) {
        const uint64_t nc = r[k] >> 63;
        r[k] = (r[k] << 1) | carry;
        carry = nc;
      }
      if ((r[gd >> 6] >> (gd & 63u)) & 1ull)
        for (uint32_t k = 0; k < gw; k++) r[k] ^= packed_gen_poly[k];
      par = 0u;
      for (uint32_t k = 0; k < gw; k++) par ^= (uint32_t)__builtin_popcountll(r[k]);
      par &= 1u;
    }
  }
  {
    int urc;
    if ((urc = upload_table(vbits, &c->d_vbits)) || (urc = upload_table(taps, &c->d_taps)) ||
        (urc = upload_table(fold, &c->d_fold))) {
      bchfft_code_destroy(c);
      return urc;
    }
  }
  // universal-LFSR syndromes: every class j must permute the positions (gcd(j, order) = 1), the
  // whole codeword must fit shared memory, and no codeword bit may sit at position >= order
  {
    auto gcd = [](uint32_t a, uint32_t b) { while (b) { const uint32_t r = a % b; a = b; b = r; } return a; };
    bool ok = M >= 9 && M <= 15 && length <= order;
    for (uint32_t ji = 0; ji < c->nodd && ok; ji++) ok = gcd((2u * ji + 1u) % order, order) == 1u;
    const size_t smem_p = 2u * sizeof(uint32_t) * ((size_t)((order + 31u) / 32u) * 32u + (size_t)c->nodd * M);
    if (ok && smem_p <= (size_t)200 * 1024) {
      const uint32_t lc_min = (order + 31u) / 32u, lc_max = lc_min + lc_min / 8u;
      auto inverse = [&](uint32_t v) -> uint32_t {   // v^-1 mod order by the extended Euclid
        int64_t a = v, b = order, x0 = 1, x1 = 0;
        while (b) { const int64_t qq = a / b, r = a % b; a = b; b = r; const int64_t x2 = x0 - qq * x1; x0 = x1; x1 = x2; }
        return (uint32_t)(((x0 % (int64_t)order) + order) % order);
      };
      // shared-memory wavefronts of the 32 lanes' reads, sampled every 4th step
      auto cost = [&](uint32_t inv, uint32_t lc) -> uint64_t {
        uint32_t idx[32], n[32];
        uint32_t longest = 0;
        for (uint32_t q = 0; q < 32u; q++) {
          const uint64_t e0 = (uint64_t)q * lc, e1 = e0 + lc < order ? e0 + lc : order;
          n[q] = e1 > e0 ? (uint32_t)(e1 - e0) : 0u;
          idx[q] = n[q] ? (uint32_t)(((e1 - 1u) * (uint64_t)inv) % order) : 0u;
          if (n[q] > longest) longest = n[q];
        }
        const uint32_t step4 = (uint32_t)(((uint64_t)inv * 4u) % order);
        uint64_t total = 0;
        for (uint32_t k = 0; k < longest; k += 4u) {
          // 64-bit loads (two slabs per position): each half-warp is one request over 16 bank pairs
          for (uint32_t half = 0; half < 2u; half++) {
            uint8_t load[16] = {0};
            uint32_t worst = 0;
            for (uint32_t q = half * 16u; q < half * 16u + 16u; q++) {
              if (k >= n[q]) continue;
              const uint32_t bank = idx[q] & 15u;
              if (++load[bank] > worst) worst = load[bank];
              idx[q] = idx[q] >= step4 ? idx[q] - step4 : idx[q] + order - step4;
            }
            total += worst;
          }
        }
        return total;
      };
      const char *awenv = getenv("BCHFFT_SYN_ALUW");
      const uint32_t alu_weight = awenv ? (uint32_t)atoi(awenv) : 4u;   // issue cost of a step in wavefront units (measured best)
      std::vector<uint32_t> jinv(c->nodd), lcs(c->nodd), start((size_t)c->nodd * 32u), foldp((size_t)c->nodd * 32u * M);
      for (uint32_t ji = 0; ji < c->nodd; ji++) {
        const uint32_t j = (2u * ji + 1u) % order;
        uint64_t best = ~0ull;
        uint32_t best_i = 0, best_lc = lc_min, best_inv = inverse(j);
        uint32_t jc = j;   // conjugate j 2^i
        for (uint32_t i = 0; i < M; i++, jc = (uint32_t)(((uint64_t)jc * 2u) % order)) {
          const uint32_t inv = inverse(jc);
          for (uint32_t lc = lc_min; lc <= lc_max; lc++) {
            // time per chunk ~ steps x (issue cost of a step + wavefronts of its load): longer
            // chunks (idle lanes at the end of the range) cost issue slots even when they remove
            // conflicts.  cost() sums the wavefronts over every 4th step.
            const uint64_t cst = cost(inv, lc) + (uint64_t)alu_weight * ((lc + 3u) / 4u);
            if (cst < best) { best = cst; best_i = i; best_lc = lc; best_inv = inv; }
          }
          if (i > 0 && jc == j) break;   // short orbit
        }
        jinv[ji] = best_inv;
        lcs[ji] = best_lc;
        // S_j = S_(j 2^i)^(2^(M-i)): raise the fold constants to that power (exponents times 2^(M-i))
        const uint32_t fpow = (M - best_i) % M;
        for (uint32_t q = 0; q < 32u; q++) {
          const uint64_t e0 = (uint64_t)q * best_lc, e1 = e0 + best_lc < order ? e0 + best_lc : order;
          start[(size_t)ji * 32u + q] = e1 > e0 ? (uint32_t)(((e1 - 1u) * (uint64_t)best_inv) % order) : 0u;
          for (uint32_t b = 0; b < M; b++) {
            uint64_t ex = ((uint64_t)q * best_lc + b) % order;
            for (uint32_t r = 0; r < fpow; r++) ex = (ex * 2u) % order;
            foldp[((size_t)ji * 32u + q) * M + b] = f->h_exp[ex];
          }
        }
      }
      int urc;
      if ((urc = upload_table(jinv, &c->d_jinv)) || (urc = upload_table(start, &c->d_start)) ||
          (urc = upload_table(foldp, &c->d_foldp)) || (urc = upload_table(lcs, &c->d_lcs))) {
        bchfft_code_destroy(c);
        return urc;
      }
      c->syn_p = 1u;
    }
  }
  // encoder tables (need the generator polynomial)
  if (packed_gen_poly && gen_poly_degree >= 8u && gen_poly_degree < length) {
    const uint32_t dg = gen_poly_degree, gw = (dg >> 6) + 1u;
    const char *wenv = getenv("BCHFFT_ENC_WIDE");   // test hook: force the warp-per-codeword encoder
    const bool wide = dg > 2048u || (wenv && atoi(wenv) != 0 && dg >= 64u);
    std::vector<uint64_t> table, r(gw + 1u);
    uint32_t nw = 1;
    // r <- r X mod g (r has degree < dg)
    auto times_x = [&]() {
      uint64_t carry = 0;
      for (uint32_t k = 0; k < gw; k++) {
        const uint64_t nc = r[k] >> 63;
        r[k] = (r[k] << 1) | carry;
        carry = nc;
      }
      if ((r[dg >> 6] >> (dg & 63u)) & 1ull)
        for (uint32_t k = 0; k < gw; k++) r[k] ^= packed_gen_poly[k];
    };
    if (wide) {
      // rows[k] = X^(dg+k) mod g: row 0 is g without its leading term
      table.assign((size_t)64 * gw, 0ull);
      for (uint32_t k = 0; k < gw; k++) r[k] = packed_gen_poly[k];
      r[dg >> 6] &= ~(1ull << (dg & 63u));
      for (uint32_t row = 0; row < 64u; row++) {
        for (uint32_t k = 0; k < gw; k++) table[(size_t)row * gw + k] = r[k];
        times_x();
      }
    } else {
      while (nw * 64u < dg) nw *= 2u;
      table.assign((size_t)256 * nw, 0ull);
      for (uint32_t v = 0; v < 256u; v++) {
        // r = v(X) X^dg mod g: feed the 8 bits of v (highest first) through the division by g
        std::fill(r.begin(), r.end(), 0ull);
        for (int bit = 7; bit >= 0; bit--) {
          const uint32_t fb = (uint32_t)((r[(dg - 1u) >> 6] >> ((dg - 1u) & 63u)) & 1ull) ^ ((v >> bit) & 1u);
          uint64_t carry = 0;
          for (uint32_t k = 0; k < gw; k++) {
            const uint64_t nc = r[k] >> 63;
            r[k] = (r[k] << 1) | carry;
            carry = nc;
          }
          r[dg >> 6] &= ~(1ull << (dg & 63u));                   // drop the bit shifted out of degree dg-1
          if (fb)
            for (uint32_t k = 0; k < gw; k++) r[k] ^= packed_gen_poly[k];
          r[dg >> 6] &= (dg & 63u) ? ((1ull << (dg & 63u)) - 1ull) : 0ull;   // g's leading coefficient
        }
        for (uint32_t k = 0; k < nw && k < gw; k++) table[(size_t)v * nw + k] = r[k];
      }
    }
    int urc = upload_table(table, &c->d_enc_table);
    if (urc) {
      bchfft_code_destroy(c);
      return urc;
    }
    c->enc_nw = wide ? gw : nw;
    c->enc_wide = wide ? 1u : 0u;
  }
  // every table is in place before the first kernel on a non-blocking stream can read it
  {
    const cudaError_t se = cudaDeviceSynchronize();
    if (se != cudaSuccess) {
      set_error("bchfft_code_create: table upload failed: %s", cudaGetErrorString(se));
      bchfft_code_destroy(c);
      return BCHFFT_ERR_CUDA;
    }
  }
  *out = c;
  return BCHFFT_OK;
}

extern "C" void bchfft_code_destroy(bchfft_code *c)
{
  if (!c) return;
  cudaSetDevice(c->f->device);
  cudaStreamSynchronize(c->f->stream);
  cudaFree(c->d_taps);
  cudaFree(c->d_fold);
  cudaFree(c->d_vbits);
  cudaFree(c->d_jinv);
  cudaFree(c->d_start);
  cudaFree(c->d_foldp);
  cudaFree(c->d_lcs);
  cudaFree(c->d_enc_table);
  for (auto &w : c->work) {
    w.scratch.release();
    w.fwd_buf.release();
    w.words.release();
    w.ok.release();
    w.corr.release();
    w.fft_ws.release();
    w.fft_io.release();
    w.descs.release();
    w.h_ok.release();
    w.h_corr.release();
    w.h_pos.release();
    if (w.results_ready) cudaEventDestroy(w.results_ready);
    if (w.stream) cudaStreamDestroy(w.stream);
  }
  delete c;
}

#define BCHFFT_DISPATCH_M(m, CALL)                       \
  switch (m) {                                           \
    BCHFFT_M_LIST(CALL)                                  \
    default:                                             \
      set_error("unsupported field degree %u", (m));     \
      return BCHFFT_ERR_UNSUPPORTED;                     \
  }

extern "C" int bchfft_bch_decode_dev(bchfft_code *c, uint64_t *d_words, size_t batch, uint8_t *d_ok,
                                     uint32_t *d_corrected, void *stream)
{
  if (!c || !d_words) {
    set_error("bchfft_bch_decode_dev: null argument");
    return BCHFFT_ERR_ARG;
  }
  if (batch == 0) return BCHFFT_OK;
  BCHFFT_CUDA_TRY(cudaSetDevice(c->f->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : c->f->stream;
#define X(M) case M: return decode_run<M>(c, c->work[0], d_words, batch, d_ok, d_corrected, st);
  BCHFFT_DISPATCH_M(c->f->m, X)
#undef X
}

// Host-pointer entry point: the batch is cut into chunks that travel through up to three lanes
// (stream + staging + scratch each), so the host->device copy of chunk i+1 and the result traffic of
// chunk i-1 overlap the decode of chunk i.  (Full overlap needs page-locked host memory; with pageable
// memory the copies are staged by the driver and mostly serialise.)
//
// Results come back in one of three ways (BCHFFT_D2H_MODE = positions | words | patch picks one):
//   positions  (default from 8192 codewords on) only the verdicts and the error-position lists cross the
//          bus (73 bytes per codeword at t = 16 instead of 1 KiB); a few host threads apply the bit flips
//          (add_bit, bch.c:539-542) to the caller's array while later chunks are still in flight.  The
//          array already holds the received words -- it is where they were copied from -- so flipping the
//          located bits completes bch_decode's in-place contract.
//   words  every decoded word is copied back (small batches: no thread start-up cost).
//   patch  page-locked caller arrays only: k_correct writes the 64-bit words it changed straight into the
//          caller's array through unified addressing.  Measured (profiles/r2c_zerocopy_scatter_bench.jsonl):
//          scattered 8-byte writes over PCIe run at 4*10^8 per second, i.e. no faster than copying whole
//          words, and they slow the concurrent host-to-device copies down: kept as an option, not a default.
namespace bchfft {
static uint64_t *device_alias_of_host(const void *host_ptr)
{
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, host_ptr) != cudaSuccess) {
    cudaGetLastError();   /* not an error for us: plain pageable memory on older runtimes */
    return nullptr;
  }
  if (attr.type != cudaMemoryTypeHost || !attr.devicePointer) return nullptr;
  return (uint64_t *)attr.devicePointer;
}

// one chunk whose verdicts and position lists are (or will be, once `ready` fires) in pinned staging
struct FlipJob {
  uint64_t *words;            // caller's array, first codeword of the chunk
  uint8_t *ok_out;            // caller's arrays (may be null)
  uint32_t *corr_out;
  const uint8_t *ok;          // staging
  const uint32_t *corr, *pos;
  size_t cnt;
  cudaEvent_t ready;
  int lane;
};

// The host threads of one bchfft_bch_decode_host call in "positions" mode.  Every worker owns a fixed
// slice of EVERY chunk, so the chunks complete in order and a lane's staging is free again as soon as all
// workers have passed its chunk.
struct FlipPool {
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cv;
  std::vector<FlipJob> jobs;          // published jobs, in order
  bool closed = false;
  std::vector<size_t> done;           // per job: workers that finished it
  int device = 0;
  uint32_t wpc = 0, pos_stride = 0;
  size_t nworkers = 0;

  void worker(size_t me)
  {
    cudaSetDevice(device);
    for (size_t k = 0;; k++) {
      FlipJob j;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return k < jobs.size() || closed; });
        if (k >= jobs.size()) return;
        j = jobs[k];
      }
      cudaEventSynchronize(j.ready);
      const size_t lo = j.cnt * me / nworkers, hi = j.cnt * (me + 1) / nworkers;
      for (size_t i = lo; i < hi; i++) {
        // pull the words the flips of codeword i + 8 will touch into the cache while i is handled
        const size_t ahead = i + 8;
        if (ahead < hi && j.ok[ahead]) {
          const uint32_t na = j.corr[ahead];
          for (uint32_t e = 0; e < na; e++) {
            const uint32_t p = j.pos[ahead * pos_stride + e];
            if ((p >> 6) < wpc) __builtin_prefetch(j.words + ahead * wpc + (p >> 6), 1, 0);
          }
        }
        if (j.ok[i]) {                                   // corrected = L roots found: flip them (bch.c:539-542)
          const uint32_t nf = j.corr[i];
          uint64_t *w = j.words + i * wpc;
          for (uint32_t e = 0; e < nf; e++) {
            const uint32_t p = j.pos[i * pos_stride + e];
            if ((p >> 6) < wpc) w[p >> 6] ^= 1ull << (p & 63u);
          }
        }
      }
      if (j.ok_out) memcpy(j.ok_out + lo, j.ok + lo, hi - lo);
      if (j.corr_out) memcpy(j.corr_out + lo, j.corr + lo, (hi - lo) * sizeof(uint32_t));
      {
        std::lock_guard<std::mutex> lk(mu);
        done[k]++;
      }
      cv.notify_all();
    }
  }
  void start(size_t n, int dev, uint32_t wpc_, uint32_t pos_stride_)
  {
    nworkers = n;
    device = dev;
    wpc = wpc_;
    pos_stride = pos_stride_;
    for (size_t i = 0; i < n; i++) threads.emplace_back([this, i] { worker(i); });
  }
  size_t publish(const FlipJob &j)
  {
    std::lock_guard<std::mutex> lk(mu);
    jobs.push_back(j);
    done.push_back(0);
    cv.notify_all();
    return jobs.size() - 1;
  }
  void wait_done(size_t k)
  {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return done[k] == nworkers; });
  }
  void finish()
  {
    {
      std::lock_guard<std::mutex> lk(mu);
      closed = true;
    }
    cv.notify_all();
    for (auto &t : threads) t.join();
    threads.clear();
  }
};

static size_t flip_threads()
{
  const char *env = getenv("BCHFFT_FLIP_THREADS");
  if (env && atoi(env) >= 1 && atoi(env) <= 64) return (size_t)atoi(env);
  // The host's hardware threads divided between the ranks of a one-process-per-GPU job (torchrun exports
  // LOCAL_WORLD_SIZE) or the devices one process drives at once (bchfft_set_concurrent_hosts), between 2 and 8
  // per call.  Measured: one flipping thread keeps up with about 15 M codewords/s (the flips are cache misses
  // into the caller's array), the host-to-device copy of one GPU with 54 M.  16 vCPUs, one GPU: 2 threads 38.7,
  // 4 threads 49.5 M codewords/s end to end; 24 vCPUs, two ranks: 2 threads each 54.8 M, 4 each 79.3 M; 32
  // vCPUs, eight devices in one process: 1 thread each 88, 2 each 90, 4 each 99, 8 each 98 M codewords/s
  // (profiles/r2d_e2e_result_modes_1gpu.jsonl, r2m_*, r2zz_*).
  const char *lws = getenv("LOCAL_WORLD_SIZE");
  size_t ranks = lws && atoi(lws) >= 1 ? (size_t)atoi(lws) : 1;
  // ... or with the other devices one process drives at once (bchfft_set_concurrent_hosts)
  const size_t peers = (size_t)g_concurrent_hosts.load();
  if (peers > ranks) ranks = peers;
  size_t n = std::thread::hardware_concurrency() / ranks;
  return n < 2 ? 2 : (n > 8 ? 8 : n);
}

template <int M>
static int decode_host_run(bchfft_code *c, uint64_t *words, size_t batch, uint8_t *ok, uint32_t *corrected)
{
  const size_t cw_bytes = (size_t)c->wpc * 8;
  const char *env = getenv("BCHFFT_E2E_CHUNK");
  size_t per = env ? (size_t)atol(env) : ((size_t)64 << 20) / cw_bytes;
  per = per < 32 ? 32 : (per / 32) * 32;
  if (per > batch) per = (batch + 31) / 32 * 32;
  int rc = BCHFFT_OK;
  const char *lenv = getenv("BCHFFT_E2E_LANES");
  const int nlanes = (lenv && atoi(lenv) >= 1 && atoi(lenv) <= 3) ? atoi(lenv) : 3;
  const char *menv = getenv("BCHFFT_D2H_MODE");
  enum { WORDS, PATCH, POSITIONS } mode = batch >= 8192 ? POSITIONS : WORDS;
  if (menv && !strcmp(menv, "words")) mode = WORDS;
  else if (menv && !strcmp(menv, "positions")) mode = POSITIONS;
  else if (menv && !strcmp(menv, "patch")) mode = PATCH;
  uint64_t *patch = mode == PATCH ? device_alias_of_host(words) : nullptr;
  if (mode == PATCH && !patch) mode = WORDS;             // pageable memory: the device cannot address it
  const uint32_t pos_stride = c->t + 1u;
  for (int l = 1; l <= nlanes; l++) {
    bchfft_code::WorkSet &w = c->work[l];
    if (!w.stream) BCHFFT_CUDA_TRY(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
    if ((rc = w.words.reserve(per * cw_bytes))) return rc;
    if ((rc = w.ok.reserve(per))) return rc;
    if ((rc = w.corr.reserve(per * 4))) return rc;
    if (mode == POSITIONS) {
      if (!w.results_ready) BCHFFT_CUDA_TRY(cudaEventCreateWithFlags(&w.results_ready, cudaEventDisableTiming));
      if ((rc = w.h_ok.reserve(per))) return rc;
      if ((rc = w.h_corr.reserve(per * 4))) return rc;
      if ((rc = w.h_pos.reserve(per * (size_t)pos_stride * 4))) return rc;
    }
  }
  FlipPool pool;
  if (mode == POSITIONS) pool.start(flip_threads(), c->f->device, c->wpc, pos_stride);
  size_t lane_job[4] = {(size_t)-1, (size_t)-1, (size_t)-1, (size_t)-1};
  auto enqueue = [&](size_t i0, int lane) -> int {
    const size_t cnt = batch - i0 < per ? batch - i0 : per;
    bchfft_code::WorkSet &w = c->work[lane];
    cudaStream_t st = w.stream;
    uint64_t *dw = (uint64_t *)w.words.p;
    if (mode == POSITIONS && lane_job[lane] != (size_t)-1) pool.wait_done(lane_job[lane]);   // staging free again
    BCHFFT_CUDA_TRY(cudaMemcpyAsync(dw, words + i0 * c->wpc, cnt * cw_bytes, cudaMemcpyHostToDevice, st));
    int r = decode_run<M>(c, w, dw, cnt, (uint8_t *)w.ok.p, (uint32_t *)w.corr.p, st,
                          patch ? patch + i0 * c->wpc : (uint64_t *)nullptr,
                          mode == POSITIONS ? (uint32_t *)w.h_pos.p : (uint32_t *)nullptr);
    if (r) return r;
    if (mode == POSITIONS) {
      BCHFFT_CUDA_TRY(cudaMemcpyAsync(w.h_ok.p, w.ok.p, cnt, cudaMemcpyDeviceToHost, st));
      BCHFFT_CUDA_TRY(cudaMemcpyAsync(w.h_corr.p, w.corr.p, cnt * 4, cudaMemcpyDeviceToHost, st));
      BCHFFT_CUDA_TRY(cudaEventRecord(w.results_ready, st));
      FlipJob j;
      j.words = words + i0 * c->wpc;
      j.ok_out = ok ? ok + i0 : nullptr;
      j.corr_out = corrected ? corrected + i0 : nullptr;
      j.ok = (const uint8_t *)w.h_ok.p;
      j.corr = (const uint32_t *)w.h_corr.p;
      j.pos = (const uint32_t *)w.h_pos.p;
      j.cnt = cnt;
      j.ready = w.results_ready;
      j.lane = lane;
      lane_job[lane] = pool.publish(j);
      return BCHFFT_OK;
    }
    if (mode == WORDS)
      BCHFFT_CUDA_TRY(cudaMemcpyAsync(words + i0 * c->wpc, dw, cnt * cw_bytes, cudaMemcpyDeviceToHost, st));
    if (ok) BCHFFT_CUDA_TRY(cudaMemcpyAsync(ok + i0, w.ok.p, cnt, cudaMemcpyDeviceToHost, st));
    if (corrected)
      BCHFFT_CUDA_TRY(cudaMemcpyAsync(corrected + i0, w.corr.p, cnt * 4, cudaMemcpyDeviceToHost, st));
    return BCHFFT_OK;
  };
  int lane = 1;
  for (size_t i0 = 0; i0 < batch && !rc; i0 += per, lane = lane % nlanes + 1) rc = enqueue(i0, lane);
  // drain every lane, also on an error: copies into the caller's buffers may still be in flight
  for (int l = 1; l <= nlanes; l++) {
    const cudaError_t e = cudaStreamSynchronize(c->work[l].stream);
    if (e != cudaSuccess && !rc) {
      set_error("bch decode: %s", cudaGetErrorString(e));
      rc = BCHFFT_ERR_CUDA;
    }
  }
  if (mode == POSITIONS) pool.finish();                  // every published chunk is flipped before we return
  return rc;
}
}  // namespace bchfft

extern "C" int bchfft_bch_decode_host(bchfft_code *c, uint64_t *words, size_t batch, uint8_t *ok,
                                      uint32_t *corrected)
{
  if (!c || !words) {
    set_error("bchfft_bch_decode_host: null argument");
    return BCHFFT_ERR_ARG;
  }
  if (batch == 0) return BCHFFT_OK;
  BCHFFT_CUDA_TRY(cudaSetDevice(c->f->device));
#define X(M) case M: return decode_host_run<M>(c, words, batch, ok, corrected);
  BCHFFT_DISPATCH_M(c->f->m, X)
#undef X
}

// ---- encoder ------------------------------------------------------------------------------------
static int encode_launch(bchfft_code *c, const uint64_t *d_data, uint32_t dwords, uint32_t data_bits,
                         uint64_t *d_cw, size_t batch, cudaStream_t st)
{
  if (c->enc_wide) {
#ifdef BCHFFT_EMU
    const unsigned nt = 32;
#else
    const unsigned nt = 256;
#endif
    const size_t smem_w = (size_t)(nt / 32) * 2u * c->enc_nw * 8;
    BCHFFT_ENSURE_SMEM(k_encode_wide, c->f->device, smem_w);
    BCHFFT_RUN(k_encode_wide, dim3((unsigned)((batch + nt / 32 - 1) / (nt / 32))), dim3(nt), smem_w, st, d_data, dwords,
               data_bits, c->gen_degree, (const uint64_t *)c->d_enc_table, d_cw, c->wpc, (uint32_t)batch);
    return BCHFFT_OK;
  }
  const size_t smem = (size_t)256 * c->enc_nw * 8 + (size_t)4 * 32 * 33 * 4;   // table + one staging tile per warp
  const dim3 grid((unsigned)((batch + 127) / 128)), block(128);
#define ENC(NW)                                                                                        \
  case NW:                                                                                             \
    BCHFFT_ENSURE_SMEM(k_encode<NW>, c->f->device, smem);                                              \
    BCHFFT_RUN(k_encode<NW>, grid, block, smem, st, d_data, dwords, data_bits, c->gen_degree,          \
               (const uint64_t *)c->d_enc_table, d_cw, c->wpc, (uint32_t)batch);                       \
    break;
  switch (c->enc_nw) {
    ENC(1) ENC(2) ENC(4) ENC(8) ENC(16) ENC(32)
    default:
      set_error("bch encode: no encoder table for this code");
      return BCHFFT_ERR_UNSUPPORTED;
  }
#undef ENC
  return BCHFFT_OK;
}

static int encode_check(bchfft_code *c, const void *data, const void *cw, uint32_t data_bits, const char *who)
{
  if (!c || !data || !cw) {
    set_error("%s: null argument", who);
    return BCHFFT_ERR_ARG;
  }
  if (!c->enc_nw) {
    set_error("%s: the code context has no encoder (generator polynomial missing, or its degree is below 8)", who);
    return BCHFFT_ERR_UNSUPPORTED;
  }
  if (data_bits > c->length - c->gen_degree) {   // bch.c:495
    set_error("%s: %u data bits do not fit a codeword of %u bits with %u parity bits", who, data_bits,
              c->length, c->gen_degree);
    return BCHFFT_ERR_ARG;
  }
  return BCHFFT_OK;
}

extern "C" int bchfft_bch_encode_dev(bchfft_code *c, const uint64_t *d_data, size_t data_words,
                                     uint32_t data_bits, uint64_t *d_codewords, size_t batch, void *stream)
{
  int rc = encode_check(c, d_data, d_codewords, data_bits, "bchfft_bch_encode_dev");
  if (rc) return rc;
  if (batch == 0) return BCHFFT_OK;
  if (data_words * 64u < data_bits) {
    set_error("bchfft_bch_encode_dev: data_words too small for data_bits");
    return BCHFFT_ERR_ARG;
  }
  BCHFFT_CUDA_TRY(cudaSetDevice(c->f->device));
  return encode_launch(c, d_data, (uint32_t)data_words, data_bits, d_codewords, batch,
                       stream ? (cudaStream_t)stream : c->f->stream);
}

extern "C" int bchfft_bch_encode_host(bchfft_code *c, const uint64_t *data, size_t data_words,
                                      uint32_t data_bits, uint64_t *codewords, size_t batch)
{
  int rc = encode_check(c, data, codewords, data_bits, "bchfft_bch_encode_host");
  if (rc) return rc;
  if (batch == 0) return BCHFFT_OK;
  if (data_words * 64u < data_bits) {
    set_error("bchfft_bch_encode_host: data_words too small for data_bits");
    return BCHFFT_ERR_ARG;
  }
  BCHFFT_CUDA_TRY(cudaSetDevice(c->f->device));
  cudaStream_t st = c->f->stream;
  const size_t cw_bytes = (size_t)c->wpc * 8, d_bytes = data_words * 8;
  size_t per = ((size_t)64 << 20) / (cw_bytes + d_bytes);
  if (per < 1) per = 1;
  if (per > batch) per = batch;
  bchfft_code::WorkSet &w = c->work[0];
  if ((rc = w.words.reserve(per * cw_bytes))) return rc;
  if ((rc = w.scratch.reserve(per * d_bytes))) return rc;
  for (size_t i0 = 0; i0 < batch; i0 += per) {
    const size_t cnt = batch - i0 < per ? batch - i0 : per;
    BCHFFT_CUDA_TRY(cudaMemcpyAsync(w.scratch.p, data + i0 * data_words, cnt * d_bytes, cudaMemcpyHostToDevice, st));
    if ((rc = encode_launch(c, (const uint64_t *)w.scratch.p, (uint32_t)data_words, data_bits,
                            (uint64_t *)w.words.p, cnt, st)))
      return rc;
    BCHFFT_CUDA_TRY(cudaMemcpyAsync(codewords + i0 * c->wpc, w.words.p, cnt * cw_bytes, cudaMemcpyDeviceToHost, st));
    BCHFFT_CUDA_TRY(cudaStreamSynchronize(st));
  }
  return BCHFFT_OK;
}

// ---- stage-level entry points ---------------------------------------------------------------
namespace bchfft {

template <int M>
static int syndrome_stage(bchfft_code *c, const uint64_t *words, size_t batch, uint32_t *syndromes,
                          int32_t *flags)
{
  cudaStream_t st = c->f->stream;
  const uint32_t pos_stride = c->d + 1u;
  const size_t cw_bytes = (size_t)c->wpc * 8;
  size_t per = ((size_t)64 << 20) / (cw_bytes + scratch_per_cw(c, pos_stride, M));
  per = per < 32 ? 32 : (per / 32) * 32;
  if (per > batch) per = batch;
  int rc;
  if ((rc = c->work[0].words.reserve(per * cw_bytes))) return rc;
  if ((rc = c->work[0].scratch.reserve(scratch_bytes(c, per, pos_stride, M)))) return rc;
  for (size_t i0 = 0; i0 < batch; i0 += per) {
    const size_t cnt = batch - i0 < per ? batch - i0 : per;
    Scratch s;
    carve(c, c->work[0].scratch.p, cnt, pos_stride, M, s);
    BCHFFT_CUDA_TRY(cudaMemsetAsync(c->work[0].scratch.p, 0, zero_bytes(c, cnt, M), st));
    BCHFFT_CUDA_TRY(cudaMemcpyAsync(c->work[0].words.p, words + i0 * c->wpc, cnt * cw_bytes,
                                    cudaMemcpyHostToDevice, st));
    if ((rc = launch_syndrome<M>(c, c->work[0], (const uint64_t *)c->work[0].words.p, cnt, s, st))) return rc;
    BCHFFT_CUDA_TRY(cudaMemcpyAsync(syndromes + i0 * c->d, s.syn, cnt * c->d * 4, cudaMemcpyDeviceToHost, st));
    if (flags) BCHFFT_CUDA_TRY(cudaMemcpyAsync(flags + i0, s.flag, cnt * 4, cudaMemcpyDeviceToHost, st));
    BCHFFT_CUDA_TRY(cudaStreamSynchronize(st));
  }
  return BCHFFT_OK;
}

template <int M>
static int elp_stage(bchfft_code *c, const uint32_t *syndromes, size_t batch, uint32_t *elp, uint32_t *L)
{
  cudaStream_t st = c->f->stream;
  const uint32_t pos_stride = c->d + 1u;
  size_t per = ((size_t)64 << 20) / scratch_per_cw(c, pos_stride, M);
  per = per < 32 ? 32 : (per / 32) * 32;
  if (per > batch) per = batch;
  int rc;
  if ((rc = c->work[0].scratch.reserve(scratch_bytes(c, per, pos_stride, M)))) return rc;
  std::vector<int32_t> ones(per, 1);
  for (size_t i0 = 0; i0 < batch; i0 += per) {
    const size_t cnt = batch - i0 < per ? batch - i0 : per;
    Scratch s;
    carve(c, c->work[0].scratch.p, cnt, pos_stride, M, s);
    BCHFFT_CUDA_TRY(cudaMemsetAsync(c->work[0].scratch.p, 0, zero_bytes(c, cnt, M), st));
    BCHFFT_CUDA_TRY(cudaMemcpyAsync(s.syn, syndromes + i0 * c->d, cnt * c->d * 4, cudaMemcpyHostToDevice, st));
    // bch_compute_elp runs unconditionally when called directly
    BCHFFT_CUDA_TRY(cudaMemcpyAsync(s.flag, ones.data(), cnt * 4, cudaMemcpyHostToDevice, st));
    if ((rc = launch_bm<M>(c, cnt, s, st))) return rc;
    BCHFFT_CUDA_TRY(cudaMemcpyAsync(elp + i0 * (c->d + 1), s.elp, cnt * (c->d + 1) * 4, cudaMemcpyDeviceToHost, st));
    BCHFFT_CUDA_TRY(cudaMemcpyAsync(L + i0, s.L, cnt * 4, cudaMemcpyDeviceToHost, st));
    BCHFFT_CUDA_TRY(cudaStreamSynchronize(st));
  }
  return BCHFFT_OK;
}

template <int M>
static int locate_stage(bchfft_code *c, const uint32_t *elp, const uint32_t *L, size_t batch,
                        uint32_t *positions, uint32_t *count)
{
  cudaStream_t st = c->f->stream;
  const uint32_t pos_stride = c->d + 1u;
  size_t per = ((size_t)64 << 20) / scratch_per_cw(c, pos_stride, M);
  per = per < 32 ? 32 : (per / 32) * 32;
  if (per > batch) per = batch;
  int rc;
  if ((rc = c->work[0].scratch.reserve(scratch_bytes(c, per, pos_stride, M)))) return rc;
  std::vector<int32_t> ones(per, 1);
  // the FFT is sized for the largest degree actually requested (bch.c:406: p_elp_degree)
  uint32_t lmax = 1;
  for (size_t i = 0; i < batch; i++)
    if (L[i] > lmax && L[i] <= c->d) lmax = L[i];
  for (size_t i0 = 0; i0 < batch; i0 += per) {
    const size_t cnt = batch - i0 < per ? batch - i0 : per;
    Scratch s;
    carve(c, c->work[0].scratch.p, cnt, pos_stride, M, s);
    BCHFFT_CUDA_TRY(cudaMemsetAsync(c->work[0].scratch.p, 0, zero_bytes(c, cnt, M), st));
    BCHFFT_CUDA_TRY(cudaMemcpyAsync(s.elp, elp + i0 * (c->d + 1), cnt * (c->d + 1) * 4, cudaMemcpyHostToDevice, st));
    BCHFFT_CUDA_TRY(cudaMemcpyAsync(s.L, L + i0, cnt * 4, cudaMemcpyHostToDevice, st));
    BCHFFT_CUDA_TRY(cudaMemcpyAsync(s.flag, ones.data(), cnt * 4, cudaMemcpyHostToDevice, st));
    if ((rc = launch_locate<M>(c, c->work[0], cnt, s, lmax, st))) return rc;
    BCHFFT_RUN(k_sort_positions, dim3((unsigned)((cnt + 127) / 128)), dim3(128), 0, st, s.positions,
                  s.pos_stride, (const uint32_t *)s.count, (uint32_t)cnt);
    BCHFFT_CUDA_TRY(cudaMemcpyAsync(positions + i0 * pos_stride, s.positions, cnt * pos_stride * 4,
                                    cudaMemcpyDeviceToHost, st));
    BCHFFT_CUDA_TRY(cudaMemcpyAsync(count + i0, s.count, cnt * 4, cudaMemcpyDeviceToHost, st));
    BCHFFT_CUDA_TRY(cudaStreamSynchronize(st));
  }
  return BCHFFT_OK;
}

}  // namespace bchfft

extern "C" int bchfft_bch_syndrome_host(bchfft_code *c, const uint64_t *words, size_t batch,
                                        uint32_t *syndromes, int32_t *flags)
{
  if (!c || !words || !syndromes) {
    set_error("bchfft_bch_syndrome_host: null argument");
    return BCHFFT_ERR_ARG;
  }
  if (batch == 0) return BCHFFT_OK;
  BCHFFT_CUDA_TRY(cudaSetDevice(c->f->device));
#define X(M) case M: return syndrome_stage<M>(c, words, batch, syndromes, flags);
  BCHFFT_DISPATCH_M(c->f->m, X)
#undef X
}

extern "C" int bchfft_bch_elp_host(bchfft_code *c, const uint32_t *syndromes, size_t batch,
                                   uint32_t *elp, uint32_t *L)
{
  if (!c || !syndromes || !elp || !L) {
    set_error("bchfft_bch_elp_host: null argument");
    return BCHFFT_ERR_ARG;
  }
  if (batch == 0) return BCHFFT_OK;
  BCHFFT_CUDA_TRY(cudaSetDevice(c->f->device));
#define X(M) case M: return elp_stage<M>(c, syndromes, batch, elp, L);
  BCHFFT_DISPATCH_M(c->f->m, X)
#undef X
}

extern "C" int bchfft_bch_locate_host(bchfft_code *c, const uint32_t *elp, const uint32_t *L,
                                      size_t batch, uint32_t *positions, uint32_t *count)
{
  if (!c || !elp || !L || !positions || !count) {
    set_error("bchfft_bch_locate_host: null argument");
    return BCHFFT_ERR_ARG;
  }
  if (batch == 0) return BCHFFT_OK;
  BCHFFT_CUDA_TRY(cudaSetDevice(c->f->device));
#define X(M) case M: return locate_stage<M>(c, elp, L, batch, positions, count);
  BCHFFT_DISPATCH_M(c->f->m, X)
#undef X
}

#ifdef BCHFFT_TRACE
// debug builds only: sweep trace of CTA (0,0) of the kernels in this translation unit
extern "C" int bchfft_debug_trace_bch(unsigned long long *out, int max_pairs)
{
  unsigned int n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, bchfft::g_trace_n, sizeof n);
  if ((int)n > max_pairs) n = max_pairs;
  cudaMemcpyFromSymbol(out, bchfft::g_trace, sizeof(unsigned long long) * 2 * n);
  unsigned int zero = 0;
  cudaMemcpyToSymbol(bchfft::g_trace_n, &zero, sizeof zero);
  return (int)n;
}
#endif
