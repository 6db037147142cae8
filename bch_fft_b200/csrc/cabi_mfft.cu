// Batched multiplicative FFT driver (see include/bchfft_cuda.h, bchfft_multiplicative_fft_*).
#include <algorithm>

#include "cabi_common.h"
#include "mfft_kernels.cuh"
#include "mfft16_kernels.cuh"

namespace bchfft {

struct MfftState {
  std::vector<uint32_t> factors;   // prime factors of 2^m - 1, descending
  uint32_t *d_perm_id = nullptr;   // identity permutation
  uint32_t *d_perm_zero = nullptr; // all zeros (constant polynomial)
  // GF(2^16) fast path (mfft16_kernels.cuh): coset table k_s 2^i mod 257 and the normal-basis
  // coordinates of the 257th roots of unity
  uint32_t *d_cos16 = nullptr, *d_ctab = nullptr;
  bool fast16 = false;
};

// tables of the cyclotomic 257-point transform; false when x^kNormalLog does not generate a normal basis
// (cannot happen for the reference's GF(2^16); checked rather than assumed)
static bool mfft16_tables(const bchfft_field *f, std::vector<uint32_t> &cos16, std::vector<uint32_t> &ctab)
{
  const uint32_t N = f->order;
  const uint32_t *ex = f->h_exp.data(), *lg = f->h_log.data();
  uint32_t conj[16];
  conj[0] = ex[mf16::kNormalLog];
  for (int b = 1; b < 16; b++) conj[b] = ex[(2u * lg[conj[b - 1]]) % N];
  // echelon form of the conjugates, each row with the combination of conjugates that produced it
  uint32_t ech_v[16] = {0}, ech_c[16] = {0};
  for (int b = 0; b < 16; b++) {
    uint32_t v = conj[b], comb = 1u << b;
    for (int bit = 15; bit >= 0 && v; bit--) {
      if (!((v >> bit) & 1u)) continue;
      if (ech_v[bit]) {
        v ^= ech_v[bit];
        comb ^= ech_c[bit];
      } else {
        ech_v[bit] = v;
        ech_c[bit] = comb;
        v = 0;
      }
    }
  }
  for (int bit = 0; bit < 16; bit++)
    if (!ech_v[bit]) return false;
  ctab.assign(257, 0u);
  for (uint32_t e = 0; e < 257u; e++) {
    uint32_t z = ex[(255u * e) % N], out = 0u;
    for (int bit = 15; bit >= 0; bit--)
      if ((z >> bit) & 1u) {
        z ^= ech_v[bit];
        out ^= ech_c[bit];
      }
    ctab[e] = out;
  }
  cos16.assign(256, 0u);
  std::vector<bool> seen(257, false);
  uint32_t s = 0;
  for (uint32_t k = 1; k < 257u && s < 16u; k++) {
    if (seen[k]) continue;
    uint32_t v = k;
    for (uint32_t i = 0; i < 16u; i++) {
      cos16[s * 16u + i] = v;
      seen[v] = true;
      v = (2u * v) % 257u;
    }
    s++;
  }
  return s == 16u;
}

// prime factorisations of 2^m - 1 for m = 8..26 (the reference's table,
// libfft/multiplicative_fft.c:11-33); other m: unknown -> the reference returns 1
static bool known_factors(uint32_t m, std::vector<uint32_t> &out)
{
  static const uint32_t tab[] = {
    3, 5, 17, 0,  7, 73, 0,  3, 11, 31, 0,  23, 89, 0,  3, 3, 5, 7, 13, 0,  8191, 0,  3, 43, 127, 0,
    7, 31, 151, 0,  3, 5, 17, 257, 0,  131071, 0,  3, 3, 3, 7, 19, 73, 0,  524287, 0,
    3, 5, 5, 11, 31, 41, 0,  7, 7, 127, 337, 0,  3, 23, 89, 683, 0,  47, 178481, 0,
    3, 3, 5, 7, 13, 17, 241, 0,  31, 601, 1801, 0,  3, 2731, 8191, 0};
  if (m < 8 || m > 26) return false;
  size_t k = 0;
  for (uint32_t cur = 8; cur < m; cur++) {
    while (tab[k]) k++;
    k++;
  }
  out.clear();
  while (tab[k]) out.push_back(tab[k++]);
  std::sort(out.begin(), out.end(), [](uint32_t a, uint32_t b) { return a > b; });
  return true;
}

template <int M>
static int mfft_run(bchfft_field *f, const uint32_t *d_polys, size_t poly_stride, uint32_t *d_evals,
                    size_t eval_stride, uint32_t num_coeffs, size_t batch, cudaStream_t stream)
{
  constexpr int PS = plane_stride(M);
  const uint32_t n = f->n, N = f->order;
  if (!f->mfft) {
    MfftState *st = new MfftState();
    if (!known_factors(f->m, st->factors)) {
      delete st;
      set_error("multiplicative FFT: factorisation of 2^%u - 1 unknown", f->m);
      return BCHFFT_ERR_UNSUPPORTED;
    }
    std::vector<uint32_t> id(n), zero(n, 0u);
    for (uint32_t i = 0; i < n; i++) id[i] = i;
    if (cudaMalloc((void **)&st->d_perm_id, (size_t)n * 4) != cudaSuccess ||
        cudaMalloc((void **)&st->d_perm_zero, (size_t)n * 4) != cudaSuccess) {
      set_error("multiplicative FFT: cudaMalloc failed");
      delete st;
      return BCHFFT_ERR_CUDA;
    }
    BCHFFT_CUDA_TRY(cudaMemcpy(st->d_perm_id, id.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    BCHFFT_CUDA_TRY(cudaMemcpy(st->d_perm_zero, zero.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    if (M == 16) {
      std::vector<uint32_t> cos16, ctab;
      if (mfft16_tables(f, cos16, ctab) &&
          cudaMalloc((void **)&st->d_cos16, cos16.size() * 4) == cudaSuccess &&
          cudaMalloc((void **)&st->d_ctab, ctab.size() * 4 + 16) == cudaSuccess) {
        BCHFFT_CUDA_TRY(cudaMemcpy(st->d_cos16, cos16.data(), cos16.size() * 4, cudaMemcpyHostToDevice));
        BCHFFT_CUDA_TRY(cudaMemcpy(st->d_ctab, ctab.data(), ctab.size() * 4, cudaMemcpyHostToDevice));
        st->fast16 = true;
      }
    }
    BCHFFT_CUDA_TRY(cudaDeviceSynchronize());   // tables in place before a non-blocking stream reads them
    f->mfft = st;
  }
  MfftState *st = (MfftState *)f->mfft;
  const uint32_t ncoef = num_coeffs < N ? num_coeffs : N;
  // levels that are not pure broadcasts
  struct Level { uint32_t S_prev, p, step; };
  std::vector<Level> levels;
  {
    uint32_t S = 1;
    for (uint32_t p : st->factors) {
      const uint32_t S_next = S * p;
      if (ncoef > N / S_next) levels.push_back(Level{S, p, N / S_next});
      S = S_next;
    }
  }
  const size_t slab_words = (size_t)n * PS;
  const size_t nslabs = (batch + 31) / 32;
  // GF(2^16) with more coefficients than the 3 x 5 x 17 stages alone cover: Good-Thomas 255 x 257 with the
  // cyclotomic 257-point transform (mfft16_kernels.cuh)
  const char *fenv = getenv("BCHFFT_MFFT_FAST");
  const bool fast16 = M == 16 && st->fast16 && ncoef > 255u && !(fenv && atoi(fenv) == 0);
  size_t chunk = f->chunk_bytes / (2 * slab_words * 4);
  if (chunk < 1) chunk = 1;
  if (chunk > nslabs) chunk = nslabs;
  if (chunk > 65535) chunk = 65535;
  int rc = f->ws.reserve(chunk * 2 * slab_words * 4);
  if (rc) return rc;
  uint32_t *buf[2] = {(uint32_t *)f->ws.p, (uint32_t *)f->ws.p + chunk * slab_words};
  for (size_t s0 = 0; s0 < nslabs; s0 += chunk) {
    const size_t cs = nslabs - s0 < chunk ? nslabs - s0 : chunk;
    const uint32_t items = (uint32_t)((batch - s0 * 32 < cs * 32) ? batch - s0 * 32 : cs * 32);
    BCHFFT_RUN(k_bitslice_in<M>, dim3((n + 127) / 128, (unsigned)cs), dim3(128), 0, stream,
               d_polys + s0 * 32 * poly_stride, poly_stride, items, ncoef, f->m, buf[0]);
    if (fast16) {
      if constexpr (M == 16) {
        const size_t smem1 = sizeof(uint32_t) * (2 * 16 * mf16::kAS + 256 + 257 + 8);
        const size_t smem2 = sizeof(uint32_t) * (2 * 16 * mf16::kAS);
#ifdef BCHFFT_EMU
        const unsigned nt = 32;
#else
        const unsigned nt = 256;
#endif
        BCHFFT_ENSURE_SMEM(mf16::k_mfft16_p257, f->device, smem1);
        BCHFFT_ENSURE_SMEM(mf16::k_mfft16_p255, f->device, smem2);
        BCHFFT_RUN(mf16::k_mfft16_p257, dim3(255, (unsigned)cs), dim3(nt), smem1, stream, (const uint32_t *)buf[0],
                   buf[1], slab_words, (const uint32_t *)st->d_cos16, (const uint32_t *)st->d_ctab);
        BCHFFT_RUN(mf16::k_mfft16_p255, dim3(257, (unsigned)cs), dim3(nt), smem2, stream, (const uint32_t *)buf[1],
                   buf[0], slab_words);
        BCHFFT_RUN(k_gather_out<M>, dim3((N + 127) / 128, (unsigned)cs), dim3(128), 0, stream,
                   (const uint32_t *)buf[0], slab_words, (const uint32_t *)st->d_perm_id, N,
                   d_evals + s0 * 32 * eval_stride, eval_stride, items);
      }
      continue;
    }
    int cur = 0;
    for (size_t li = 0; li < levels.size(); li++) {
      const Level &lv = levels[li];
      BCHFFT_RUN(k_mfft_stage<M>, dim3((N + 255) / 256, (unsigned)cs), dim3(256), 0, stream,
                 (const uint32_t *)buf[cur], buf[cur ^ 1], slab_words, N, lv.S_prev, lv.p, lv.step,
                 (uint32_t)(li == 0), ncoef, (const uint32_t *)f->d_exp);
      cur ^= 1;
    }
    // no level left: the polynomial is the constant a_0 (multiplicative_fft.c:278-283)
    const uint32_t *perm = levels.empty() ? st->d_perm_zero : st->d_perm_id;
    BCHFFT_RUN(k_gather_out<M>, dim3((N + 127) / 128, (unsigned)cs), dim3(128), 0, stream,
               (const uint32_t *)buf[cur], slab_words, perm, N, d_evals + s0 * 32 * eval_stride,
               eval_stride, items);
  }
  return BCHFFT_OK;
}

}  // namespace bchfft

using namespace bchfft;

extern "C" void bchfft_mfft_release(bchfft_field *f)
{
  if (!f || !f->mfft) return;
  MfftState *st = (MfftState *)f->mfft;
  cudaFree(st->d_perm_id);
  cudaFree(st->d_perm_zero);
  cudaFree(st->d_cos16);
  cudaFree(st->d_ctab);
  delete st;
  f->mfft = nullptr;
}

extern "C" int bchfft_multiplicative_fft_dev(bchfft_field *f, const uint32_t *d_polys, size_t poly_stride,
                                             uint32_t *d_evals, size_t eval_stride, uint32_t num_coeffs,
                                             size_t batch, void *stream)
{
  if (!f || !d_polys || !d_evals) {
    set_error("bchfft_multiplicative_fft_dev: null argument");
    return BCHFFT_ERR_ARG;
  }
  if (poly_stride < f->order || eval_stride < f->order) {
    set_error("bchfft_multiplicative_fft_dev: stride smaller than 2^m - 1");
    return BCHFFT_ERR_ARG;
  }
  if (batch == 0) return BCHFFT_OK;
  BCHFFT_CUDA_TRY(cudaSetDevice(f->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : f->stream;
  switch (f->m) {
#define X(M) case M: return mfft_run<M>(f, d_polys, poly_stride, d_evals, eval_stride, num_coeffs, batch, st);
    BCHFFT_M_LIST(X)
#undef X
  }
  set_error("bchfft_multiplicative_fft_dev: unsupported field degree %u", f->m);
  return BCHFFT_ERR_UNSUPPORTED;
}

extern "C" int bchfft_multiplicative_fft_host(bchfft_field *f, const uint32_t *polys, size_t poly_stride,
                                              uint32_t *evals, size_t eval_stride, uint32_t num_coeffs,
                                              size_t batch)
{
  if (!f || !polys || !evals) {
    set_error("bchfft_multiplicative_fft_host: null argument");
    return BCHFFT_ERR_ARG;
  }
  if (batch == 0) return BCHFFT_OK;
  BCHFFT_CUDA_TRY(cudaSetDevice(f->device));
  const uint32_t n = f->n, N = f->order;
  size_t per = ((size_t)256 << 20) / ((size_t)n * 4);
  per = per < 32 ? 32 : (per / 32) * 32;
  if (per > batch) per = batch;
  int rc;
  if ((rc = f->stage_in.reserve(per * (size_t)n * 4))) return rc;
  if ((rc = f->stage_out.reserve(per * (size_t)n * 4))) return rc;
  for (size_t i0 = 0; i0 < batch; i0 += per) {
    const size_t cnt = batch - i0 < per ? batch - i0 : per;
    uint32_t *d_in = (uint32_t *)f->stage_in.p, *d_out = (uint32_t *)f->stage_out.p;
    // rows may be only N = n-1 words long (the reference's buffers hold `order` entries)
    BCHFFT_CUDA_TRY(cudaMemcpy2DAsync(d_in, (size_t)n * 4, polys + i0 * poly_stride, poly_stride * 4,
                                      (size_t)N * 4, cnt, cudaMemcpyHostToDevice, f->stream));
    rc = bchfft_multiplicative_fft_dev(f, d_in, n, d_out, n, num_coeffs, cnt, f->stream);
    if (rc) return rc;
    BCHFFT_CUDA_TRY(cudaMemcpy2DAsync(evals + i0 * eval_stride, eval_stride * 4, d_out, (size_t)n * 4,
                                      (size_t)N * 4, cnt, cudaMemcpyDeviceToHost, f->stream));
    BCHFFT_CUDA_TRY(cudaStreamSynchronize(f->stream));
  }
  return BCHFFT_OK;
}
