/*
 * libfft entry points (reference: include/additive_fft.h, include/multiplicative_fft.h) on top
 * of the C-ABI CUDA layer.  Every transform runs on the GPU; the only host-side arithmetic left
 * here is the reference's host-only helpers (naive DFT_direct, the staged CPU building blocks of
 * the multiplicative transform, the subspace-polynomial self-check) and the factor tables.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "additive_fft.h"
#include "multiplicative_fft.h"
#include "host_internal.h"

/* ---- additive FFT ----------------------------------------------------------------------- */

typedef struct
{
  TfiniteField* f;
  const uint32_t* polys;
  uint32_t* results;
  uint32_t* natural;
  uint32_t num_terms;
  int mult;              /* 1: multiplicative transform, num_terms = number of coefficients */
} fft_batch_args;

static int fft_shard(int slot, size_t lo, size_t hi, void* p)
{
  fft_batch_args* a = (fft_batch_args*) p;
  const size_t n = a->f->n;
  int status = 0;
  bchfft_host_lock_slot(slot);
  bchfft_field* ctx = bchfft_host_field(a->f, slot, &status);
  if (ctx && a->mult)
    status = bchfft_multiplicative_fft_host(ctx, a->polys + lo * n, n, a->results + lo * n, n, a->num_terms, hi - lo);
  else if (ctx)
    status = bchfft_additive_fft_host(ctx, a->polys + lo * n, n, a->results ? a->results + lo * n : NULL, n,
                                      a->natural ? a->natural + lo * n : NULL, n, a->num_terms, hi - lo);
  bchfft_host_unlock_slot(slot);
  return status;
}

/* one contiguous range of polynomials per device (a device is worth its own thread from about 32 MiB
 * of coefficients on); single transforms run on the first device */
static size_t fft_min_per_device(const TfiniteField* p_f)
{
  size_t min_per = ((size_t) 32 << 20) / ((size_t) p_f->n * sizeof(uint32_t));
  return min_per < 32 ? 32 : min_per;
}

static int additive_run(TfiniteField* p_f, const uint32_t* polys, uint32_t* results, uint32_t* natural,
                        uint32_t num_terms, size_t count)
{
  fft_batch_args a = {p_f, polys, results, natural, num_terms, 0};
  return bchfft_host_run_sharded(count, fft_min_per_device(p_f), fft_shard, &a);
}

/* replaces libfft/additive_fft.c:13-21 */
void evaluate_polynomial_additive_FFT(TfiniteField* p_f, uint32_t* p_poly, uint32_t* po_result,
                                      uint32_t p_num_terms)
{
  /* p_poly is both the input and, like in the reference, the natural-order output */
  const int status = additive_run(p_f, p_poly, po_result, p_poly, p_num_terms, 1);
  if (status) bchfft_host_die("evaluate_polynomial_additive_FFT", status);
}

/* replaces libfft/additive_fft.c:23-178 (and the two equivalent variants :264-484) */
void additive_DFT_opt(TfiniteField* p_f, uint32_t* p_poly, uint32_t p_poly_degree)
{
  const int status = additive_run(p_f, p_poly, NULL, p_poly, p_poly_degree, 1);
  if (status) bchfft_host_die("additive_DFT", status);
}

void additive_DFT(TfiniteField* p_f, uint32_t* p_poly, uint32_t p_poly_degree)
{
  additive_DFT_opt(p_f, p_poly, p_poly_degree);
}

void additive_DFT_ref(TfiniteField* p_f, uint32_t* p_poly, uint32_t p_poly_degree)
{
  additive_DFT_opt(p_f, p_poly, p_poly_degree);
}

int additive_fft_batch(TfiniteField* p_f, const uint32_t* p_polys, uint32_t* po_results,
                       uint32_t p_num_terms, size_t p_count)
{
  return additive_run(p_f, p_polys, po_results, NULL, p_num_terms, p_count);
}

/* Host-only self-check of a table of subspace polynomials (libfft/additive_fft.c:181-262; the
 * reference calls it from additive_DFT / additive_DFT_ref only): row i of `div` ((m+1) entries)
 * holds the coefficients of P_i on the monomials X^degrees[0..i+1].  Checks that P_i vanishes
 * exactly on {0 .. 2^i - 1} and that P_i(k) = P_(i-1)(k) P_(i-1)(k ^ 2^(i-1)).  Returns 1 when
 * every row passes, and reports the failing rows on stdout. */
int check_fft_polynomials(TfiniteField* p_f, uint32_t* div, uint32_t* degrees, uint32_t* y)
{
  (void) y;
  const uint32_t m = p_f->m, n = p_f->n;
  uint32_t* prev = (uint32_t*) calloc(n, sizeof(uint32_t));
  uint32_t* cur = (uint32_t*) calloc(n, sizeof(uint32_t));
  if (!prev || !cur) { free(prev); free(cur); return 0; }
  int all_ok = 1;
  for (uint32_t i = 0; i < m; i++)
  {
    uint32_t* t = prev; prev = cur; cur = t;
    for (uint32_t k = 0; k < n; k++)
      cur[k] = evaluate_sparse_polynomial(p_f, div + (size_t) (m + 1) * i, degrees, i + 2, k);
    uint32_t bad_roots = 0, bad_products = 0;
    for (uint32_t k = 0; k < n; k++)
    {
      if ((cur[k] == 0) != (k < (1u << i))) bad_roots++;
      if (i && multiply(p_f, prev[k], prev[k ^ (1u << (i - 1))]) != cur[k]) bad_products++;
    }
    if (bad_roots) printf("nok: Polynomial P_%u has wrong roots (%u points)\n", i, bad_roots);
    if (bad_products)
      printf("nok: P_%u != P_%u(X) x P_%u(X-x_%u) (%u discrepancies)\n", i, i - 1, i - 1, i - 1,
             bad_products);
    /* the reference tests P_0's roots but leaves them out of its return value */
    if (i && (bad_roots || bad_products)) all_ok = 0;
  }
  free(prev);
  free(cur);
  return all_ok;
}

/* ---- multiplicative FFT ----------------------------------------------------------------- */

/* prime factors of 2^m - 1, m = 8 .. 26 (libfft/multiplicative_fft.c:11-33) */
static uint32_t factor_table[] = {
  3, 5, 17,  7, 73,  3, 11, 31,  23, 89,  3, 3, 5, 7, 13,  8191,  3, 43, 127,  7, 31, 151,
  3, 5, 17, 257,  131071,  3, 3, 3, 7, 19, 73,  524287,  3, 5, 5, 11, 31, 41,  7, 7, 127, 337,
  3, 23, 89, 683,  47, 178481,  3, 3, 5, 7, 13, 17, 241,  31, 601, 1801,  3, 2731, 8191};
static const uint32_t factor_start[] = {
  0, 3, 5, 8, 10, 15, 16, 19, 22, 26, 27, 33, 34, 40, 44, 48, 50, 57, 60, 63};
#define MIN_LOG 8u
#define MAX_LOG 26u

uint32_t factorOffset(uint32_t p_log) { return factor_start[p_log - MIN_LOG]; }

uint32_t numFactors(uint32_t m)
{
  return (m >= MIN_LOG && m <= MAX_LOG) ? factorOffset(m + 1) - factorOffset(m) : 1;
}

uint32_t* factors(uint32_t m)
{
  return (m >= MIN_LOG && m <= MAX_LOG) ? factor_table + factorOffset(m) : NULL;
}

uint32_t getFactorSum(TfiniteField* p_f)
{
  if (p_f->m < MIN_LOG || p_f->m > MAX_LOG) return p_f->n - 1;
  uint32_t sum = 0;
  for (uint32_t i = 0; i < numFactors(p_f->m); i++) sum += factors(p_f->m)[i];
  return sum;
}

/* host-only O(n^2) definition of the transform (libfft/multiplicative_fft.c:84-104); the
 * reference uses it as a test oracle for small fields, and so do its demos */
void DFT_direct(TfiniteField* p_f, uint32_t* pio_data, uint32_t* p_buffer)
{
  const uint32_t order = p_f->n - 1;
  for (uint32_t i = 0; i < order; i++)
  {
    uint32_t acc = 0;
    for (uint32_t j = 0; j < order; j++)
      if (pio_data[j])
        acc ^= p_f->exp[(uint32_t) ((p_f->log[pio_data[j]] + (uint64_t) i * j) % order)];
    p_buffer[i] = acc;
  }
  memcpy(pio_data, p_buffer, order * sizeof(uint32_t));
}

/* ---- host-only building blocks of the reference's CPU transform --------------------------
 * Exported by the reference's libfft (`nm`: T) and declared in its header, so a program that
 * drives the stages by hand still links.  They are not on the device path:
 * multiplicative_FFT() above runs k_mfft_stage. */

/* `count` interleaved direct DFTs of size N = order / p_local_exponent on log-form input
 * (libfft/multiplicative_fft.c:124-173): po_buffer[s + count i] = sum_j x^(d[s + count j] +
 * p_local_exponent i j), entries equal to logzero contribute nothing.  The reference walks the
 * logs forward by j p_local_exponent per output row, which returns them to their initial
 * values after N rows; here the exponent is formed per term, p_logdata is only read. */
void multiplicative_FFT_compute_step(TfiniteField *p_f, uint32_t* p_logdata, uint32_t* po_buffer,
                                     uint32_t p_exponent, uint32_t p_initial_exponent,
                                     uint32_t p_stride_glob)
{
  const uint32_t order = p_f->n - 1;
  const uint32_t count = p_stride_glob * (p_exponent / p_initial_exponent);
  const uint32_t N = order / p_exponent;
  for (uint32_t i = 0; i < N; i++)
    for (uint32_t s = 0; s < count; s++)
    {
      uint32_t acc = 0;
      uint64_t twist = 0;                           /* p_exponent i j mod order */
      const uint64_t step = ((uint64_t) p_exponent * i) % order;
      for (uint32_t j = 0; j < N; j++)
      {
        const uint32_t d = p_logdata[s + count * j];
        if (d != p_f->logzero) acc ^= p_f->exp[(d + twist) % order];
        twist += step;
        if (twist >= order) twist -= order;
      }
      po_buffer[s + count * i] = acc;
    }
}

/* Back to log form with the twiddle of the next stage applied and the sub-transforms regrouped
 * (libfft/multiplicative_fft.c:175-231): element (row i, column s) with s = s_int group + s_ext
 * goes to s_ext + group (i + s_int N) and is multiplied by x^(p_stride_ext p_initial_exponent
 * s_int i). */
void multiplicative_FFT_reorder_step(TfiniteField* p_f, uint32_t* p_buffer, uint32_t* po_logdata,
                                     uint32_t p_local_exponent, uint32_t p_initial_exponent,
                                     uint32_t p_stride_ext, uint32_t p_stride_glob)
{
  const uint32_t order = p_f->n - 1;
  const uint32_t count = p_stride_glob * (p_local_exponent / p_initial_exponent);
  const uint32_t N = order / p_local_exponent;
  const uint32_t group = p_stride_ext * p_stride_glob;
  const uint64_t unit = (uint64_t) p_stride_ext * p_initial_exponent;
  for (uint32_t i = 0; i < N; i++)
    for (uint32_t s = 0; s < count; s++)
    {
      const uint32_t s_int = s / group, s_ext = s - s_int * group;
      uint32_t lg = p_f->log[p_buffer[(size_t) i * count + s]];
      if (lg != p_f->logzero) lg = (uint32_t) ((lg + unit * s_int * i) % order);
      po_logdata[s_ext + (size_t) group * (i + (size_t) s_int * N)] = lg;
    }
}

/* One compute step per prime factor, a reorder step between two of them; the last compute
 * step's output is the result (libfft/multiplicative_fft.c:233-257). */
void multiplicative_FFT_step_chaining(TfiniteField* p_f, uint32_t* pio_data, uint32_t* p_buffer,
                                      const uint32_t *p_factors, uint32_t p_num_factors,
                                      uint32_t order, uint32_t e)
{
  uint32_t done = 1;                                /* product of the factors already processed */
  for (uint32_t i = 0; i < p_num_factors; i++)
  {
    const uint32_t local = order / p_factors[i];
    multiplicative_FFT_compute_step(p_f, pio_data, p_buffer, local, e, e);
    if (i + 1 < p_num_factors)
      multiplicative_FFT_reorder_step(p_f, p_buffer, pio_data, local, e, done, e);
    else
      memcpy(pio_data, p_buffer, (size_t) (p_f->n - 1) * sizeof(uint32_t));
    done *= p_factors[i];
  }
}

static int multiplicative_run(TfiniteField* p_f, const uint32_t* polys, uint32_t* evals,
                              uint32_t num_coeffs, size_t count)
{
  fft_batch_args a = {p_f, polys, evals, NULL, num_coeffs, 1};
  return bchfft_host_run_sharded(count, fft_min_per_device(p_f), fft_shard, &a);
}

/* replaces libfft/multiplicative_fft.c:106-122 */
int multiplicative_FFT(TfiniteField* p_f, uint32_t* pio_data, uint32_t* p_buffer)
{
  (void) p_buffer;
  if (!factors(p_f->m)) return 1;
  const int status = multiplicative_run(p_f, pio_data, pio_data, p_f->n - 1, 1);
  if (status == BCHFFT_ERR_UNSUPPORTED) return 1;
  if (status) bchfft_host_die("multiplicative_FFT", status);
  return 0;
}

/* replaces libfft/multiplicative_fft.c:259-324 */
int evaluate_polynomial_multiplicative_FFT(TfiniteField* p_f, uint32_t* p_polynomial,
                                           uint32_t* p_eval, uint32_t p_numCoeffs, uint32_t* p_buffer)
{
  (void) p_buffer;
  if (p_numCoeffs > 1 && !factors(p_f->m)) return 1;
  const int status = multiplicative_run(p_f, p_polynomial, p_eval, p_numCoeffs, 1);
  if (status == BCHFFT_ERR_UNSUPPORTED) return 1;
  if (status) bchfft_host_die("evaluate_polynomial_multiplicative_FFT", status);
  return 0;
}

int multiplicative_fft_batch(TfiniteField* p_f, const uint32_t* p_polys, uint32_t* po_evals,
                             uint32_t p_numCoeffs, size_t p_count)
{
  return multiplicative_run(p_f, p_polys, po_evals, p_numCoeffs, p_count);
}
