/*
 * libfft entry points (reference: include/additive_fft.h, include/multiplicative_fft.h) on top
 * of the C-ABI CUDA layer.  Every transform runs on the GPU; the only host-side arithmetic left
 * here is the reference's naive DFT_direct test helper and the factor tables.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "additive_fft.h"
#include "multiplicative_fft.h"
#include "host_internal.h"

/* ---- additive FFT ----------------------------------------------------------------------- */

static int additive_run(TfiniteField* p_f, const uint32_t* polys, uint32_t* results, uint32_t* natural,
                        uint32_t num_terms, size_t count)
{
  int status = 0;
  bchfft_host_lock();
  bchfft_field* ctx = bchfft_host_field(p_f, &status);
  if (ctx)
    status = bchfft_additive_fft_host(ctx, polys, p_f->n, results, p_f->n, natural, p_f->n,
                                      num_terms, count);
  bchfft_host_unlock();
  return status;
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

static int multiplicative_run(TfiniteField* p_f, const uint32_t* polys, uint32_t* evals,
                              uint32_t num_coeffs, size_t count)
{
  int status = 0;
  bchfft_host_lock();
  bchfft_field* ctx = bchfft_host_field(p_f, &status);
  if (ctx)
    status = bchfft_multiplicative_fft_host(ctx, polys, p_f->n, evals, p_f->n, num_coeffs, count);
  bchfft_host_unlock();
  return status;
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
