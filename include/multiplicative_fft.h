/* Multiplicative (mixed-radix) FFT over the cyclic group of GF(2^m).  Same entry points as
 * the reference's include/multiplicative_fft.h:11-190. */
#pragma once
#include <stdint.h>
#include <stddef.h>
#include "finite_field.h"

#ifdef __cplusplus
extern "C" {
#endif

/* prime factors of 2^m - 1 (ascending) for m in [8, 26]; NULL / 1 outside */
uint32_t numFactors(uint32_t m);
uint32_t* factors(uint32_t m);
uint32_t factorOffset(uint32_t p_log);
uint32_t getFactorSum(TfiniteField* p_f);

/* naive O(n^2) transform, host only (test helper) */
void DFT_direct(TfiniteField* p_f, uint32_t* pio_data, uint32_t* p_buffer);

/* pio_data[i] <- P(x^i), i < n-1; returns 1 when the factorisation of 2^m-1 is unknown */
int multiplicative_FFT(TfiniteField* p_f, uint32_t* pio_data, uint32_t* p_buffer);

/* evaluation of a polynomial with p_numCoeffs coefficients (skips the FFT stages a low-degree
 * input makes unnecessary); p_polynomial == p_eval is allowed */
int evaluate_polynomial_multiplicative_FFT(TfiniteField* p_f, uint32_t* p_polynomial,
                                           uint32_t* p_eval, uint32_t p_numCoeffs,
                                           uint32_t* p_buffer);

/* host-only building blocks of the reference's CPU implementation, kept for link compatibility */
void multiplicative_FFT_compute_step(TfiniteField *p_f, uint32_t* p_logdata, uint32_t* po_buffer,
                                     uint32_t p_exponent, uint32_t p_initial_exponent,
                                     uint32_t p_stride_glob);
void multiplicative_FFT_reorder_step(TfiniteField* p_f, uint32_t* p_buffer, uint32_t* po_logdata,
                                     uint32_t p_local_exponent, uint32_t p_initial_exponent,
                                     uint32_t p_stride_ext, uint32_t p_stride_glob);
void multiplicative_FFT_step_chaining(TfiniteField* p_f, uint32_t* pio_data, uint32_t* p_buffer,
                                      const uint32_t *p_factors, uint32_t p_num_factors,
                                      uint32_t order, uint32_t e);

/* Batch extension: p_count polynomials, rows of n words in, rows of n words out (n-1 written) */
int multiplicative_fft_batch(TfiniteField* p_f, const uint32_t* p_polys, uint32_t* po_evals,
                             uint32_t p_numCoeffs, size_t p_count);

#ifdef __cplusplus
}
#endif
