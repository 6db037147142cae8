/* Additive FFT over GF(2^m): evaluation of a polynomial at every field element.
 * Same entry points as the reference's include/additive_fft.h:8-55; the work runs in CUDA
 * kernels (bch_fft_b200/csrc) reached through include/bchfft_cuda.h.  There is no CPU path:
 * a CUDA failure prints the error and aborts (these functions return void in the reference). */
#pragma once
#include <stddef.h>
#include "finite_field.h"
#include "gf2_extension_polynomials.h"

#ifdef __cplusplus
extern "C" {
#endif

/* po_result[i] = P(x^i), i = 0 .. n-2; coefficients p_poly[0 .. p_num_terms] take part;
 * p_poly is clobbered with P(k), k = 0 .. n-1 (natural order); buffers must be distinct. */
void evaluate_polynomial_additive_FFT(TfiniteField* p_f, uint32_t* p_poly, uint32_t* po_result,
                                      uint32_t p_num_terms);

/* in place: p_poly[k] <- P(k).  The three names are one device path (the reference keeps three
 * host implementations of the same transform). */
void additive_DFT_opt(TfiniteField* p_f, uint32_t* p_poly, uint32_t p_poly_degree);
void additive_DFT(TfiniteField* p_f, uint32_t* p_poly, uint32_t p_poly_degree);
void additive_DFT_ref(TfiniteField* p_f, uint32_t* p_poly, uint32_t p_poly_degree);

/* Batch extension (not in the reference): p_count polynomials of n coefficients each, rows
 * p_polys[i*n ..]; po_results rows of n words (entries 0..n-2 written).  Inputs are left
 * untouched.  Returns 0, or a negative bchfft status (see bchfft_last_error()). */
int additive_fft_batch(TfiniteField* p_f, const uint32_t* p_polys, uint32_t* po_results,
                       uint32_t p_num_terms, size_t p_count);

/* host-only diagnostic exported by the reference's libfft (libfft/additive_fft.c:181; not in its
 * header): 1 when rows P_0 .. P_(m-1) of `div` (m+1 coefficients each, on X^degrees[]) are the
 * subspace polynomials of span{1, 2, .., 2^(i-1)} */
int check_fft_polynomials(TfiniteField* p_f, uint32_t* div, uint32_t* degrees, uint32_t* y);

#ifdef __cplusplus
}
#endif
