/* Sparse polynomials over GF(2^m) (reference: include/gf2_extension_polynomials.h:30-177).
 * A sparse polynomial is a pair of arrays (coefficients, degrees), degrees ascending. */
#pragma once
#include <stdint.h>
#include "finite_field.h"

#ifdef __cplusplus
extern "C" {
#endif

uint32_t evaluate_sparse_polynomial(TfiniteField* p_f, uint32_t* p_coeffs, uint32_t* p_degrees,
                                    uint32_t p_num_monomials, uint32_t p_y);
/* same value, computed without the 64-bit modular product */
uint32_t evaluate_sparse_polynomial_ref(TfiniteField* p_f, uint32_t* p_coeffs, uint32_t* p_degrees,
                                        uint32_t p_num_monomials, uint32_t p_y);
/* coefficients given as logs (logzero marks a zero coefficient) */
uint32_t evaluate_sparse_polynomial_log(TfiniteField* p_f, uint32_t* p_log_coeffs,
                                        uint32_t* p_degrees, uint32_t num_coeffs, uint32_t p_y);
/* pio_poly (degree p_degree) <- pio_poly * (X + x^p_nonzero_root_log) */
void multiply_polynomial_by_monomial(TfiniteField* p_f, uint32_t* pio_poly, uint32_t p_degree,
                                     uint32_t p_nonzero_root_log);
/* dense reductee / monic sparse reductor; po_dividend receives quotient * x^p_dividend_factor_log */
void euclidean_division_sparse_reductor(TfiniteField* p_f, const uint32_t *p_reductee,
    uint32_t p_reductee_degree, uint32_t* p_reductor_coeffs, uint32_t* p_reductor_degrees,
    uint32_t p_reductor_num_terms, uint32_t* po_remainder, uint32_t* po_dividend,
    uint32_t p_dividend_factor_log);
/* same with the quotient scaled by the field element p_dividend_factor */
void euclidean_division_sparse_reductor_ref(TfiniteField* p_f, const uint32_t *p_reductee,
    uint32_t p_reductee_degree, uint32_t* p_reductor_coeffs, uint32_t* p_reductor_degrees,
    uint32_t p_reductor_num_terms, uint32_t* po_remainder, uint32_t* po_dividend,
    uint32_t p_dividend_factor);

#ifdef __cplusplus
}
#endif
