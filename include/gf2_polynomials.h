/* Division of packed binary polynomials (reference: include/gf2_polynomials.h:35-40). */
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* pio_reductee <- pio_reductee mod p_reductor (in place); quotient bits to po_dividend when
 * it is not NULL */
void euclidean_division_packed_inplace_gf2(
    unsigned long* pio_reductee, uint32_t p_reductee_degree,
    unsigned long* p_reductor, uint32_t p_reductor_degree,
    unsigned long* po_dividend);

#ifdef __cplusplus
}
#endif
