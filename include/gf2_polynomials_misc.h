/* Packed GF(2)[x] helpers: coefficient i lives in bit (i mod W) of word i / W, W = bits of
 * unsigned long.  Same symbols as the reference's include/gf2_polynomials_misc.h:17-46. */
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned char binary_coeff;

uint32_t word_size();
size_t wordsize_to_max_degree(uint32_t p_word_size);
size_t degree_packed(unsigned long* p_polynomial, size_t p_max_degree);     /* -1u for zero */
uint32_t degree_packed_u32(unsigned long* p_polynomial, uint32_t p_max_degree);
size_t degree_to_packed_size(size_t p_degree);                              /* words needed */
uint32_t degree_to_packed_size_u32(uint32_t p_degree);
void degree_to_packed_pos(size_t p_degree, size_t* po_word_idx, uint32_t* po_word_offset);
void degree_to_packed_pos_u32(uint32_t p_degree, uint32_t* po_word_idx, uint32_t* po_word_offset);
void set_bit(unsigned long* p_polynomial, size_t p_degree, uint32_t p_val);
uint32_t get_bit(unsigned long* p_polynomial, size_t p_degree);
void add_bit(unsigned long* p_polynomial, size_t p_degree);                 /* coefficient += 1 */

#ifdef __cplusplus
}
#endif
