/* Binary BCH codes: construction, systematic encoding (host) and decoding (CUDA).
 * Same types and entry points as the reference's include/bch.h:10-211. */
#pragma once
#include <stddef.h>
#include "finite_field.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct _TbchCode
{
  TfiniteField m_field;
  uint32_t m_distance;               /* designed distance d = 2t+1 */
  unsigned long* m_packed_gen_poly;  /* generator, packed */
  uint32_t m_gen_poly_degree;
  uint32_t m_length;                 /* codeword length in bits */
} TbchCode;

typedef struct _TbchCodeBuffers
{
  uint32_t* m_syndrome;  /* n words */
  uint32_t* m_buffer_1;  /* n words */
  uint32_t* m_buffer_2;  /* n words */
  uint32_t* m_A;         /* d+1 words each: Berlekamp-Massey work polynomials */
  uint32_t* m_B;
  uint32_t* m_C_elp;     /* error-locator polynomial */
  uint32_t* m_error;     /* error positions */
} TbchCodeBuffers;

/* m_packed_gen_poly == NULL on failure */
TbchCode bch_gen_code(uint32_t p_codeword_size, uint32_t p_minimum_distance);
TbchCode bch_build_polynomial(TfiniteField* p_f, uint32_t p_minimumDistance);
void bch_explore_cycles(TfiniteField *p_f, uint32_t p_minimumDistance, uint32_t p_offset,
    uint32_t* p_remainingPtr, unsigned long *p_binaryPolynomials, uint32_t *po_numPolys,
    uint32_t *po_totalDegree);
void bch_find_best_offset(TfiniteField* p_f, uint32_t p_minimum_distance);
void bch_reduce_syndrome(TbchCode *p_code, unsigned long* pio_encoded_data);
/* data bit i -> codeword bit i + deg g; returns 1 when there are too many data bits */
int bch_encode(TbchCode *p_code, unsigned long* po_codeword, uint32_t* p_data, uint32_t p_data_size);

/* decode stages; each leaves its result in the caller's host buffers */
int32_t bch_eval_syndrome(TbchCode* p_code, TbchCodeBuffers *p_buffer, unsigned long* p_packed_syndrome);
uint32_t bch_compute_elp(TbchCode* p_code, TbchCodeBuffers* p_buffer);
uint32_t bch_locate_errors(TbchCode* p_code, TbchCodeBuffers *p_buffer, uint32_t p_elp_degree);
/* 1 = clean or corrected in place, 0 = uncorrectable (word untouched) */
uint32_t bch_decode(TbchCode* p_code, unsigned long* pio_packed_data,
    TbchCodeBuffers* p_buffer, uint32_t *po_corrected);

TbchCodeBuffers bch_alloc_buffers(TbchCode* p_code);
void bch_free_code(TbchCode *p_code);
void bch_free_buffers(TbchCodeBuffers* p_buffer);

/* Batch extension: p_count codewords of degree_to_packed_size(m_length-1) words each, decoded in
 * place; po_ok[i] / po_corrected[i] = what bch_decode returns for word i.  Returns 0 or a
 * negative bchfft status. */
int bch_decode_batch(TbchCode* p_code, unsigned long* pio_words, size_t p_count,
                     unsigned char* po_ok, uint32_t* po_corrected);

/* Batched systematic encoder: word i of po_codewords ([p_count][W], W = words per codeword) is what
 * bch_encode writes for the data bits of item i, given packed: p_data is [p_count][p_data_words],
 * bit k of an item = bit k % 64 of word k / 64, p_data_bits bits per item.  Returns 0, or a negative
 * bchfft status (too many data bits, generator degree outside what the device encoder supports). */
int bch_encode_batch(TbchCode* p_code, unsigned long* po_codewords, const unsigned long* p_data,
                     size_t p_data_words, uint32_t p_data_bits, size_t p_count);

#ifdef __cplusplus
}
#endif
