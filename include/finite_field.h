/* GF(2^m) in polynomial basis with log/exp tables -- same types and symbols as the
 * reference's include/finite_field.h:13-109 so that callers compile unchanged.
 * Host-side setup only; the kernels do field arithmetic bitsliced (bch_fft_b200/csrc). */
#pragma once
#include <stdint.h>
#include <malloc.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#define g_maxDegree 30

/* element = m-bit integer, bit i = coefficient of x^i; generator x = 2;
 * exp[i] = x^i, log[x^i] = i, log[0] = logzero = n-1, exp[logzero] = 0 */
typedef struct
{
  uint32_t m;        /* extension degree  */
  uint32_t n;        /* field size 2^m    */
  uint32_t logzero;  /* n - 1             */
  uint32_t* exp;
  uint32_t* log;
} TfiniteField;

/* exp = log = NULL when p_degree is outside [2, g_maxDegree] */
TfiniteField create_binary_finite_field(uint32_t p_degree);
void free_finite_field(TfiniteField * p_f);

uint32_t inverse(TfiniteField * p_f, uint32_t p_a);                      /* 0 -> 0 */
uint32_t square(TfiniteField * p_f, uint32_t p_a);
uint32_t multiply(TfiniteField * p_f, uint32_t p_a, uint32_t p_b);
uint32_t multiply_by_log(TfiniteField * p_f, uint32_t p_a, uint32_t p_b_log); /* b = x^p_b_log != 0 */

#ifdef __cplusplus
}
#endif
