/*
 * Host-side base layer: field tables, packed GF(2)[x] helpers, sparse-polynomial helpers and
 * harness utilities.  These are the one-time setup / utility symbols of the reference's
 * libbase (libbase/finite_field.c, gf2_polynomials*.c, gf2_extension_polynomials.c, utils.c),
 * written independently; nothing here is on the hot path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "finite_field.h"
#include "gf2_extension_polynomials.h"
#include "gf2_polynomials.h"
#include "gf2_polynomials_misc.h"
#include "utils.h"

/* ---- GF(2^m) -------------------------------------------------------------------------- */

/* reduction polynomials (x^m term included) equal to the reference's table
 * (libbase/finite_field.c:9-18 decoded), index m - 2 */
static const uint32_t reduction_poly[29] = {
  0x7,        0xB,        0x13,       0x25,       0x43,       0x83,       0x171,
  0x211,      0x409,      0x805,      0x1099,     0x201B,     0x5803,     0x8003,
  0x1002D,    0x20009,    0x40081,    0x80063,    0x100009,   0x200005,   0x400003,
  0x800021,   0x100001B,  0x2000009,  0x400001B,  0x8000003,  0x10000005, 0x20000021,
  0x40000003};

TfiniteField create_binary_finite_field(uint32_t p_degree)
{
  TfiniteField f;
  f.m = p_degree;
  f.n = 1u << p_degree;
  f.logzero = f.n - 1;
  f.exp = NULL;
  f.log = NULL;
  if (p_degree < 2 || p_degree > g_maxDegree) return f;
  const uint32_t poly = reduction_poly[p_degree - 2];
  f.exp = (uint32_t*) calloc(f.n, sizeof(uint32_t));
  f.log = (uint32_t*) calloc(f.n, sizeof(uint32_t));
  uint32_t v = 1;
  for (uint32_t i = 0; i + 1 < f.n; i++)
  {
    f.exp[i] = v;
    f.log[v] = i;
    v <<= 1;
    if (v >> p_degree) v ^= poly;
  }
  f.exp[f.logzero] = 0;
  f.log[0] = f.logzero;
  f.log[1] = 0;
  return f;
}

void free_finite_field(TfiniteField* p_f)
{
  free(p_f->exp);
  free(p_f->log);
  p_f->exp = NULL;
  p_f->log = NULL;
}

static inline uint32_t log_add(const TfiniteField* f, uint32_t la, uint32_t lb)
{
  const uint32_t s = la + lb, order = f->n - 1;
  return s >= order ? s - order : s;
}

uint32_t multiply(TfiniteField* p_f, uint32_t p_a, uint32_t p_b)
{
  if (!p_a || !p_b) return 0;
  return p_f->exp[log_add(p_f, p_f->log[p_a], p_f->log[p_b])];
}

uint32_t multiply_by_log(TfiniteField* p_f, uint32_t p_a, uint32_t p_b_log)
{
  if (!p_a) return 0;
  return p_f->exp[log_add(p_f, p_f->log[p_a], p_b_log)];
}

uint32_t square(TfiniteField* p_f, uint32_t p_a)
{
  if (!p_a) return 0;
  return p_f->exp[log_add(p_f, p_f->log[p_a], p_f->log[p_a])];
}

uint32_t inverse(TfiniteField* p_f, uint32_t p_a)
{
  if (p_a < 2) return p_a;
  return p_f->exp[p_f->n - 1 - p_f->log[p_a]];
}

/* ---- packed GF(2)[x] ------------------------------------------------------------------- */

#define WORD_BITS ((uint32_t) (8 * sizeof(unsigned long)))

uint32_t word_size() { return WORD_BITS; }

size_t wordsize_to_max_degree(uint32_t p_word_size) { return (size_t) p_word_size * WORD_BITS - 1; }

size_t degree_to_packed_size(size_t p_degree)
{
  return p_degree == -1u ? 0 : p_degree / WORD_BITS + 1;
}

uint32_t degree_to_packed_size_u32(uint32_t p_degree)
{
  return p_degree == -1u ? 0 : p_degree / WORD_BITS + 1;
}

void degree_to_packed_pos(size_t p_degree, size_t* po_word_idx, uint32_t* po_word_offset)
{
  *po_word_idx = p_degree / WORD_BITS;
  if (po_word_offset) *po_word_offset = (uint32_t) (p_degree % WORD_BITS);
}

void degree_to_packed_pos_u32(uint32_t p_degree, uint32_t* po_word_idx, uint32_t* po_word_offset)
{
  *po_word_idx = p_degree / WORD_BITS;
  if (po_word_offset) *po_word_offset = p_degree % WORD_BITS;
}

size_t degree_packed(unsigned long* p_polynomial, size_t p_max_degree)
{
  if (p_max_degree == -1u) return -1u;
  size_t w = p_max_degree / WORD_BITS;
  while (w > 0 && p_polynomial[w] == 0) w--;
  if (p_polynomial[w] == 0) return -1u;
  return w * WORD_BITS + (WORD_BITS - 1 - (uint32_t) __builtin_clzl(p_polynomial[w]));
}

uint32_t degree_packed_u32(unsigned long* p_polynomial, uint32_t p_max_degree)
{
  return (uint32_t) degree_packed(p_polynomial, (size_t) p_max_degree);
}

void set_bit(unsigned long* p_polynomial, size_t p_degree, uint32_t p_val)
{
  const unsigned long bit = 1ul << (p_degree % WORD_BITS);
  if (p_val) p_polynomial[p_degree / WORD_BITS] |= bit;
  else       p_polynomial[p_degree / WORD_BITS] &= ~bit;
}

uint32_t get_bit(unsigned long* p_polynomial, size_t p_degree)
{
  return (uint32_t) ((p_polynomial[p_degree / WORD_BITS] >> (p_degree % WORD_BITS)) & 1ul);
}

void add_bit(unsigned long* p_polynomial, size_t p_degree)
{
  p_polynomial[p_degree / WORD_BITS] ^= 1ul << (p_degree % WORD_BITS);
}

/* bit-serial long division from the top coefficient down (gf2_polynomials.c:10-70) */
void euclidean_division_packed_inplace_gf2(
    unsigned long* pio_reductee, uint32_t p_reductee_degree,
    unsigned long* p_reductor, uint32_t p_reductor_degree,
    unsigned long* po_dividend)
{
  if (po_dividend)
    memset(po_dividend, 0,
           degree_to_packed_size(p_reductee_degree - p_reductor_degree) * sizeof(unsigned long));
  if (p_reductor_degree > p_reductee_degree) return;
  const uint32_t rw = p_reductor_degree / WORD_BITS + 1;
  const uint32_t tw = p_reductee_degree / WORD_BITS + 1;
  for (uint32_t i = p_reductee_degree; i + 1 > p_reductor_degree; i--)
  {
    if (get_bit(pio_reductee, i))
    {
      const uint32_t shift = i - p_reductor_degree, w0 = shift / WORD_BITS, s = shift % WORD_BITS;
      if (po_dividend) po_dividend[w0] |= 1ul << s;
      for (uint32_t k = 0; k < rw; k++)
      {
        pio_reductee[w0 + k] ^= p_reductor[k] << s;
        if (s && w0 + k + 1 < tw) pio_reductee[w0 + k + 1] ^= p_reductor[k] >> (WORD_BITS - s);
      }
    }
    if (i == 0) break;
  }
}

/* ---- sparse polynomials over GF(2^m) ---------------------------------------------------- */

static uint32_t field_pow(TfiniteField* f, uint32_t a, uint32_t e)
{
  if (e == 0) return 1;
  if (a == 0) return 0;
  const uint64_t l = ((uint64_t) f->log[a] * e) % (f->n - 1);
  return f->exp[l];
}

uint32_t evaluate_sparse_polynomial(TfiniteField* p_f, uint32_t* p_coeffs, uint32_t* p_degrees,
                                    uint32_t p_num_monomials, uint32_t p_y)
{
  uint32_t acc = 0;
  if (p_y == 0) return (p_num_monomials > 0 && p_degrees[0] == 0) ? p_coeffs[0] : 0;
  for (uint32_t j = 0; j < p_num_monomials; j++)
    acc ^= multiply(p_f, p_coeffs[j], field_pow(p_f, p_y, p_degrees[j]));
  return acc;
}

uint32_t evaluate_sparse_polynomial_ref(TfiniteField* p_f, uint32_t* p_coeffs, uint32_t* p_degrees,
                                        uint32_t p_num_monomials, uint32_t p_y)
{
  /* table-light variant: repeated squaring on the degree gaps */
  if (p_y == 0) return (p_num_monomials > 0 && p_degrees[0] == 0) ? p_coeffs[0] : 0;
  uint32_t acc = 0, cur = 1, cur_deg = 0;
  for (uint32_t j = 0; j < p_num_monomials; j++)
  {
    if (!p_coeffs[j]) continue;
    uint32_t gap = p_degrees[j] - cur_deg, base = p_y, part = 1;
    while (gap)
    {
      if (gap & 1) part = multiply(p_f, part, base);
      base = square(p_f, base);
      gap >>= 1;
    }
    cur = multiply(p_f, cur, part);
    cur_deg = p_degrees[j];
    acc ^= multiply(p_f, cur, p_coeffs[j]);
  }
  return acc;
}

uint32_t evaluate_sparse_polynomial_log(TfiniteField* p_f, uint32_t* p_log_coeffs,
                                        uint32_t* p_degrees, uint32_t num_coeffs, uint32_t p_y)
{
  const uint32_t order = p_f->n - 1;
  if (p_y == 0)
    return (num_coeffs > 0 && p_degrees[0] == 0 && p_log_coeffs[0] != p_f->logzero)
           ? p_f->exp[p_log_coeffs[0]] : 0;
  uint32_t acc = 0;
  const uint64_t ly = p_f->log[p_y];
  for (uint32_t j = 0; j < num_coeffs; j++)
    if (p_log_coeffs[j] != p_f->logzero)
      acc ^= p_f->exp[(p_log_coeffs[j] + ly * p_degrees[j]) % order];
  return acc;
}

void multiply_polynomial_by_monomial(TfiniteField* p_f, uint32_t* pio_poly, uint32_t p_degree,
                                     uint32_t p_nonzero_root_log)
{
  pio_poly[p_degree + 1] = pio_poly[p_degree];
  for (uint32_t i = p_degree; i > 0; i--)
    pio_poly[i] = pio_poly[i - 1] ^ multiply_by_log(p_f, pio_poly[i], p_nonzero_root_log);
  pio_poly[0] = multiply_by_log(p_f, pio_poly[0], p_nonzero_root_log);
}

/* schoolbook division by a monic sparse polynomial whose last term is the leading one */
static void sparse_divide(TfiniteField* f, const uint32_t* reductee, uint32_t reductee_degree,
                          uint32_t* rc, uint32_t* rd, uint32_t nterms, uint32_t* rem, uint32_t* quo,
                          uint32_t factor)
{
  const uint32_t lead = rd[nterms - 1];
  uint32_t* work = (uint32_t*) malloc(sizeof(uint32_t) * (reductee_degree + 1));
  memcpy(work, reductee, sizeof(uint32_t) * (reductee_degree + 1));
  for (uint32_t k = 0; k < lead; k++) quo[k] = 0;
  for (uint32_t d = reductee_degree; d + 1 > lead; d--)
  {
    const uint32_t c = work[d];
    if (c)
    {
      if (d - lead < lead) quo[d - lead] = multiply(f, c, factor);
      for (uint32_t k = 0; k + 1 < nterms; k++) work[d - lead + rd[k]] ^= multiply(f, c, rc[k]);
    }
    if (d == 0) break;
  }
  for (uint32_t k = 0; k < lead; k++) rem[k] = k <= reductee_degree ? work[k] : 0;
  free(work);
}

void euclidean_division_sparse_reductor(TfiniteField* p_f, const uint32_t *p_reductee,
    uint32_t p_reductee_degree, uint32_t* p_reductor_coeffs, uint32_t* p_reductor_degrees,
    uint32_t p_reductor_num_terms, uint32_t* po_remainder, uint32_t* po_dividend,
    uint32_t p_dividend_factor_log)
{
  sparse_divide(p_f, p_reductee, p_reductee_degree, p_reductor_coeffs, p_reductor_degrees,
                p_reductor_num_terms, po_remainder, po_dividend, p_f->exp[p_dividend_factor_log]);
}

void euclidean_division_sparse_reductor_ref(TfiniteField* p_f, const uint32_t *p_reductee,
    uint32_t p_reductee_degree, uint32_t* p_reductor_coeffs, uint32_t* p_reductor_degrees,
    uint32_t p_reductor_num_terms, uint32_t* po_remainder, uint32_t* po_dividend,
    uint32_t p_dividend_factor)
{
  sparse_divide(p_f, p_reductee, p_reductee_degree, p_reductor_coeffs, p_reductor_degrees,
                p_reductor_num_terms, po_remainder, po_dividend, p_dividend_factor);
}

/* ---- harness utilities ------------------------------------------------------------------ */

static long time_origin_sec = 0;

void init_time()
{
  struct timeval tv;
  gettimeofday(&tv, NULL);
  time_origin_sec = tv.tv_sec;
}

double absolute_time()
{
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return (double) (tv.tv_sec - time_origin_sec) + 1e-6 * (double) tv.tv_usec;
}

uint32_t urand() { return (uint32_t) rand(); }

int compare_results(uint32_t* p_dataA, uint32_t* p_dataB, uint32_t p_groupSize, uint32_t szLog)
{
  int mismatches = 0;
  uint32_t sumA = 0, sumB = 0;
  for (uint32_t i = 0; i < p_groupSize; i++)
  {
    sumA += p_dataA[i];
    sumB += p_dataB[i];
    mismatches += p_dataA[i] != p_dataB[i];
  }
  if (mismatches)
  {
    const int digits = (szLog >= 4 && szLog <= 32) ? (int) ((szLog + 3) / 4) : 4;
    const uint32_t show = p_groupSize > 20 ? 10 : p_groupSize;
    printf("************* %d errors ********************\n", mismatches);
    for (int side = 0; side < 2; side++)
    {
      const uint32_t* v = side ? p_dataB : p_dataA;
      for (uint32_t i = 0; i < show; i++) printf("%0*x ", digits, v[i]);
      if (p_groupSize > 20)
      {
        printf("... ");
        for (uint32_t i = p_groupSize - show; i < p_groupSize; i++) printf("%0*x ", digits, v[i]);
      }
      printf("\n");
    }
    printf("checksums: %8x / %8x\n", sumA, sumB);
  }
  return mismatches;
}
