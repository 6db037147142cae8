/*
 * libbch entry points (reference: include/bch.h, libbch/bch.c).
 *
 *  - code construction and systematic encoding stay on the host (one-time setup; the
 *    reference needs the external gf2x library for the polynomial products, here a plain
 *    carry-less product does it -- products in GF(2)[x] are unique, so g is identical);
 *  - every decode stage runs on the GPU through include/bchfft_cuda.h and leaves its result in
 *    the caller's host TbchCodeBuffers, like the reference's stages do.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bch.h"
#include "gf2_extension_polynomials.h"
#include "gf2_polynomials.h"
#include "gf2_polynomials_misc.h"
#include "host_internal.h"

#define WORD_BITS ((uint32_t) (8 * sizeof(unsigned long)))

/* ---- construction ------------------------------------------------------------------------- */

/* Cyclotomic-coset walk over the exponents 1 .. n-2 (reference: bch.c:143-220).  A coset is
 * kept when it contains an exponent in [p_offset, p_offset + p_numValues); its minimal
 * polynomial (degree < word size) is packed into one word of p_binaryPolynomials. */
void bch_explore_cycles(TfiniteField* p_f, uint32_t p_numValues, uint32_t p_offset,
    uint32_t* p_remainingPtr, unsigned long* p_binaryPolynomials, uint32_t* po_numPolys,
    uint32_t* po_totalDegree)
{
  const uint32_t order = p_f->n - 1;
  uint32_t* visited = p_remainingPtr;          /* caller-provided scratch of n words */
  uint32_t found = 0, total_degree = 0, npolys = 0;
  uint32_t coset_poly[g_maxDegree + 2];
  memset(visited, 0, sizeof(uint32_t) * p_f->n);
  for (uint32_t start = 1; start < order && found < p_numValues; start++)
  {
    if (visited[start]) continue;
    uint32_t degree = 0, e = start, tagged = 0;
    coset_poly[0] = 1;
    do
    {
      visited[e] = 1;
      if (e >= p_offset && e < p_offset + p_numValues) { tagged = 1; found++; }
      multiply_polynomial_by_monomial(p_f, coset_poly, degree++, e);
      e = (uint32_t) (((uint64_t) e * 2) % order);
    }
    while (e != start);
    if (!tagged) continue;
    if (p_binaryPolynomials)
    {
      unsigned long w = 0;
      for (uint32_t i = 0; i <= degree; i++) if (coset_poly[i] & 1u) w |= 1ul << i;
      p_binaryPolynomials[npolys] = w;
    }
    npolys++;
    total_degree += degree;
  }
  if (po_totalDegree) *po_totalDegree = total_degree;
  if (po_numPolys) *po_numPolys = npolys;
}

/* c <- a * b in GF(2)[x]; a has na words, b fits one word */
static void clmul_word(unsigned long* c, const unsigned long* a, size_t na, unsigned long b)
{
  memset(c, 0, (na + 1) * sizeof(unsigned long));
  while (b)
  {
    const int s = __builtin_ctzl(b);
    b &= b - 1;
    for (size_t k = 0; k < na; k++)
    {
      c[k] ^= a[k] << s;
      if (s) c[k + 1] ^= a[k] >> (WORD_BITS - s);
    }
  }
}

/* reference: bch.c:64-119 */
TbchCode bch_build_polynomial(TfiniteField* p_f, uint32_t p_minimumDistance)
{
  TbchCode code;
  code.m_field = *p_f;
  code.m_distance = p_minimumDistance;
  code.m_length = 0;
  code.m_gen_poly_degree = 0;
  unsigned long* minimal = (unsigned long*) calloc(p_minimumDistance + 1, sizeof(unsigned long));
  uint32_t* scratch = (uint32_t*) malloc(sizeof(uint32_t) * p_f->n);
  uint32_t npolys = 0;
  bch_explore_cycles(p_f, p_minimumDistance, 0, scratch, minimal, &npolys, &code.m_gen_poly_degree);
  free(scratch);
  const size_t words = code.m_gen_poly_degree / WORD_BITS + 2;
  unsigned long* g = (unsigned long*) calloc(words + 1, sizeof(unsigned long));
  unsigned long* tmp = (unsigned long*) calloc(words + 1, sizeof(unsigned long));
  g[0] = 1;
  size_t used = 1;
  for (uint32_t i = 0; i < npolys; i++)
  {
    clmul_word(tmp, g, used, minimal[i]);
    if (used + 1 <= words) used++;
    memcpy(g, tmp, used * sizeof(unsigned long));
  }
  const uint32_t packed = degree_to_packed_size_u32(code.m_gen_poly_degree);
  code.m_packed_gen_poly = (unsigned long*) calloc(packed, sizeof(unsigned long));
  memcpy(code.m_packed_gen_poly, g, packed * sizeof(unsigned long));
  free(tmp);
  free(g);
  free(minimal);
  return code;
}

/* reference: bch.c:30-54 */
TbchCode bch_gen_code(uint32_t p_codeword_size, uint32_t p_minimum_distance)
{
  TbchCode code;
  memset(&code, 0, sizeof code);
  if (p_minimum_distance >= p_codeword_size || p_codeword_size > (1u << g_maxDegree)) return code;
  uint32_t m = 2;
  while ((1u << m) < p_codeword_size) m++;
  /* Parameters the device decoder has no kernels for give the reference's failure value (no
   * generator) here, instead of a code whose first decode call would have to abort: fields above
   * GF(2^25) (the reference's own tables for m = 26..30 are not primitive) and
   * designed distances below 3 (nothing to correct). */
  if (m > BCHFFT_MAX_FIELD_DEGREE || p_minimum_distance < 3) return code;
  TfiniteField f = create_binary_finite_field(m);
  code = bch_build_polynomial(&f, p_minimum_distance);
  if (code.m_gen_poly_degree >= p_codeword_size) bch_free_code(&code);   /* no data bits left */
  else code.m_length = p_codeword_size;
  return code;
}

void bch_free_code(TbchCode* p_code)
{
  p_code->m_distance = 0;
  free_finite_field(&p_code->m_field);
  free(p_code->m_packed_gen_poly);
  p_code->m_packed_gen_poly = NULL;
}

/* exploratory utility of the reference (bch.c:121-140): generator degree for every offset */
void bch_find_best_offset(TfiniteField* p_f, uint32_t p_minimum_distance)
{
  uint32_t* scratch = (uint32_t*) malloc(sizeof(uint32_t) * p_f->n);
  uint32_t best = 0, best_degree = 0, degree = 0;
  for (uint32_t offset = 0; offset + 1 + p_minimum_distance < p_f->n; offset++)
  {
    bch_explore_cycles(p_f, p_minimum_distance, offset, scratch, NULL, NULL, &degree);
    printf("Offset %d, degree %d\n", offset, degree);
    if (offset == 0 || degree < best_degree) { best_degree = degree; best = offset; }
  }
  printf("Degree with offset %d: %d\n", best, best_degree);
  free(scratch);
}

/* ---- encoding ----------------------------------------------------------------------------- */

void bch_reduce_syndrome(TbchCode* p_code, unsigned long* pio_encoded_data)
{
  euclidean_division_packed_inplace_gf2(pio_encoded_data, p_code->m_length - 1,
                                        p_code->m_packed_gen_poly, p_code->m_gen_poly_degree, NULL);
}

/* reference: bch.c:492-506 */
int bch_encode(TbchCode* p_code, unsigned long* po_codeword, uint32_t* p_data, uint32_t p_data_size)
{
  const uint32_t shift = p_code->m_gen_poly_degree;
  if (p_data_size > p_code->m_length - shift) return 1;
  memset(po_codeword, 0, degree_to_packed_size(p_code->m_length - 1) * sizeof(unsigned long));
  for (uint32_t i = 0; i < p_data_size; i++) set_bit(po_codeword, i + shift, p_data[i]);
  bch_reduce_syndrome(p_code, po_codeword);           /* parity = data * X^deg(g) mod g */
  for (uint32_t i = 0; i < p_data_size; i++) set_bit(po_codeword, i + shift, p_data[i]);
  return 0;
}

/* ---- buffers ------------------------------------------------------------------------------ */

TbchCodeBuffers bch_alloc_buffers(TbchCode* p_code)
{
  TbchCodeBuffers b;
  const size_t big = (size_t) p_code->m_field.n * sizeof(uint32_t);
  const size_t small = (size_t) (p_code->m_distance + 1) * sizeof(uint32_t);
  b.m_syndrome = (uint32_t*) malloc(big);
  b.m_buffer_1 = (uint32_t*) malloc(big);
  b.m_buffer_2 = (uint32_t*) malloc(big);
  b.m_A = (uint32_t*) calloc(1, small);
  b.m_B = (uint32_t*) calloc(1, small);
  b.m_C_elp = (uint32_t*) calloc(1, small);
  b.m_error = (uint32_t*) calloc(1, small);
  return b;
}

void bch_free_buffers(TbchCodeBuffers* p_buffer)
{
  uint32_t** all[] = {&p_buffer->m_syndrome, &p_buffer->m_buffer_1, &p_buffer->m_buffer_2,
                      &p_buffer->m_A, &p_buffer->m_B, &p_buffer->m_C_elp, &p_buffer->m_error};
  for (size_t i = 0; i < sizeof all / sizeof all[0]; i++) { free(*all[i]); *all[i] = NULL; }
}

/* ---- decoding (GPU) ----------------------------------------------------------------------- */

/* single-item calls run on the first device of the set; returns with the slot lock held */
static bchfft_code* code_context(TbchCode* p_code, const char* who)
{
  int status = 0;
  bchfft_host_lock_slot(0);
  bchfft_code* ctx = bchfft_host_code(p_code, 0, &status);
  if (!ctx) { bchfft_host_unlock_slot(0); bchfft_host_die(who, status); }
  return ctx;
}

/* reference: bch.c:228-326.  syndrome[0 .. d-1] into p_buffer->m_syndrome */
int32_t bch_eval_syndrome(TbchCode* p_code, TbchCodeBuffers* p_buffer, unsigned long* p_packed_syndrome)
{
  int32_t flag = 0;
  bchfft_code* ctx = code_context(p_code, "bch_eval_syndrome");
  const int status = bchfft_bch_syndrome_host(ctx, (const uint64_t*) p_packed_syndrome, 1,
                                              p_buffer->m_syndrome, &flag);
  bchfft_host_unlock_slot(0);
  if (status) bchfft_host_die("bch_eval_syndrome", status);
  return flag;
}

/* reference: bch.c:328-404.  C[0 .. L] into p_buffer->m_C_elp, returns L */
uint32_t bch_compute_elp(TbchCode* p_code, TbchCodeBuffers* p_buffer)
{
  uint32_t L = 0;
  bchfft_code* ctx = code_context(p_code, "bch_compute_elp");
  const int status = bchfft_bch_elp_host(ctx, p_buffer->m_syndrome, 1, p_buffer->m_C_elp, &L);
  bchfft_host_unlock_slot(0);
  if (status) bchfft_host_die("bch_compute_elp", status);
  return L;
}

/* reference: bch.c:406-490.  positions (descending) into p_buffer->m_error, returns the count */
uint32_t bch_locate_errors(TbchCode* p_code, TbchCodeBuffers* p_buffer, uint32_t p_elp_degree)
{
  uint32_t count = 0;
  bchfft_code* ctx = code_context(p_code, "bch_locate_errors");
  const int status = bchfft_bch_locate_host(ctx, p_buffer->m_C_elp, &p_elp_degree, 1,
                                            p_buffer->m_error, &count);
  bchfft_host_unlock_slot(0);
  if (status) bchfft_host_die("bch_locate_errors", status);
  return count;
}

/* reference: bch.c:508-562 */
uint32_t bch_decode(TbchCode* p_code, unsigned long* pio_packed_data, TbchCodeBuffers* p_buffer,
                    uint32_t* po_corrected)
{
  (void) p_buffer;   /* scratch lives on the device */
  unsigned char ok = 0;
  uint32_t corrected = 0;
  bchfft_code* ctx = code_context(p_code, "bch_decode");
  const int status = bchfft_bch_decode_host(ctx, (uint64_t*) pio_packed_data, 1, &ok, &corrected);
  bchfft_host_unlock_slot(0);
  if (status) bchfft_host_die("bch_decode", status);
  if (po_corrected) *po_corrected = corrected;
  return ok;
}

/* ---- batch entry points: one contiguous range of the caller's arrays per device ------------- */
typedef struct
{
  TbchCode* code;
  unsigned long* words;
  unsigned char* ok;
  uint32_t* corrected;
  size_t wpc;
} decode_batch_args;

static int decode_shard(int slot, size_t lo, size_t hi, void* p)
{
  decode_batch_args* a = (decode_batch_args*) p;
  int status = 0;
  bchfft_host_lock_slot(slot);
  bchfft_code* ctx = bchfft_host_code(a->code, slot, &status);
  if (ctx)
    status = bchfft_bch_decode_host(ctx, (uint64_t*) (a->words + lo * a->wpc), hi - lo,
                                    a->ok ? a->ok + lo : NULL, a->corrected ? a->corrected + lo : NULL);
  bchfft_host_unlock_slot(slot);
  return status;
}

int bch_decode_batch(TbchCode* p_code, unsigned long* pio_words, size_t p_count,
                     unsigned char* po_ok, uint32_t* po_corrected)
{
  decode_batch_args a = {p_code, pio_words, po_ok, po_corrected,
                         degree_to_packed_size(p_code->m_length - 1)};
  /* a device is worth a thread of its own from about 16 MiB of codewords on */
  size_t min_per = ((size_t) 16 << 20) / (a.wpc * sizeof(unsigned long));
  if (min_per < 1024) min_per = 1024;
  return bchfft_host_run_sharded(p_count, min_per, decode_shard, &a);
}

typedef struct
{
  TbchCode* code;
  unsigned long* codewords;
  const unsigned long* data;
  size_t data_words, wpc;
  uint32_t data_bits;
} encode_batch_args;

static int encode_shard(int slot, size_t lo, size_t hi, void* p)
{
  encode_batch_args* a = (encode_batch_args*) p;
  int status = 0;
  bchfft_host_lock_slot(slot);
  bchfft_code* ctx = bchfft_host_code(a->code, slot, &status);
  if (ctx)
    status = bchfft_bch_encode_host(ctx, (const uint64_t*) (a->data + lo * a->data_words), a->data_words,
                                    a->data_bits, (uint64_t*) (a->codewords + lo * a->wpc), hi - lo);
  bchfft_host_unlock_slot(slot);
  return status;
}

int bch_encode_batch(TbchCode* p_code, unsigned long* po_codewords, const unsigned long* p_data,
                     size_t p_data_words, uint32_t p_data_bits, size_t p_count)
{
  encode_batch_args a = {p_code, po_codewords, p_data, p_data_words,
                         degree_to_packed_size(p_code->m_length - 1), p_data_bits};
  size_t min_per = ((size_t) 16 << 20) / (a.wpc * sizeof(unsigned long));
  if (min_per < 1024) min_per = 1024;
  return bchfft_host_run_sharded(p_count, min_per, encode_shard, &a);
}
