/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement ("port") of the reference algorithms on the
 * hot path of kunzjacq/bch_fft, used as the parity checker by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
 * Nothing under bch_fft_b200/ or host/ may include, link or call this.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below against
 *   (1) the known answers of SURVEY.md Appendix C (produced by the unmodified reference),
 *   (2) oracle/_ref/libref_oracle.so = the reference's own sources compiled in place from
 *       /root/reference by oracle/Makefile (when that file is present), and
 *   (3) the committed fixtures under tests/golden/ (generated from oracle/_ref by
 *       tests/golden/make_golden.py).
 */
#ifndef BCHFFT_ORACLE_H
#define BCHFFT_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  uint32_t m;      /* extension degree                        */
  uint32_t n;      /* 2^m                                      */
  uint32_t order;  /* 2^m - 1 (also the "log of zero" marker)  */
  uint32_t *exp;   /* exp[i] = x^i, exp[order] = 0             */
  uint32_t *log;   /* log[x^i] = i, log[0] = order             */
} orc_field;

typedef struct {
  orc_field f;
  uint32_t distance;   /* designed distance d = 2t+1 */
  uint32_t gen_degree; /* deg g                      */
  uint32_t length;     /* codeword length in bits    */
  uint64_t *gen;       /* packed generator, bit i of word i/64 = coeff of X^i */
} orc_code;

/* field (libbase/finite_field.c:5-93) */
uint32_t orc_field_poly(uint32_t m);               /* reduction polynomial incl. x^m; 0 if unsupported */
int      orc_field_init(orc_field *f, uint32_t m); /* 0 ok, 1 unsupported m */
void     orc_field_free(orc_field *f);
uint32_t orc_mul(const orc_field *f, uint32_t a, uint32_t b);
uint32_t orc_inv(const orc_field *f, uint32_t a);

/* additive FFT (libfft/additive_fft.c:13-178) */
void orc_subspace_polys(const orc_field *f, uint32_t *coef /* m*(m+1) */, uint32_t *y /* m */);
void orc_additive_dft(const orc_field *f, uint32_t *poly /* n, in place */, uint32_t degree);
void orc_eval_additive_fft(const orc_field *f, uint32_t *poly /* n, clobbered */,
                           uint32_t *result /* n-1 */, uint32_t num_terms);

/* multiplicative FFT (libfft/multiplicative_fft.c:8-324) */
int orc_mult_factors(uint32_t m, uint32_t *out /* >=8 */);   /* returns #factors, 0 if unknown */
int orc_multiplicative_fft(const orc_field *f, uint32_t *data /* order, in place */,
                           uint32_t *scratch /* order */);
int orc_eval_multiplicative_fft(const orc_field *f, const uint32_t *poly, uint32_t *eval,
                                uint32_t num_coeffs, uint32_t *scratch);

/* BCH (libbch/bch.c) */
int      orc_bch_gen(orc_code *c, uint32_t length, uint32_t distance);  /* 0 ok */
void     orc_bch_free(orc_code *c);
size_t   orc_bch_words(const orc_code *c);                               /* 64-bit words per codeword */
int      orc_bch_encode(const orc_code *c, uint64_t *codeword, const uint32_t *bits, uint32_t nbits);
/* syndrome[0..d-1]; *used_fft (may be NULL) reports which reference method the rule picked */
int32_t  orc_bch_syndrome(const orc_code *c, const uint64_t *word, uint32_t *syndrome, int *used_fft);
/* Berlekamp-Massey on syndrome[1..d-1]; bufs = 3*(d+1) zero-initialised u32; *elp points into bufs */
uint32_t orc_bch_elp(const orc_code *c, const uint32_t *syndrome, uint32_t *bufs, uint32_t **elp);
/* roots of elp[0..L] -> positions (order - i), descending; returns count */
uint32_t orc_bch_locate(const orc_code *c, const uint32_t *elp, uint32_t L, uint32_t *positions);
uint32_t orc_bch_decode(const orc_code *c, uint64_t *word, uint32_t *corrected);
/* convenience: decode `count` words of orc_bch_words() each; ok[i] in {0,1} */
void     orc_bch_decode_batch(const orc_code *c, uint64_t *words, size_t count,
                              uint8_t *ok, uint32_t *corrected);

uint64_t orc_fnv64(const void *data, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif
