/*
 * TEST INFRASTRUCTURE ONLY -- see bchfft_oracle.h.  Parity status: PINNED (tests/test_oracle.py).
 *
 * Scalar, single-threaded CPU restatement of the algorithms the reference runs on the hot
 * path.  Every function names the reference lines it follows (paths relative to
 * /root/reference).  The code is written independently: same arithmetic, same index
 * conventions, same corner-case behaviour, different program text.
 */
#include "bchfft_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------ */
/* GF(2^m): log/exp tables.  Follows libbase/finite_field.c:5-93.                        */
/* ------------------------------------------------------------------------------------ */

/* Reduction polynomials decoded from the monomial table at libbase/finite_field.c:9-18
 * (SURVEY.md Appendix B).  Entry [m-2] includes the x^m and the constant term.  For m>=26
 * x is not primitive in the reference's table; kept for table parity, never used. */
static const uint32_t k_field_poly[29] = {
  0x7,        0xB,        0x13,       0x25,       0x43,       0x83,       0x171,
  0x211,      0x409,      0x805,      0x1099,     0x201B,     0x5803,     0x8003,
  0x1002D,    0x20009,    0x40081,    0x80063,    0x100009,   0x200005,   0x400003,
  0x800021,   0x100001B,  0x2000009,  0x400001B,  0x8000003,  0x10000005, 0x20000021,
  0x40000003};

uint32_t orc_field_poly(uint32_t m)
{
  return (m < 2 || m > 30) ? 0 : k_field_poly[m - 2];
}

int orc_field_init(orc_field *f, uint32_t m)
{
  memset(f, 0, sizeof *f);
  f->m = m;
  f->n = 1u << m;
  f->order = f->n - 1;
  const uint32_t poly = orc_field_poly(m);
  if (!poly) return 1;                               /* finite_field.c:26-27 */
  f->exp = (uint32_t *)calloc(f->n, sizeof(uint32_t));
  f->log = (uint32_t *)calloc(f->n, sizeof(uint32_t));
  uint32_t v = 1;
  for (uint32_t i = 0; i < f->order; i++) {          /* finite_field.c:36-44 */
    f->exp[i] = v;
    v <<= 1;
    if (v & f->n) v ^= poly;
  }
  f->exp[f->order] = 0;                              /* exp[logzero] = 0 */
  f->log[0] = f->order;                              /* log[0] = logzero */
  for (uint32_t i = 0; i < f->order; i++) f->log[f->exp[i]] = i;
  return 0;
}

void orc_field_free(orc_field *f)
{
  free(f->exp);
  free(f->log);
  f->exp = f->log = NULL;
}

/* finite_field.c:77-84 */
uint32_t orc_mul(const orc_field *f, uint32_t a, uint32_t b)
{
  if (!a || !b) return 0;
  uint32_t s = f->log[a] + f->log[b];
  if (s >= f->order) s -= f->order;
  return f->exp[s];
}

/* finite_field.c:59-66 (inverse(0) = 0 by convention) */
uint32_t orc_inv(const orc_field *f, uint32_t a)
{
  if (a < 2) return a;
  return f->exp[f->order - f->log[a]];
}

static uint32_t orc_sqr(const orc_field *f, uint32_t a) { return orc_mul(f, a, a); }

/* ------------------------------------------------------------------------------------ */
/* Additive FFT (von zur Gathen - Gerhard).  Follows libfft/additive_fft.c:13-178 and the */
/* linearized-polynomial evaluation of libbase/gf2_extension_polynomials.c:81-115.        */
/* ------------------------------------------------------------------------------------ */

/* Evaluate L(y) = c[0]*y + c[1]*y^2 + c[2]*y^4 + ... (nterms terms, no constant term). */
static uint32_t lin_eval(const orc_field *f, const uint32_t *c, uint32_t nterms, uint32_t y)
{
  uint32_t acc = 0, p = y;
  for (uint32_t j = 0; j < nterms; j++) {
    acc ^= orc_mul(f, c[j], p);
    p = orc_sqr(f, p);
  }
  return acc;
}

/* Subspace polynomials P_i(X) = prod_{w < 2^i} (X - w) = sum_{j<=i} coef[i][j] X^(2^j),
 * through P_i = P_{i-1}^2 + y_{i-1} P_{i-1}, y_i = P_i(2^i).  additive_fft.c:41-78.
 * coef is m rows of (m+1) entries, row i holds the coefficients of X^(2^j), j = 0..m. */
void orc_subspace_polys(const orc_field *f, uint32_t *coef, uint32_t *y)
{
  const uint32_t m = f->m, w = m + 1;
  memset(coef, 0, sizeof(uint32_t) * m * w);
  coef[0] = 1;                 /* P_0 = X        */
  y[0] = 1;                    /* P_0(2^0) = 1   */
  for (uint32_t i = 1; i < m; i++) {
    const uint32_t *prev = coef + (i - 1) * w;
    uint32_t *cur = coef + i * w;
    for (uint32_t j = 0; j < i; j++) cur[j] = orc_mul(f, prev[j], y[i - 1]);
    for (uint32_t j = 0; j < i; j++) cur[j + 1] ^= orc_sqr(f, prev[j]);
    y[i] = lin_eval(f, cur, i + 1, 1u << i);
  }
}

/* In-place transform: on return poly[k] = P(k) for every field element k (natural order).
 * Only coefficients 0..degree (inclusive) of the input take part.  additive_fft.c:23-178. */
void orc_additive_dft(const orc_field *f, uint32_t *poly, uint32_t degree)
{
  const uint32_t m = f->m, w = m + 1;
  uint32_t *coef = (uint32_t *)malloc(sizeof(uint32_t) * m * w);
  uint32_t *y = (uint32_t *)malloc(sizeof(uint32_t) * m);
  uint32_t *work = (uint32_t *)malloc(sizeof(uint32_t) * f->n);
  uint32_t *quot = (uint32_t *)malloc(sizeof(uint32_t) * f->n);
  orc_subspace_polys(f, coef, y);

  for (uint32_t stage = 0; stage < m; stage++) {                  /* additive_fft.c:85 */
    const uint32_t lvl = m - 1 - stage;                           /* reductor = P_lvl + const */
    const uint32_t half = 1u << lvl, span = half << 1, blocks = 1u << stage;
    const uint32_t *red = coef + lvl * w;                         /* lvl+1 linear terms; X^half is monic */
    const uint32_t v = y[lvl];                                    /* P_lvl(2^lvl), :96-97 */
    const uint32_t top = degree < span - 1 ? degree : span - 1;   /* :99 */

    if (degree > 0 && degree < half) {                            /* low-degree shortcut, :103-110 */
      for (uint32_t b = 0; b < blocks; b++)
        memcpy(poly + b * span + half, poly, sizeof(uint32_t) * half);
      continue;
    }
    for (uint32_t b = 0; b < blocks; b++) {                       /* :112 */
      uint32_t *blk = poly + b * span;
      /* constant term making the reductor vanish on the coset b*span + <1..2^(lvl-1)>, :117-120 */
      const uint32_t c0 = lin_eval(f, red, lvl + 1, b * span);
      /* schoolbook division of blk[0..top] by X^half + sum_{j<lvl} red[j] X^(2^j) + c0, :131-160 */
      memcpy(work, blk, sizeof(uint32_t) * (top + 1));
      memset(quot, 0, sizeof(uint32_t) * half);
      for (uint32_t d = top; d >= half && d != UINT32_MAX; d--) {
        const uint32_t lead = work[d];
        if (!lead) continue;
        const uint32_t base = d - half;
        quot[base] = lead;
        work[base] ^= orc_mul(f, lead, c0);
        for (uint32_t j = 0; j < lvl; j++) work[base + (1u << j)] ^= orc_mul(f, lead, red[j]);
      }
      /* remainder r in work[0..half); store r | r + v*q, :161-165.  Coefficients of blk
       * below `half` but above `top` are kept as they are (the reference xors the
       * accumulated remainder into the untouched low half). */
      for (uint32_t k = 0; k < half; k++) {
        const uint32_t r = (k <= top) ? work[k] : blk[k];
        blk[k] = r;
        blk[k + half] = r ^ orc_mul(f, v, quot[k]);
      }
    }
  }
  free(quot);
  free(work);
  free(y);
  free(coef);
}

/* additive_fft.c:13-21: transform, then gather in exp order: result[i] = P(x^i). */
void orc_eval_additive_fft(const orc_field *f, uint32_t *poly, uint32_t *result, uint32_t num_terms)
{
  orc_additive_dft(f, poly, num_terms);
  for (uint32_t i = 0; i < f->order; i++) result[i] = poly[f->exp[i]];
}

/* ------------------------------------------------------------------------------------ */
/* Multiplicative FFT (mixed radix over the cyclic group).  Follows                       */
/* libfft/multiplicative_fft.c:8-324.                                                     */
/* ------------------------------------------------------------------------------------ */

/* Prime factorisations of 2^m - 1 for m = 8..26, ascending (multiplicative_fft.c:11-33). */
static const uint16_t k_fact_cnt[19] = {3, 2, 3, 2, 5, 1, 3, 3, 4, 1, 6, 1, 6, 4, 4, 2, 7, 3, 3};
static const uint32_t k_fact[] = {
  3, 5, 17,   7, 73,   3, 11, 31,   23, 89,   3, 3, 5, 7, 13,   8191,   3, 43, 127,
  7, 31, 151,   3, 5, 17, 257,   131071,   3, 3, 3, 7, 19, 73,   524287,
  3, 5, 5, 11, 31, 41,   7, 7, 127, 337,   3, 23, 89, 683,   47, 178481,
  3, 3, 5, 7, 13, 17, 241,   31, 601, 1801,   3, 2731, 8191};

int orc_mult_factors(uint32_t m, uint32_t *out)
{
  if (m < 8 || m > 26) return 0;
  uint32_t off = 0;
  for (uint32_t i = 8; i < m; i++) off += k_fact_cnt[i - 8];
  const int cnt = k_fact_cnt[m - 8];
  for (int i = 0; i < cnt; i++) out[i] = k_fact[off + i];
  return cnt;
}

/* One radix-p stage on data held as logs: `cols` interleaved p-point DFTs with root
 * x^(order/p); out[s + cols*i] = sum_j in[s + cols*j] * x^((order/p)*i*j).
 * multiplicative_fft.c:124-173 (there the exponent is advanced incrementally). */
static void mfft_stage(const orc_field *f, const uint32_t *lg, uint32_t *out, uint32_t p)
{
  const uint32_t order = f->order, cols = order / p, step = order / p;
  for (uint32_t i = 0; i < p; i++)
    for (uint32_t s = 0; s < cols; s++) {
      uint32_t acc = 0;
      for (uint32_t j = 0; j < p; j++) {
        const uint32_t l = lg[s + cols * j];
        if (l == order) continue;
        acc ^= f->exp[(uint32_t)((l + (uint64_t)step * i % order * j) % order)];
      }
      out[s + cols * i] = acc;
    }
}

/* Twiddle + digit permutation between stages, back to log form.
 * multiplicative_fft.c:175-231: element (i, s), s = s_ext + group*s_int, goes to
 * s_ext + group*(i + s_int*p) after multiplication by x^(group*s_int*i);
 * group = (product of the factors already processed) * e. */
static void mfft_reorder(const orc_field *f, const uint32_t *val, uint32_t *lg, uint32_t p,
                         uint32_t group)
{
  const uint32_t order = f->order, cols = order / p;
  for (uint32_t i = 0; i < p; i++)
    for (uint32_t s = 0; s < cols; s++) {
      const uint32_t s_ext = s % group, s_int = s / group;
      uint32_t l = f->log[val[i * cols + s]];
      if (l != order) l = (uint32_t)((l + (uint64_t)group * s_int * i) % order);
      lg[s_ext + group * (i + s_int * p)] = l;
    }
}

/* multiplicative_fft.c:233-257.  lg: logs in / values out; buf: scratch. */
static void mfft_chain(const orc_field *f, uint32_t *lg, uint32_t *buf, const uint32_t *fac,
                       int nfac, uint32_t e)
{
  uint32_t done = 1;
  for (int k = 0; k < nfac; k++) {
    mfft_stage(f, lg, buf, fac[k]);
    if (k != nfac - 1) mfft_reorder(f, buf, lg, fac[k], done * e);
    else memcpy(lg, buf, sizeof(uint32_t) * f->order);
    done *= fac[k];
  }
}

/* multiplicative_fft.c:106-122: data[i] <- P(x^i), i < order.  1 = factorisation unknown. */
int orc_multiplicative_fft(const orc_field *f, uint32_t *data, uint32_t *scratch)
{
  uint32_t fac[8];
  const int nfac = orc_mult_factors(f->m, fac);
  if (!nfac) return 1;
  for (uint32_t i = 0; i < f->order; i++) data[i] = f->log[data[i]];
  mfft_chain(f, data, scratch, fac, nfac, 1);
  return 0;
}

/* multiplicative_fft.c:259-324: evaluation of a polynomial with num_coeffs coefficients,
 * skipping the trailing factors whose product e satisfies e*num_coeffs <= order.
 * poly and eval must be distinct here. */
int orc_eval_multiplicative_fft(const orc_field *f, const uint32_t *poly, uint32_t *eval,
                                uint32_t num_coeffs, uint32_t *scratch)
{
  const uint32_t order = f->order;
  if (num_coeffs <= 1) {                                           /* :278-283 */
    const uint32_t v = num_coeffs ? poly[0] : 0;
    for (uint32_t i = 0; i < order; i++) eval[i] = v;
    return 0;
  }
  uint32_t fac[8];
  int nfac = orc_mult_factors(f->m, fac);
  if (!nfac) return 1;
  uint32_t e = 1;
  while (nfac > 0 && (uint64_t)e * fac[nfac - 1] * num_coeffs <= order) e *= fac[--nfac];  /* :290-295 */
  for (uint32_t i = 0; i < order / e; i++) {                       /* pre-twist, :306-320 */
    const uint32_t l = i < num_coeffs ? f->log[poly[i]] : order;
    for (uint32_t j = 0; j < e; j++)
      eval[e * i + j] = (l == order) ? order : (uint32_t)((l + (uint64_t)i * j) % order);
  }
  /* :322 -> step chaining with stride_glob = initial exponent = e: the stage/reorder
   * geometry is the same as the full transform because cols = order/p in both. */
  mfft_chain(f, eval, scratch, fac, nfac, e);
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* BCH code construction / encode.  Follows libbch/bch.c:30-226, 492-506.                 */
/* ------------------------------------------------------------------------------------ */

static inline uint32_t bit_get(const uint64_t *p, size_t i) { return (uint32_t)(p[i >> 6] >> (i & 63)) & 1u; }
static inline void bit_flip(uint64_t *p, size_t i) { p[i >> 6] ^= 1ull << (i & 63); }

size_t orc_bch_words(const orc_code *c) { return (c->length - 1) / 64 + 1; }

/* r <- r mod g, bit-serial from the top (libbase/gf2_polynomials.c:10-70). */
static void gf2_mod(uint64_t *r, uint32_t rdeg, const uint64_t *g, uint32_t gdeg)
{
  if (gdeg > rdeg) return;
  const uint32_t gw = gdeg / 64 + 1, rw = rdeg / 64 + 1;
  for (uint32_t i = rdeg; i >= gdeg && i != UINT32_MAX; i--) {
    if (!bit_get(r, i)) continue;
    const uint32_t sh = i - gdeg, w0 = sh >> 6, s = sh & 63;
    for (uint32_t k = 0; k < gw; k++) {
      r[w0 + k] ^= g[k] << s;
      if (s && w0 + k + 1 < rw) r[w0 + k + 1] ^= g[k] >> (64 - s);
    }
  }
}

/* Generator = product of the minimal polynomials of x^1 .. x^(d-1): walk the cyclotomic
 * cosets of the exponents (bch.c:143-220), build each minimal polynomial over GF(2^m)
 * (gf2_extension_polynomials.c:117-129), multiply them over GF(2) (bch.c:95-113). */
int orc_bch_gen(orc_code *c, uint32_t length, uint32_t distance)
{
  memset(c, 0, sizeof *c);
  if (distance >= length || length > (1u << 30)) return 1;        /* bch.c:37 */
  uint32_t m = 2;
  while ((1u << m) < length) m++;                                  /* bch.c:39-40 */
  if (orc_field_init(&c->f, m)) return 1;
  const orc_field *f = &c->f;
  const uint32_t order = f->order;
  uint8_t *seen = (uint8_t *)calloc(f->n, 1);
  const size_t cap = (size_t)distance * m / 64 + 4;
  uint64_t *g = (uint64_t *)calloc(2 * cap, sizeof(uint64_t));
  uint64_t *tmp = g + cap;
  g[0] = 1;
  uint32_t gdeg = 0;
  uint32_t mp[32];
  for (uint32_t start = 1; start < order; start++) {
    if (seen[start]) continue;
    int tagged = 0;
    uint32_t deg = 0, e = start;
    mp[0] = 1;
    do {                                                           /* one coset */
      seen[e] = 1;
      if (e < distance) tagged = 1;                                /* bch.c:178, offset 0 */
      /* mp <- mp * (X + x^e) */
      mp[deg + 1] = mp[deg];
      for (uint32_t i = deg; i > 0; i--) mp[i] = mp[i - 1] ^ orc_mul(f, mp[i], f->exp[e]);
      mp[0] = orc_mul(f, mp[0], f->exp[e]);
      deg++;
      e = (uint32_t)(((uint64_t)e * 2) % order);
    } while (e != start);
    if (!tagged) continue;
    /* g <- g * mp (mp has coefficients in {0,1}) */
    const size_t gw = gdeg / 64 + 1;
    memset(tmp, 0, cap * sizeof(uint64_t));
    for (uint32_t i = 0; i <= deg; i++) {
      if (!(mp[i] & 1)) continue;
      const uint32_t w0 = i >> 6, s = i & 63;
      for (size_t k = 0; k < gw; k++) {
        tmp[w0 + k] ^= g[k] << s;
        if (s) tmp[w0 + k + 1] ^= g[k] >> (64 - s);
      }
    }
    memcpy(g, tmp, cap * sizeof(uint64_t));
    gdeg += deg;
  }
  free(seen);
  if (gdeg >= length) {                                            /* bch.c:44-48 */
    free(g);
    orc_field_free(&c->f);
    return 1;
  }
  c->gen = (uint64_t *)calloc(gdeg / 64 + 1, sizeof(uint64_t));
  memcpy(c->gen, g, (gdeg / 64 + 1) * sizeof(uint64_t));
  free(g);
  c->distance = distance;
  c->gen_degree = gdeg;
  c->length = length;
  return 0;
}

void orc_bch_free(orc_code *c)
{
  orc_field_free(&c->f);
  free(c->gen);
  c->gen = NULL;
}

/* Systematic encoding, bch.c:492-506: data bit i -> codeword bit i + deg g; parity =
 * (data * X^deg g) mod g in bits 0..deg g - 1. */
int orc_bch_encode(const orc_code *c, uint64_t *cw, const uint32_t *bits, uint32_t nbits)
{
  const uint32_t gd = c->gen_degree;
  if (nbits > c->length - gd) return 1;
  memset(cw, 0, orc_bch_words(c) * sizeof(uint64_t));
  for (uint32_t i = 0; i < nbits; i++)
    if (bits[i]) cw[(i + gd) >> 6] |= 1ull << ((i + gd) & 63);
  gf2_mod(cw, c->length - 1, c->gen, gd);
  for (uint32_t i = 0; i < nbits; i++) {
    const uint64_t b = 1ull << ((i + gd) & 63);
    if (bits[i]) cw[(i + gd) >> 6] |= b; else cw[(i + gd) >> 6] &= ~b;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* BCH decode.  Follows libbch/bch.c:228-562.                                             */
/* ------------------------------------------------------------------------------------ */

/* bch.c:228-326.  syndrome[i] = c(x^i), i = 0..d-1.  Method rule of bch.c:240-254:
 * FFT when d*deg(g)/2^m > 0.5*m^2, otherwise reduce mod g and accumulate powers. */
int32_t orc_bch_syndrome(const orc_code *c, const uint64_t *word, uint32_t *syn, int *used_fft)
{
  const orc_field *f = &c->f;
  const uint32_t d = c->distance, order = f->order;
  const float threshold = 0.5f * f->m * f->m;
  const float cost = ((float)d) * ((float)c->gen_degree) / ((float)f->n);
  const int fft = cost > threshold;
  if (used_fft) *used_fft = fft;
  if (fft) {
    uint32_t *coef = (uint32_t *)calloc(f->n, sizeof(uint32_t));
    uint32_t *res = (uint32_t *)calloc(f->n, sizeof(uint32_t));
    for (uint32_t i = 0; i < c->length; i++) coef[i] = bit_get(word, i);       /* :278-279 */
    orc_eval_additive_fft(f, coef, res, c->length);                             /* :292 */
    for (uint32_t i = 0; i < d; i++) syn[i] = res[i];
    free(res);
    free(coef);
  } else {
    const size_t words = orc_bch_words(c);
    uint64_t *r = (uint64_t *)malloc(words * sizeof(uint64_t));
    memcpy(r, word, words * sizeof(uint64_t));
    gf2_mod(r, c->length - 1, c->gen, c->gen_degree);                           /* :261-262 */
    for (uint32_t i = 0; i < d; i++) syn[i] = 0;
    for (uint32_t j = 0; j < c->gen_degree; j++) {                              /* :301-314 */
      if (!bit_get(r, j)) continue;
      uint32_t e = 0;
      for (uint32_t i = 0; i < d; i++) {
        syn[i] ^= f->exp[e];
        e += j;
        if (e >= order) e -= order;
      }
    }
    free(r);
  }
  for (uint32_t i = 0; i < d; i++)
    if (syn[i]) return 1;                                                        /* :317-324 */
  return 0;
}

/* Berlekamp-Massey, bch.c:328-404.  Three rotating buffers of d+1 words (A scratch,
 * B previous connection polynomial, C current one); the caller zero-initialises them,
 * which makes the otherwise unspecified tail C[deg C + 1 .. L] deterministic.  Returns the
 * LFSR length L and leaves *elp = C. */
uint32_t orc_bch_elp(const orc_code *c, const uint32_t *syn, uint32_t *bufs, uint32_t **elp)
{
  const orc_field *f = &c->f;
  const uint32_t order = f->order, d = c->distance;
  uint32_t *A = bufs, *B = bufs + (d + 1), *C = bufs + 2 * (d + 1);
  uint32_t degA = 0, degB = 0, degC = 0;
  uint32_t L = 0, gap = 1, inv_b_log = 0;
  B[0] = 1;
  C[0] = 1;
  for (uint32_t i = 0; i + 1 < d; i++) {
    uint32_t disc = syn[i + 1];                                                  /* :346-348 */
    for (uint32_t j = 1; j <= degC; j++) disc ^= orc_mul(f, C[j], syn[i + 1 - j]);
    if (!disc) { gap++; continue; }
    const uint32_t dlog = f->log[disc];
    uint32_t scale = dlog + inv_b_log;
    if (scale >= order) scale -= order;
    /* A = C - (disc/b) X^gap B, written over [0, max(degC, degB+gap)], :352-375 */
    const uint32_t hiB = degB + gap;
    uint32_t j = 0;
    const uint32_t keep = gap < degC + 1 ? gap : degC + 1;
    for (; j < keep; j++) A[j] = C[j];
    for (; j < gap; j++) A[j] = 0;
    for (; j <= hiB; j++) {
      const uint32_t bv = B[j - gap];
      uint32_t s = bv ? f->log[bv] + scale : 0;
      if (s >= order) s -= order;
      A[j] = bv ? f->exp[s] : 0;
    }
    for (; j <= degC; j++) A[j] = 0;
    if (degC != hiB) {
      degA = degC > hiB ? degC : hiB;
      for (j = gap; j <= degC; j++) A[j] ^= C[j];
    } else {
      /* equal degrees: leading terms may cancel; degA keeps its previous value when every
       * coefficient in [gap, degC] cancels (it then equals degC: an upper bound). */
      for (j = gap; j <= degC; j++) {
        A[j] ^= C[j];
        if (A[j]) degA = j;
      }
    }
    if (2 * L <= i) {                                                            /* :377-389 */
      L = i + 1 - L;
      degB = degC;
      uint32_t *t = B; B = C; C = t;
      inv_b_log = order - dlog;
      gap = 1;
    } else {
      gap++;
    }
    { uint32_t *t = C; C = A; A = t; }                                          /* :394-395 */
    degC = degA;
  }
  *elp = C;
  return L;
}

/* bch.c:406-490.  Evaluate elp[0..L] at x^i, i < order; each zero yields position order-i.
 * Method rule of bch.c:429-435 (FFT when L > 0.25 m^2, else log-domain Chien search). */
uint32_t orc_bch_locate(const orc_code *c, const uint32_t *elp, uint32_t L, uint32_t *pos)
{
  const orc_field *f = &c->f;
  const uint32_t order = f->order;
  uint32_t *val = (uint32_t *)calloc(f->n, sizeof(uint32_t));
  if ((float)L > 0.25f * f->m * f->m) {
    uint32_t *coef = (uint32_t *)calloc(f->n, sizeof(uint32_t));
    for (uint32_t i = 0; i <= L; i++) coef[i] = elp[i];                          /* :444-445 */
    orc_eval_additive_fft(f, coef, val, L + 1);                                  /* :450 */
    free(coef);
  } else {
    uint32_t *lg = (uint32_t *)malloc(sizeof(uint32_t) * (L + 1));               /* :459-477 */
    for (uint32_t j = 0; j <= L; j++) lg[j] = f->log[elp[j]];
    for (uint32_t i = 0; i < order; i++) {
      uint32_t v = elp[0];
      for (uint32_t j = 1; j <= L; j++) {
        if (lg[j] == order) continue;
        v ^= f->exp[lg[j]];
        lg[j] += j;
        if (lg[j] >= order) lg[j] -= order;
      }
      val[i] = v;
    }
    free(lg);
  }
  uint32_t count = 0;
  for (uint32_t i = 0; i < order; i++)
    if (val[i] == 0) pos[count++] = order - i;                                   /* :479-488 */
  free(val);
  return count;
}

/* bch.c:508-562 */
uint32_t orc_bch_decode(const orc_code *c, uint64_t *word, uint32_t *corrected)
{
  const uint32_t d = c->distance;
  uint32_t *syn = (uint32_t *)calloc(d + 1, sizeof(uint32_t));
  uint32_t found = 0, ok = 0;
  if (orc_bch_syndrome(c, word, syn, NULL)) {
    uint32_t *bufs = (uint32_t *)calloc(3 * (size_t)(d + 1), sizeof(uint32_t));
    uint32_t *elp = NULL;
    const uint32_t L = orc_bch_elp(c, syn, bufs, &elp);
    if (L <= (d - 1) / 2) {
      /* up to order roots can be reported when the stale-extended polynomial is zero-ish;
       * size the list for the worst case */
      uint32_t *pos = (uint32_t *)malloc(sizeof(uint32_t) * c->f.n);
      found = orc_bch_locate(c, elp, L, pos);
      if (found == L) {
        ok = 1;
        /* :539-542.  The bit-0 quirk reports position `order`; for a shortened code that
         * bit lies outside the packed array and the reference's add_bit writes out of
         * bounds (undefined behaviour).  Defined here as: counted, not flipped. */
        const size_t nbits = orc_bch_words(c) * 64;
        for (uint32_t i = 0; i < found; i++)
          if (pos[i] < nbits) bit_flip(word, pos[i]);
      }
      free(pos);
    }
    free(bufs);
  } else {
    ok = 1;
  }
  free(syn);
  if (corrected) *corrected = found;
  return ok;
}

void orc_bch_decode_batch(const orc_code *c, uint64_t *words, size_t count, uint8_t *ok,
                          uint32_t *corrected)
{
  const size_t w = orc_bch_words(c);
  for (size_t i = 0; i < count; i++) {
    uint32_t corr = 0;
    const uint32_t r = orc_bch_decode(c, words + i * w, &corr);
    if (ok) ok[i] = (uint8_t)r;
    if (corrected) corrected[i] = corr;
  }
}

/* FNV-1a 64 over bytes (SURVEY.md Appendix C fingerprints). */
uint64_t orc_fnv64(const void *data, size_t bytes)
{
  const unsigned char *p = (const unsigned char *)data;
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < bytes; i++) h = (h ^ p[i]) * 1099511628211ull;
  return h;
}
