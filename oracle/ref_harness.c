/*
 * TEST / BASELINE INFRASTRUCTURE ONLY (oracle/): multi-threaded batch drivers around the
 * UNMODIFIED reference entry points, linked into oracle/_ref/libref_oracle.so.  The reference is
 * single-threaded but re-entrant ("Threading"): each worker gets private
 * buffers and a disjoint contiguous range, the code / field structs are shared read-only.
 * Used by bench.py's cpu_baseline and `--impl reference` legs; wall time is measured here with
 * clock_gettime(CLOCK_MONOTONIC) around the whole job.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "additive_fft.h"
#include "bch.h"
#include "gf2_polynomials_misc.h"

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

typedef struct
{
  TbchCode* code;
  unsigned long* words;
  size_t lo, hi, wpc;
  unsigned char* ok;
  uint32_t* corrected;
} decode_job;

static void* decode_worker(void* arg)
{
  decode_job* j = (decode_job*) arg;
  TbchCodeBuffers buf = bch_alloc_buffers(j->code);
  const size_t bm_bytes = (size_t) (j->code->m_distance + 1) * sizeof(uint32_t);
  for (size_t i = j->lo; i < j->hi; i++)
  {
    /* the reference mallocs these (bch.c:568-575); zero them so that words with more than t
     * errors decode deterministically */
    memset(buf.m_A, 0, bm_bytes);
    memset(buf.m_B, 0, bm_bytes);
    memset(buf.m_C_elp, 0, bm_bytes);
    memset(buf.m_error, 0, bm_bytes);
    uint32_t corr = 0;
    const uint32_t ok = bch_decode(j->code, j->words + i * j->wpc, &buf, &corr);
    if (j->ok) j->ok[i] = (unsigned char) ok;
    if (j->corrected) j->corrected[i] = corr;
  }
  bch_free_buffers(&buf);
  return NULL;
}

/* returns wall seconds */
double ref_bch_decode_batch_mt(TbchCode* code, unsigned long* words, size_t count,
                               unsigned char* ok, uint32_t* corrected, int nthreads)
{
  if (nthreads < 1) nthreads = 1;
  pthread_t* th = (pthread_t*) malloc(sizeof(pthread_t) * nthreads);
  decode_job* jobs = (decode_job*) malloc(sizeof(decode_job) * nthreads);
  const size_t per = (count + nthreads - 1) / nthreads;
  const double t0 = now_s();
  for (int k = 0; k < nthreads; k++)
  {
    jobs[k].code = code;
    jobs[k].words = words;
    jobs[k].wpc = degree_to_packed_size(code->m_length - 1);
    jobs[k].lo = (size_t) k * per < count ? (size_t) k * per : count;
    jobs[k].hi = jobs[k].lo + per < count ? jobs[k].lo + per : count;
    jobs[k].ok = ok;
    jobs[k].corrected = corrected;
    pthread_create(&th[k], NULL, decode_worker, &jobs[k]);
  }
  for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
  const double dt = now_s() - t0;
  free(jobs);
  free(th);
  return dt;
}

typedef struct
{
  TfiniteField* f;
  uint32_t* polys;
  uint32_t* results;
  uint32_t num_terms;
  size_t lo, hi;
} fft_job;

static void* fft_worker(void* arg)
{
  fft_job* j = (fft_job*) arg;
  const size_t n = j->f->n;
  for (size_t i = j->lo; i < j->hi; i++)
    evaluate_polynomial_additive_FFT(j->f, j->polys + i * n, j->results + i * n, j->num_terms);
  return NULL;
}

/* polys: [count][n] (clobbered like the reference does), results: [count][n]; wall seconds */
double ref_additive_fft_batch_mt(TfiniteField* f, uint32_t* polys, uint32_t* results,
                                 uint32_t num_terms, size_t count, int nthreads)
{
  if (nthreads < 1) nthreads = 1;
  pthread_t* th = (pthread_t*) malloc(sizeof(pthread_t) * nthreads);
  fft_job* jobs = (fft_job*) malloc(sizeof(fft_job) * nthreads);
  const size_t per = (count + nthreads - 1) / nthreads;
  const double t0 = now_s();
  for (int k = 0; k < nthreads; k++)
  {
    jobs[k].f = f;
    jobs[k].polys = polys;
    jobs[k].results = results;
    jobs[k].num_terms = num_terms;
    jobs[k].lo = (size_t) k * per < count ? (size_t) k * per : count;
    jobs[k].hi = jobs[k].lo + per < count ? jobs[k].lo + per : count;
    pthread_create(&th[k], NULL, fft_worker, &jobs[k]);
  }
  for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
  const double dt = now_s() - t0;
  free(jobs);
  free(th);
  return dt;
}

/* ---- multiplicative FFT -------------------------------------------------------------------- */
#include "multiplicative_fft.h"

typedef struct
{
  TfiniteField* f;
  uint32_t* data;
  size_t lo, hi;
  int status;
} mfft_job;

static void* mfft_worker(void* arg)
{
  mfft_job* j = (mfft_job*) arg;
  const size_t n = j->f->n;
  uint32_t* scratch = (uint32_t*) malloc(sizeof(uint32_t) * n);
  j->status = 0;
  for (size_t i = j->lo; i < j->hi; i++)
    j->status |= multiplicative_FFT(j->f, j->data + i * n, scratch);
  free(scratch);
  return NULL;
}

/* data: [count][n], rows transformed in place (entries 0 .. n-2); wall seconds, < 0 when the
 * reference reports an unknown factorisation */
double ref_multiplicative_fft_batch_mt(TfiniteField* f, uint32_t* data, size_t count, int nthreads)
{
  if (nthreads < 1) nthreads = 1;
  pthread_t* th = (pthread_t*) malloc(sizeof(pthread_t) * nthreads);
  mfft_job* jobs = (mfft_job*) malloc(sizeof(mfft_job) * nthreads);
  const size_t per = (count + nthreads - 1) / nthreads;
  const double t0 = now_s();
  for (int k = 0; k < nthreads; k++)
  {
    jobs[k].f = f;
    jobs[k].data = data;
    jobs[k].lo = (size_t) k * per < count ? (size_t) k * per : count;
    jobs[k].hi = jobs[k].lo + per < count ? jobs[k].lo + per : count;
    pthread_create(&th[k], NULL, mfft_worker, &jobs[k]);
  }
  int bad = 0;
  for (int k = 0; k < nthreads; k++) { pthread_join(th[k], NULL); bad |= jobs[k].status; }
  const double dt = now_s() - t0;
  free(jobs);
  free(th);
  return bad ? -1.0 : dt;
}

/* ---- encoder --------------------------------------------------------------------------------- */
typedef struct
{
  TbchCode* code;
  const unsigned long* data;   /* [count][dwords] packed data bits */
  unsigned long* codewords;    /* [count][wpc] */
  size_t lo, hi, dwords, wpc;
  uint32_t data_bits;
} encode_job;

static void* encode_worker(void* arg)
{
  encode_job* j = (encode_job*) arg;
  uint32_t* bits = (uint32_t*) malloc(sizeof(uint32_t) * (j->data_bits + 1));
  for (size_t i = j->lo; i < j->hi; i++)
  {
    const unsigned long* d = j->data + i * j->dwords;
    for (uint32_t b = 0; b < j->data_bits; b++) bits[b] = (uint32_t) ((d[b >> 6] >> (b & 63)) & 1ul);
    bch_encode(j->code, j->codewords + i * j->wpc, bits, j->data_bits);
  }
  free(bits);
  return NULL;
}

/* bch_encode (bch.c:492-506) per item from packed data bits; wall seconds */
double ref_bch_encode_batch_mt(TbchCode* code, const unsigned long* data, size_t dwords, uint32_t data_bits,
                               unsigned long* codewords, size_t count, int nthreads)
{
  if (nthreads < 1) nthreads = 1;
  pthread_t* th = (pthread_t*) malloc(sizeof(pthread_t) * nthreads);
  encode_job* jobs = (encode_job*) malloc(sizeof(encode_job) * nthreads);
  const size_t per = (count + nthreads - 1) / nthreads;
  const double t0 = now_s();
  for (int k = 0; k < nthreads; k++)
  {
    jobs[k].code = code;
    jobs[k].data = data;
    jobs[k].codewords = codewords;
    jobs[k].dwords = dwords;
    jobs[k].data_bits = data_bits;
    jobs[k].wpc = degree_to_packed_size(code->m_length - 1);
    jobs[k].lo = (size_t) k * per < count ? (size_t) k * per : count;
    jobs[k].hi = jobs[k].lo + per < count ? jobs[k].lo + per : count;
    pthread_create(&th[k], NULL, encode_worker, &jobs[k]);
  }
  for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
  const double dt = now_s() - t0;
  free(jobs);
  free(th);
  return dt;
}
