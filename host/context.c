/*
 * Lazily created device contexts behind the reference-shaped host API.
 *
 * The reference's structs (TfiniteField, TbchCode) are passed by value and have no slot for
 * device state, so the contexts live in a small cache keyed by the field degree (and, for
 * codes, by distance / length / generator degree).  Access is serialised by one mutex: the
 * reference API is re-entrant for distinct buffers, and so is this one, at the price of
 * serialising concurrent callers on the GPU (use the batch entry points for throughput).
 *
 * Device selection: environment variable BCHFFT_DEVICE (default 0).
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

#include "host_internal.h"

#define MAX_FIELDS 32
#define MAX_CODES 64

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

static struct { uint32_t m; bchfft_field* ctx; } g_fields[MAX_FIELDS];
static int g_num_fields = 0;

static struct
{
  uint32_t m, distance, length, gen_degree;
  bchfft_code* ctx;
} g_codes[MAX_CODES];
static int g_num_codes = 0;

void bchfft_host_lock(void) { pthread_mutex_lock(&g_lock); }
void bchfft_host_unlock(void) { pthread_mutex_unlock(&g_lock); }

void bchfft_host_die(const char* what, int status)
{
  fprintf(stderr, "bch_fft_b200: %s failed (%d): %s\n"
          "bch_fft_b200: this library has no CPU fallback; a CUDA device and "
          "libbchfft_cuda.so are required.\n", what, status, bchfft_last_error());
  abort();
}

static int device_ordinal(void)
{
  const char* e = getenv("BCHFFT_DEVICE");
  return e ? atoi(e) : 0;
}

/* caller holds the lock */
bchfft_field* bchfft_host_field(TfiniteField* p_f, int* status)
{
  for (int i = 0; i < g_num_fields; i++)
    if (g_fields[i].m == p_f->m) { *status = 0; return g_fields[i].ctx; }
  bchfft_field* ctx = NULL;
  *status = bchfft_field_create(device_ordinal(), p_f->m, p_f->exp, p_f->log, &ctx);
  if (*status) return NULL;
  if (g_num_fields < MAX_FIELDS)
  {
    g_fields[g_num_fields].m = p_f->m;
    g_fields[g_num_fields].ctx = ctx;
    g_num_fields++;
  }
  return ctx;
}

/* caller holds the lock */
bchfft_code* bchfft_host_code(TbchCode* p_code, int* status)
{
  const uint32_t m = p_code->m_field.m;
  for (int i = 0; i < g_num_codes; i++)
    if (g_codes[i].m == m && g_codes[i].distance == p_code->m_distance &&
        g_codes[i].length == p_code->m_length && g_codes[i].gen_degree == p_code->m_gen_poly_degree)
    { *status = 0; return g_codes[i].ctx; }
  bchfft_field* f = bchfft_host_field(&p_code->m_field, status);
  if (!f) return NULL;
  bchfft_code* ctx = NULL;
  *status = bchfft_code_create(f, p_code->m_distance, p_code->m_length,
                               (const uint64_t*) p_code->m_packed_gen_poly,
                               p_code->m_gen_poly_degree, &ctx);
  if (*status) return NULL;
  if (g_num_codes == MAX_CODES)
  {
    /* recycle the oldest entry */
    bchfft_code_destroy(g_codes[0].ctx);
    for (int i = 1; i < MAX_CODES; i++) g_codes[i - 1] = g_codes[i];
    g_num_codes--;
  }
  g_codes[g_num_codes].m = m;
  g_codes[g_num_codes].distance = p_code->m_distance;
  g_codes[g_num_codes].length = p_code->m_length;
  g_codes[g_num_codes].gen_degree = p_code->m_gen_poly_degree;
  g_codes[g_num_codes].ctx = ctx;
  g_num_codes++;
  return ctx;
}

/* release every cached device context (optional; for leak checkers and tests) */
void bchfft_host_shutdown(void)
{
  pthread_mutex_lock(&g_lock);
  for (int i = 0; i < g_num_codes; i++) bchfft_code_destroy(g_codes[i].ctx);
  g_num_codes = 0;
  for (int i = 0; i < g_num_fields; i++) bchfft_field_destroy(g_fields[i].ctx);
  g_num_fields = 0;
  pthread_mutex_unlock(&g_lock);
}
