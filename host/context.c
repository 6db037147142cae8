/*
 * Lazily created device contexts behind the reference-shaped host API, and the multi-device
 * driver of the batch entry points.
 *
 * The reference's structs (TfiniteField, TbchCode) are passed by value and have no slot for
 * device state, so the contexts live in a small cache keyed by (device slot, field degree) and,
 * for codes, by (device slot, distance, length, generator degree).
 *
 * Device set (the batch shards by codeword / polynomial range, no collective):
 *   BCHFFT_DEVICES=0,2,5 | all    explicit list of CUDA ordinals
 *   BCHFFT_DEVICE=3               one device (what a one-process-per-GPU launcher sets per rank)
 *   neither                       every visible device
 * Batch calls (bch_decode_batch, bch_encode_batch, additive_fft_batch, multiplicative_fft_batch)
 * cut the caller's array into one contiguous range per device and drive each range from its own
 * host thread; the caller's array is the gathered result.  Single-item calls use the first device.
 *
 * Locking: one mutex per device slot, held for the duration of a device call (the GPU is a
 * serial resource at that granularity, and contexts of one device share streams and workspaces);
 * calls on different devices run concurrently.  A short global mutex guards the cache tables.
 * The reference API is re-entrant for distinct buffers, and so is this one.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "host_internal.h"

#define MAX_FIELDS 64
#define MAX_CODES 128

static pthread_mutex_t g_cache_lock = PTHREAD_MUTEX_INITIALIZER;
static pthread_mutex_t g_slot_lock[BCHFFT_HOST_MAX_DEVICES];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static int g_devices[BCHFFT_HOST_MAX_DEVICES];
static int g_num_devices = 0;

static struct { int slot; uint32_t m; bchfft_field* ctx; } g_fields[MAX_FIELDS];
static int g_num_fields = 0;

static struct
{
  int slot;
  uint32_t m, distance, length, gen_degree;
  bchfft_code* ctx;
} g_codes[MAX_CODES];
static int g_num_codes = 0;

void bchfft_host_die(const char* what, int status)
{
  fprintf(stderr, "bch_fft_b200: %s failed (%d): %s\n"
          "bch_fft_b200: this library has no CPU fallback; a CUDA device and "
          "libbchfft_cuda.so are required.\n", what, status, bchfft_last_error());
  abort();
}

static void resolve_devices(void)
{
  for (int i = 0; i < BCHFFT_HOST_MAX_DEVICES; i++) pthread_mutex_init(&g_slot_lock[i], NULL);
  int visible = bchfft_device_count();
  if (visible < 0) visible = 0;
  const char* list = getenv("BCHFFT_DEVICES");
  const char* one = getenv("BCHFFT_DEVICE");
  g_num_devices = 0;
  if (list && strcmp(list, "all") != 0 && *list)
  {
    const char* p = list;
    while (*p && g_num_devices < BCHFFT_HOST_MAX_DEVICES)
    {
      char* end = NULL;
      const long v = strtol(p, &end, 10);
      if (end == p) break;
      g_devices[g_num_devices++] = (int) v;   /* validated by bchfft_field_create */
      p = (*end == ',') ? end + 1 : end;
    }
  }
  else if (!list && one && *one)
  {
    g_devices[g_num_devices++] = atoi(one);
  }
  else
  {
    for (int i = 0; i < visible && i < BCHFFT_HOST_MAX_DEVICES; i++) g_devices[g_num_devices++] = i;
  }
  if (g_num_devices == 0) g_devices[g_num_devices++] = 0;   /* the create call reports "no device" */
}

int bchfft_host_num_devices(void)
{
  pthread_once(&g_once, resolve_devices);
  return g_num_devices;
}

int bchfft_host_device(int slot)
{
  pthread_once(&g_once, resolve_devices);
  return g_devices[slot];
}

void bchfft_host_lock_slot(int slot)
{
  pthread_once(&g_once, resolve_devices);
  pthread_mutex_lock(&g_slot_lock[slot]);
}

void bchfft_host_unlock_slot(int slot) { pthread_mutex_unlock(&g_slot_lock[slot]); }

/* caller holds the slot's lock */
bchfft_field* bchfft_host_field(TfiniteField* p_f, int slot, int* status)
{
  bchfft_field* ctx = NULL;
  pthread_mutex_lock(&g_cache_lock);
  for (int i = 0; i < g_num_fields; i++)
    if (g_fields[i].slot == slot && g_fields[i].m == p_f->m) { ctx = g_fields[i].ctx; break; }
  if (ctx) { pthread_mutex_unlock(&g_cache_lock); *status = 0; return ctx; }
  *status = bchfft_field_create(bchfft_host_device(slot), p_f->m, p_f->exp, p_f->log, &ctx);
  if (*status == 0 && g_num_fields < MAX_FIELDS)
  {
    g_fields[g_num_fields].slot = slot;
    g_fields[g_num_fields].m = p_f->m;
    g_fields[g_num_fields].ctx = ctx;
    g_num_fields++;
  }
  pthread_mutex_unlock(&g_cache_lock);
  return *status ? NULL : ctx;
}

/* caller holds the slot's lock */
bchfft_code* bchfft_host_code(TbchCode* p_code, int slot, int* status)
{
  const uint32_t m = p_code->m_field.m;
  pthread_mutex_lock(&g_cache_lock);
  for (int i = 0; i < g_num_codes; i++)
    if (g_codes[i].slot == slot && g_codes[i].m == m && g_codes[i].distance == p_code->m_distance &&
        g_codes[i].length == p_code->m_length && g_codes[i].gen_degree == p_code->m_gen_poly_degree)
    {
      bchfft_code* hit = g_codes[i].ctx;
      pthread_mutex_unlock(&g_cache_lock);
      *status = 0;
      return hit;
    }
  pthread_mutex_unlock(&g_cache_lock);
  bchfft_field* f = bchfft_host_field(&p_code->m_field, slot, status);
  if (!f) return NULL;
  bchfft_code* ctx = NULL;
  *status = bchfft_code_create(f, p_code->m_distance, p_code->m_length,
                               (const uint64_t*) p_code->m_packed_gen_poly,
                               p_code->m_gen_poly_degree, &ctx);
  if (*status) return NULL;
  pthread_mutex_lock(&g_cache_lock);
  if (g_num_codes == MAX_CODES)
  {
    /* recycle the oldest entry of this slot (its lock is held by the caller, so nobody uses it) */
    int victim = -1;
    for (int i = 0; i < g_num_codes && victim < 0; i++) if (g_codes[i].slot == slot) victim = i;
    if (victim < 0)
    {
      /* table full of other devices' codes: use the context uncached, the caller destroys nothing */
      pthread_mutex_unlock(&g_cache_lock);
      return ctx;
    }
    bchfft_code_destroy(g_codes[victim].ctx);
    for (int i = victim + 1; i < g_num_codes; i++) g_codes[i - 1] = g_codes[i];
    g_num_codes--;
  }
  g_codes[g_num_codes].slot = slot;
  g_codes[g_num_codes].m = m;
  g_codes[g_num_codes].distance = p_code->m_distance;
  g_codes[g_num_codes].length = p_code->m_length;
  g_codes[g_num_codes].gen_degree = p_code->m_gen_poly_degree;
  g_codes[g_num_codes].ctx = ctx;
  g_num_codes++;
  pthread_mutex_unlock(&g_cache_lock);
  return ctx;
}

/* ---- one contiguous range per device, one host thread per range ------------------------------ */
typedef struct
{
  bchfft_shard_fn fn;
  void* arg;
  int slot, status;
  size_t lo, hi;
  char err[256];
} shard_job;

static void* shard_worker(void* p)
{
  shard_job* j = (shard_job*) p;
  j->status = j->fn(j->slot, j->lo, j->hi, j->arg);
  if (j->status)
  {
    /* the C-ABI's error text is thread-local: carry it back to the caller's thread */
    strncpy(j->err, bchfft_last_error(), sizeof j->err - 1);
    j->err[sizeof j->err - 1] = 0;
  }
  return NULL;
}

int bchfft_host_run_sharded(size_t count, size_t min_per_device, bchfft_shard_fn fn, void* arg)
{
  int g = bchfft_host_num_devices();
  const char* menv = getenv("BCHFFT_MIN_PER_DEVICE");   /* test hook: split small batches too */
  if (menv && atol(menv) >= 1) min_per_device = (size_t) atol(menv);
  if (min_per_device < 1) min_per_device = 1;
  if ((size_t) g > count / min_per_device) g = (int) (count / min_per_device);
  if (g <= 1) return fn(0, 0, count, arg);
  bchfft_set_concurrent_hosts(g);   /* host helper threads of the g concurrent device calls share this machine */
  /* ranges in multiples of 32 items (one slab), so that no device sees a ragged slab in the middle */
  size_t per = (count + (size_t) g - 1) / (size_t) g;
  per = (per + 31) / 32 * 32;
  shard_job jobs[BCHFFT_HOST_MAX_DEVICES];
  pthread_t th[BCHFFT_HOST_MAX_DEVICES];
  int started[BCHFFT_HOST_MAX_DEVICES];
  int n = 0;
  for (int k = 0; k < g; k++)
  {
    const size_t lo = (size_t) k * per < count ? (size_t) k * per : count;
    const size_t hi = lo + per < count ? lo + per : count;
    if (lo >= hi) break;
    jobs[n].fn = fn;
    jobs[n].arg = arg;
    jobs[n].slot = k;
    jobs[n].lo = lo;
    jobs[n].hi = hi;
    jobs[n].status = 0;
    jobs[n].err[0] = 0;
    n++;
  }
  for (int k = 1; k < n; k++)
  {
    started[k] = pthread_create(&th[k], NULL, shard_worker, &jobs[k]) == 0;
    if (!started[k]) shard_worker(&jobs[k]);      /* no thread: run the range here */
  }
  shard_worker(&jobs[0]);
  int status = jobs[0].status;
  const char* err = jobs[0].err;
  for (int k = 1; k < n; k++)
  {
    if (started[k]) pthread_join(th[k], NULL);
    if (!status && jobs[k].status) { status = jobs[k].status; err = jobs[k].err; }
  }
  bchfft_set_concurrent_hosts(1);
  if (status && err != jobs[0].err) bchfft_set_last_error(err);
  return status;
}

/* number of cached device contexts (fields + codes, all devices) */
int bchfft_host_num_contexts(void)
{
  pthread_mutex_lock(&g_cache_lock);
  const int n = g_num_fields + g_num_codes;
  pthread_mutex_unlock(&g_cache_lock);
  return n;
}

/* release every cached device context (optional; for leak checkers and tests) */
void bchfft_host_shutdown(void)
{
  const int g = bchfft_host_num_devices();
  for (int s = 0; s < g; s++) pthread_mutex_lock(&g_slot_lock[s]);
  pthread_mutex_lock(&g_cache_lock);
  for (int i = 0; i < g_num_codes; i++) bchfft_code_destroy(g_codes[i].ctx);
  g_num_codes = 0;
  for (int i = 0; i < g_num_fields; i++) bchfft_field_destroy(g_fields[i].ctx);
  g_num_fields = 0;
  pthread_mutex_unlock(&g_cache_lock);
  for (int s = g - 1; s >= 0; s--) pthread_mutex_unlock(&g_slot_lock[s]);
}
