/*
 * bchstream: streaming BCH decoder on top of the drop-in host library.
 *
 *   bchstream <codeword bits n> <t> [-b words per batch] [-s status file] < received > corrected
 *
 * stdin / stdout carry packed codewords in the reference's own in-memory format
 * (libbase/gf2_polynomials_misc.c:56-98): degree_to_packed_size(n-1) `unsigned long` words per
 * codeword, bit j of the codeword = bit j%64 of word j/64.  Every word read is written back,
 * corrected in place when bch_decode says ok (libbch/bch.c:508-562), untouched otherwise.  The
 * optional status file receives one record per codeword: uint32 ok, uint32 corrected.
 *
 * Three page-locked batch buffers rotate through a reader thread, the decoder (this thread,
 * bch_decode_batch -> CUDA) and a writer thread, so file IO overlaps the GPU work; inside
 * bch_decode_batch the copies of neighbouring chunks overlap the kernels.  There is no CPU
 * decode path: without a CUDA device the first batch fails and the tool exits non-zero.
 */
#include <pthread.h>
#include <semaphore.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bch.h"
#include "bchfft_cuda.h"
#include "gf2_polynomials_misc.h"
#include "utils.h"

#define NBUF 3

typedef struct
{
  unsigned long* words;
  unsigned char* ok;
  uint32_t* corrected;
  size_t count; /* codewords in the buffer; 0 = end of stream */
} Tbatch;

static Tbatch batch[NBUF];
static sem_t sem_free, sem_filled, sem_decoded;
static size_t words_per_cw, batch_cap;
static FILE* status_file;
static int io_error;

static void* reader_main(void* arg)
{
  (void) arg;
  const size_t cw_bytes = words_per_cw * sizeof(unsigned long);
  for (unsigned i = 0;; i = (i + 1) % NBUF)
  {
    sem_wait(&sem_free);
    const size_t got = fread(batch[i].words, 1, batch_cap * cw_bytes, stdin);
    if (got % cw_bytes)
    {
      fprintf(stderr, "bchstream: input ends inside a codeword (%zu stray bytes dropped)\n",
              got % cw_bytes);
      io_error = 1;
    }
    batch[i].count = got / cw_bytes;
    sem_post(&sem_filled);
    if (batch[i].count == 0) return NULL;
  }
}

static void* writer_main(void* arg)
{
  (void) arg;
  for (unsigned i = 0;; i = (i + 1) % NBUF)
  {
    sem_wait(&sem_decoded);
    const size_t n = batch[i].count;
    if (n == 0) return NULL;
    if (fwrite(batch[i].words, words_per_cw * sizeof(unsigned long), n, stdout) != n) io_error = 1;
    if (status_file)
      for (size_t k = 0; k < n; k++)
      {
        const uint32_t rec[2] = {batch[i].ok[k], batch[i].corrected[k]};
        if (fwrite(rec, sizeof rec, 1, status_file) != 1) io_error = 1;
      }
    sem_post(&sem_free);
  }
}

static void* pinned(size_t bytes)
{
  void* p = NULL;
  if (bchfft_pinned_alloc(bytes, &p))
  {
    fprintf(stderr, "bchstream: %s\n", bchfft_last_error());
    exit(2);
  }
  return p;
}

int main(int argc, char** argv)
{
  if (argc < 3)
  {
    fprintf(stderr, "usage: %s <codeword bits> <t> [-b words per batch] [-s status file]\n", argv[0]);
    return 2;
  }
  const uint32_t length = (uint32_t) strtoul(argv[1], NULL, 0);
  const uint32_t t = (uint32_t) strtoul(argv[2], NULL, 0);
  batch_cap = (size_t) 1 << 18;
  for (int a = 3; a + 1 < argc; a += 2)
  {
    if (!strcmp(argv[a], "-b")) batch_cap = strtoull(argv[a + 1], NULL, 0);
    else if (!strcmp(argv[a], "-s")) status_file = fopen(argv[a + 1], "wb");
  }
  if (batch_cap == 0) batch_cap = 1;

  TbchCode code = bch_gen_code(length, 2 * t + 1);
  if (!code.m_packed_gen_poly)
  {
    fprintf(stderr, "bchstream: no BCH code with length %u and t = %u\n", length, t);
    return 2;
  }
  words_per_cw = degree_to_packed_size(code.m_length - 1);
  for (int i = 0; i < NBUF; i++)
  {
    batch[i].words = (unsigned long*) pinned(batch_cap * words_per_cw * sizeof(unsigned long));
    batch[i].ok = (unsigned char*) pinned(batch_cap);
    batch[i].corrected = (uint32_t*) pinned(batch_cap * sizeof(uint32_t));
  }
  sem_init(&sem_free, 0, NBUF);
  sem_init(&sem_filled, 0, 0);
  sem_init(&sem_decoded, 0, 0);
  pthread_t reader, writer;
  pthread_create(&reader, NULL, reader_main, NULL);
  pthread_create(&writer, NULL, writer_main, NULL);

  size_t total = 0, total_ok = 0, total_bits = 0;
  double busy = 0.;
  init_time();
  const double t_start = absolute_time();
  int rc = 0;
  for (unsigned i = 0;; i = (i + 1) % NBUF)
  {
    sem_wait(&sem_filled);
    const size_t n = batch[i].count;
    if (n)
    {
      const double t0 = absolute_time();
      const int status = bch_decode_batch(&code, batch[i].words, n, batch[i].ok, batch[i].corrected);
      busy += absolute_time() - t0;
      if (status)
      {
        /* no fallback: report and stop (the writer is released with an empty batch) */
        fprintf(stderr, "bchstream: bch_decode_batch failed (%d): %s\n", status, bchfft_last_error());
        batch[i].count = 0;
        rc = 1;
      }
      else
        for (size_t k = 0; k < n; k++)
        {
          total_ok += batch[i].ok[k];
          if (batch[i].ok[k]) total_bits += batch[i].corrected[k];
        }
      total += batch[i].count;
    }
    sem_post(&sem_decoded);
    if (batch[i].count == 0) break;
  }
  pthread_join(writer, NULL);
  if (rc == 0) pthread_join(reader, NULL);
  fflush(stdout);
  if (status_file) fclose(status_file);
  const double wall = absolute_time() - t_start;
  fprintf(stderr,
          "bchstream: %zu codewords (n = %u, t = %u), %zu decoded ok, %zu bits corrected; "
          "%.3f s wall (%.3e codewords/s), %.3f s inside bch_decode_batch (%.3e codewords/s)\n",
          total, code.m_length, t, total_ok, total_bits, wall, wall > 0 ? total / wall : 0.,
          busy, busy > 0 ? total / busy : 0.);
  if (rc)
    return 1;   /* the reader was not joined and may still be inside fread() into a batch buffer:
                   leave the buffers alone, the process ends here */
  for (int i = 0; i < NBUF; i++)
  {
    bchfft_pinned_free(batch[i].words);
    bchfft_pinned_free(batch[i].ok);
    bchfft_pinned_free(batch[i].corrected);
  }
  bch_free_code(&code);
  return io_error;
}
