/*
 * bchfft_cuda.h -- the thin C-ABI between the host C libraries (host/, same entry points as
 * the reference's libfft / libbch) and the hand-written sm_100a kernels in
 * bch_fft_b200/csrc.  Plain C: opaque handles, raw pointers and sizes, int status
 * (0 = ok, < 0 = error, text via bchfft_last_error()).  There is no CPU fallback behind any
 * of these calls: without a CUDA device they fail.
 *
 * Each entry point names the reference interface it stands in for (paths relative to the
 * reference tree, kunzjacq/bch_fft).
 *
 * Batch conventions: item i of a batch is the i-th contiguous record.  Results for item i are
 * bit-identical to one call of the corresponding single-item reference function on item i.
 * "_host" variants take host pointers and do the host<->device copies themselves on the
 * context's stream; "_dev" variants take device pointers (HBM-resident data) and only enqueue
 * work on the given stream.
 *
 * Threading: contexts that live on DIFFERENT devices may be used concurrently from different host
 * threads (the host library drives one thread per GPU that way).  Calls on contexts of the SAME device
 * -- a field and the codes built on it share streams, workspaces and plan caches -- must be serialised
 * by the caller.  The error text is thread-local; the launch counter is atomic; the bchfft_profile_*
 * calls are a single-threaded measurement aid.
 */
#ifndef BCHFFT_CUDA_H
#define BCHFFT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bchfft_field bchfft_field; /* per (device, GF(2^m)) state: tables, constants, workspace */
typedef struct bchfft_code bchfft_code;   /* per BCH code state on top of a field */

/* largest field degree the kernels are compiled for (GF(2^2) .. GF(2^25): every field of the
 * reference's table whose x is primitive, libbase/finite_field.c:9-18) */
#define BCHFFT_MAX_FIELD_DEGREE 25

#define BCHFFT_OK 0
#define BCHFFT_ERR_CUDA (-1)     /* CUDA runtime error (no device, launch failure, ...) */
#define BCHFFT_ERR_ARG (-2)      /* invalid argument */
#define BCHFFT_ERR_UNSUPPORTED (-3) /* field degree / size outside what was compiled */

/* thread-local text of the last error returned on this thread */
const char *bchfft_last_error(void);
/* set this thread's error text (the host library moves a worker thread's message to its caller) */
void bchfft_set_last_error(const char *text);
/* number of visible CUDA devices (0 or negative status if none) */
int bchfft_device_count(void);
/* 1 when the library is the CPU emulator build used by the unit tests, 0 for the real CUDA build */
int bchfft_is_emulated(void);

/* ---- field context ------------------------------------------------------------------------
 * Stands in for the per-call setup of libfft/additive_fft.c:41-78 (subspace-polynomial
 * constants) and uploads what libbase/finite_field.c:5-49 built: `exp`/`log` are the caller's
 * TfiniteField tables (n entries each, exp[n-1] = 0, log[0] = n-1).  `device` is the CUDA
 * ordinal the context lives on. */
int bchfft_field_create(int device, uint32_t m, const uint32_t *exp, const uint32_t *log,
                        bchfft_field **out);
void bchfft_field_destroy(bchfft_field *f);
/* block until everything enqueued on the context's own stream has finished */
int bchfft_field_sync(bchfft_field *f);

/* ---- additive FFT ---------------------------------------------------------------------------
 * Replaces evaluate_polynomial_additive_FFT (include/additive_fft.h:8-12,
 * libfft/additive_fft.c:13-21) run `batch` times.  polys: [batch][poly_stride] coefficients
 * (poly_stride >= 2^m; only indices 0..num_terms are used, as in additive_fft.c:99);
 * results: [batch][result_stride], entries 0..2^m-2 = P(x^i) (entry 2^m-1 is not written,
 * like the reference); natural (may be NULL): [batch][poly_stride'] = P(k) for k = 0..2^m-1,
 * i.e. what the reference leaves in the clobbered p_poly (additive_DFT_opt, :23-178).
 * The input array is NOT modified. */
int bchfft_additive_fft_host(bchfft_field *f, const uint32_t *polys, size_t poly_stride,
                             uint32_t *results, size_t result_stride, uint32_t *natural,
                             size_t natural_stride, uint32_t num_terms, size_t batch);
int bchfft_additive_fft_dev(bchfft_field *f, const uint32_t *d_polys, size_t poly_stride,
                            uint32_t *d_results, size_t result_stride, uint32_t *d_natural,
                            size_t natural_stride, uint32_t num_terms, size_t batch, void *stream);

/* ---- multiplicative FFT (secondary) -------------------------------------------------------
 * Replaces multiplicative_FFT / evaluate_polynomial_multiplicative_FFT
 * (include/multiplicative_fft.h:47,185-190; libfft/multiplicative_fft.c:106-122,259-324) run
 * `batch` times: evals[b][i] = P_b(x^i), i < 2^m-1, for polynomials with num_coeffs
 * coefficients.  Returns BCHFFT_ERR_UNSUPPORTED where the reference returns 1 (m outside
 * 8..26: factorisation unknown). */
int bchfft_multiplicative_fft_host(bchfft_field *f, const uint32_t *polys, size_t poly_stride,
                                   uint32_t *evals, size_t eval_stride, uint32_t num_coeffs,
                                   size_t batch);
int bchfft_multiplicative_fft_dev(bchfft_field *f, const uint32_t *d_polys, size_t poly_stride,
                                  uint32_t *d_evals, size_t eval_stride, uint32_t num_coeffs,
                                  size_t batch, void *stream);

/* ---- BCH decode -----------------------------------------------------------------------------
 * A code context is built from what bch_gen_code (libbch/bch.c:30-54) produced: designed
 * distance d = 2t+1, code length in bits, packed generator polynomial and its degree
 * (include/bch.h:10-22).  Codewords are `words_per_cw = (length-1)/64+1` 64-bit words, bit i
 * of the polynomial in bit (i mod 64) of word i/64 (libbase/gf2_polynomials_misc.c:56-98). */
int bchfft_code_create(bchfft_field *f, uint32_t distance, uint32_t length,
                       const uint64_t *packed_gen_poly, uint32_t gen_poly_degree,
                       bchfft_code **out);
void bchfft_code_destroy(bchfft_code *c);

/* Replaces bch_decode (include/bch.h:170-171, libbch/bch.c:508-562) run `batch` times, in
 * place: ok[i] = return value (1 corrected or clean, 0 uncorrectable and word untouched),
 * corrected[i] = *po_corrected (number of error-locator roots found). */
int bchfft_bch_decode_host(bchfft_code *c, uint64_t *words, size_t batch, uint8_t *ok,
                           uint32_t *corrected);
int bchfft_bch_decode_dev(bchfft_code *c, uint64_t *d_words, size_t batch, uint8_t *d_ok,
                          uint32_t *d_corrected, void *stream);

/* Systematic encoder: bch_encode (include/bch.h:174-188, libbch/bch.c:492-506) run `batch` times.
 * data: [batch][data_words] packed data bits (bit i of an item = bit i % 64 of word i / 64; only the
 * first data_bits count, the reference's p_data[i] = that bit); codewords: [batch][words_per_cw],
 * data bit i at codeword bit i + deg(g), parity (data X^deg(g) mod g) in bits 0 .. deg(g)-1, the
 * rest zero.  BCHFFT_ERR_ARG when data_bits > length - deg(g) (where the reference returns 1);
 * BCHFFT_ERR_UNSUPPORTED when the code context was created without its generator polynomial or
 * deg(g) < 8. */
int bchfft_bch_encode_host(bchfft_code *c, const uint64_t *data, size_t data_words, uint32_t data_bits,
                           uint64_t *codewords, size_t batch);
int bchfft_bch_encode_dev(bchfft_code *c, const uint64_t *d_data, size_t data_words, uint32_t data_bits,
                          uint64_t *d_codewords, size_t batch, void *stream);

/* Stage-level entry points, for callers that drive the stages one by one like bch/main.c:103-122.
 * syndromes: [batch][distance] = c(x^i), i = 0..d-1 (bch_eval_syndrome, bch.c:228-326; entry 0
 * follows the reference's method rule: c(1) on its FFT path, (c mod g)(1) on its direct path);
 * flags[i] = its return value. */
int bchfft_bch_syndrome_host(bchfft_code *c, const uint64_t *words, size_t batch,
                             uint32_t *syndromes, int32_t *flags);
/* bch_compute_elp (bch.c:328-404): elp: [batch][distance+1] (C[0..L], zero-initialised
 * scratch semantics), L[i] = returned LFSR length */
int bchfft_bch_elp_host(bchfft_code *c, const uint32_t *syndromes, size_t batch, uint32_t *elp,
                        uint32_t *L);
/* bch_locate_errors (bch.c:406-490): positions: [batch][distance+1] = order - i for every root
 * x^i of elp[0..L], descending; count[i] = number of roots */
int bchfft_bch_locate_host(bchfft_code *c, const uint32_t *elp, const uint32_t *L, size_t batch,
                           uint32_t *positions, uint32_t *count);

/* ---- raw device memory helpers (so that C callers need no CUDA headers) -------------------*/
int bchfft_dev_alloc(int device, size_t bytes, void **out);
int bchfft_dev_free(int device, void *p);
int bchfft_dev_upload(int device, void *d_dst, const void *h_src, size_t bytes);
int bchfft_dev_download(int device, void *h_dst, const void *d_src, size_t bytes);

/* page-locked host memory: buffers handed to the *_host entry points should come from here, so
 * that the copy-in / decode / copy-out lanes really overlap (pageable memory is staged by the
 * driver and the copies serialise).  Usable from any thread; freed with bchfft_pinned_free. */
int bchfft_pinned_alloc(size_t bytes, void **out);
int bchfft_pinned_free(void *p);

/* Hint from a host driver that runs `n` *_host calls at once, one per device, from one process (the drop-in
 * library's sharded batch calls): the helper threads a call starts on the host (the bit flips of the positions
 * result mode) are then sized for a host shared between n such calls.  n < 1 counts as 1. */
void bchfft_set_concurrent_hosts(int n);

/* launch counter: number of kernels this library has launched since the last reset
 * (bench.py reports it as gpu_launches) */
uint64_t bchfft_launch_count(void);
void bchfft_launch_count_reset(void);

/* per-kernel device timing: between begin and end every kernel launch of this library is
 * bracketed by CUDA events on its stream.  bchfft_profile_end() synchronises the device,
 * returns the number of distinct kernels seen; bchfft_profile_get(i, ...) reads kernel i's
 * name, summed device time and launch count.  (bench.py uses it for the roofline of the
 * dominant kernel; keep it off in timed regions.) */
int bchfft_profile_begin(void);
int bchfft_profile_end(void);
int bchfft_profile_get(int idx, const char **name, double *total_ms, uint64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* BCHFFT_CUDA_H */
