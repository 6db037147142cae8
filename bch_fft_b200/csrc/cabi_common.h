// Internal declarations shared by the translation units behind include/bchfft_cuda.h.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <map>
#include <string>
#include <vector>

#include "bs_common.cuh"
#include "gm_plan.h"
#include "fft_kernels.cuh"
#include "../../include/bchfft_cuda.h"

namespace bchfft {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;
extern std::atomic<int> g_concurrent_hosts;   // *_host calls a host driver runs at once in this process

#define BCHFFT_CUDA_TRY(expr)                                                              \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      bchfft::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__,  \
                        __LINE__);                                                         \
      return BCHFFT_ERR_CUDA;                                                              \
    }                                                                                      \
  } while (0)

// per-kernel device timing (see bchfft_profile_begin): events around a launch while enabled
void prof_before(const char *kernel_name, cudaStream_t stream);
void prof_after();

// launch + bookkeeping: counts the launch, times it when profiling, surfaces launch errors
#define BCHFFT_RUN(kernel, grid, block, smem, stream, ...)                                 \
  do {                                                                                     \
    bchfft::prof_before(#kernel, (stream));                                                \
    BCHFFT_LAUNCH(kernel, grid, block, smem, stream, __VA_ARGS__);                         \
    bchfft::prof_after();                                                                  \
    bchfft::g_launches++;                                                                  \
    BCHFFT_CUDA_TRY(cudaGetLastError());                                                   \
  } while (0)

// Opt a kernel in to `bytes` of dynamic shared memory.  The attribute belongs to the function
// (per device), not to one of our contexts or call sites, so one process-wide table keeps the
// high-water mark per (function, device) and the attribute is only ever raised.
int ensure_smem(const void *kernel, int device, size_t bytes);
#define BCHFFT_ENSURE_SMEM(kernel, device, bytes)                                          \
  do {                                                                                     \
    int rc_smem_ = bchfft::ensure_smem((const void *)(kernel), (device), (size_t)(bytes)); \
    if (rc_smem_) return rc_smem_;                                                         \
  } while (0)

// field degrees the kernels are instantiated for
#ifndef BCHFFT_M_LIST
#define BCHFFT_M_LIST(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17) X(18) X(19) X(20) X(21) X(22) X(23) X(24) X(25)
#endif

struct DeviceBuffer {
  void *p = nullptr;
  size_t bytes = 0;
  int reserve(size_t need);  // grow-only
  void release();
};

// page-locked host staging (grow-only)
struct PinnedBuffer {
  void *p = nullptr;
  size_t bytes = 0;
  int reserve(size_t need);
  void release();
};

}  // namespace bchfft

struct bchfft_field {
  int device = 0;
  uint32_t m = 0, n = 0, order = 0;
  int ps = 0;      // plane words per element in HBM
  int t_max = 0;   // log2(tile positions) that fits shared memory
  cudaStream_t stream = nullptr;
  std::vector<uint32_t> h_exp, h_log;
  uint32_t *d_tw = nullptr, *d_gam = nullptr, *d_perm_exp = nullptr, *d_perm_nat = nullptr;
  uint32_t *d_exp = nullptr, *d_log = nullptr, *d_errloc = nullptr;
  uint32_t *d_qsolve = nullptr;   // {trace mask, h_0 .. h_(m-1)}: y^2 + y = u solver of the degree-2 root finder
  uint32_t twist_free = 0;               // depths whose twist is the identity (gm_plan.h)
  bchfft::FftTables tab;
  std::map<int, bchfft::GmPlan> plans;   // keyed by K
  bchfft::GmPlanPM pm_plan;              // plane-major schedule (fully twist-free fields)
  bool pm_checked = false, pm_ok = false;
  bchfft::DeviceBuffer ws;               // slab workspace
  bchfft::DeviceBuffer stage_in, stage_out, stage_aux;  // device staging for the _host variants
  // two lanes of the pipelined additive-FFT host entry point (copy-in / transform / copy-out overlap)
  struct Lane {
    bchfft::DeviceBuffer in, out, aux, ws;
    cudaStream_t stream = nullptr;
  } lanes[2];
  size_t chunk_bytes = 0;                // workspace budget per chunk of slabs
  // multiplicative FFT state (lazily built)
  void *mfft = nullptr;
};

// bchfft_additive_fft_dev with an explicit slab workspace (cabi_fft.cu); the decode pipeline uses it
// with a workspace of its own lane for the syndromes of few long codewords
int bchfft_additive_fft_dev_ws(bchfft_field *f, bchfft::DeviceBuffer &ws, const uint32_t *d_polys, size_t poly_stride,
                               uint32_t *d_results, size_t result_stride, uint32_t *d_natural,
                               size_t natural_stride, uint32_t num_terms, size_t batch, void *stream);
