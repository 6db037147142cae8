// Batched multiplicative FFT (secondary kernel; see include/bchfft_cuda.h).
#include "cabi_common.h"
using namespace bchfft;

extern "C" void bchfft_mfft_release(bchfft_field *) {}

extern "C" int bchfft_multiplicative_fft_dev(bchfft_field *, const uint32_t *, size_t, uint32_t *, size_t,
                                             uint32_t, size_t, void *)
{
  set_error("bchfft_multiplicative_fft_dev: not implemented yet");
  return BCHFFT_ERR_UNSUPPORTED;
}

extern "C" int bchfft_multiplicative_fft_host(bchfft_field *, const uint32_t *, size_t, uint32_t *, size_t,
                                              uint32_t, size_t)
{
  set_error("bchfft_multiplicative_fft_host: not implemented yet");
  return BCHFFT_ERR_UNSUPPORTED;
}
