// temporary: entry points not implemented yet
#include "cabi_common.h"
using namespace bchfft;
extern "C" void bchfft_mfft_release(bchfft_field *) {}
