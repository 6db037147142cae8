/* internal glue between the reference-shaped host API and the C-ABI CUDA layer */
#pragma once
#include "finite_field.h"
#include "bch.h"
#include "bchfft_cuda.h"

void bchfft_host_lock(void);
void bchfft_host_unlock(void);
/* print the last C-ABI error and abort: used by entry points that cannot report errors */
void bchfft_host_die(const char* what, int status) __attribute__((noreturn));
bchfft_field* bchfft_host_field(TfiniteField* p_f, int* status);
bchfft_code* bchfft_host_code(TbchCode* p_code, int* status);
void bchfft_host_shutdown(void);
