/* internal glue between the reference-shaped host API and the C-ABI CUDA layer */
#pragma once
#include "finite_field.h"
#include "bch.h"
#include "bchfft_cuda.h"

#define BCHFFT_HOST_MAX_DEVICES 16

/* print the last C-ABI error and abort: used by entry points that cannot report errors */
void bchfft_host_die(const char* what, int status) __attribute__((noreturn));

/* device set of this process (BCHFFT_DEVICES / BCHFFT_DEVICE / all visible), see context.c */
int bchfft_host_num_devices(void);
int bchfft_host_device(int slot);
/* one lock per device slot, held around every device call on that slot */
void bchfft_host_lock_slot(int slot);
void bchfft_host_unlock_slot(int slot);
/* cached contexts of a slot; the caller holds the slot's lock */
bchfft_field* bchfft_host_field(TfiniteField* p_f, int slot, int* status);
bchfft_code* bchfft_host_code(TbchCode* p_code, int slot, int* status);
void bchfft_host_shutdown(void);
int bchfft_host_num_contexts(void);

/* Batch driver: cuts [0, count) into one contiguous range per device (multiples of 32 items, at
 * least min_per_device each) and calls fn(slot, lo, hi, arg) for every range from its own host thread
 * (range 0 on the calling thread).  Returns the first non-zero status; its error text is moved to
 * the calling thread. */
typedef int (*bchfft_shard_fn)(int slot, size_t lo, size_t hi, void* arg);
int bchfft_host_run_sharded(size_t count, size_t min_per_device, bchfft_shard_fn fn, void* arg);
