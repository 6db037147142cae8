// TEST INFRASTRUCTURE ONLY: a tiny single-threaded CUDA execution emulator.
//
// It lets the *actual* kernel sources under bch_fft_b200/csrc be compiled with g++ (-x c++,
// -DBCHFFT_EMU, this header force-included) and executed on the CPU, block by block, with one
// ucontext fiber per CUDA thread so that __syncthreads() has real barrier semantics.  The
// build container has no GPU; running the kernels' real index arithmetic against the oracle
// here catches logic errors before GPU minutes are spent.  The product never loads the
// emulated library (bch_fft_b200/_lib.py only opens the nvcc-built one and fails loudly if it
// is missing); only tests/test_emu_*.py do.
//
// Supported: __global__/__device__/__shared__ (static and dynamic), threadIdx/blockIdx/
// blockDim/gridDim, __syncthreads, atomicAdd/atomicXor/atomicOr (trivially atomic: fibers are
// cooperative), __popc/__ffs/__brev/__clz, and the slice of the runtime API the C-ABI layer
// uses (malloc/free/memcpy/memset/streams/events as no-ops), __shfl_xor_sync between lanes that run in
// lock step.  Not supported: other shuffles, votes, cooperative groups, textures.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ucontext.h>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__
#define __launch_bounds__(...)
#define __restrict__ __restrict
// per host thread: the multi-device driver of the host library runs one emulated device per thread
#define __shared__ static thread_local
#define __constant__ static
#define __align__(x) __attribute__((aligned(x)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

namespace cudaemu {
struct Fiber {
  ucontext_t ctx;
  uint3 tid;
  bool done;
  char *stack;
  unsigned shfl_phase;   // parity of this thread's next emulated warp shuffle
};
struct State {
  ucontext_t sched;
  Fiber *cur = nullptr;
  uint3 block_idx{0, 0, 0};
  dim3 block_dim, grid_dim;
  char *dyn_smem = nullptr;
  void (*entry)(void *) = nullptr;
  void *arg = nullptr;
};
inline State &st() { static thread_local State s; return s; }

inline void trampoline()
{
  State &s = st();
  s.entry(s.arg);
  s.cur->done = true;
  swapcontext(&s.cur->ctx, &s.sched);
}

// Run `body` once per (block, thread); threads of a block are fibers stepped round-robin, so
// every thread reaches barrier k before any thread runs past it.
template <class Body>
void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, Body body)
{
  State &s = st();
  const size_t nthreads = (size_t)block.x * block.y * block.z;
  const size_t stack_bytes = 256 * 1024;
  std::vector<Fiber> fibers(nthreads);
  for (auto &f : fibers) f.stack = (char *)malloc(stack_bytes);
  std::vector<char> smem(dyn_smem_bytes + 64);
  s.dyn_smem = smem.data();
  s.block_dim = block;
  s.grid_dim = grid;
  s.entry = [](void *p) { (*(Body *)p)(); };
  s.arg = &body;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        s.block_idx = uint3{bx, by, bz};
        size_t k = 0;
        for (unsigned tz = 0; tz < block.z; tz++)
          for (unsigned ty = 0; ty < block.y; ty++)
            for (unsigned tx = 0; tx < block.x; tx++, k++) {
              Fiber &f = fibers[k];
              f.tid = uint3{tx, ty, tz};
              f.done = false;
              f.shfl_phase = 0u;
              getcontext(&f.ctx);
              f.ctx.uc_stack.ss_sp = f.stack;
              f.ctx.uc_stack.ss_size = stack_bytes;
              f.ctx.uc_link = nullptr;
              makecontext(&f.ctx, (void (*)())trampoline, 0);
            }
        size_t live = nthreads;
        while (live) {
          live = 0;
          for (auto &f : fibers) {
            if (f.done) continue;
            s.cur = &f;
            swapcontext(&s.sched, &f.ctx);
            if (!f.done) live++;
          }
        }
      }
  for (auto &f : fibers) free(f.stack);
  s.cur = nullptr;
}
}  // namespace cudaemu

#define threadIdx (cudaemu::st().cur->tid)
#define blockIdx (cudaemu::st().block_idx)
#define blockDim (cudaemu::st().block_dim)
#define gridDim (cudaemu::st().grid_dim)

static inline void __syncthreads()
{
  cudaemu::State &s = cudaemu::st();
  swapcontext(&s.cur->ctx, &s.sched);
}
// fibers are cooperative: a warp barrier yields like a block barrier (stronger than needed, fine for
// kernels whose warps all reach the same number of barriers)
static inline void __syncwarp(unsigned = 0xffffffffu) { __syncthreads(); }
static inline void __threadfence() {}
// Warp shuffle among lanes that run in lock step (the same sequence of yields, e.g. the lane group of
// one codeword in k_bm): every lane posts its value, yields once, and reads the partner's post.  Two
// post buffers alternate, because a lower-numbered lane resumes first and may already have posted
// its next value by the time the partner reads this one.
namespace cudaemu {
inline uint32_t *shfl_post() { static thread_local uint32_t post[2][1024]; return &post[0][0]; }
}
static inline uint32_t __shfl_xor_sync(unsigned, uint32_t v, int lane_mask)
{
  cudaemu::State &s = cudaemu::st();
  const unsigned me = s.cur->tid.x, ph = s.cur->shfl_phase;
  s.cur->shfl_phase = ph ^ 1u;
  uint32_t *post = cudaemu::shfl_post() + ph * 1024u;
  post[me] = v;
  __syncthreads();
  return post[me ^ (unsigned)lane_mask];
}
static inline int __shfl_xor_sync(unsigned m, int v, int lane_mask)
{
  return (int)__shfl_xor_sync(m, (uint32_t)v, lane_mask);
}

template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicXor(T *p, T v) { T o = *p; *p = o ^ v; return o; }
template <class T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __brev(unsigned x)
{
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
  return __builtin_bswap32(x);
}
template <class T> static inline T __ldg(const T *p) { return *p; }

// ---- runtime API subset -------------------------------------------------------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0 };
struct cudaDeviceProp { int multiProcessorCount; size_t sharedMemPerBlockOptin; char name[64]; int major, minor; };
static inline const char *cudaGetErrorString(cudaError_t e) { return e ? "emulated CUDA error" : "no error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
// BCHFFT_EMU_DEVICES=k: pretend there are k devices (all of them this CPU), so that the host library's
// multi-device sharding can be exercised without a GPU
static inline cudaError_t cudaGetDeviceCount(int *n)
{
  const char *e = getenv("BCHFFT_EMU_DEVICES");
  *n = (e && atoi(e) >= 1) ? atoi(e) : 1;
  return cudaSuccess;
}
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
  memset(p, 0, sizeof *p);
  p->multiProcessorCount = 4;
  p->sharedMemPerBlockOptin = 227 * 1024;
  strcpy(p->name, "cudaemu");
  p->major = 10;
  return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { return cudaFree(p); }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = 0)
{
  for (size_t r = 0; r < h; r++) memmove((char *)d + r * dp, (const char *)s + r * sp, w);
  return cudaSuccess;
}
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (cudaEvent_t)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
// emulated "device" memory is host memory: every pointer is its own device alias
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void *devicePointer; void *hostPointer; };
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p)
{
  a->type = cudaMemoryTypeHost;
  a->device = 0;
  a->devicePointer = const_cast<void *>(p);
  a->hostPointer = const_cast<void *>(p);
  return cudaSuccess;
}
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

// kernel<<<grid, block, smem, stream>>>(args...)  is written  BCHFFT_LAUNCH(kernel, grid, block, smem, stream, args...)
#define BCHFFT_LAUNCH(kernel, grid, block, smem, stream, ...) \
  cudaemu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define BCHFFT_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>( \
    (reinterpret_cast<uintptr_t>(cudaemu::st().dyn_smem) + 63) & ~uintptr_t(63))
