// Copy-only microbenchmark: what the host's pinned-memory DMA path sustains with 1, 2, 4, .. GPUs
// copying at once -- the ceiling of every end-to-end (host buffer in, host buffer out) number.
// One process, one host thread + two streams per device (like the library's multi-device entry points);
// per device a private pinned buffer.  Prints one JSON line per (devices, mode).
//   nvcc -O2 -o copybench copybench.cu -lpthread ;  ./copybench [MiB per device = 1024] [reps = 5]
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

static double now() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

struct Dev {
  int id;
  size_t bytes;
  char *h_in, *h_out, *d_a, *d_b;
  cudaStream_t s0, s1;
  int mode, reps;       // 0: H2D, 1: D2H, 2: both directions at once
  pthread_barrier_t *bar;
  double t0, t1;
};

static void *worker(void *p)
{
  Dev *d = (Dev *)p;
  CK(cudaSetDevice(d->id));
  const size_t chunk = (size_t)64 << 20;
  pthread_barrier_wait(d->bar);
  d->t0 = now();
  for (int r = 0; r < d->reps; r++)
    for (size_t o = 0; o < d->bytes; o += chunk) {
      const size_t n = d->bytes - o < chunk ? d->bytes - o : chunk;
      if (d->mode != 1) CK(cudaMemcpyAsync(d->d_a + o, d->h_in + o, n, cudaMemcpyHostToDevice, d->s0));
      if (d->mode != 0) CK(cudaMemcpyAsync(d->h_out + o, d->d_b + o, n, cudaMemcpyDeviceToHost, d->s1));
    }
  CK(cudaStreamSynchronize(d->s0));
  CK(cudaStreamSynchronize(d->s1));
  d->t1 = now();
  return 0;
}

int main(int argc, char **argv)
{
  const size_t mib = argc > 1 ? (size_t)atol(argv[1]) : 1024;
  const int reps = argc > 2 ? atoi(argv[2]) : 5;
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  Dev devs[16];
  for (int i = 0; i < ndev && i < 16; i++) {
    Dev &d = devs[i];
    d.id = i;
    d.bytes = mib << 20;
    CK(cudaSetDevice(i));
    CK(cudaMallocHost((void **)&d.h_in, d.bytes));
    CK(cudaMallocHost((void **)&d.h_out, d.bytes));
    memset(d.h_in, i + 1, d.bytes);
    memset(d.h_out, 0, d.bytes);
    CK(cudaMalloc((void **)&d.d_a, d.bytes));
    CK(cudaMalloc((void **)&d.d_b, d.bytes));
    CK(cudaMemset(d.d_b, 7, d.bytes));
    CK(cudaStreamCreateWithFlags(&d.s0, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&d.s1, cudaStreamNonBlocking));
    CK(cudaDeviceSynchronize());
  }
  const char *names[3] = {"h2d", "d2h", "h2d+d2h"};
  for (int g = 1; g <= ndev; g *= 2)
    for (int mode = 0; mode < 3; mode++) {
      pthread_barrier_t bar;
      pthread_barrier_init(&bar, 0, g);
      pthread_t th[16];
      for (int i = 0; i < g; i++) {
        devs[i].mode = mode;
        devs[i].reps = reps;
        devs[i].bar = &bar;
        pthread_create(&th[i], 0, worker, &devs[i]);
      }
      double t0 = 1e300, t1 = 0;
      for (int i = 0; i < g; i++) {
        pthread_join(th[i], 0);
        if (devs[i].t0 < t0) t0 = devs[i].t0;
        if (devs[i].t1 > t1) t1 = devs[i].t1;
      }
      pthread_barrier_destroy(&bar);
      const double bytes = (double)g * reps * (mib << 20) * (mode == 2 ? 2 : 1);
      printf("{\"devices\": %d, \"mode\": \"%s\", \"GB_per_s_total\": %.1f, \"GB_per_s_per_device_per_direction\": %.1f, "
             "\"MiB_per_device\": %zu, \"reps\": %d}\n",
             g, names[mode], bytes / (t1 - t0) / 1e9, bytes / (t1 - t0) / 1e9 / g / (mode == 2 ? 2 : 1), mib, reps);
      fflush(stdout);
    }
  return 0;
}
