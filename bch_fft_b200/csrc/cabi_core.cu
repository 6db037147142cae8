// C-ABI plumbing: error text, device helpers, field contexts (see include/bchfft_cuda.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>

#include "cabi_common.h"

namespace bchfft {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_concurrent_hosts{1};

void set_error(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

int DeviceBuffer::reserve(size_t need)
{
  if (need <= bytes) return BCHFFT_OK;
  if (p) cudaFree(p);
  p = nullptr;
  bytes = 0;
  BCHFFT_CUDA_TRY(cudaMalloc(&p, need));
  bytes = need;
  return BCHFFT_OK;
}

void DeviceBuffer::release()
{
  if (p) cudaFree(p);
  p = nullptr;
  bytes = 0;
}

int PinnedBuffer::reserve(size_t need)
{
  if (need <= bytes) return BCHFFT_OK;
  if (p) cudaFreeHost(p);
  p = nullptr;
  bytes = 0;
  BCHFFT_CUDA_TRY(cudaMallocHost(&p, need));
  bytes = need;
  return BCHFFT_OK;
}

void PinnedBuffer::release()
{
  if (p) cudaFreeHost(p);
  p = nullptr;
  bytes = 0;
}

int ensure_smem(const void *kernel, int device, size_t bytes)
{
  // process-wide table: callers on different contexts / devices may run concurrently
  static std::mutex mu;
  static std::map<std::pair<const void *, int>, size_t> high_water;
  std::lock_guard<std::mutex> guard(mu);
  size_t &hw = high_water[std::make_pair(kernel, device)];
  if (bytes <= hw) return BCHFFT_OK;
#ifdef BCHFFT_EMU
  hw = bytes;
  return BCHFFT_OK;
#else
  BCHFFT_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  hw = bytes;
  return BCHFFT_OK;
#endif
}

// ---- per-kernel profiling -------------------------------------------------------------------
struct ProfRec {
  const char *name;
  cudaEvent_t a, b;
};
// (the profiling API is a single-threaded measurement tool: bench.py drives it from one thread with
// no other caller active -- see bchfft_cuda.h)
static bool g_prof = false;
static std::vector<ProfRec> g_prof_recs;
static cudaStream_t g_prof_stream = nullptr;
struct ProfSum {
  std::string name;
  double ms = 0.0;
  uint64_t launches = 0;
};
static std::vector<ProfSum> g_prof_sums;

void prof_before(const char *kernel_name, cudaStream_t stream)
{
  if (!g_prof) return;
  ProfRec r;
  r.name = kernel_name;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, stream);
  g_prof_stream = stream;
  g_prof_recs.push_back(r);
}

void prof_after()
{
  if (!g_prof || g_prof_recs.empty()) return;
  cudaEventRecord(g_prof_recs.back().b, g_prof_stream);
}

template <class T>
static int upload_vec(const std::vector<T> &v, T **out)
{
  *out = nullptr;
  BCHFFT_CUDA_TRY(cudaMalloc((void **)out, v.size() * sizeof(T) + 16));
  BCHFFT_CUDA_TRY(cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return BCHFFT_OK;
}

}  // namespace bchfft

using namespace bchfft;

extern "C" const char *bchfft_last_error(void) { return g_err; }

extern "C" void bchfft_set_last_error(const char *text) { set_error("%s", text ? text : ""); }

extern "C" int bchfft_is_emulated(void)
{
#ifdef BCHFFT_EMU
  return 1;
#else
  return 0;
#endif
}

extern "C" int bchfft_device_count(void)
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    return BCHFFT_ERR_CUDA;
  }
  return n;
}

extern "C" int bchfft_profile_begin(void)
{
  g_prof_recs.clear();
  g_prof_sums.clear();
  g_prof = true;
  return BCHFFT_OK;
}

extern "C" int bchfft_profile_end(void)
{
  g_prof = false;
  BCHFFT_CUDA_TRY(cudaDeviceSynchronize());
  for (auto &r : g_prof_recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
    std::string nm(r.name);
    if (!nm.empty() && nm[0] == '(') nm = nm.substr(1);   // "(kernel<A, B>)": template-id with a comma
    const size_t lt = nm.find('<');
    if (lt != std::string::npos) nm = nm.substr(0, lt);
    ProfSum *slot = nullptr;
    for (auto &s : g_prof_sums)
      if (s.name == nm) slot = &s;
    if (!slot) {
      g_prof_sums.push_back(ProfSum());
      slot = &g_prof_sums.back();
      slot->name = nm;
    }
    slot->ms += ms;
    slot->launches++;
  }
  g_prof_recs.clear();
  return (int)g_prof_sums.size();
}

extern "C" int bchfft_profile_get(int idx, const char **name, double *total_ms, uint64_t *launches)
{
  if (idx < 0 || idx >= (int)g_prof_sums.size()) return BCHFFT_ERR_ARG;
  if (name) *name = g_prof_sums[idx].name.c_str();
  if (total_ms) *total_ms = g_prof_sums[idx].ms;
  if (launches) *launches = g_prof_sums[idx].launches;
  return BCHFFT_OK;
}

extern "C" uint64_t bchfft_launch_count(void) { return g_launches; }
extern "C" void bchfft_launch_count_reset(void) { g_launches = 0; }

extern "C" int bchfft_dev_alloc(int device, size_t bytes, void **out)
{
  BCHFFT_CUDA_TRY(cudaSetDevice(device));
  BCHFFT_CUDA_TRY(cudaMalloc(out, bytes ? bytes : 1));
  return BCHFFT_OK;
}

extern "C" int bchfft_dev_free(int device, void *p)
{
  BCHFFT_CUDA_TRY(cudaSetDevice(device));
  BCHFFT_CUDA_TRY(cudaFree(p));
  return BCHFFT_OK;
}

extern "C" int bchfft_pinned_alloc(size_t bytes, void **out)
{
  if (!out) {
    set_error("bchfft_pinned_alloc: null argument");
    return BCHFFT_ERR_ARG;
  }
  BCHFFT_CUDA_TRY(cudaMallocHost(out, bytes ? bytes : 1));
  return BCHFFT_OK;
}

extern "C" void bchfft_set_concurrent_hosts(int n) { bchfft::g_concurrent_hosts.store(n < 1 ? 1 : n); }

extern "C" int bchfft_pinned_free(void *p)
{
  BCHFFT_CUDA_TRY(cudaFreeHost(p));
  return BCHFFT_OK;
}

extern "C" int bchfft_dev_upload(int device, void *d_dst, const void *h_src, size_t bytes)
{
  BCHFFT_CUDA_TRY(cudaSetDevice(device));
  BCHFFT_CUDA_TRY(cudaMemcpy(d_dst, h_src, bytes, cudaMemcpyHostToDevice));
  return BCHFFT_OK;
}

extern "C" int bchfft_dev_download(int device, void *h_dst, const void *d_src, size_t bytes)
{
  BCHFFT_CUDA_TRY(cudaSetDevice(device));
  BCHFFT_CUDA_TRY(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
  return BCHFFT_OK;
}

static bool m_supported(uint32_t m)
{
#define X(M) if (m == M) return true;
  BCHFFT_M_LIST(X)
#undef X
  return false;
}

extern "C" int bchfft_field_create(int device, uint32_t m, const uint32_t *exp, const uint32_t *log,
                                   bchfft_field **out)
{
  if (!out || !exp || !log) {
    set_error("bchfft_field_create: null argument");
    return BCHFFT_ERR_ARG;
  }
  *out = nullptr;
  if (!m_supported(m)) {
    set_error("bchfft_field_create: GF(2^%u) is not among the compiled field degrees", m);
    return BCHFFT_ERR_UNSUPPORTED;
  }
  int ndev = bchfft_device_count();
  if (ndev <= 0) {
    if (ndev == 0) set_error("bchfft_field_create: no CUDA device (there is no CPU fallback)");
    return BCHFFT_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    set_error("bchfft_field_create: device %d out of range (%d visible)", device, ndev);
    return BCHFFT_ERR_ARG;
  }
  const uint32_t n = 1u << m;
  // sanity-check the caller's tables against the reference convention (finite_field.c:36-48)
  if (exp[0] != 1u || exp[1] != 2u || log[0] != n - 1u || log[1] != 0u || exp[n - 1] != 0u ||
      ((exp[m] ^ n) != field_poly((int)m))) {
    set_error("bchfft_field_create: exp/log tables do not describe GF(2^%u) with the reference's "
              "reduction polynomial 0x%x", m, field_poly((int)m));
    return BCHFFT_ERR_ARG;
  }
  BCHFFT_CUDA_TRY(cudaSetDevice(device));
  bchfft_field *f = new bchfft_field();
  f->device = device;
  f->m = m;
  f->n = n;
  f->order = n - 1u;
  f->ps = plane_stride((int)m);
  {
    // largest power-of-two tile whose M planes fit the 227 KB a CTA may use on sm_100a
    int t = 0;
    while ((size_t)(4u * m) << (t + 1) <= (size_t)227 * 1024) t++;
    f->t_max = t;
    // test hook: smaller tiles force the multi-pass schedule on small fields
    const char *tenv = getenv("BCHFFT_TMAX");
    if (tenv && atoi(tenv) >= 2 && atoi(tenv) < t) f->t_max = atoi(tenv);
  }
  f->h_exp.assign(exp, exp + n);
  f->h_log.assign(log, log + n);
  const char *env = getenv("BCHFFT_CHUNK_MB");
  f->chunk_bytes = (size_t)(env ? atoi(env) : 1024) << 20;

  GmConstants c = gm_constants((int)m, exp, log);
  // position of every evaluation point: natural order (clobbered p_poly) and exp order (po_result)
  std::vector<uint32_t> perm_exp(n), perm_nat(c.pos_of_elem), errloc(n);
  for (uint32_t i = 0; i < n - 1u; i++) perm_exp[i] = perm_nat[exp[i]];
  perm_exp[n - 1] = 0;
  // bch_locate_errors: a root at field element k = x^i is reported as position order - i
  // (bch.c:479-488); per tile position, 0xFFFFFFFF for the point 0
  for (uint32_t p = 0; p < n; p++) {
    const uint32_t k = c.elem_of_pos[p];
    errloc[p] = k ? (n - 1u) - log[k] : 0xFFFFFFFFu;
  }
  f->twist_free = c.twist_free;
  // y^2 + y = u, solved bit by bit (k_bucket_count's closed-form roots of degree-2 locators): trace of every
  // x^k, one trace-one element tau, h_k with h_k^2 + h_k = x^k + Tr(x^k) tau
  std::vector<uint32_t> qsolve(1u + m, 0u);
  {
    auto mul = [&](uint32_t a, uint32_t b) -> uint32_t {
      if (!a || !b) return 0u;
      uint32_t sum = log[a] + log[b];
      if (sum >= n - 1u) sum -= n - 1u;
      return exp[sum];
    };
    uint32_t tau = 0u;
    for (uint32_t k = 0; k < m; k++) {
      uint32_t v = 1u << k, tr = 0u;
      for (uint32_t r = 0; r < m; r++) {
        tr ^= v;
        v = mul(v, v);
      }
      if (tr & 1u) {
        qsolve[0] |= 1u << k;
        if (!tau) tau = 1u << k;
      }
    }
    for (uint32_t k = 0; k < m; k++) {
      uint32_t y = 0u;
      solve_artin_schreier((int)m, mul, (1u << k) ^ (((qsolve[0] >> k) & 1u) ? tau : 0u), &y);
      qsolve[1u + k] = y;
    }
  }
  int rc = BCHFFT_OK;
  if ((rc = upload_vec(c.tw, &f->d_tw)) || (rc = upload_vec(c.gam, &f->d_gam)) ||
      (rc = upload_vec(perm_exp, &f->d_perm_exp)) || (rc = upload_vec(perm_nat, &f->d_perm_nat)) ||
      (rc = upload_vec(errloc, &f->d_errloc)) || (rc = upload_vec(qsolve, &f->d_qsolve)) ||
      (rc = upload_vec(f->h_exp, &f->d_exp)) || (rc = upload_vec(f->h_log, &f->d_log))) {
    bchfft_field_destroy(f);
    return rc;
  }
  memset(&f->tab, 0, sizeof f->tab);
  f->tab.tw = f->d_tw;
  f->tab.gam = f->d_gam;
  for (uint32_t d = 0; d < m; d++) {
    f->tab.tw_off[d] = c.tw_off[d];
    f->tab.gam_off[d] = c.gam_off[d];
  }
  cudaError_t e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
    bchfft_field_destroy(f);
    return BCHFFT_ERR_CUDA;
  }
  // the uploads came from pageable memory on the legacy stream; kernels run on non-blocking streams
  // that do not order against it: make sure every table has landed before the context is handed out
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("bchfft_field_create: table upload failed: %s", cudaGetErrorString(e));
    bchfft_field_destroy(f);
    return BCHFFT_ERR_CUDA;
  }
  *out = f;
  return BCHFFT_OK;
}

extern "C" void bchfft_mfft_release(bchfft_field *f);

extern "C" void bchfft_field_destroy(bchfft_field *f)
{
  if (!f) return;
  cudaSetDevice(f->device);
  if (f->stream) cudaStreamSynchronize(f->stream);
  bchfft_mfft_release(f);
  cudaFree(f->d_tw);
  cudaFree(f->d_gam);
  cudaFree(f->d_perm_exp);
  cudaFree(f->d_perm_nat);
  cudaFree(f->d_errloc);
  cudaFree(f->d_qsolve);
  cudaFree(f->d_exp);
  cudaFree(f->d_log);
  f->ws.release();
  f->stage_in.release();
  f->stage_out.release();
  f->stage_aux.release();
  for (auto &l : f->lanes) {
    l.in.release();
    l.out.release();
    l.aux.release();
    l.ws.release();
    if (l.stream) cudaStreamDestroy(l.stream);
  }
  if (f->stream) cudaStreamDestroy(f->stream);
  delete f;
}

extern "C" int bchfft_field_sync(bchfft_field *f)
{
  if (!f) return BCHFFT_ERR_ARG;
  BCHFFT_CUDA_TRY(cudaSetDevice(f->device));
  BCHFFT_CUDA_TRY(cudaStreamSynchronize(f->stream));
  return BCHFFT_OK;
}
