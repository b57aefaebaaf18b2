// anm_lp.cu -- C ABI of the batched LP solver (include/anm_lp.h); the solver itself is anm_lp.cuh.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/anm_lp.h"
#include "anm_lp.cuh"
#include "anm_lp_warp.cuh"
#include <stdlib.h>

/* NVTX range around the solve (as anm_capi.cu does for reset / step / rollout); -DANM_NO_NVTX compiles it out. */
#ifndef ANM_NO_NVTX
#include <nvtx3/nvToolsExt.h>
namespace {
struct LpNvtxRange {
  explicit LpNvtxRange(const char* name) { nvtxRangePushA(name); }
  ~LpNvtxRange() { nvtxRangePop(); }
};
}  // namespace
#define ANM_LP_NVTX(name) LpNvtxRange lp_nvtx_range_(name)
#else
#define ANM_LP_NVTX(name)
#endif

namespace {

thread_local char g_lp_err[512] = "";

int lp_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_lp_err, sizeof g_lp_err, fmt, ap);
  va_end(ap);
  return code;
}

#define LP_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t err_ = (call);                                                                     \
    if (err_ != cudaSuccess) return lp_fail(ANM_LP_E_CUDA, "%s: %s", #call, cudaGetErrorString(err_)); \
  } while (0)

// the caller's current device is restored on return (as in anm_capi.cu)
struct LpDeviceGuard {
  int prev = -1;
  explicit LpDeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~LpDeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// byte offsets of the per-batch arrays inside one allocation (device handle and host test state alike)
struct Layout {
  int64_t T, d, xB, rval, bl, bu, nl, nu, bid, nid, ridx, atup, total;
};

Layout layout_of(int64_t n, int64_t m, int64_t stride) {
  Layout L;
  int64_t o = 0;
  auto take = [&](int64_t count, int64_t elem) {
    const int64_t at = o;
    o += (count * stride * elem + 255) / 256 * 256;
    return at;
  };
  L.T = take(m * n, 8);
  L.d = take(n, 8);
  L.xB = take(m, 8);
  L.rval = take(n, 8);
  L.bl = take(m, 8);
  L.bu = take(m, 8);
  L.nl = take(n, 8);
  L.nu = take(n, 8);
  L.bid = take(m, 4);
  L.nid = take(n, 4);
  L.ridx = take(n, 4);
  L.atup = take(n, 1);
  L.total = o;
  return L;
}

void bind(anm_lp::Batch& b, char* base, const Layout& L) {
  b.T = reinterpret_cast<double*>(base + L.T);
  b.d = reinterpret_cast<double*>(base + L.d);
  b.xB = reinterpret_cast<double*>(base + L.xB);
  b.rval = reinterpret_cast<double*>(base + L.rval);
  b.bl = reinterpret_cast<double*>(base + L.bl);
  b.bu = reinterpret_cast<double*>(base + L.bu);
  b.nl = reinterpret_cast<double*>(base + L.nl);
  b.nu = reinterpret_cast<double*>(base + L.nu);
  b.bid = reinterpret_cast<int32_t*>(base + L.bid);
  b.nid = reinterpret_cast<int32_t*>(base + L.nid);
  b.ridx = reinterpret_cast<int32_t*>(base + L.ridx);
  b.atup = reinterpret_cast<uint8_t*>(base + L.atup);
}

int check_dims(int32_t n, int32_t m, int64_t batch, int64_t stride) {
  if (n <= 0 || m <= 0 || n > 4096 || m > 4096) return lp_fail(ANM_LP_E_INVALID, "n, m must be in 1..4096 (got %d, %d)", n, m);
  if (batch <= 0 || stride < batch) return lp_fail(ANM_LP_E_INVALID, "need 0 < batch <= stride (got %lld, %lld)", (long long)batch, (long long)stride);
  return 0;
}

}  // namespace

struct anm_lp_batch {
  anm_lp::Batch b;
  anm_lp::WarpBatch w;
  int mode;  // ANM_LP_KERNEL_THREAD / ANM_LP_KERNEL_WARP
  int64_t batch, bytes;
  int device;
  bool first;
  char* state;   // device: tableaux, bases, scratch
  double* consts;  // device: A then c
};

extern "C" {

const char* anm_lp_last_error(void) { return g_lp_err; }

int anm_lp_create(int32_t n, int32_t m, const double* a_host, const double* c_host, int64_t batch, int64_t stride,
                  int32_t max_iter, int device, anm_lp_handle* out) {
  if (!out || !a_host || !c_host) return lp_fail(ANM_LP_E_INVALID, "null argument");
  *out = nullptr;
  if (int rc = check_dims(n, m, batch, stride)) return rc;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
    return lp_fail(ANM_LP_E_CUDA, "no usable CUDA device %d (%d visible): the LP solver has no CPU fallback", device, count);
  LpDeviceGuard guard(device);
  anm_lp_batch* h = new anm_lp_batch();
  memset(&h->b, 0, sizeof h->b);
  memset(&h->w, 0, sizeof h->w);
  const char* km = getenv("ANM_LP_KERNEL");  // "thread" | "warp" (A/B switch; default below)
  h->mode = ANM_LP_KERNEL_DEFAULT;
  if (km && !strcmp(km, "thread")) h->mode = ANM_LP_KERNEL_THREAD;
  if (km && !strcmp(km, "warp")) h->mode = ANM_LP_KERNEL_WARP;
  const Layout L = layout_of(n, m, stride);
  const int64_t wblock = anm_lp::warp_block_doubles(n, m);
  const int64_t total = h->mode == ANM_LP_KERNEL_WARP ? wblock * batch * 8 : L.total;
  h->batch = batch, h->device = device, h->first = true, h->bytes = total + (int64_t)(m * n + n) * 8;
  h->state = nullptr, h->consts = nullptr;
  if (cudaMalloc(&h->state, total) != cudaSuccess || cudaMalloc(&h->consts, (size_t)(m * n + n) * 8) != cudaSuccess) {
    cudaGetLastError();
    if (h->state) cudaFree(h->state);
    delete h;
    return lp_fail(ANM_LP_E_NOMEM, "cudaMalloc of %lld bytes failed", (long long)total);
  }
  cudaError_t e1 = cudaMemcpy(h->consts, a_host, (size_t)m * n * 8, cudaMemcpyHostToDevice);
  cudaError_t e2 = cudaMemcpy(h->consts + (size_t)m * n, c_host, (size_t)n * 8, cudaMemcpyHostToDevice);
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    cudaFree(h->state), cudaFree(h->consts);
    delete h;
    return lp_fail(ANM_LP_E_CUDA, "upload of A / c failed");
  }
  h->b.n = n, h->b.m = m, h->b.stride = stride;
  h->b.max_iter = max_iter > 0 ? max_iter : 4 * (n + m) + 50;
  h->b.A = h->consts, h->b.c = h->consts + (size_t)m * n;
  bind(h->b, h->state, L);
  h->w.n = n, h->w.m = m, h->w.stride = stride, h->w.max_iter = h->b.max_iter, h->w.block = wblock;
  h->w.A = h->b.A, h->w.c = h->b.c, h->w.mem = reinterpret_cast<double*>(h->state);
  *out = h;
  return ANM_LP_OK;
}

int anm_lp_destroy(anm_lp_handle h) {
  if (!h) return ANM_LP_OK;
  LpDeviceGuard guard(h->device);
  cudaFree(h->state);
  cudaFree(h->consts);
  delete h;
  return ANM_LP_OK;
}

int64_t anm_lp_bytes(anm_lp_handle h) { return h ? h->bytes : 0; }

int anm_lp_kernel(anm_lp_handle h) { return h ? h->mode : ANM_LP_E_INVALID; }

int anm_lp_solve(anm_lp_handle h, const double* lo_dev, const double* up_dev, const uint8_t* restart_dev_or_null,
                 double* x_dev, double* obj_dev_or_null, int32_t* status_dev_or_null, int32_t* iters_dev_or_null,
                 void* stream) {
  if (!h || !lo_dev || !up_dev || !x_dev) return lp_fail(ANM_LP_E_INVALID, "null argument");
  ANM_LP_NVTX("anm_lp_solve");
  LpDeviceGuard guard(h->device);
  const int threads = 32;
  if (h->mode == ANM_LP_KERNEL_WARP) {  // one warp (= one block) per program
    anm_lp::lp_solve_warp_kernel<<<(unsigned)h->batch, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        h->w, h->batch, lo_dev, up_dev, restart_dev_or_null, h->first ? 1 : 0, x_dev, obj_dev_or_null,
        status_dev_or_null, iters_dev_or_null);
  } else {  // one thread per program
    const unsigned blocks = (unsigned)((h->batch + threads - 1) / threads);
    anm_lp::lp_solve_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        h->b, h->batch, lo_dev, up_dev, restart_dev_or_null, h->first ? 1 : 0, x_dev, obj_dev_or_null,
        status_dev_or_null, iters_dev_or_null);
  }
  LP_CUDA(cudaGetLastError());
  h->first = false;
  return ANM_LP_OK;
}

int64_t anm_debug_lp_state_bytes(int32_t n, int32_t m, int64_t stride) {
  if (n <= 0 || m <= 0 || stride <= 0) return 0;
  return layout_of(n, m, stride).total;
}

int anm_debug_lp_solve_host(int32_t n, int32_t m, const double* a_host, const double* c_host, int64_t batch,
                            int64_t stride, int32_t max_iter, void* state_host, int32_t first, const double* lo,
                            const double* up, const uint8_t* restart_or_null, double* x, double* obj_or_null,
                            int32_t* status_or_null, int32_t* iters_or_null) {
  if (!a_host || !c_host || !state_host || !lo || !up || !x) return lp_fail(ANM_LP_E_INVALID, "null argument");
  if (int rc = check_dims(n, m, batch, stride)) return rc;
  anm_lp::Batch b;
  memset(&b, 0, sizeof b);
  b.n = n, b.m = m, b.stride = stride, b.max_iter = max_iter > 0 ? max_iter : 4 * (n + m) + 50;
  b.A = a_host, b.c = c_host;
  bind(b, static_cast<char*>(state_host), layout_of(n, m, stride));
  for (int64_t e = 0; e < batch; ++e) {
    double z;
    int32_t it;
    const bool rs = first || (restart_or_null && restart_or_null[e]);
    const int st = anm_lp::solve_one(b, e, lo, up, rs, x, &z, &it);
    if (obj_or_null) obj_or_null[e] = z;
    if (status_or_null) status_or_null[e] = st;
    if (iters_or_null) iters_or_null[e] = it;
  }
  return ANM_LP_OK;
}

int64_t anm_debug_lp_warp_state_bytes(int32_t n, int32_t m, int64_t batch) {
  if (n <= 0 || m <= 0 || batch <= 0) return 0;
  return anm_lp::warp_block_doubles(n, m) * batch * 8;
}

int anm_debug_lp_solve_host_warp(int32_t n, int32_t m, const double* a_host, const double* c_host, int64_t batch,
                                 int64_t stride, int32_t max_iter, void* state_host, int32_t first, const double* lo,
                                 const double* up, const uint8_t* restart_or_null, double* x, double* obj_or_null,
                                 int32_t* status_or_null, int32_t* iters_or_null, int32_t reverse_lanes) {
  if (!a_host || !c_host || !state_host || !lo || !up || !x) return lp_fail(ANM_LP_E_INVALID, "null argument");
  if (int rc = check_dims(n, m, batch, stride)) return rc;
  anm_lp::WarpBatch w;
  memset(&w, 0, sizeof w);
  w.n = n, w.m = m, w.stride = stride, w.max_iter = max_iter > 0 ? max_iter : 4 * (n + m) + 50;
  w.block = anm_lp::warp_block_doubles(n, m);
  w.A = a_host, w.c = c_host, w.mem = static_cast<double*>(state_host);
  double tile[32 * 33], rv[32], col[anm_lp::kColScratch];
  int32_t ri[32], rows[anm_lp::kColScratch];
  anm_lp::WarpScratch s{tile, rv, ri, col, rows, reverse_lanes != 0};
  for (int64_t e = 0; e < batch; ++e) {
    double z;
    int32_t it;
    const bool rs = first || (restart_or_null && restart_or_null[e]);
    const int st = anm_lp::solve_one_warp(w, e, lo, up, rs, x, &z, &it, s);
    if (obj_or_null) obj_or_null[e] = z;
    if (status_or_null) status_or_null[e] = st;
    if (iters_or_null) iters_or_null[e] = it;
  }
  return ANM_LP_OK;
}

}  // extern "C"
