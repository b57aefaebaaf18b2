/*
 * anm_capi.cu -- host side of the C ABI declared in include/anm_b200.h.
 *
 * Builds the per-network constant blob (layout: anm_layout.h) from the caller's
 * anm_network_desc / anm_env_desc, owns the carried per-env state (SoC, aux, terminated)
 * and launches the fused step kernel (anm_kernels.cuh).  No torch types, no exceptions
 * across the ABI, no host synchronisation except in the *_host convenience calls.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <vector>

#include "anm_kernels.cuh"
#include "anm_seeded_reset.cuh"

/* NVTX ranges around every entry point that enqueues device work (reset / step / rollout / host copies), so that a
 * timeline (Nsight Systems, or anything that injects NVTX) shows the calls by name; header-only NVTX3, a no-op when no
 * tool is attached.  -DANM_NO_NVTX compiles them out. */
#ifndef ANM_NO_NVTX
#include <nvtx3/nvToolsExt.h>
namespace {
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace
#define ANM_NVTX(name) NvtxRange nvtx_range_(name)
#else
#define ANM_NVTX(name)
#endif

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess) return fail(ANM_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e_));   \
  } while (0)

typedef std::complex<double> cd;
#ifndef ANM_RESET_ROUNDS
#define ANM_RESET_ROUNDS 4 /* retry rounds of anm_reset_seeded enqueued per host synchronisation */
#endif

struct BlobBuilder {
  std::vector<unsigned char> buf;
  explicit BlobBuilder(size_t header) : buf((header + 15) / 16 * 16, 0) {}
  template <typename T>
  int32_t add(const std::vector<T>& v) {
    size_t off = buf.size();
    size_t bytes = (v.size() * sizeof(T) + 15) / 16 * 16;
    if (bytes == 0) bytes = 16;
    buf.resize(off + bytes, 0);
    if (!v.empty()) memcpy(buf.data() + off, v.data(), v.size() * sizeof(T));
    return (int32_t)off;
  }
};

int full_offset(const AnmConstHeader& H, int quantity) {
  const int N = H.n_bus, D = H.n_dev, L = H.n_branch;
  if (quantity <= ANM_Q_BUS_I_ANG) return quantity * N;
  if (quantity <= ANM_Q_GEN_P_MAX) return 6 * N + (quantity - ANM_Q_DEV_P) * D;
  if (quantity <= ANM_Q_BRANCH_I_ANG) return 6 * N + 4 * D + (quantity - ANM_Q_BRANCH_P) * L;
  return 6 * N + 4 * D + 5 * L;
}

int var_limit(const AnmConstHeader& H, int quantity) {
  if (quantity <= ANM_Q_BUS_I_ANG) return H.n_bus;
  if (quantity <= ANM_Q_GEN_P_MAX) return H.n_dev;
  if (quantity <= ANM_Q_BRANCH_I_ANG) return H.n_branch;
  return H.K;
}

}  // namespace

namespace anm {
/* test hook (anm_debug_math): the kernel's own fast math routines evaluated element-wise */
__global__ void debug_math_kernel(int kind, int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                                  double* __restrict__ a, double* __restrict__ b) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (kind == 0) {
    double sn, cs;
    sincos_fast(x[i], sn, cs);
    a[i] = sn;
    b[i] = cs;
  } else if (kind == 1) {
    a[i] = cabs2(x[i], y[i]);
  } else {
    a[i] = fast_rcp(x[i]);
  }
}
}  // namespace anm

namespace anm {
/* fp64 FMA throughput of the device (the denominator of bench.py's fp64 roofline leg): eight independent DFMA
 * chains per thread, no memory traffic */
__global__ void fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 12345.678) out[0] = s; /* keeps the chains alive */
}
}  // namespace anm

struct anm_handle_s {
  int device = 0;
  int64_t B = 0;
  AnmConstHeader H;
  unsigned char* d_blob = nullptr;
  int blob_bytes = 0;
  double* d_soc = nullptr;
  double* d_aux = nullptr;
  uint8_t* d_term = nullptr;
  uint32_t* d_episode = nullptr;
  uint32_t* d_seq = nullptr; /* [B] per-instance launch ordinals (launch chaining, anm_kernels.cuh) */
  unsigned long long* d_ticket = nullptr; /* [16] (its own 128-byte line) CTAs started so far on this handle */
  AnmPcg64* d_rng = nullptr; /* [B] per-instance random streams (anm_seed / anm_reset_seeded) */
  bool seeded = false;
  uint8_t *d_need = nullptr, *d_conv_tmp = nullptr; /* seeded reset: still looking for an initial state / last attempt */
  int *d_remaining = nullptr, *h_remaining = nullptr;
  uint32_t* wd_host = nullptr; /* [ANM_WD_WORDS] mapped host memory: chaining watchdog record */
  uint32_t* wd_dev = nullptr;
  const double* pool = nullptr;
  int64_t pool_size = 0;
  double* reset_full = nullptr; /* optional [B, n_full] output of every reset launch (anm_set_reset_full_state) */
  double* d_dense_scratch = nullptr; /* block-sparse solver: [grid x gpb][M x (M+1)] (singular-block guard's redo) */
  /* fused observation all-gather (anm_gather_*): this rank's buffer [flags 1 KB | slots x rows x (O + 2) doubles], the
   * peers' buffers opened over CUDA IPC, and the device arrays of pointers the kernel reads */
  struct Gather {
    int world = 0, rank = 0, slots = 0;
    int64_t rows = 0, row0 = 0;
    unsigned char* base = nullptr;            /* own allocation */
    std::vector<void*> peer_base;             /* [world] mapped base pointers (own: base) */
    double** d_peers = nullptr;               /* device: [world] data regions */
    unsigned long long** d_flags = nullptr;   /* device: [world] flag regions */
    unsigned long long* d_state = nullptr;    /* device: [0] steps completed, [8..72) CTA counters */
    int64_t steps_host = 0;                   /* gather steps launched so far (host mirror of d_state[0]) */
    bool attached = false;
    bool last_waited_inline = false;          /* the most recent gather step waited for the arrival inside its kernel */
  } g;
  /* launch geometry */
  int lpe = 32, gpb = 4, grid = 1, smem = 0, num_sms = 1;
  /* host-buffer path */
  cudaStream_t stream = nullptr;
  double *s_action = nullptr, *s_nv = nullptr, *s_obs = nullptr, *s_reward = nullptr, *s_s0 = nullptr, *s_state = nullptr;
  uint8_t *s_term = nullptr, *s_mask = nullptr;
  /* queued host rollouts (anm_rollout_host_async): two sets of device staging buffers, inputs uploaded and outputs
   * downloaded by the copy engines on their own streams while the compute stream runs the kernels back to back */
  struct HostSet {
    double *act = nullptr, *nv = nullptr, *obs = nullptr, *rew = nullptr;
    uint8_t* term = nullptr;
    int64_t cap_rows = 0; /* T * B rows allocated */
    cudaEvent_t ev_in = nullptr, ev_k = nullptr, ev_out = nullptr;
    bool used = false;
  } hset[2];
  int hset_next = 0;
  cudaStream_t st_in = nullptr, st_out = nullptr;
  int64_t wd_last = 0;  /* the most recent launch's own share of wd_steps */
  int64_t wd_steps = 0; /* steps x passes of every launch since the last fully ordered one: a chained launch may have
                           to wait for all of them (chained waits are transitive) -- the watchdog limit of the next */
  bool st_last_was_rollout = false; /* the compute stream's last operation was one of our rollout kernels */
  int64_t launches = 0;
};

namespace {

/* Solver selection (anm_kernels.cuh), decided when the handle is created; ANM_SOLVER (environment) overrides:
 *   radial  (default for tree networks <= 9 buses)   2x2-block elimination along the tree, one bus per lane  RadialNR
 *   dense   (default for meshed networks <= 9 buses) Jacobian rows in registers, natural-order Gauss-Jordan   SmallNR
 *   sparse  (default for > 9 buses)                  block-sparse LU on the filled Y-bus pattern, host-side
 *                                                    symbolic factorisation                                   nr_sparse
 *   generic                                          dense Jacobian in shared memory, partial pivoting        nr_generic
 * B200, ANM6Easy bench: radial 0.147 ms/step, dense 0.198 ms/step; synthetic 30-bus: see profiles/.
 * ANM_FORCE_GENERIC=1 / ANM_FORCE_DENSE=1 are kept as aliases (tests). */
static int solver_env() { /* -1: no override */
  static const int v = [] {
    const char* e = getenv("ANM_SOLVER");
    if (e && !strcmp(e, "generic")) return 0;
    if (e && !strcmp(e, "dense")) return 1;
    if (e && !strcmp(e, "radial")) return 2;
    if (e && !strcmp(e, "sparse")) return 4;
    if (getenv("ANM_FORCE_GENERIC") && atoi(getenv("ANM_FORCE_GENERIC")) != 0) return 0;
    if (getenv("ANM_FORCE_DENSE") && atoi(getenv("ANM_FORCE_DENSE")) != 0) return 1;
    return -1;
  }();
  return v;
}
static int choose_solver(int n_bus, bool is_radial) {
  const int want = solver_env();
  if (want == 0 || want == 4) return want;
  if (n_bus > 9) return 4;
  if (want != 1 && is_radial) return 2;
  return 1;
}
static int solver_for(const AnmConstHeader& H) { return H.solver; }

static int lanes_for(const AnmConstHeader& H); /* below, with the kernel table */

/* The candidates of the exact projection on the polygon {a_k x + b_k y <= h_k, k < R} (project_polygon,
 * anm_kernels.cuh) as affine maps of (p, q, h[s1], h[s2]): the point itself, its projection on each row's line, each
 * pairwise intersection, in the order of oracle/shims/cvxpy/_projection.py; parallel pairs are dropped.  Appends
 * info = s1 | s2 << 8 | need << 16 and the eight coefficients kx[4], ky[4] of every candidate.
 *
 * Pruning (exact: no candidate that can ever win is dropped).  Rows whose bit is not in `dyn_mask` are STATIC: their
 * right-hand side h_static[k] never changes (+-inf = unused).  The intersection of two static rows is a fixed point;
 * if it violates a third static row by more than 1e-9 it is outside the polygon for every step and is dropped (for
 * ANM6Easy: 25 -> 18 candidates per generator, 47 -> 31 for the storage unit).  `dom_row` >= 0: a static row that
 * the dynamic row `dom_by` (same normal) always dominates when it is finite -- a generator's p <= p_max against
 * p <= p_pot, p_pot being clipped to p_max (simulator.py:511): the candidates on `dom_row` go last, and the kernel
 * leaves them out (n_short candidates) whenever `dom_by` is finite; they would be the same points, bit for bit. */
static void build_candidates(const double* a, const double* b, int R, std::vector<int>& cinfo, std::vector<double>& ccoef,
                             unsigned dyn_mask = ~0u, const double* h_static = nullptr, int dom_row = -1,
                             int* n_short = nullptr) {
  std::vector<int> tail_info;
  std::vector<double> tail_coef;
  const size_t first = cinfo.size();
  auto push = [&](int s1, int s2, unsigned need, const double (&kx)[4], const double (&ky)[4]) {
    const bool last = dom_row >= 0 && ((need >> dom_row) & 1u);
    std::vector<int>& ci = last ? tail_info : cinfo;
    std::vector<double>& cc = last ? tail_coef : ccoef;
    ci.push_back(s1 | (s2 << 8) | (int)(need << 16));
    cc.insert(cc.end(), kx, kx + 4);
    cc.insert(cc.end(), ky, ky + 4);
  };
  auto is_static = [&](int k) { return h_static && !((dyn_mask >> k) & 1u); };
  push(0, 0, 0u, {1, 0, 0, 0}, {0, 1, 0, 0});
  for (int k = 0; k < R; ++k) {
    if (a[k] == 0.0 && b[k] == 0.0) continue;
    if (is_static(k) && !std::isfinite(h_static[k])) continue; /* a static row that is never used */
    if (b[k] == 0.0) {
      push(k, k, 1u << k, {0, 0, 1.0 / a[k], 0}, {0, 1, 0, 0});
    } else if (a[k] == 0.0) {
      push(k, k, 1u << k, {1, 0, 0, 0}, {0, 0, 1.0 / b[k], 0});
    } else {
      const double w = 1.0 / (a[k] * a[k] + b[k] * b[k]);
      push(k, k, 1u << k, {1.0 - a[k] * a[k] * w, -a[k] * b[k] * w, a[k] * w, 0},
           {-a[k] * b[k] * w, 1.0 - b[k] * b[k] * w, b[k] * w, 0});
    }
  }
  for (int j = 1; j < R; ++j)
    for (int i = 0; i < j; ++i) {
      const double det = a[i] * b[j] - a[j] * b[i];
      if (det == 0.0) continue;
      if ((is_static(i) && !std::isfinite(h_static[i])) || (is_static(j) && !std::isfinite(h_static[j]))) continue;
      if (is_static(i) && is_static(j)) { /* a fixed point: outside the static polygon for good? */
        const double x = (b[j] * h_static[i] - b[i] * h_static[j]) / det, y = (a[i] * h_static[j] - a[j] * h_static[i]) / det;
        bool outside = false;
        for (int k = 0; k < R && !outside; ++k) {
          if (k == i || k == j || !is_static(k) || !std::isfinite(h_static[k])) continue;
          const double scale = std::max(1.0, std::max(std::fabs(h_static[k]), std::fabs(a[k] * x) + std::fabs(b[k] * y)));
          outside = (a[k] * x + b[k] * y - h_static[k]) > 1e-9 * scale;
        }
        if (outside) continue;
      }
      push(i, j, (1u << i) | (1u << j), {0, 0, b[j] / det, -b[i] / det}, {0, 0, -a[j] / det, a[i] / det});
    }
  if (n_short) *n_short = (int)(cinfo.size() - first);
  cinfo.insert(cinfo.end(), tail_info.begin(), tail_info.end());
  ccoef.insert(ccoef.end(), tail_coef.begin(), tail_coef.end());
}

int build_blob(const anm_network_desc* net, const anm_env_desc* env, AnmConstHeader& H, std::vector<unsigned char>& out) {
  const int N = net->n_bus, D = net->n_dev, L = net->n_branch, K = env->K;
  if (N < 2 || N > 64) return fail(ANM_E_UNSUPPORTED, "n_bus=%d not in [2, 64]", N);
  if (D < 1 || D > 255 || L < 1 || L > 1024) return fail(ANM_E_UNSUPPORTED, "n_dev=%d / n_branch=%d out of range", D, L);
  if (K < 0 || K > 16) return fail(ANM_E_UNSUPPORTED, "K=%d not in [0, 16]", K);
  if (!(net->base_mva > 0) || !(net->delta_t > 0)) return fail(ANM_E_INVALID, "baseMVA and delta_t must be > 0");
  memset(&H, 0, sizeof(H));
  H.n_bus = N; H.n_dev = D; H.n_branch = L; H.K = K;
  H.n_unk = 2 * (N - 1);
  H.base_mva = net->base_mva; H.delta_t = net->delta_t; H.lamb = net->lamb; H.gamma = env->gamma;
  H.clip_e = env->clip_e_loss; H.clip_pen = env->clip_penalty;
  H.term_reward = -env->clip_penalty / (1.0 - env->gamma); /* anm_env.py:430 */
  H.nr_maxit = ANM_NR_MAXIT;
  if (const char* e = getenv("ANM_DEBUG_NR_MAXIT")) { /* single-iteration tests of the solvers; never set in production */
    const int v = atoi(e);
    if (v >= 0 && v < ANM_NR_MAXIT) H.nr_maxit = v;
  }

  std::vector<int> dev_bus(D), dev_type(D), dev_slot(D);
  std::vector<int> gens, dess;
  int n_slack = 0;
  for (int d = 0; d < D; ++d) {
    dev_bus[d] = net->dev_bus[d];
    dev_type[d] = net->dev_type[d];
    if (dev_bus[d] < 0 || dev_bus[d] >= N) return fail(ANM_E_INVALID, "device %d: bad bus %d", d, dev_bus[d]);
    switch (dev_type[d]) {
      case ANM_DEV_LOAD: dev_slot[d] = H.n_load++; break;
      case ANM_DEV_GEN:
      case ANM_DEV_RENEWABLE: dev_slot[d] = H.n_gen++; gens.push_back(d); break;
      case ANM_DEV_STORAGE: dev_slot[d] = H.n_des++; dess.push_back(d); break;
      case ANM_DEV_SLACK:
        dev_slot[d] = 0; ++n_slack;
        if (dev_bus[d] != 0) return fail(ANM_E_INVALID, "the slack device must sit at bus 0 (solve_load_flow.py:64)");
        break;
      default: return fail(ANM_E_INVALID, "device %d: unknown DEV_TYPE %d", d, dev_type[d]);
    }
  }
  if (n_slack != 1) return fail(ANM_E_INVALID, "exactly one slack device expected, got %d", n_slack);
  H.n_ctrl = H.n_gen + H.n_des;
  H.n_action = 2 * H.n_ctrl;
  H.n_next_vars = H.n_load + H.n_gen + K;
  H.n_full = 6 * N + 4 * D + 5 * L + K;
  H.n_state = env->n_state; H.n_obs = env->n_obs;
  if (H.n_state != 2 * D + H.n_des + H.n_gen + K)
    return fail(ANM_E_INVALID, "n_state=%d but 2*n_dev+n_des+n_gen+K=%d (anm_env.py:274)", H.n_state,
                2 * D + H.n_des + H.n_gen + K);
  if (H.n_obs < 1) return fail(ANM_E_INVALID, "n_obs must be >= 1");
  H.table_len = env->table_len;
  if (H.table_len < 0 || (H.table_len > 0 && (K < 1 || !env->table)))
    return fail(ANM_E_INVALID, "built-in next_vars table needs K >= 1 and a table pointer");

  BlobBuilder bb(sizeof(AnmConstHeader));
  H.o_vmin = bb.add(std::vector<double>(net->bus_vmin, net->bus_vmin + N));
  H.o_vmax = bb.add(std::vector<double>(net->bus_vmax, net->bus_vmax + N));
  H.o_dev_bus = bb.add(dev_bus);
  H.o_dev_type = bb.add(dev_type);
  H.o_dev_slot = bb.add(dev_slot);
  H.o_dev_param = bb.add(std::vector<double>(net->dev_param, net->dev_param + (size_t)D * ANM_DEV_NPARAM));
  {
    std::vector<int> ptr(N + 1, 0), idx;
    for (int b = 0; b < N; ++b) {
      for (int d = 0; d < D; ++d)
        if (dev_bus[d] == b) idx.push_back(d);
      ptr[b + 1] = (int)idx.size();
    }
    H.o_bus_dev_ptr = bb.add(ptr);
    H.o_bus_dev_idx = bb.add(idx);
  }
  {
    std::vector<int> bf(L), bt(L);
    std::vector<double> coef((size_t)L * 10, 0.0);
    for (int l = 0; l < L; ++l) {
      bf[l] = net->br_from[l]; bt[l] = net->br_to[l];
      if (bf[l] < 0 || bf[l] >= N || bt[l] < 0 || bt[l] >= N) return fail(ANM_E_INVALID, "branch %d: bad endpoints", l);
      const double* B = net->br_param + (size_t)l * ANM_BR_NPARAM;
      const cd y(B[ANM_BP_SERIES_RE], B[ANM_BP_SERIES_IM]), ysh(B[ANM_BP_SHUNT_RE], B[ANM_BP_SHUNT_IM]);
      const cd tap(B[ANM_BP_TAP_RE], B[ANM_BP_TAP_IM]);
      const double t2 = std::abs(tap) * std::abs(tap);
      const cd aff = (y + ysh) / t2, aft = -y / std::conj(tap), atf = -y / tap, att = y + ysh; /* branch.py:165-173 */
      double* c = &coef[(size_t)l * 10];
      c[0] = aff.real(); c[1] = aff.imag(); c[2] = aft.real(); c[3] = aft.imag();
      c[4] = atf.real(); c[5] = atf.imag(); c[6] = att.real(); c[7] = att.imag();
      c[8] = B[ANM_BP_RATE];
    }
    H.o_br_from = bb.add(bf);
    H.o_br_to = bb.add(bt);
    H.o_br_coef = bb.add(coef);
  }
  {
    std::vector<int> ptr(N + 1, 0), col, jr, jc, jy;
    std::vector<double> val;
    for (int i = 0; i < N; ++i) {
      for (int j = 0; j < N; ++j) {
        const double re = net->ybus[2 * ((size_t)i * N + j)], im = net->ybus[2 * ((size_t)i * N + j) + 1];
        if (i != j && re == 0.0 && im == 0.0) continue;
        if (i >= 1 && j >= 1) { jr.push_back(i); jc.push_back(j); jy.push_back((int)col.size()); }
        col.push_back(j);
        val.push_back(re);
        val.push_back(im);
      }
      ptr[i + 1] = (int)col.size();
    }
    H.y_nnz = (int)col.size();
    H.n_jac = (int)jr.size();
    H.o_y_ptr = bb.add(ptr);
    H.o_y_col = bb.add(col);
    H.o_y_val = bb.add(val);
    H.o_y_dense = bb.add(std::vector<double>(net->ybus, net->ybus + (size_t)2 * N * N));
    H.o_jac_row = bb.add(jr);
    H.o_jac_col = bb.add(jc);
    H.o_jac_y = bb.add(jy);
  }
  {
    const double inf = std::numeric_limits<double>::infinity();
    std::vector<int> ctrl;
    ctrl.insert(ctrl.end(), gens.begin(), gens.end());
    ctrl.insert(ctrl.end(), dess.begin(), dess.end());
    std::vector<double> rows((size_t)H.n_ctrl * 3 * ANM_MAX_ROWS, 0.0);
    for (int c = 0; c < H.n_ctrl; ++c) {
      const double* P = net->dev_param + (size_t)ctrl[c] * ANM_DEV_NPARAM;
      double* a = &rows[(size_t)c * 3 * ANM_MAX_ROWS];
      double* b = a + ANM_MAX_ROWS;
      double* h = b + ANM_MAX_ROWS;
      for (int k = 0; k < ANM_MAX_ROWS; ++k) h[k] = inf;
      if (c < H.n_gen) { /* Generator.map_pq rows, devices.py:294-296 */
        const double A[7][3] = {{-1, 0, -P[ANM_DP_PMIN]}, {1, 0, P[ANM_DP_PMAX]}, {1, 0, P[ANM_DP_PMAX]},
                                {0, -1, -P[ANM_DP_QMIN]}, {0, 1, P[ANM_DP_QMAX]},
                                {-P[ANM_DP_TAU1], 1, P[ANM_DP_RHO1]}, {P[ANM_DP_TAU2], -1, -P[ANM_DP_RHO2]}};
        for (int k = 0; k < 7; ++k) a[k] = A[k][0], b[k] = A[k][1], h[k] = A[k][2];
      } else { /* StorageUnit.map_pq rows, devices.py:486-514 (rows 8, 9 depend on the SoC) */
        const double A[10][3] = {{-1, 0, -P[ANM_DP_PMIN]}, {1, 0, P[ANM_DP_PMAX]}, {0, -1, -P[ANM_DP_QMIN]},
                                 {0, 1, P[ANM_DP_QMAX]}, {-P[ANM_DP_TAU1], 1, P[ANM_DP_RHO1]},
                                 {P[ANM_DP_TAU2], -1, -P[ANM_DP_RHO2]}, {P[ANM_DP_TAU3], -1, -P[ANM_DP_RHO3]},
                                 {-P[ANM_DP_TAU4], 1, P[ANM_DP_RHO4]}, {-1, 0, 0.0}, {1, 0, 0.0}};
        for (int k = 0; k < 10; ++k) a[k] = A[k][0], b[k] = A[k][1], h[k] = A[k][2];
      }
    }
    H.o_ctrl_dev = bb.add(ctrl);
    H.o_ctrl_rows = bb.add(rows);
    {
      std::vector<int> fin(H.n_ctrl, 0);
      for (int c = 0; c < H.n_ctrl; ++c) {
        const double* h = &rows[(size_t)c * 3 * ANM_MAX_ROWS + 2 * ANM_MAX_ROWS];
        for (int k = 0; k < ANM_MAX_ROWS; ++k) {
          const bool dynamic = (c < H.n_gen) ? (k == 2) : (k == 8 || k == 9); /* rewritten every step by the kernel */
          if (!dynamic && std::isfinite(h[k])) fin[c] |= 1 << k;
        }
      }
      H.o_ctrl_fin = bb.add(fin);
    }
    /* candidate_table: build_candidates() for every controllable device */
    std::vector<int> cptr(1, 0), cinfo, cshort;
    std::vector<double> ccoef;
    const bool prune = !(getenv("ANM_CAND_PRUNE") && atoi(getenv("ANM_CAND_PRUNE")) == 0); /* A/B switch */
    for (int c = 0; c < H.n_ctrl; ++c) {
      const double* a = &rows[(size_t)c * 3 * ANM_MAX_ROWS];
      const bool gen = c < H.n_gen;
      const double* P = net->dev_param + (size_t)ctrl[c] * ANM_DEV_NPARAM;
      int ns_ = 0;
      /* generator: row 2 (p <= p_pot, dynamic) dominates row 1 (p <= p_max) whenever it is finite, because the
       * kernel clips p_pot to [p_min, p_max] (needs a finite p_max: otherwise row 1 is unused anyway) */
      const int dom = (prune && gen && std::isfinite(P[ANM_DP_PMAX])) ? 1 : -1;
      if (prune) build_candidates(a, a + ANM_MAX_ROWS, gen ? 7 : 10, cinfo, ccoef, gen ? 4u : 0x300u, a + 2 * ANM_MAX_ROWS, dom, &ns_);
      else build_candidates(a, a + ANM_MAX_ROWS, gen ? 7 : 10, cinfo, ccoef, ~0u, nullptr, -1, &ns_);
      cptr.push_back((int)cinfo.size());
      cshort.push_back(ns_);               /* candidates to evaluate when ... */
      cshort.push_back(dom >= 0 ? 4 : 0);  /* ... this finite-row bit (the dominating dynamic row) is set */
    }
    H.o_cand_short = bb.add(cshort);
    H.o_cand_ptr = bb.add(cptr);
    H.o_cand_info = bb.add(cinfo);
    H.o_cand_coef = bb.add(ccoef);
  }
  {
    auto pack = [&](int n, const anm_var_spec* v, std::vector<int>& off, std::vector<double>& mul,
                    std::vector<double>& dv, std::vector<double>& lo, std::vector<double>& hi) -> int {
      for (int k = 0; k < n; ++k) {
        if (v[k].quantity < 0 || v[k].quantity >= ANM_NQUANTITY) return fail(ANM_E_INVALID, "var %d: bad quantity", k);
        if (v[k].index < 0 || v[k].index >= var_limit(H, v[k].quantity)) return fail(ANM_E_INVALID, "var %d: bad index", k);
        if (v[k].quantity == ANM_Q_BUS_V_ANG || v[k].quantity == ANM_Q_BUS_I_ANG || v[k].quantity == ANM_Q_BRANCH_I_ANG)
          H.need_angles = 1, H.need_mask |= ANM_NEED_ANGLES;
        if (v[k].quantity <= ANM_Q_BUS_I_ANG) H.need_mask |= ANM_NEED_BUS;
        if (v[k].quantity >= ANM_Q_BRANCH_P && v[k].quantity <= ANM_Q_BRANCH_I_ANG) H.need_mask |= ANM_NEED_BRANCH;
        if (v[k].quantity == ANM_Q_BUS_V_MAGN) H.need_mask |= ANM_NEED_BUS_V;
        if (v[k].quantity == ANM_Q_BUS_I_MAGN) H.need_mask |= ANM_NEED_BUS_I;
        if (v[k].quantity == ANM_Q_BRANCH_I_MAGN) H.need_mask |= ANM_NEED_BRANCH_I;
        off.push_back(full_offset(H, v[k].quantity) + v[k].index);
        mul.push_back(v[k].mul); dv.push_back(v[k].div); lo.push_back(v[k].low); hi.push_back(v[k].high);
      }
      return 0;
    };
    std::vector<int> so, oo;
    std::vector<double> sm, sd, sl, sh, om, od, ol, oh;
    if (int rc = pack(env->n_state, env->state_vars, so, sm, sd, sl, sh)) return rc;
    if (int rc = pack(env->n_obs, env->obs_vars, oo, om, od, ol, oh)) return rc;
    H.o_sv_off = bb.add(so); H.o_sv_mul = bb.add(sm); H.o_sv_div = bb.add(sd);
    H.o_ov_off = bb.add(oo); H.o_ov_mul = bb.add(om); H.o_ov_div = bb.add(od);
    H.o_ov_low = bb.add(ol); H.o_ov_high = bb.add(oh);
  }
  {
    std::vector<double> table;
    if (H.table_len > 0) table.assign(env->table, env->table + (size_t)H.table_len * (H.n_load + H.n_gen));
    H.o_table = bb.add(table);
    H.o_pair_i = H.o_pair_j = 0; /* superseded by the candidate table */
  }
  {
    /* Radial network?  The bus graph (branches) must be a tree rooted at the slack bus and no bus may have
     * more than ANM_RAD_MAXC children; then the Jacobian has the tree's block sparsity and the kernel can
     * eliminate it leaf-to-root (RadialNR in anm_kernels.cuh). */
    const int n = N - 1;
    std::vector<int> parent(N, -2), depth(N, 0), order;
    std::vector<std::vector<int>> adj(N);
    bool ok = (L == N - 1);
    for (int l = 0; l < L && ok; ++l) {
      const int f = net->br_from[l], t = net->br_to[l];
      if (f == t) ok = false;
      adj[f].push_back(t);
      adj[t].push_back(f);
    }
    if (ok) {
      parent[0] = -1;
      order.push_back(0);
      for (size_t q = 0; q < order.size(); ++q) {
        const int u = order[q];
        for (int v : adj[u])
          if (parent[v] == -2) { parent[v] = u; depth[v] = depth[u] + 1; order.push_back(v); }
      }
      ok = ((int)order.size() == N);
    }
    std::vector<int> rp(n > 0 ? n : 1, -1), rd(n > 0 ? n : 1, 0), rc((size_t)(n > 0 ? n : 1) * ANM_RAD_MAXC, -1);
    std::vector<double> ry((size_t)(n > 0 ? n : 1) * 6, 0.0);
    int maxc = 0, maxd = 0;
    if (ok) {
      std::vector<int> nchild(N, 0);
      for (int b = 1; b < N && ok; ++b) {
        const int p = parent[b];
        rp[b - 1] = (p == 0) ? -1 : p - 1;
        rd[b - 1] = depth[b];
        if (depth[b] > maxd) maxd = depth[b];
        if (p > 0) {
          if (nchild[p] >= ANM_RAD_MAXC) { ok = false; break; }
          rc[(size_t)(p - 1) * ANM_RAD_MAXC + nchild[p]++] = b - 1;
          if (nchild[p] > maxc) maxc = nchild[p];
        }
        const double* Y = net->ybus;
        auto y = [&](int i, int j, int c) { return Y[2 * ((size_t)i * N + j) + c]; };
        double* o = &ry[(size_t)(b - 1) * 6];
        o[0] = y(b, b, 0); o[1] = y(b, b, 1); o[2] = y(b, p, 0); o[3] = y(b, p, 1); o[4] = y(p, b, 0); o[5] = y(p, b, 1);
      }
      /* the Y-bus must not couple buses that are not tree neighbours */
      for (int i = 0; i < N && ok; ++i)
        for (int j = 0; j < N && ok; ++j) {
          if (i == j || parent[i] == j || parent[j] == i) continue;
          if (net->ybus[2 * ((size_t)i * N + j)] != 0.0 || net->ybus[2 * ((size_t)i * N + j) + 1] != 0.0) ok = false;
        }
    }
    H.is_radial = ok ? 1 : 0;
    H.rad_maxc = maxc;
    H.rad_maxdepth = maxd;
    H.o_rad_parent = bb.add(rp);
    H.o_rad_depth = bb.add(rd);
    H.o_rad_child = bb.add(rc);
    H.o_rad_y = bb.add(ry);
  }
  H.solver = choose_solver(N, H.is_radial != 0);
  {
    /* Symbolic factorisation for the block-sparse solver: unknowns grouped per non-slack bus (2x2 blocks);
     * greedy minimum-degree elimination order on the Y-bus graph, fill blocks added; per elimination step the
     * blocks of the pivot row (b, j), of the pivot column (i, b) and the degree^2 targets (i, j). */
    const int n = N - 1;
    auto coupled = [&](int i, int j) {
      return net->ybus[2 * ((size_t)i * N + j)] != 0.0 || net->ybus[2 * ((size_t)i * N + j) + 1] != 0.0 ||
             net->ybus[2 * ((size_t)j * N + i)] != 0.0 || net->ybus[2 * ((size_t)j * N + i) + 1] != 0.0;
    };
    std::vector<std::vector<char>> A(N, std::vector<char>(N, 0));
    for (int i = 1; i < N; ++i)
      for (int j = 1; j < N; ++j) A[i][j] = (i == j) || coupled(i, j);
    std::vector<int> bi, bj, by, diag(N, -1);
    std::vector<std::vector<int>> idx(N, std::vector<int>(N, -1));
    auto y_index = [&](int i, int j) { /* position of (i, j) in the CSR built above (diagonal always stored) */
      int k = 0;
      for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) {
          const double re = net->ybus[2 * ((size_t)r * N + c)], im = net->ybus[2 * ((size_t)r * N + c) + 1];
          if (r != c && re == 0.0 && im == 0.0) continue;
          if (r == i && c == j) return k;
          ++k;
        }
      return -1;
    };
    auto add_block = [&](int i, int j, bool original) {
      if (idx[i][j] >= 0) return;
      idx[i][j] = (int)bi.size();
      bi.push_back(i); bj.push_back(j);
      by.push_back(original ? y_index(i, j) : -1);
      if (i == j) diag[i] = idx[i][j];
    };
    for (int i = 1; i < N; ++i)
      for (int j = 1; j < N; ++j)
        if (A[i][j]) add_block(i, j, true);
    std::vector<char> gone(N, 0);
    std::vector<int> step, rows, cols, nbrs, tgts;
    for (int k = 0; k < n; ++k) {
      int best = -1, bestdeg = 1 << 30;
      for (int b = 1; b < N; ++b) {
        if (gone[b]) continue;
        int deg = 0;
        for (int j = 1; j < N; ++j) deg += (!gone[j] && j != b && A[b][j]);
        if (deg < bestdeg) bestdeg = deg, best = b;
      }
      const int b = best;
      std::vector<int> nb;
      for (int j = 1; j < N; ++j)
        if (!gone[j] && j != b && A[b][j]) nb.push_back(j);
      for (int i : nb)
        for (int j : nb) {
          if (!A[i][j]) A[i][j] = 1;
          add_block(i, j, false); /* no-op when the block exists */
        }
      step.push_back(b); step.push_back(idx[b][b]); step.push_back((int)nb.size());
      step.push_back((int)rows.size()); step.push_back((int)tgts.size());
      for (int j : nb) { rows.push_back(idx[b][j]); cols.push_back(idx[j][b]); nbrs.push_back(j); }
      for (int i : nb)
        for (int j : nb) tgts.push_back(idx[i][j]);
      gone[b] = 1;
    }
    H.sp_nblk = (int)bi.size();
    H.sp_nsteps = n;
    H.o_sp_blk_i = bb.add(bi); H.o_sp_blk_j = bb.add(bj); H.o_sp_blk_y = bb.add(by);
    H.o_sp_step = bb.add(step);
    H.o_sp_row = bb.add(rows); H.o_sp_col = bb.add(cols); H.o_sp_nbr = bb.add(nbrs);
    H.o_sp_tgt = bb.add(tgts);
    H.o_sp_diag = bb.add(diag);
  }
  /* per-env shared-memory workspace (doubles) */
  {
    int w = 0;
    auto take = [&](int n) { int o = w; w += (n + 1) / 2 * 2; return o; };
    const int M = H.n_unk;
    H.w_in_pl = take(H.n_load); H.w_in_pp = take(H.n_gen); H.w_in_ps = take(H.n_ctrl); H.w_in_qs = take(H.n_ctrl);
    H.w_soc = take(H.n_des); H.w_aux = take(K);
    H.w_devp = take(D); H.w_devq = take(D); H.w_ppot = take(D); H.w_busp = take(N); H.w_busq = take(N);
    H.w_x = take(M); H.w_vre = take(N); H.w_vim = take(N); H.w_ere = take(N); H.w_eim = take(N);
    H.w_ire = take(N); H.w_iim = take(N);
    /* RadialNR: two sets of exchange slots; its singular-block guard redoes an iteration on the dense M x (M+1) system */
    H.w_J = take(H.solver == 4 ? 0 : (H.solver == 2 ? std::max(2 * 6 * lanes_for(H), M * (M + 1)) : M * (M + 1))); H.w_rowh = take(2 * H.n_ctrl * ANM_MAX_ROWS + H.n_ctrl); /* -h / -inf, h / 0, finite-row mask per device */
    H.w_brp = take(L); H.w_brq = take(L); H.w_brs = take(L); H.w_brire = take(L); H.w_briim = take(L);
    H.w_full = take(H.n_full); H.w_s0 = take(H.n_state > K ? H.n_state : K);
    H.w_vx = take(4 * N);
    H.w_dx = take(M);
    H.w_blk = take(H.solver == 4 ? 4 * H.sp_nblk : 0);
    /* stride = 32 bytes mod 128: the lane groups of a warp read the same workspace entry at the same time, and
     * their four copies then sit in four different bank groups instead of one */
    H.ws_doubles = (w + 15) / 16 * 16 + 4;
  }
  bb.buf.resize((bb.buf.size() + 127) / 128 * 128, 0);
  H.blob_bytes = (int)bb.buf.size();
  memcpy(bb.buf.data(), &H, sizeof(H));
  out.swap(bb.buf);
  return 0;
}

#ifndef ANM_HOST_IO_DEFAULT
#define ANM_HOST_IO_DEFAULT 2 /* zero-copy when the caller's buffers are pinned: +14 % e2e on B200 (profiles/) */
#endif
typedef void (*kernel_fn)(const AnmLaunch);
/* ANM_LANES=16 (environment): 16 lanes per instance for the radial solver (two instances per warp instead of four:
 * half the lock-step penalty of a divergent instance, half the projection trips, twice the issue slots). */
static int radial_lanes() {
  static const int v = [] {
    const char* e = getenv("ANM_LANES");
    return (e && atoi(e) == 16) ? 16 : 8;
  }();
  return v;
}
static int lanes_for(const AnmConstHeader& H) {
  const int M = H.n_unk;
  switch (solver_for(H)) {
    case 2: return (H.n_bus >= 9) ? 16 : radial_lanes(); /* the radial solver wants an idle lane: n_bus - 1 < lanes */
    case 1: return (H.n_bus <= 5) ? 8 : 16;
    case 4: return 32;
    default: return (M <= 8) ? 8 : (M <= 16 ? 16 : 32);
  }
}
static int threads_for(const AnmConstHeader& H) {
  const int sv = solver_for(H);
  return (sv == 1 || sv == 2) ? ANM_VAR_THREADS : ANM_THREADS;
}

kernel_fn kernel_for(const AnmConstHeader& H) {
  const int solver = solver_for(H);
  if (solver == 2 && lanes_for(H) == 16) {
    switch (H.n_bus) {
      case 2: return anm::anm_env_kernel<16, 2, 2>;
      case 3: return anm::anm_env_kernel<16, 3, 2>;
      case 4: return anm::anm_env_kernel<16, 4, 2>;
      case 5: return anm::anm_env_kernel<16, 5, 2>;
      case 6: return anm::anm_env_kernel<16, 6, 2>;
      case 7: return anm::anm_env_kernel<16, 7, 2>;
      case 8: return anm::anm_env_kernel<16, 8, 2>;
      default: return anm::anm_env_kernel<16, 9, 2>;
    }
  }
  if (solver == 2) {
    switch (H.n_bus) {
      case 2: return anm::anm_env_kernel<8, 2, 2>;
      case 3: return anm::anm_env_kernel<8, 3, 2>;
      case 4: return anm::anm_env_kernel<8, 4, 2>;
      case 5: return anm::anm_env_kernel<8, 5, 2>;
      case 6: return anm::anm_env_kernel<8, 6, 2>;
      case 7: return anm::anm_env_kernel<8, 7, 2>;
      default: return anm::anm_env_kernel<8, 8, 2>;
    }
  }
  if (solver == 1) {
    switch (H.n_bus) {
      case 2: return anm::anm_env_kernel<8, 2, 1>;
      case 3: return anm::anm_env_kernel<8, 3, 1>;
      case 4: return anm::anm_env_kernel<8, 4, 1>;
      case 5: return anm::anm_env_kernel<8, 5, 1>;
      case 6: return anm::anm_env_kernel<16, 6, 1>;
      case 7: return anm::anm_env_kernel<16, 7, 1>;
      case 8: return anm::anm_env_kernel<16, 8, 1>;
      default: return anm::anm_env_kernel<16, 9, 1>;
    }
  }
  if (solver == 4) return anm::anm_env_kernel<32, 0, 4>;
  switch (lanes_for(H)) {
    case 8: return anm::anm_env_kernel<8, 0, 0>;
    case 16: return anm::anm_env_kernel<16, 0, 0>;
    default: return anm::anm_env_kernel<32, 0, 0>;
  }
}

int choose_geometry(anm_handle h) {
  h->lpe = lanes_for(h->H);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, h->device));
  h->num_sms = prop.multiProcessorCount;
  const size_t max_smem = prop.sharedMemPerBlockOptin;
  const size_t fixed = ANM_BLOB_SMEM_OFF + (size_t)h->blob_bytes;
  const size_t per_env = (size_t)h->H.ws_doubles * sizeof(double);
  int gpb = threads_for(h->H) / h->lpe;
  while (gpb > 1 && fixed + gpb * per_env > max_smem) gpb /= 2;
  if (fixed + gpb * per_env > max_smem)
    return fail(ANM_E_UNSUPPORTED, "network too large: %zu B of shared memory per CTA needed, %zu available",
                fixed + gpb * per_env, max_smem);
  h->gpb = gpb;
  h->smem = (int)(fixed + gpb * per_env);
  kernel_fn fn = kernel_for(h->H);
  CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem));
  int per_sm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, gpb * h->lpe, h->smem));
  if (per_sm < 1) return fail(ANM_E_UNSUPPORTED, "kernel does not fit on an SM (smem %d B)", h->smem);
  const int64_t need = (h->B + gpb - 1) / gpb;
  const int64_t resident = (int64_t)per_sm * h->num_sms;
  h->grid = (int)(need < resident ? need : resident);
  return 0;
}

/* ANM_PDL=0 (environment) turns programmatic dependent launch off: every launch then waits for the
 * complete previous one, as in a plain stream (A/B switch; results are identical either way). */
static bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("ANM_PDL");
    return !(e && atoi(e) == 0);
  }();
  return on;
}

/* base of the chaining watchdog's limit; ANM_WD_LIMIT_MS (environment) shortens it for the tests of the time-out path */
static uint64_t wd_base_ns() {
  static const uint64_t v = [] {
    const char* e = getenv("ANM_WD_LIMIT_MS");
    const long ms = e ? atol(e) : 2000;
    return (uint64_t)(ms >= 1 ? ms : 2000) * 1000000ull;
  }();
  return v;
}

/* a launch-chaining time-out is sticky: the kernel that hit it skipped the instance, so the handle's state is no
 * longer what the caller thinks it is -- every later call fails loudly (the CUDA context is fine) */
static int check_watchdog(anm_handle h) {
  if (h->wd_host && ((volatile uint32_t*)h->wd_host)[0] != 0u) {
    const volatile uint32_t* w = h->wd_host;
    return fail(ANM_E_TIMEOUT, "launch-chaining time-out: instance %u waited for launch ordinal %u but saw %u (CTA %u of %u); "
                "the handle is unusable, destroy it", w[1], w[2], w[3], w[4], w[5]);
  }
  return 0;
}

int launch(anm_handle h, AnmLaunch& p, cudaStream_t st, uint32_t flags = 0) {
  if (int rc = check_watchdog(h)) return rc;
  p.blob = h->d_blob;
  p.blob_bytes = h->blob_bytes;
  p.B = h->B;
  p.soc = h->d_soc; p.aux = h->d_aux; p.terminated = h->d_term; p.episode = h->d_episode;
  p.pool = h->pool; p.pool_size = h->pool_size;
  /* launch chaining: wait for the previous launch per instance, publish this one (anm_kernels.cuh) */
  p.seq = h->d_seq;
  p.ticket = h->d_ticket;
  {
    /* Watchdog limit: 2 s plus 2 ms for every step x pass of the launches this one may (transitively) have to wait
     * for.  A launch that is not chained executes griddepcontrol.wait, i.e. runs after everything earlier has
     * completed: the sum starts again with it. */
    const int64_t passes = (h->B + (int64_t)h->grid * h->gpb - 1) / ((int64_t)h->grid * h->gpb);
    const bool chained = pdl_enabled() && (flags & ANM_LF_CHAINED);
    if (!chained) h->wd_steps = 0;
    p.wd_limit_ns = wd_base_ns() + 2000000ull * (uint64_t)h->wd_steps;
    h->wd_last = (int64_t)(p.T > 1 ? p.T : 1) * passes;
    h->wd_steps += h->wd_last;
  }
  p.watchdog = h->wd_dev;
  p.dense_scratch = h->d_dense_scratch;
  const bool pdl = pdl_enabled();
  p.flags = pdl ? flags : (flags & ~ANM_LF_CHAINED);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)h->grid);
  cfg.blockDim = dim3((unsigned)(h->gpb * h->lpe));
  cfg.dynamicSmemBytes = (size_t)h->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  const AnmLaunch arg = p;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel_for(h->H), arg);
  ++h->launches;
  if (e == cudaSuccess) e = cudaPeekAtLastError();
  if (e != cudaSuccess) return fail(ANM_E_CUDA, "kernel launch: %s", cudaGetErrorString(e));
  return 0;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};


/* Host-buffer I/O mode of anm_step_host (environment variable ANM_HOST_IO):
 *   copy    stage through device buffers with cudaMemcpyAsync (works for any host memory)
 *   zc      zero-copy: if a buffer is pinned (cudaHostAlloc / cudaHostRegister, e.g. a torch pinned tensor) the
 *           kernel reads the actions from it / writes obs, reward, terminated into it directly over PCIe
 *   zc_out  zero-copy for the outputs only
 * Pageable buffers always take the copy path. */
static int host_io_mode() {
  static const int m = [] {
    const char* e = getenv("ANM_HOST_IO");
    if (e && !strcmp(e, "copy")) return 0;
    if (e && !strcmp(e, "zc")) return 2;
    if (e && !strcmp(e, "zc_out")) return 1;
    return ANM_HOST_IO_DEFAULT;
  }();
  return m;
}
template <typename T>
static T* mapped_device_pointer(T* host_ptr) { /* NULL unless the host buffer is pinned and mapped */
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, (const void*)host_ptr) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return (attr.type == cudaMemoryTypeHost) ? (T*)attr.devicePointer : nullptr;
}

}  // namespace

extern "C" {

int anm_abi_version(void) { return ANM_ABI_VERSION; }
const char* anm_last_error(void) { return g_err; }

int anm_create(const anm_network_desc* net, const anm_env_desc* env, int64_t num_envs, int device, anm_handle* out) {
  if (!net || !env || !out) return fail(ANM_E_INVALID, "null argument");
  if (num_envs < 1) return fail(ANM_E_INVALID, "num_envs must be >= 1");
  *out = nullptr;
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(ANM_E_INVALID, "device %d out of range (%d visible)", device, ndev);
  anm_handle h = new (std::nothrow) anm_handle_s();
  if (!h) return fail(ANM_E_NOMEM, "out of host memory");
  h->device = device;
  h->B = num_envs;
  std::vector<unsigned char> blob;
  int rc = build_blob(net, env, h->H, blob);
  if (rc) { delete h; return rc; }
  h->blob_bytes = (int)blob.size();
  DeviceGuard guard(device);
  const AnmConstHeader& H = h->H;
  const size_t B = (size_t)num_envs;
#define ALLOC(ptr, bytes)                                                                   \
  do {                                                                                      \
    cudaError_t e_ = cudaMalloc((void**)&(ptr), (bytes) ? (bytes) : 16);                    \
    if (e_ != cudaSuccess) { anm_destroy(h); return fail(ANM_E_CUDA, "cudaMalloc(%s): %s", #ptr, cudaGetErrorString(e_)); } \
  } while (0)
  ALLOC(h->d_blob, blob.size());
  ALLOC(h->d_soc, B * H.n_des * sizeof(double));
  ALLOC(h->d_aux, B * H.K * sizeof(double));
  ALLOC(h->d_term, B);
  ALLOC(h->d_episode, B * sizeof(uint32_t));
  ALLOC(h->d_seq, B * sizeof(uint32_t));
  ALLOC(h->d_ticket, 16 * sizeof(unsigned long long));
  ALLOC(h->s_action, B * H.n_action * sizeof(double));
  ALLOC(h->s_nv, B * H.n_next_vars * sizeof(double));
  ALLOC(h->s_obs, B * H.n_obs * sizeof(double));
  ALLOC(h->s_reward, B * sizeof(double));
  ALLOC(h->s_s0, B * H.n_state * sizeof(double));
  ALLOC(h->s_state, B * H.n_state * sizeof(double));
  ALLOC(h->s_term, B);
  ALLOC(h->s_mask, B);
#undef ALLOC
  cudaError_t e = cudaMemcpy(h->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(h->d_soc, 0, B * H.n_des * sizeof(double) + (H.n_des ? 0 : 16));
  if (e == cudaSuccess) e = cudaMemset(h->d_aux, 0, B * H.K * sizeof(double) + (H.K ? 0 : 16));
  if (e == cudaSuccess) e = cudaMemset(h->d_term, 1, B); /* nothing is runnable before the first reset */
  if (e == cudaSuccess) e = cudaMemset(h->d_episode, 0, B * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(h->d_seq, 0, B * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(h->d_ticket, 0, 16 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&h->wd_host, ANM_WD_WORDS * sizeof(uint32_t), cudaHostAllocMapped);
  if (e == cudaSuccess) {
    memset(h->wd_host, 0, ANM_WD_WORDS * sizeof(uint32_t));
    e = cudaHostGetDevicePointer((void**)&h->wd_dev, h->wd_host, 0);
  }
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy); /* the memsets above, before any other stream uses them */
  if (e != cudaSuccess) { anm_destroy(h); return fail(ANM_E_CUDA, "handle init: %s", cudaGetErrorString(e)); }
  rc = choose_geometry(h);
  if (rc) { anm_destroy(h); return rc; }
  if (h->H.solver == 4) { /* the block-sparse solver's cold redo works on a dense system in global memory */
    const size_t per = (size_t)h->H.n_unk * (size_t)(h->H.n_unk + 1) * sizeof(double);
    cudaError_t e2 = cudaMalloc((void**)&h->d_dense_scratch, per * (size_t)h->grid * (size_t)h->gpb);
    if (e2 != cudaSuccess) { anm_destroy(h); return fail(ANM_E_CUDA, "dense scratch: %s", cudaGetErrorString(e2)); }
  }
  *out = h;
  return ANM_OK;
}

int anm_destroy(anm_handle h) {
  if (!h) return ANM_OK;
  anm_gather_destroy(h);
  DeviceGuard guard(h->device);
  cudaFree(h->d_blob); cudaFree(h->d_soc); cudaFree(h->d_aux); cudaFree(h->d_term); cudaFree(h->d_episode);
  cudaFree(h->d_seq); cudaFree(h->d_ticket); cudaFree(h->d_dense_scratch);
  cudaFree(h->d_rng); cudaFree(h->d_need); cudaFree(h->d_conv_tmp); cudaFree(h->d_remaining);
  if (h->h_remaining) cudaFreeHost(h->h_remaining);
  if (h->wd_host) cudaFreeHost(h->wd_host);
  cudaFree(h->s_action); cudaFree(h->s_nv); cudaFree(h->s_obs); cudaFree(h->s_reward); cudaFree(h->s_s0);
  cudaFree(h->s_state); cudaFree(h->s_term); cudaFree(h->s_mask);
  for (auto& hs : h->hset) {
    cudaFree(hs.act); cudaFree(hs.nv); cudaFree(hs.obs); cudaFree(hs.rew); cudaFree(hs.term);
    if (hs.ev_in) cudaEventDestroy(hs.ev_in);
    if (hs.ev_k) cudaEventDestroy(hs.ev_k);
    if (hs.ev_out) cudaEventDestroy(hs.ev_out);
  }
  if (h->st_in) cudaStreamDestroy(h->st_in);
  if (h->st_out) cudaStreamDestroy(h->st_out);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return ANM_OK;
}

int anm_get_sizes(anm_handle h, anm_sizes* o) {
  if (!h || !o) return fail(ANM_E_INVALID, "null argument");
  const AnmConstHeader& H = h->H;
  o->num_envs = h->B;
  o->n_bus = H.n_bus; o->n_dev = H.n_dev; o->n_branch = H.n_branch;
  o->n_load = H.n_load; o->n_gen = H.n_gen; o->n_des = H.n_des;
  o->n_action = H.n_action; o->n_state = H.n_state; o->n_obs = H.n_obs;
  o->n_next_vars = H.n_next_vars; o->n_full_state = H.n_full; o->K = H.K;
  o->lanes_per_env = h->lpe; o->envs_per_block = h->gpb; o->smem_bytes = h->smem;
  return ANM_OK;
}

int anm_reset(anm_handle h, const double* s0, const uint8_t* mask, double* obs, double* state, uint8_t* converged,
              void* stream) {
  ANM_NVTX("anm_reset");
  if (!h || !s0 || !obs || !converged) return fail(ANM_E_INVALID, "anm_reset: null argument");
  DeviceGuard guard(h->device);
  AnmLaunch p;
  memset(&p, 0, sizeof(p));
  p.mode = ANM_MODE_RESET;
  p.s0 = s0; p.mask = mask; p.obs = obs; p.state = state; p.converged = converged;
  p.full_state = h->reset_full;
  return launch(h, p, (cudaStream_t)stream);
}

int anm_step(anm_handle h, const double* action, const double* next_vars, double* obs, double* reward,
             uint8_t* terminated, const anm_step_extras* ex, void* stream) {
  ANM_NVTX("anm_step");
  if (!h || !action || !obs || !reward || !terminated) return fail(ANM_E_INVALID, "anm_step: null argument");
  if (!next_vars && h->H.table_len == 0)
    return fail(ANM_E_INVALID, "anm_step: next_vars is NULL but the environment has no built-in table");
  DeviceGuard guard(h->device);
  AnmLaunch p;
  memset(&p, 0, sizeof(p));
  p.mode = ANM_MODE_STEP;
  p.action = action; p.next_vars = next_vars; p.obs = obs; p.reward = reward; p.term_out = terminated;
  uint32_t lf = 0;
  if (ex) {
    p.state = ex->state; p.e_loss = ex->e_loss; p.penalty = ex->penalty; p.n_iter = ex->n_iter;
    p.full_state = ex->full_state; p.solver_stats = ex->solver_stats;
#if ANM_DIAG
    if (ex->solver_stats) p.phase_stats = ex->solver_stats + 4 * h->B; /* diagnostic builds: the buffer is [B, 4 + 16] */
#endif
    if (ex->flags & ANM_STEP_CHAINED) lf |= ANM_LF_CHAINED;
  }
  return launch(h, p, (cudaStream_t)stream, lf);
}

int anm_rollout(anm_handle h, int64_t T, const double* action, const double* next_vars, double* obs, double* reward,
                uint8_t* terminated, uint32_t flags, void* stream) {
  ANM_NVTX("anm_rollout");
  if (!h || !action || !obs || !reward || !terminated || T < 0) return fail(ANM_E_INVALID, "anm_rollout: bad argument");
  if (!next_vars && h->H.table_len == 0)
    return fail(ANM_E_INVALID, "anm_rollout: next_vars is NULL but the environment has no built-in table");
  if (T > INT32_MAX) return fail(ANM_E_INVALID, "anm_rollout: T too large");
  if (T == 0) return ANM_OK;
  DeviceGuard guard(h->device);
  /* ONE launch: every lane group takes its instance through the T steps (its carried state stays on chip), so an
   * instance whose Newton iteration diverges delays nobody but the instances of its own warp */
  AnmLaunch p;
  memset(&p, 0, sizeof(p));
  p.mode = ANM_MODE_STEP;
  p.T = (int32_t)T;
  p.action = action; p.next_vars = next_vars; p.obs = obs; p.reward = reward; p.term_out = terminated;
  return launch(h, p, (cudaStream_t)stream, (flags & ANM_STEP_CHAINED) ? ANM_LF_CHAINED : 0u);
}

int anm_seed(anm_handle h, uint64_t seed_first) { /* the blocking form: the streams are seeded when it returns */
  if (int rc = anm_seed_async(h, seed_first, nullptr)) return rc;
  CUDA_TRY(cudaStreamSynchronize(nullptr));
  return ANM_OK;
}

int anm_seed_async(anm_handle h, uint64_t seed_first, void* stream) {
  if (!h) return fail(ANM_E_INVALID, "null handle");
  DeviceGuard guard(h->device);
  if (!h->d_rng) CUDA_TRY(cudaMalloc((void**)&h->d_rng, (size_t)h->B * sizeof(AnmPcg64)));
  /* SeedSequence -> PCG64 expansion on the device, ordered on the caller's stream like every other call: no host
   * loop over the batch, no device-wide synchronisation */
  const int threads = 128, blocks = (int)((h->B + threads - 1) / threads);
  anm::seed_streams_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(h->d_rng, seed_first, h->B);
  CUDA_TRY(cudaPeekAtLastError());
  h->seeded = true;
  return ANM_OK;
}

int anm_reset_seeded(anm_handle h, const uint8_t* mask, int32_t max_tries, int32_t date_draw, double* obs, double* state,
                     uint8_t* converged, void* stream) {
  ANM_NVTX("anm_reset_seeded");
  if (!h || !obs || !converged) return fail(ANM_E_INVALID, "anm_reset_seeded: null argument");
  if (!h->seeded) return fail(ANM_E_INVALID, "anm_reset_seeded: call anm_seed first");
  if (h->H.table_len == 0 || h->H.K < 1)
    return fail(ANM_E_UNSUPPORTED, "anm_reset_seeded: needs the built-in next_vars table (ANM6Easy-style init_state)");
  if (max_tries < 1) return fail(ANM_E_INVALID, "anm_reset_seeded: max_tries must be >= 1");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t B = (size_t)h->B;
  if (!h->d_need) {
    CUDA_TRY(cudaMalloc((void**)&h->d_need, B));
    CUDA_TRY(cudaMalloc((void**)&h->d_conv_tmp, B));
    CUDA_TRY(cudaMalloc((void**)&h->d_remaining, sizeof(int)));
    CUDA_TRY(cudaHostAlloc((void**)&h->h_remaining, sizeof(int), cudaHostAllocDefault));
  }
  const int threads = 128, blocks = (int)((B + threads - 1) / threads);
  /* Round k: instances whose previous attempt converged are finished (date draw), the others draw a new initial
   * state; then the ordinary reset launch applies the drawn states of the instances that drew.  Both kernels do
   * nothing for an instance that is done, so the rounds are enqueued ANM_RESET_ROUNDS at a time (the loop needs two or
   * three for ANM6Easy) and the host looks at the number of draws of the batch's LAST round only: one stream
   * synchronisation per reset instead of one per round. */
  const int batch = ANM_RESET_ROUNDS;
  for (int round = 0; round <= max_tries;) {
    const int stop = std::min(round + batch, max_tries + 1);
    for (; round < stop; ++round) {
      CUDA_TRY(cudaMemsetAsync(h->d_remaining, 0, sizeof(int), st));
      anm::seeded_draw_kernel<<<blocks, threads, 0, st>>>(h->d_blob, h->B, h->d_rng, h->d_need, mask, round == 0,
                                                          round == max_tries, h->d_conv_tmp, converged, date_draw, h->s_s0,
                                                          h->d_remaining);
      CUDA_TRY(cudaPeekAtLastError());
      if (round < max_tries) {
        int rc = anm_reset(h, h->s_s0, h->d_need, obs, state, h->d_conv_tmp, stream);
        if (rc) return rc;
      }
    }
    /* the draws of one more (draw-only) look: did anybody still need a state after the batch's last reset? */
    CUDA_TRY(cudaMemcpyAsync(h->h_remaining, h->d_remaining, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (*h->h_remaining == 0) break;
  }
  return ANM_OK;
}

int anm_debug_project(const double* a, const double* b, const double* h, int32_t R, double p, double q, double* out2) {
  return anm_debug_project_ex(a, b, h, R, ~0u, -1, -1, p, q, out2);
}

int anm_debug_project_ex(const double* a, const double* b, const double* h, int32_t R, uint32_t dyn_mask, int32_t dom_row,
                         int32_t dom_by, double p, double q, double* out2) {
  if (!a || !b || !h || !out2 || R < 1 || R > ANM_MAX_ROWS) return fail(ANM_E_INVALID, "anm_debug_project: bad argument");
  std::vector<int> cinfo;
  std::vector<double> ccoef;
  int n_short = 0;
  build_candidates(a, b, R, cinfo, ccoef, dyn_mask, h, dom_row, &n_short);
  /* the kernel leaves the dominated row's candidates out when the dominating dynamic row is finite */
  if (dom_row >= 0 && dom_by >= 0 && dom_by < R && std::isfinite(h[dom_by])) {
    cinfo.resize((size_t)n_short);
    ccoef.resize((size_t)n_short * 8);
  }
  /* the evaluation of project_polygon (anm_kernels.cuh), one candidate after the other */
  const double inf = std::numeric_limits<double>::infinity();
  unsigned fin = 0u;
  double nh[ANM_MAX_ROWS], he[ANM_MAX_ROWS];
  for (int k = 0; k < R; ++k) {
    const bool f = std::fabs(h[k]) < inf;
    fin |= f ? (1u << k) : 0u;
    nh[k] = f ? -h[k] : -inf;
    he[k] = f ? h[k] : 0.0;
  }
  double best = inf, bx = std::nan(""), by = std::nan("");
  for (size_t c = 0; c < cinfo.size(); ++c) {
    const int nf = cinfo[c];
    const unsigned need = (unsigned)nf >> 16;
    const double h1 = he[nf & 0xff], h2 = he[(nf >> 8) & 0xff];
    const double* kx = &ccoef[8 * c];
    const double* ky = kx + 4;
    const double x = std::fma(kx[0], p, std::fma(kx[1], q, std::fma(kx[2], h1, kx[3] * h2)));
    const double y = std::fma(ky[0], p, std::fma(ky[1], q, std::fma(ky[2], h1, ky[3] * h2)));
    unsigned viol = 0u;
    for (int k = 0; k < R; ++k) viol |= (std::fma(a[k], x, std::fma(b[k], y, nh[k])) > ANM_FEAS_TOL) ? (1u << k) : 0u;
    const double dx = x - p, dy = y - q;
    const double d = std::fma(dx, dx, dy * dy);
    if (((fin & need) == need) && ((viol & ~need) == 0u) && d < best) best = d, bx = x, by = y;
  }
  out2[0] = bx;
  out2[1] = by;
  return ANM_OK;
}

int64_t anm_debug_blob(const anm_network_desc* net, const anm_env_desc* env, unsigned char* out, int64_t cap) {
  if (!net || !env) return fail(ANM_E_INVALID, "null argument");
  AnmConstHeader H;
  std::vector<unsigned char> blob;
  int rc = build_blob(net, env, H, blob);
  if (rc) return rc;
  if (out && cap > 0) memcpy(out, blob.data(), (size_t)std::min<int64_t>(cap, (int64_t)blob.size()));
  return (int64_t)blob.size();
}

int anm_debug_rng(uint64_t seed, int32_t n, const int32_t* kind, const double* lo, const double* hi, double* out) {
  if (n < 0 || (n > 0 && (!kind || !lo || !hi || !out))) return fail(ANM_E_INVALID, "anm_debug_rng: bad argument");
  AnmPcg64 r;
  anm_pcg_seed(r, seed);
  for (int32_t i = 0; i < n; ++i)
    out[i] = kind[i] ? anm_rng_uniform(r, lo[i], hi[i]) : (double)anm_rng_integers(r, (int64_t)lo[i], (int64_t)hi[i]);
  return ANM_OK;
}

int anm_debug_set_launch_ordinal(anm_handle h, uint64_t k) {
  if (!h) return fail(ANM_E_INVALID, "null handle");
  DeviceGuard guard(h->device);
  CUDA_TRY(cudaDeviceSynchronize());
  const unsigned long long t = (unsigned long long)k * (unsigned long long)h->grid;
  std::vector<uint32_t> seq((size_t)h->B, (uint32_t)k);
  CUDA_TRY(cudaMemcpy(h->d_ticket, &t, sizeof(t), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(h->d_seq, seq.data(), seq.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  return ANM_OK;
}

int anm_debug_stall_instance(anm_handle h, int64_t e) {
  if (!h || e < 0 || e >= h->B) return fail(ANM_E_INVALID, "anm_debug_stall_instance: bad argument");
  DeviceGuard guard(h->device);
  CUDA_TRY(cudaDeviceSynchronize());
  uint32_t v = 0;
  CUDA_TRY(cudaMemcpy(&v, h->d_seq + e, sizeof(v), cudaMemcpyDeviceToHost));
  v -= 3u; /* as if the last three launches had never finished with this instance */
  CUDA_TRY(cudaMemcpy(h->d_seq + e, &v, sizeof(v), cudaMemcpyHostToDevice));
  return ANM_OK;
}

int64_t anm_rng_state_bytes(anm_handle h) { return h ? (int64_t)h->B * (int64_t)sizeof(AnmPcg64) : 0; }

int anm_get_rng(anm_handle h, void* out, void* stream) {
  if (!h || !out) return fail(ANM_E_INVALID, "anm_get_rng: null argument");
  if (!h->seeded) return fail(ANM_E_INVALID, "anm_get_rng: call anm_seed first");
  DeviceGuard guard(h->device);
  CUDA_TRY(cudaMemcpyAsync(out, h->d_rng, (size_t)h->B * sizeof(AnmPcg64), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return ANM_OK;
}

int anm_set_rng(anm_handle h, const void* in, void* stream) {
  if (!h || !in) return fail(ANM_E_INVALID, "anm_set_rng: null argument");
  DeviceGuard guard(h->device);
  if (!h->d_rng) CUDA_TRY(cudaMalloc((void**)&h->d_rng, (size_t)h->B * sizeof(AnmPcg64)));
  CUDA_TRY(cudaMemcpyAsync(h->d_rng, in, (size_t)h->B * sizeof(AnmPcg64), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  h->seeded = true;
  return ANM_OK;
}

int anm_debug_math(int32_t kind, int64_t n, const double* x, const double* y, double* a, double* b, void* stream) {
  if (kind < 0 || kind > 2 || n < 0 || !x || !a || (kind == 0 && !b) || (kind == 1 && !y))
    return fail(ANM_E_INVALID, "anm_debug_math: bad argument");
  if (n == 0) return ANM_OK;
  anm::debug_math_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kind, n, x, y, a, b);
  CUDA_TRY(cudaPeekAtLastError());
  return ANM_OK;
}

#define ANM_GATHER_FLAG_BYTES 1024
#define ANM_GATHER_MAX_WORLD 64

int anm_gather_create(anm_handle h, int32_t world, int32_t rank, int64_t row0, int64_t rows_global, int32_t slots,
                      void* ipc_handle_out) {
  if (!h || !ipc_handle_out) return fail(ANM_E_INVALID, "anm_gather_create: null argument");
  if (world < 1 || world > ANM_GATHER_MAX_WORLD || rank < 0 || rank >= world || slots < 1 || slots > 64)
    return fail(ANM_E_INVALID, "anm_gather_create: world %d / rank %d / slots %d out of range", world, rank, slots);
  if (row0 < 0 || rows_global < row0 + h->B) return fail(ANM_E_INVALID, "anm_gather_create: rows [%lld, %lld) do not fit %lld global rows",
                                                         (long long)row0, (long long)(row0 + h->B), (long long)rows_global);
  if (h->g.base) return fail(ANM_E_INVALID, "anm_gather_create: a gather is already attached to this handle");
  DeviceGuard guard(h->device);
  auto& g = h->g;
  g.world = world; g.rank = rank; g.slots = slots; g.rows = rows_global; g.row0 = row0;
  const size_t W = (size_t)h->H.n_obs + 2;
  const size_t bytes = ANM_GATHER_FLAG_BYTES + (size_t)slots * (size_t)rows_global * W * sizeof(double);
  CUDA_TRY(cudaMalloc((void**)&g.base, bytes));
  CUDA_TRY(cudaMemset(g.base, 0, bytes));
  CUDA_TRY(cudaMalloc((void**)&g.d_state, 128 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemset(g.d_state, 0, 128 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMalloc((void**)&g.d_peers, ANM_GATHER_MAX_WORLD * sizeof(void*)));
  CUDA_TRY(cudaMalloc((void**)&g.d_flags, ANM_GATHER_MAX_WORLD * sizeof(void*)));
  CUDA_TRY(cudaDeviceSynchronize());
  cudaIpcMemHandle_t mh;
  CUDA_TRY(cudaIpcGetMemHandle(&mh, g.base));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  memcpy(ipc_handle_out, &mh, sizeof(mh));
  return ANM_OK;
}

int anm_gather_attach(anm_handle h, const void* ipc_handles) {
  if (!h || !ipc_handles) return fail(ANM_E_INVALID, "anm_gather_attach: null argument");
  auto& g = h->g;
  if (!g.base || g.attached) return fail(ANM_E_INVALID, "anm_gather_attach: call anm_gather_create first (once)");
  DeviceGuard guard(h->device);
  g.peer_base.assign((size_t)g.world, nullptr);
  std::vector<double*> peers((size_t)g.world);
  std::vector<unsigned long long*> flags((size_t)g.world);
  for (int p = 0; p < g.world; ++p) {
    void* b = g.base;
    if (p != g.rank) {
      cudaIpcMemHandle_t mh;
      memcpy(&mh, (const unsigned char*)ipc_handles + (size_t)p * sizeof(mh), sizeof(mh));
      cudaError_t e = cudaIpcOpenMemHandle(&b, mh, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) return fail(ANM_E_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e));
    }
    g.peer_base[(size_t)p] = b;
    flags[(size_t)p] = (unsigned long long*)b;
    peers[(size_t)p] = (double*)((unsigned char*)b + ANM_GATHER_FLAG_BYTES);
  }
  CUDA_TRY(cudaMemcpy(g.d_peers, peers.data(), peers.size() * sizeof(void*), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(g.d_flags, flags.data(), flags.size() * sizeof(void*), cudaMemcpyHostToDevice));
  g.attached = true;
  return ANM_OK;
}

int anm_gather_destroy(anm_handle h) {
  if (!h) return ANM_OK;
  auto& g = h->g;
  if (!g.base) return ANM_OK;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  for (int p = 0; p < (int)g.peer_base.size(); ++p)
    if (p != g.rank && g.peer_base[(size_t)p]) cudaIpcCloseMemHandle(g.peer_base[(size_t)p]);
  cudaFree(g.base); cudaFree(g.d_state); cudaFree(g.d_peers); cudaFree(g.d_flags);
  g = anm_handle_s::Gather();
  return ANM_OK;
}

int anm_step_packed(anm_handle h, const double* action, const double* next_vars, double* obs, double* reward,
                    uint8_t* terminated, double* packed, int32_t gather, const anm_step_extras* ex, void* stream) {
  ANM_NVTX("anm_step_packed");
  if (!h || !action || !obs || !reward || !terminated) return fail(ANM_E_INVALID, "anm_step_packed: null argument");
  if (!next_vars && h->H.table_len == 0)
    return fail(ANM_E_INVALID, "anm_step_packed: next_vars is NULL but the environment has no built-in table");
  if (gather && !h->g.attached) return fail(ANM_E_INVALID, "anm_step_packed: no gather attached (anm_gather_create / _attach)");
  if (!gather && !packed) return fail(ANM_E_INVALID, "anm_step_packed: neither a packed output nor the gather was asked for");
  DeviceGuard guard(h->device);
  AnmLaunch p;
  memset(&p, 0, sizeof(p));
  p.mode = ANM_MODE_STEP;
  p.action = action; p.next_vars = next_vars; p.obs = obs; p.reward = reward; p.term_out = terminated;
  p.packed = packed;
  if (ex) {
    p.state = ex->state; p.e_loss = ex->e_loss; p.penalty = ex->penalty; p.n_iter = ex->n_iter;
    p.full_state = ex->full_state;
  }
  if (gather) {
    auto& g = h->g;
    p.g_peers = g.d_peers; p.g_flags = g.d_flags; p.g_state = g.d_state;
    p.g_rows = g.rows; p.g_row0 = g.row0; p.g_world = g.world; p.g_rank = g.rank; p.g_slots = g.slots;
  }
  /* never chained: the gather slot is derived from the completed-step count at kernel entry; remote stores need the
   * system-scope fence */
  const int rc = launch(h, p, (cudaStream_t)stream, gather ? (ANM_LF_SYSOUT | (gather > 1 ? ANM_LF_GATHER_WAIT : 0u)) : 0u);
  if (rc == ANM_OK && gather) {
    h->g.last_waited_inline = gather > 1;
    ++h->g.steps_host;
  }
  return rc;
}

int anm_gather_wait(anm_handle h, double** rows_out, void* stream) {
  if (!h) return fail(ANM_E_INVALID, "null handle");
  auto& g = h->g;
  if (!g.attached || g.steps_host < 1) return fail(ANM_E_INVALID, "anm_gather_wait: no gather step has been launched");
  DeviceGuard guard(h->device);
  if (!g.last_waited_inline) {
    anm::gather_wait_kernel<<<1, 64, 0, (cudaStream_t)stream>>>((const unsigned long long*)g.base, g.d_state, g.world,
                                                                30000000000ull);
    CUDA_TRY(cudaPeekAtLastError());
  }
  if (rows_out) {
    const size_t W = (size_t)h->H.n_obs + 2;
    const int64_t slot = (g.steps_host - 1) % g.slots;
    *rows_out = (double*)(g.base + ANM_GATHER_FLAG_BYTES) + (size_t)slot * (size_t)g.rows * W;
  }
  return ANM_OK;
}

int anm_debug_fp64_peak(int device, double* tflops_out) {
  if (!tflops_out) return fail(ANM_E_INVALID, "null argument");
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  double* d = nullptr;
  CUDA_TRY(cudaMalloc((void**)&d, 64));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 2048;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) { /* the first pass warms up; best of the rest */
    CUDA_TRY(cudaEventRecord(e0, 0));
    anm::fp64_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    CUDA_TRY(cudaEventRecord(e1, 0));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    if (rep > 0 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops_out = best;
  return ANM_OK;
}

int anm_set_reset_full_state(anm_handle h, double* full_state) {
  if (!h) return fail(ANM_E_INVALID, "null handle");
  h->reset_full = full_state;
  return ANM_OK;
}

int anm_set_autoreset_pool(anm_handle h, const double* pool, int64_t pool_size) {
  if (!h || pool_size < 0 || (pool_size > 0 && !pool)) return fail(ANM_E_INVALID, "anm_set_autoreset_pool: bad argument");
  h->pool = pool_size ? pool : nullptr;
  h->pool_size = pool_size;
  return ANM_OK;
}

int anm_transition(anm_handle h, const double* p_load, const double* p_pot, const double* p_set, const double* q_set,
                   double* full_state, double* reward, double* e_loss, double* penalty, uint8_t* converged,
                   void* stream) {
  ANM_NVTX("anm_transition");
  if (!h) return fail(ANM_E_INVALID, "anm_transition: null handle");
  const AnmConstHeader& H = h->H;
  if ((H.n_load && !p_load) || (H.n_gen && !p_pot) || (H.n_ctrl && (!p_set || !q_set)))
    return fail(ANM_E_INVALID, "anm_transition: missing input");
  DeviceGuard guard(h->device);
  AnmLaunch p;
  memset(&p, 0, sizeof(p));
  p.mode = ANM_MODE_TRANSITION;
  p.p_load = p_load; p.p_pot = p_pot; p.p_set = p_set; p.q_set = q_set;
  p.full_state = full_state; p.reward = reward; p.e_loss = e_loss; p.penalty = penalty; p.converged = converged;
  return launch(h, p, (cudaStream_t)stream);
}

int anm_get_state(anm_handle h, double* soc, double* aux, uint8_t* term, void* stream) {
  if (!h) return fail(ANM_E_INVALID, "null handle");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t B = (size_t)h->B;
  if (soc && h->H.n_des) CUDA_TRY(cudaMemcpyAsync(soc, h->d_soc, B * h->H.n_des * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (aux && h->H.K) CUDA_TRY(cudaMemcpyAsync(aux, h->d_aux, B * h->H.K * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (term) CUDA_TRY(cudaMemcpyAsync(term, h->d_term, B, cudaMemcpyDeviceToDevice, st));
  return ANM_OK;
}

int anm_set_state(anm_handle h, const double* soc, const double* aux, const uint8_t* term, void* stream) {
  if (!h) return fail(ANM_E_INVALID, "null handle");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t B = (size_t)h->B;
  if (soc && h->H.n_des) CUDA_TRY(cudaMemcpyAsync(h->d_soc, soc, B * h->H.n_des * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (aux && h->H.K) CUDA_TRY(cudaMemcpyAsync(h->d_aux, aux, B * h->H.K * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (term) CUDA_TRY(cudaMemcpyAsync(h->d_term, term, B, cudaMemcpyDeviceToDevice, st));
  return ANM_OK;
}

static int step_host_enqueue(anm_handle h, int64_t T, const double* action, const double* next_vars, double* obs,
                             double* reward, uint8_t* terminated, bool queued) {
  DeviceGuard guard(h->device);
  h->st_last_was_rollout = false;
  const AnmConstHeader& H = h->H;
  const size_t B = (size_t)h->B * (size_t)T; /* rows of the [T, B, .] host arrays */
  if (T > 1 && host_io_mode() < 2)
    return fail(ANM_E_UNSUPPORTED, "anm_rollout_host_async needs the zero-copy path (ANM_HOST_IO=zc)");
  cudaStream_t st = h->stream;
  const int mode = host_io_mode();
  const double* d_action = (mode >= 2) ? mapped_device_pointer(action) : nullptr;
  const double* d_nv = (mode >= 2 && next_vars) ? mapped_device_pointer(next_vars) : nullptr;
  double* d_obs = (mode >= 1) ? mapped_device_pointer(obs) : nullptr;
  double* d_reward = (mode >= 1) ? mapped_device_pointer(reward) : nullptr;
  uint8_t* d_term = (mode >= 1) ? mapped_device_pointer(terminated) : nullptr;
  const bool zc_in = d_action && (!next_vars || d_nv); /* every input is read straight from mapped host memory */
  if (T > 1 && !(zc_in && d_obs && d_reward && d_term))
    return fail(ANM_E_INVALID, "anm_rollout_host_async: every buffer must be pinned (page-locked, mapped) host memory");
  if (!d_action) {
    CUDA_TRY(cudaMemcpyAsync(h->s_action, action, B * H.n_action * sizeof(double), cudaMemcpyHostToDevice, st));
    d_action = h->s_action;
  }
  if (next_vars && !d_nv) {
    CUDA_TRY(cudaMemcpyAsync(h->s_nv, next_vars, B * H.n_next_vars * sizeof(double), cudaMemcpyHostToDevice, st));
    d_nv = h->s_nv;
  }
  AnmLaunch p;
  memset(&p, 0, sizeof(p));
  p.mode = ANM_MODE_STEP;
  p.T = (int32_t)T;
  p.action = d_action; p.next_vars = next_vars ? d_nv : nullptr;
  p.obs = d_obs ? d_obs : h->s_obs; p.reward = d_reward ? d_reward : h->s_reward; p.term_out = d_term ? d_term : h->s_term;
  /* Inputs that the kernel reads straight from mapped host memory cannot depend on earlier device work:
   * a queued step is chained to the previous launch per instance.  Staged inputs are ordered by the copies. */
  uint32_t lf = 0;
  if (queued && zc_in) lf |= ANM_LF_CHAINED;
  if (d_obs || d_reward || d_term) lf |= ANM_LF_SYSOUT;
  int rc = launch(h, p, st, lf);
  if (rc) return rc;
  if (!d_obs) CUDA_TRY(cudaMemcpyAsync(obs, h->s_obs, B * H.n_obs * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (!d_reward) CUDA_TRY(cudaMemcpyAsync(reward, h->s_reward, B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (!d_term) CUDA_TRY(cudaMemcpyAsync(terminated, h->s_term, B, cudaMemcpyDeviceToHost, st));
  return ANM_OK;
}

int anm_step_host(anm_handle h, const double* action, const double* next_vars, double* obs, double* reward,
                  uint8_t* terminated) {
  ANM_NVTX("anm_step_host");
  if (!h || !action || !obs || !reward || !terminated) return fail(ANM_E_INVALID, "anm_step_host: null argument");
  if (!next_vars && h->H.table_len == 0)
    return fail(ANM_E_INVALID, "anm_step_host: next_vars is NULL but the environment has no built-in table");
  int rc = step_host_enqueue(h, 1, action, next_vars, obs, reward, terminated, false);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return ANM_OK;
}

int anm_step_host_async(anm_handle h, const double* action, const double* next_vars, double* obs, double* reward,
                        uint8_t* terminated) {
  if (!h || !action || !obs || !reward || !terminated) return fail(ANM_E_INVALID, "anm_step_host_async: null argument");
  if (!next_vars && h->H.table_len == 0)
    return fail(ANM_E_INVALID, "anm_step_host_async: next_vars is NULL but the environment has no built-in table");
  return step_host_enqueue(h, 1, action, next_vars, obs, reward, terminated, true);
}

/* ANM_HOST_ROLLOUT=zc (environment): queued host rollouts read / write the caller's pinned buffers from the kernel
 * (zero-copy) instead of staging through device memory with the copy engines (default; measured faster on B200:
 * 144-byte observation rows make poor PCIe writes, bulk DMA of [T, B, .] arrays runs at link speed). */
static bool host_rollout_zero_copy() {
  static const bool v = [] {
    const char* e = getenv("ANM_HOST_ROLLOUT");
    return e && !strcmp(e, "zc");
  }();
  return v;
}

static int hostset_reserve(anm_handle h, anm_handle_s::HostSet& hs, int64_t rows, bool need_nv) {
  const AnmConstHeader& H = h->H;
  if (!hs.ev_in) {
    CUDA_TRY(cudaEventCreateWithFlags(&hs.ev_in, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&hs.ev_k, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&hs.ev_out, cudaEventDisableTiming));
  }
  if (rows > hs.cap_rows) {
    cudaFree(hs.act); cudaFree(hs.nv); cudaFree(hs.obs); cudaFree(hs.rew); cudaFree(hs.term);
    hs.act = hs.nv = hs.obs = hs.rew = nullptr; hs.term = nullptr; hs.cap_rows = 0;
    CUDA_TRY(cudaMalloc((void**)&hs.act, (size_t)rows * H.n_action * sizeof(double)));
    CUDA_TRY(cudaMalloc((void**)&hs.obs, (size_t)rows * H.n_obs * sizeof(double)));
    CUDA_TRY(cudaMalloc((void**)&hs.rew, (size_t)rows * sizeof(double)));
    CUDA_TRY(cudaMalloc((void**)&hs.term, (size_t)rows));
    hs.cap_rows = rows;
  }
  if (need_nv && !hs.nv) CUDA_TRY(cudaMalloc((void**)&hs.nv, (size_t)hs.cap_rows * H.n_next_vars * sizeof(double)));
  return 0;
}

int anm_rollout_host_async(anm_handle h, int64_t T, const double* action, const double* next_vars, double* obs,
                           double* reward, uint8_t* terminated) {
  ANM_NVTX("anm_rollout_host_async");
  if (!h || !action || !obs || !reward || !terminated || T < 1 || T > INT32_MAX)
    return fail(ANM_E_INVALID, "anm_rollout_host_async: bad argument");
  if (!next_vars && h->H.table_len == 0)
    return fail(ANM_E_INVALID, "anm_rollout_host_async: next_vars is NULL but the environment has no built-in table");
  if (host_rollout_zero_copy()) {
    h->st_last_was_rollout = false;
    return step_host_enqueue(h, T, action, next_vars, obs, reward, terminated, true);
  }
  DeviceGuard guard(h->device);
  const AnmConstHeader& H = h->H;
  const int64_t rows = T * h->B;
  if (!h->st_in) {
    CUDA_TRY(cudaStreamCreateWithFlags(&h->st_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->st_out, cudaStreamNonBlocking));
  }
  anm_handle_s::HostSet& hs = h->hset[h->hset_next];
  h->hset_next ^= 1;
  if (hs.used) { /* the set is free once its previous kernel has read the inputs and its outputs are downloaded */
    CUDA_TRY(cudaEventSynchronize(hs.ev_k));
    CUDA_TRY(cudaEventSynchronize(hs.ev_out));
    h->wd_steps = h->wd_last; /* everything but the most recent launch is known to be complete */
  }
  if (int rc = hostset_reserve(h, hs, rows, next_vars != nullptr)) return rc;
  /* inputs: copy engine, own stream; the HOST waits for the upload (the kernels of the earlier calls are still
   * running meanwhile), so that the compute stream holds nothing but kernels and consecutive rollouts stay chained */
  CUDA_TRY(cudaMemcpyAsync(hs.act, action, (size_t)rows * H.n_action * sizeof(double), cudaMemcpyHostToDevice, h->st_in));
  if (next_vars)
    CUDA_TRY(cudaMemcpyAsync(hs.nv, next_vars, (size_t)rows * H.n_next_vars * sizeof(double), cudaMemcpyHostToDevice, h->st_in));
  CUDA_TRY(cudaEventRecord(hs.ev_in, h->st_in));
  CUDA_TRY(cudaEventSynchronize(hs.ev_in));
  AnmLaunch p;
  memset(&p, 0, sizeof(p));
  p.mode = ANM_MODE_STEP;
  p.T = (int32_t)T;
  p.action = hs.act; p.next_vars = next_vars ? hs.nv : nullptr;
  p.obs = hs.obs; p.reward = hs.rew; p.term_out = hs.term;
  int rc = launch(h, p, h->stream, h->st_last_was_rollout ? ANM_LF_CHAINED : 0u);
  if (rc) return rc;
  h->st_last_was_rollout = true;
  hs.used = true;
  CUDA_TRY(cudaEventRecord(hs.ev_k, h->stream));
  /* outputs: copy engine, own stream, after this kernel */
  CUDA_TRY(cudaStreamWaitEvent(h->st_out, hs.ev_k, 0));
  CUDA_TRY(cudaMemcpyAsync(obs, hs.obs, (size_t)rows * H.n_obs * sizeof(double), cudaMemcpyDeviceToHost, h->st_out));
  CUDA_TRY(cudaMemcpyAsync(reward, hs.rew, (size_t)rows * sizeof(double), cudaMemcpyDeviceToHost, h->st_out));
  CUDA_TRY(cudaMemcpyAsync(terminated, hs.term, (size_t)rows, cudaMemcpyDeviceToHost, h->st_out));
  CUDA_TRY(cudaEventRecord(hs.ev_out, h->st_out));
  return ANM_OK;
}

int anm_host_sync(anm_handle h) {
  if (!h) return fail(ANM_E_INVALID, "null handle");
  DeviceGuard guard(h->device);
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (h->st_out) CUDA_TRY(cudaStreamSynchronize(h->st_out));
  return check_watchdog(h);
}

int anm_host_sync_previous(anm_handle h) {
  if (!h) return fail(ANM_E_INVALID, "null handle");
  DeviceGuard guard(h->device);
  if (host_rollout_zero_copy() || !h->st_out) return anm_host_sync(h);
  anm_handle_s::HostSet& prev = h->hset[h->hset_next]; /* the set of the call before the most recent one */
  if (prev.used) CUDA_TRY(cudaEventSynchronize(prev.ev_out));
  return ANM_OK;
}

int anm_reset_host(anm_handle h, const double* s0, const uint8_t* mask, double* obs, double* state, uint8_t* converged) {
  ANM_NVTX("anm_reset_host");
  if (!h || !s0 || !obs || !converged) return fail(ANM_E_INVALID, "anm_reset_host: null argument");
  DeviceGuard guard(h->device);
  h->st_last_was_rollout = false;
  const AnmConstHeader& H = h->H;
  const size_t B = (size_t)h->B;
  cudaStream_t st = h->stream;
  CUDA_TRY(cudaMemcpyAsync(h->s_s0, s0, B * H.n_state * sizeof(double), cudaMemcpyHostToDevice, st));
  if (mask) CUDA_TRY(cudaMemcpyAsync(h->s_mask, mask, B, cudaMemcpyHostToDevice, st));
  /* rows of unselected envs are left untouched: pre-load the caller's buffers */
  CUDA_TRY(cudaMemcpyAsync(h->s_obs, obs, B * H.n_obs * sizeof(double), cudaMemcpyHostToDevice, st));
  if (state) CUDA_TRY(cudaMemcpyAsync(h->s_state, state, B * H.n_state * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(h->s_term, 0, B, st));
  int rc = anm_reset(h, h->s_s0, mask ? h->s_mask : nullptr, h->s_obs, state ? h->s_state : nullptr, h->s_term, st);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(obs, h->s_obs, B * H.n_obs * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (state) CUDA_TRY(cudaMemcpyAsync(state, h->s_state, B * H.n_state * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(converged, h->s_term, B, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return ANM_OK;
}

void* anm_host_stream(anm_handle h) { return h ? (void*)h->stream : nullptr; }

int64_t anm_launch_count(anm_handle h) { return h ? h->launches : 0; }

int anm_watchdog(anm_handle h, uint32_t* out8) {
  if (!h || !out8) return fail(ANM_E_INVALID, "null argument");
  for (int i = 0; i < ANM_WD_WORDS; ++i) out8[i] = h->wd_host ? ((volatile uint32_t*)h->wd_host)[i] : 0u;
  return ANM_OK;
}

}  // extern "C"
