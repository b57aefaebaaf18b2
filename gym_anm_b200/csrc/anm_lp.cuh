// anm_lp.cuh -- bounded dual simplex on a condensed tableau, one program per thread (include/anm_lp.h).
//
// The programs of a batch share A [m, n] and c [n]; an instance's tableau T [m, n], reduced costs d [n], basis and
// bound flags stay in HBM between solves (warm start: MPC steps change bounds only, reference
// gym_anm/agents/mpc.py:395-420 `_update_parameters`).  The same code is compiled for the host (test hook
// anm_debug_lp_solve_host) and for the device (lp_solve_kernel): no warp intrinsics, no shared memory -- a thread
// owns its program, arrays are interleaved over the batch ([k * stride + e]) so that neighbouring threads touch
// neighbouring addresses.
//
// Formulation.  Variables 0..n-1 are the columns, n..n+m-1 the row activities s = A x ("logicals"); every
// variable v has bounds lo[v] <= . <= up[v].  A basis holds m variables; the condensed tableau states the basic
// variables as a HOMOGENEOUS linear function of the non-basic ones, x_B = T x_N (A x - s = 0 has no right-hand
// side: bounds enter through the values of the non-basic variables only, which is why a change of bounds leaves
// T, d and the dual feasibility of the basis untouched).  A non-basic variable sits at its lower or upper bound
// (flag `atup`); dual feasible means d_j >= 0 at a lower bound, d_j <= 0 at an upper bound (fixed variables: any).
//
// Iteration (textbook dual simplex with bounds, Harris two-pass ratio test):
//   leaving row r   = the basic variable with the largest bound violation (none: optimal);
//   entering col q  = argmin |d_j| / |T_rj| over the non-basic j that can move x_B[r] towards its violated bound,
//                     ties (within the dual tolerance) broken towards the largest |T_rj|;
//   pivot on T_rq   : row r, the other rows with T_iq != 0, and d -- only over the non-zeros of row r (the
//                     stage structure of the MPC program keeps rows sparse).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LP_HD __host__ __device__
#else
#define LP_HD
#endif

namespace anm_lp {

constexpr double kFeasTol = 1e-9;   // bound violation that makes a basic variable leave (scaled by 1 + |bound|)
constexpr double kDualTol = 1e-9;   // Harris slack on the reduced costs
constexpr double kPivRel = 1e-7;    // pivot candidates below kPivRel * max |row| are ignored
constexpr double kPivAbs = 1e-11;
constexpr int kPivotsPerEval = 24;  // x_B is re-evaluated from the tableau before "optimal" after more pivots than this
constexpr double kInf = 1e300;      // bounds beyond +-kInf/2 count as infinite (HUGE_VAL included)

struct Batch {
  int32_t n, m, max_iter;
  int64_t stride;
  const double* A;  // [m, n] shared
  const double* c;  // [n]    shared
  double* T;        // [m * n, stride]
  double* d;        // [n, stride]
  int32_t* bid;     // [m, stride]  variable that is basic in row i
  int32_t* nid;     // [n, stride]  variable that is non-basic in column j
  uint8_t* atup;    // [n, stride]  non-basic at its upper bound
  double* xB;       // [m, stride]  scratch: values of the basic variables
  double* rval;     // [n, stride]  scratch: non-zeros of the pivot row
  int32_t* ridx;    // [n, stride]
  double* bl;       // [m, stride]  scratch: bounds of the basic variables, in row order (gathered per solve, swapped
  double* bu;       // [m, stride]           with the column's on a pivot: no lo[bid[i]] indirection in the loops)
  double* nl;       // [n, stride]  scratch: bounds of the non-basic variables, in column order
  double* nu;       // [n, stride]
};

LP_HD inline bool is_inf(double v) { return v > 0.5 * kInf || v < -0.5 * kInf; }

// value of a non-basic variable: the bound its flag names (the finite one if that bound is infinite, 0 if free)
LP_HD inline double nonbasic_value(double l, double u, bool up) {
  if (up) return !is_inf(u) ? u : (!is_inf(l) ? l : 0.0);
  return !is_inf(l) ? l : (!is_inf(u) ? u : 0.0);
}

#define LP_AT(arr, k) (arr)[(int64_t)(k) * S + e]

// Bound flags follow the sign of the reduced costs (boxed variables: always possible).  Returns false when a
// column's cost pulls towards an infinite bound.
LP_HD inline bool restore_dual_feasibility(const Batch& b, int64_t e, bool* changed) {
  const int64_t S = b.stride;
  bool ok = true;
  *changed = false;
  for (int j = 0; j < b.n; ++j) {
    const double l = LP_AT(b.nl, j), u = LP_AT(b.nu, j);
    if (!(l < u)) continue;  // fixed
    const double dj = LP_AT(b.d, j);
    uint8_t want = LP_AT(b.atup, j);
    if (dj > kDualTol) {
      want = 0;
      if (is_inf(l)) ok = false;
    } else if (dj < -kDualTol) {
      want = 1;
      if (is_inf(u)) ok = false;
    } else {
      if (want && is_inf(u)) want = 0;
      if (!want && is_inf(l) && !is_inf(u)) want = 1;
    }
    if (want != LP_AT(b.atup, j)) {
      LP_AT(b.atup, j) = want;
      *changed = true;
    }
  }
  return ok;
}

LP_HD inline void compute_basic_values(const Batch& b, int64_t e) {
  const int64_t S = b.stride;
  const int n = b.n, m = b.m;
  for (int j = 0; j < n; ++j)  // rval doubles as the vector of non-basic values here
    LP_AT(b.rval, j) = nonbasic_value(LP_AT(b.nl, j), LP_AT(b.nu, j), LP_AT(b.atup, j) != 0);
  for (int i = 0; i < m; ++i) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;  // four independent chains: loads in flight, not FMA latency
    const double* Ti = b.T + (int64_t)i * n * S + e;
    int j = 0;
    for (; j + 3 < n; j += 4) {
      s0 = fma(Ti[(int64_t)j * S], LP_AT(b.rval, j), s0);
      s1 = fma(Ti[(int64_t)(j + 1) * S], LP_AT(b.rval, j + 1), s1);
      s2 = fma(Ti[(int64_t)(j + 2) * S], LP_AT(b.rval, j + 2), s2);
      s3 = fma(Ti[(int64_t)(j + 3) * S], LP_AT(b.rval, j + 3), s3);
    }
    for (; j < n; ++j) s0 = fma(Ti[(int64_t)j * S], LP_AT(b.rval, j), s0);
    LP_AT(b.xB, i) = (s0 + s1) + (s2 + s3);
  }
}

// lo / up ([variable, stride]) gathered into row order (basic) and column order (non-basic)
LP_HD inline void gather_bounds(const Batch& b, int64_t e, const double* lo, const double* up) {
  const int64_t S = b.stride;
  for (int i = 0; i < b.m; ++i) {
    const int v = LP_AT(b.bid, i);
    LP_AT(b.bl, i) = LP_AT(lo, v);
    LP_AT(b.bu, i) = LP_AT(up, v);
  }
  for (int j = 0; j < b.n; ++j) {
    const int v = LP_AT(b.nid, j);
    LP_AT(b.nl, j) = LP_AT(lo, v);
    LP_AT(b.nu, j) = LP_AT(up, v);
  }
}

// One program.  Returns the status; *iters = pivots made.
LP_HD inline int solve_one(const Batch& b, int64_t e, const double* lo, const double* up, bool restart, double* x,
                           double* obj, int32_t* iters) {
  const int64_t S = b.stride;
  const int n = b.n, m = b.m;
  int status = ANM_LP_OPTIMAL, it = 0;
  bool changed;
  for (int attempt = 0;; ++attempt) {
    if (restart) {
      for (int i = 0; i < m; ++i) {
        LP_AT(b.bid, i) = n + i;
        for (int j = 0; j < n; ++j) LP_AT(b.T, i * n + j) = b.A[i * n + j];
      }
      for (int j = 0; j < n; ++j) {
        LP_AT(b.d, j) = b.c[j];
        LP_AT(b.nid, j) = j;
        LP_AT(b.atup, j) = 0;
      }
    }
    gather_bounds(b, e, lo, up);
    if (restore_dual_feasibility(b, e, &changed)) break;
    // A kept basis can lose its dual feasibility when a bound that a non-basic variable sits at has become
    // infinite since the last solve: start again from the all-slack basis (only columns are non-basic there).
    if (restart || attempt > 0) {
      status = ANM_LP_DUAL_INFEASIBLE;
      break;
    }
    restart = true;
  }
  compute_basic_values(b, e);
  bool verified = true;  // nothing has changed since the bound flags were last checked against d
  int it_eval = 0;       // pivot count at the last evaluation of x_B from scratch
  while (status == ANM_LP_OPTIMAL) {
    // ---- leaving row: largest scaled bound violation
    int r = -1;
    double worst = 0.0, target = 0.0, sgn = 0.0;
    for (int i = 0; i < m; ++i) {
      const double xi = LP_AT(b.xB, i), l = LP_AT(b.bl, i), u = LP_AT(b.bu, i);
      const double below = (l - xi) / (1.0 + fabs(l)), above = (xi - u) / (1.0 + fabs(u));
      if (below > kFeasTol && below > worst) worst = below, r = i, target = l, sgn = 1.0;
      if (above > kFeasTol && above > worst) worst = above, r = i, target = u, sgn = -1.0;
    }
    if (r < 0) {
      if (verified) break;
      // Bound flags against the signs of d once more; x_B from scratch if a flag moved or many pivots have been
      // accumulated into it since the solve's initial evaluation (rounding drift), otherwise the updated values stand.
      restore_dual_feasibility(b, e, &changed);
      verified = true;
      if (changed || it - it_eval > kPivotsPerEval || (restart && it_eval == 0 && it > 0)) {
        compute_basic_values(b, e);  // (a cold solve always: its first pivots move the big-M boxes' huge values)
        it_eval = it;
        continue;
      }
      break;
    }
    if (it >= b.max_iter) {
      status = ANM_LP_ITER_LIMIT;
      break;
    }
    verified = false;
    // ---- non-zeros of row r, Harris ratio test
    const double* Tr = b.T + (int64_t)r * n * S + e;
    int nnz = 0;
    double amax = 0.0;
    for (int j = 0; j < n; ++j) {
      const double a = Tr[(int64_t)j * S];
      if (a != 0.0) {
        LP_AT(b.ridx, nnz) = j;
        LP_AT(b.rval, nnz) = a;
        ++nnz;
        amax = fmax(amax, fabs(a));
      }
    }
    const double piv = fmax(kPivAbs, kPivRel * amax);
    double tmax = kInf;
    for (int k = 0; k < nnz; ++k) {
      const int j = LP_AT(b.ridx, k);
      const double l = LP_AT(b.nl, j), u = LP_AT(b.nu, j);
      if (!(l < u)) continue;
      const double as = sgn * LP_AT(b.rval, k);
      const bool atu = LP_AT(b.atup, j) != 0, fr = is_inf(l) && is_inf(u);
      if ((fr && fabs(as) > piv) || (!atu && as > piv) || (atu && as < -piv))
        tmax = fmin(tmax, (fabs(LP_AT(b.d, j)) + kDualTol) / fabs(as));
    }
    if (tmax >= kInf) {
      status = ANM_LP_INFEASIBLE;
      break;
    }
    int q = -1, kq = -1;
    double best = 0.0;
    for (int k = 0; k < nnz; ++k) {
      const int j = LP_AT(b.ridx, k);
      const double l = LP_AT(b.nl, j), u = LP_AT(b.nu, j);
      if (!(l < u)) continue;
      const double a = LP_AT(b.rval, k), as = sgn * a;
      const bool atu = LP_AT(b.atup, j) != 0, fr = is_inf(l) && is_inf(u);
      if (!((fr && fabs(as) > piv) || (!atu && as > piv) || (atu && as < -piv))) continue;
      if (fabs(LP_AT(b.d, j)) <= tmax * fabs(a) && fabs(a) > best) best = fabs(a), q = j, kq = k;
    }
    // (q exists: the candidate that set tmax satisfies |d| <= tmax |a|)
    const double p = LP_AT(b.rval, kq), inv_p = 1.0 / p;
    const double delta = (target - LP_AT(b.xB, r)) * inv_p;  // change of the entering variable
    const int ev = LP_AT(b.nid, q), lv = LP_AT(b.bid, r);
    const double el = LP_AT(b.nl, q), eu = LP_AT(b.nu, q);
    const double x_enter = nonbasic_value(el, eu, LP_AT(b.atup, q) != 0) + delta;
    // ---- pivot: the other rows
    for (int i = 0; i < m; ++i) {
      if (i == r) continue;
      double* Ti = b.T + (int64_t)i * n * S + e;
      const double tiq = Ti[(int64_t)q * S];
      if (tiq == 0.0) continue;
      LP_AT(b.xB, i) = fma(tiq, delta, LP_AT(b.xB, i));
      const double f = tiq * inv_p;
      for (int k = 0; k < nnz; ++k) {
        const int j = LP_AT(b.ridx, k);
        if (j != q) Ti[(int64_t)j * S] = fma(-f, LP_AT(b.rval, k), Ti[(int64_t)j * S]);
      }
      Ti[(int64_t)q * S] = f;
    }
    // ---- reduced costs and row r
    const double fd = LP_AT(b.d, q) * inv_p;
    double* Trw = b.T + (int64_t)r * n * S + e;
    for (int k = 0; k < nnz; ++k) {
      const int j = LP_AT(b.ridx, k);
      if (j == q) continue;
      const double a = LP_AT(b.rval, k);
      Trw[(int64_t)j * S] = -a * inv_p;
      double dj = fma(-fd, a, LP_AT(b.d, j));
      if (LP_AT(b.nl, j) < LP_AT(b.nu, j)) {  // Harris: a sign lost within the tolerance is a zero
        const bool atu = LP_AT(b.atup, j) != 0;
        if ((!atu && dj < 0.0 && dj > -16.0 * kDualTol) || (atu && dj > 0.0 && dj < 16.0 * kDualTol)) dj = 0.0;
      }
      LP_AT(b.d, j) = dj;
    }
    Trw[(int64_t)q * S] = inv_p;
    LP_AT(b.d, q) = fd;
    LP_AT(b.bid, r) = ev;
    LP_AT(b.nid, q) = lv;
    LP_AT(b.nl, q) = LP_AT(b.bl, r), LP_AT(b.nu, q) = LP_AT(b.bu, r);
    LP_AT(b.bl, r) = el, LP_AT(b.bu, r) = eu;
    LP_AT(b.atup, q) = sgn < 0.0 ? 1 : 0;
    LP_AT(b.xB, r) = x_enter;
    ++it;
  }
  // ---- the columns' values and the objective (original costs)
  for (int j = 0; j < n; ++j) {
    const int v = LP_AT(b.nid, j);
    if (v < n) LP_AT(x, v) = nonbasic_value(LP_AT(b.nl, j), LP_AT(b.nu, j), LP_AT(b.atup, j) != 0);
  }
  for (int i = 0; i < m; ++i) {
    const int v = LP_AT(b.bid, i);
    if (v < n) LP_AT(x, v) = LP_AT(b.xB, i);
  }
  double z = 0.0;
  for (int j = 0; j < n; ++j) z = fma(b.c[j], LP_AT(x, j), z);
  *obj = z;
  *iters = it;
  return status;
}

#undef LP_AT

#if defined(__CUDACC__)
__global__ void __launch_bounds__(32) lp_solve_kernel(Batch b, int64_t batch, const double* __restrict__ lo,
                                                      const double* __restrict__ up,
                                                      const uint8_t* __restrict__ restart, int restart_all, double* x,
                                                      double* obj, int32_t* status, int32_t* iters) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= batch) return;
  double z;
  int32_t it;
  const bool rs = restart_all || (restart != nullptr && restart[e] != 0);
  const int st = solve_one(b, e, lo, up, rs, x, &z, &it);
  if (obj) obj[e] = z;
  if (status) status[e] = st;
  if (iters) iters[e] = it;
}
#endif

}  // namespace anm_lp
