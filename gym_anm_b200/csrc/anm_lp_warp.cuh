// anm_lp_warp.cuh -- the bounded dual simplex of anm_lp.cuh with one WARP per program (include/anm_lp.h).
//
// Same algorithm, same tolerances, same pivoting rules (ties towards the smaller index, so that both kernels walk the
// same vertices up to the rounding of the x_B evaluations); what changes is the mapping.  With one thread per program a
// warp executes the union of its 32 programs' pivots on thread-private, uncoalesced addresses
// (profiles/r02_config5_gpu_lp.md: 120-150 ms per solve of 16 384 programs).  Here a program's tableau is ROW-MAJOR
// and contiguous, the 32 lanes of its warp sweep the columns of a row (coalesced 256-byte requests), a pivot touches
// only the rows whose pivot-column entry is non-zero, and nothing diverges: every decision (leaving row, entering
// column, status) is warp-uniform.
//
// Programming model (so that the SAME source also runs on the host, as the test hook): the code alternates between
//   * lane phases   LPW_LANES ... LPW_END  -- on the device the body runs once per lane and ends in __syncwarp(); on
//     the host it is a loop over the 32 lanes (forwards or backwards: a phase whose lanes depended on each other
//     would give different results in the two orders, which the CPU suite checks);
//   * uniform code -- scalars every lane computes identically from memory (read-only), once on the host.
// Lanes keep NO private state across phases: everything a later phase needs goes through the program's memory block
// or the warp's scratch (shared memory on the device).  Reductions: every lane leaves its partial result in the
// scratch, then all lanes fold the 32 partials in lane order (deterministic; ties resolved by index).
#pragma once
#include "anm_lp.cuh"

namespace anm_lp {

struct WarpBatch {
  int32_t n, m, max_iter;
  int64_t stride;      // of the interleaved lo / up / x arrays the caller passes
  int64_t block;       // doubles per program block
  const double* A;
  const double* c;
  double* mem;         // [batch, block]: T[m*n] d[n] xB[m] bl[m] bu[m] nl[n] nu[n] rval[n] cq[m] | bid[m] nid[n] atup[n] rows[m] (int32)
};

LP_HD inline int64_t warp_block_doubles(int64_t n, int64_t m) {
  const int64_t dbl = m * n + 4 * n + 4 * m, ints = 2 * m + 2 * n;
  return (dbl + (ints + 1) / 2 + 31) / 32 * 32;  // 256-byte multiple
}

constexpr int kColScratch = 256;  // programs with m <= this keep the pivot column and its row list in the warp's scratch

struct WarpScratch {   // device: shared memory of the program's warp; host: local arrays
  double* tile;        // [32 * 33]
  double* rv;          // [32]
  int32_t* ri;         // [32]
  double* col;         // [kColScratch]: the pivot column (read by every lane for every row of the update)
  int32_t* rows;       // [kColScratch]: the rows a pivot touches
  bool rev;            // host only: run the lanes of a phase backwards
};

#if defined(__CUDA_ARCH__)
#define LPW_LANES { const int lane = (int)(threadIdx.x & 31u);
#define LPW_END } __syncwarp();
#define LPW_SYNC() __syncwarp()
#else
#define LPW_LANES for (int l_ = 0; l_ < 32; ++l_) { const int lane = s.rev ? 31 - l_ : l_;
#define LPW_END }
#define LPW_SYNC() ((void)0)
#endif

struct WarpView {
  double *T, *d, *xB, *bl, *bu, *nl, *nu, *rval, *cq;
  int32_t *bid, *nid, *atup, *rows;
};

LP_HD inline WarpView warp_view(const WarpBatch& b, int64_t e) {
  WarpView v;
  double* p = b.mem + e * b.block;
  const int64_t n = b.n, m = b.m;
  v.T = p, p += m * n;
  v.d = p, p += n;
  v.xB = p, p += m;
  v.bl = p, p += m;
  v.bu = p, p += m;
  v.nl = p, p += n;
  v.nu = p, p += n;
  v.rval = p, p += n;
  v.cq = p, p += m;
  int32_t* q = reinterpret_cast<int32_t*>(p);
  v.bid = q, q += m;
  v.nid = q, q += n;
  v.atup = q, q += n;
  v.rows = q;
  return v;
}

// bit 0: a flag changed; bit 1: a column's cost pulls towards an infinite bound
LP_HD inline int warp_restore_dual_feasibility(const WarpBatch& b, const WarpView& v, WarpScratch& s) {
  const int n = b.n;
  LPW_LANES
    int fl = 0;
    for (int j = lane; j < n; j += 32) {
      const double l = v.nl[j], u = v.nu[j];
      if (l < u) {
        const double dj = v.d[j];
        int want = v.atup[j];
        if (dj > kDualTol) {
          want = 0;
          if (is_inf(l)) fl |= 2;
        } else if (dj < -kDualTol) {
          want = 1;
          if (is_inf(u)) fl |= 2;
        } else {
          if (want && is_inf(u)) want = 0;
          if (!want && is_inf(l) && !is_inf(u)) want = 1;
        }
        if (want != v.atup[j]) {
          v.atup[j] = want;
          fl |= 1;
        }
      }
    }
    s.ri[lane] = fl;
  LPW_END
  int flags = 0;
  for (int l = 0; l < 32; ++l) flags |= s.ri[l];
  LPW_SYNC();
  return flags;
}

LP_HD inline void warp_compute_basic_values(const WarpBatch& b, const WarpView& v, WarpScratch& s) {
  const int n = b.n, m = b.m;
  LPW_LANES
    for (int j = lane; j < n; j += 32) v.rval[j] = nonbasic_value(v.nl[j], v.nu[j], v.atup[j] != 0);
  LPW_END
  for (int i0 = 0; i0 < m; i0 += 32) {
    const int rows = m - i0 < 32 ? m - i0 : 32;
    LPW_LANES
      // This lane's share of each of the 32 rows' dot products, eight rows at a time: eight independent loads per
      // column in flight (a row at a time, every row waits for its own loads: the evaluation was latency-bound).
      for (int i1 = 0; i1 < rows; i1 += 8) {
        const int cnt = rows - i1 < 8 ? rows - i1 : 8;
        const double* T0 = v.T + (int64_t)(i0 + i1) * n;
        double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (cnt == 8) {
          for (int j = lane; j < n; j += 32) {
            const double xv = v.rval[j];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fma(T0[(int64_t)k * n + j], xv, acc[k]);
          }
        } else {
          for (int j = lane; j < n; j += 32) {
            const double xv = v.rval[j];
            for (int k = 0; k < cnt; ++k) acc[k] = fma(T0[(int64_t)k * n + j], xv, acc[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < cnt) s.tile[(i1 + k) * 33 + lane] = acc[k];
      }
    LPW_END
    LPW_LANES
      if (lane < rows) {  // lane k folds row i0 + k
        double acc = 0.0;
        for (int l = 0; l < 32; ++l) acc += s.tile[lane * 33 + l];
        v.xB[i0 + lane] = acc;
      }
    LPW_END
  }
}

LP_HD inline void warp_gather_bounds(const WarpBatch& b, const WarpView& v, int64_t e, const double* lo, const double* up,
                                     WarpScratch& s) {
  const int64_t S = b.stride;
  LPW_LANES
    for (int i = lane; i < b.m; i += 32) {
      const int64_t var = v.bid[i];
      v.bl[i] = lo[var * S + e];
      v.bu[i] = up[var * S + e];
    }
    for (int j = lane; j < b.n; j += 32) {
      const int64_t var = v.nid[j];
      v.nl[j] = lo[var * S + e];
      v.nu[j] = up[var * S + e];
    }
  LPW_END
}

LP_HD inline bool warp_eligible(double l, double u, bool atu, double as, double piv) {
  const bool fr = is_inf(l) && is_inf(u);
  return (fr && fabs(as) > piv) || (!atu && as > piv) || (atu && as < -piv);
}

// One program, solved by the 32 lanes of its warp.  Returns the status (uniform); *obj, *iters are uniform too.
LP_HD inline int solve_one_warp(const WarpBatch& b, int64_t e, const double* lo, const double* up, bool restart,
                                double* x, double* obj, int32_t* iters, WarpScratch& s) {
  const int n = b.n, m = b.m;
  const int64_t S = b.stride;
  const WarpView v = warp_view(b, e);
  int status = ANM_LP_OPTIMAL, it = 0;
  for (int attempt = 0;; ++attempt) {
    if (restart) {
      LPW_LANES
        for (int k = lane; k < m * n; k += 32) v.T[k] = b.A[k];
        for (int i = lane; i < m; i += 32) v.bid[i] = n + i;
        for (int j = lane; j < n; j += 32) {
          v.d[j] = b.c[j];
          v.nid[j] = j;
          v.atup[j] = 0;
        }
      LPW_END
    }
    warp_gather_bounds(b, v, e, lo, up, s);
    const int flags = warp_restore_dual_feasibility(b, v, s);
    if (!(flags & 2)) break;
    if (restart || attempt > 0) {  // (see solve_one: a kept basis whose bounds have become infinite restarts once)
      status = ANM_LP_DUAL_INFEASIBLE;
      break;
    }
    restart = true;
  }
  warp_compute_basic_values(b, v, s);
  bool verified = true;
  int it_eval = 0;
  while (status == ANM_LP_OPTIMAL) {
    // ---- leaving row: largest scaled bound violation, ties towards the smaller row
    LPW_LANES
      double worst = 0.0;
      int r_l = -1;
      for (int i = lane; i < m; i += 32) {
        const double xi = v.xB[i], l = v.bl[i], u = v.bu[i];
        const double below = (l - xi) / (1.0 + fabs(l)), above = (xi - u) / (1.0 + fabs(u));
        if (below > kFeasTol && below > worst) worst = below, r_l = i;
        if (above > kFeasTol && above > worst) worst = above, r_l = i;
      }
      s.rv[lane] = worst;
      s.ri[lane] = r_l;
    LPW_END
    int r = -1;
    {
      double worst = 0.0;
      for (int l = 0; l < 32; ++l) {
        const double w = s.rv[l];
        const int i = s.ri[l];
        if (i >= 0 && (w > worst || (w == worst && (r < 0 || i < r)))) worst = w, r = i;
      }
    }
    LPW_SYNC();
    if (r < 0) {
      if (verified) break;
      const int flags = warp_restore_dual_feasibility(b, v, s);
      verified = true;
      if ((flags & 1) || it - it_eval > kPivotsPerEval || (restart && it_eval == 0 && it > 0)) {
        warp_compute_basic_values(b, v, s);
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
    double target, sgn;
    {
      const double xr = v.xB[r], l = v.bl[r], u = v.bu[r];
      const double below = (l - xr) / (1.0 + fabs(l)), above = (xr - u) / (1.0 + fabs(u));
      if (below > kFeasTol && !(above > below)) target = l, sgn = 1.0;
      else target = u, sgn = -1.0;
    }
    // ---- the pivot row (copied into rval) and its largest entry
    double* Tr = v.T + (int64_t)r * n;
    LPW_LANES
      double am = 0.0;
      for (int j = lane; j < n; j += 32) {
        const double a = Tr[j];
        v.rval[j] = a;
        am = fmax(am, fabs(a));
      }
      s.rv[lane] = am;
    LPW_END
    double amax = 0.0;
    for (int l = 0; l < 32; ++l) amax = fmax(amax, s.rv[l]);
    LPW_SYNC();
    const double piv = fmax(kPivAbs, kPivRel * amax);
    // ---- Harris pass 1: the largest dual step that keeps every reduced cost within the tolerance
    LPW_LANES
      double tm = kInf;
      for (int j = lane; j < n; j += 32) {
        const double a = v.rval[j];
        const double l = v.nl[j], u = v.nu[j];
        if (a != 0.0 && l < u && warp_eligible(l, u, v.atup[j] != 0, sgn * a, piv))
          tm = fmin(tm, (fabs(v.d[j]) + kDualTol) / fabs(a));
      }
      s.rv[lane] = tm;
    LPW_END
    double tmax = kInf;
    for (int l = 0; l < 32; ++l) tmax = fmin(tmax, s.rv[l]);
    LPW_SYNC();
    if (tmax >= kInf) {
      status = ANM_LP_INFEASIBLE;
      break;
    }
    // ---- Harris pass 2: among the columns within that step, the largest pivot (ties towards the smaller column)
    LPW_LANES
      double best = 0.0;
      int q_l = -1;
      for (int j = lane; j < n; j += 32) {
        const double a = v.rval[j];
        const double l = v.nl[j], u = v.nu[j];
        if (a != 0.0 && l < u && warp_eligible(l, u, v.atup[j] != 0, sgn * a, piv) &&
            fabs(v.d[j]) <= tmax * fabs(a) && fabs(a) > best)
          best = fabs(a), q_l = j;
      }
      s.rv[lane] = best;
      s.ri[lane] = q_l;
    LPW_END
    int q = -1;
    {
      double best = 0.0;
      for (int l = 0; l < 32; ++l) {
        const double w = s.rv[l];
        const int j = s.ri[l];
        if (j >= 0 && (w > best || (w == best && (q < 0 || j < q)))) best = w, q = j;
      }
    }
    // ---- the pivot's scalars (read before any lane changes what they come from)
    const double p = v.rval[q], inv_p = 1.0 / p;
    const double delta = (target - v.xB[r]) * inv_p;
    const int ev = v.nid[q], lv = v.bid[r];
    const double el = v.nl[q], eu = v.nu[q], blr = v.bl[r], bur = v.bu[r];
    const double x_enter = nonbasic_value(el, eu, v.atup[q] != 0) + delta;
    const double fd = v.d[q] * inv_p;
    LPW_SYNC();
    // ---- the pivot column (into the warp's scratch when it fits) and the list of the rows it touches
    double* const cq = m <= kColScratch ? s.col : v.cq;
    int32_t* const rows = m <= kColScratch ? s.rows : v.rows;
    LPW_LANES
      int cnt = 0;
      for (int i = lane; i < m; i += 32) {
        const double tiq = v.T[(int64_t)i * n + q];
        cq[i] = tiq;
        cnt += (i != r && tiq != 0.0) ? 1 : 0;
      }
      s.ri[lane] = cnt;
    LPW_END
    LPW_LANES
      int at = 0;  // this lane's rows go behind those of the lanes before it
      for (int l = 0; l < lane; ++l) at += s.ri[l];
      for (int i = lane; i < m; i += 32)
        if (i != r && cq[i] != 0.0) rows[at++] = i;
    LPW_END
    int touched = 0;
    for (int l = 0; l < 32; ++l) touched += s.ri[l];
    LPW_SYNC();
    // ---- the update of those rows, four at a time: the four rows' loads are in flight together (row by row, each
    //      row's read-modify-write waited for its own loads: ~1 us per touched row)
    LPW_LANES
      int t = 0;
      for (; t + 3 < touched; t += 4) {
        const int i0 = rows[t], i1 = rows[t + 1], i2 = rows[t + 2], i3 = rows[t + 3];
        const double q0 = cq[i0], q1 = cq[i1], q2 = cq[i2], q3 = cq[i3];
        const double f0 = q0 * inv_p, f1 = q1 * inv_p, f2 = q2 * inv_p, f3 = q3 * inv_p;
        double* T0 = v.T + (int64_t)i0 * n;
        double* T1 = v.T + (int64_t)i1 * n;
        double* T2 = v.T + (int64_t)i2 * n;
        double* T3 = v.T + (int64_t)i3 * n;
        for (int j = lane; j < n; j += 32) {
          if (j == q) {
            T0[j] = f0, T1[j] = f1, T2[j] = f2, T3[j] = f3;
          } else {
            const double a = v.rval[j];
            if (a != 0.0) {
              const double t0 = T0[j], t1 = T1[j], t2 = T2[j], t3 = T3[j];
              T0[j] = fma(-f0, a, t0);
              T1[j] = fma(-f1, a, t1);
              T2[j] = fma(-f2, a, t2);
              T3[j] = fma(-f3, a, t3);
            }
          }
        }
        if ((i0 & 31) == lane) v.xB[i0] = fma(q0, delta, v.xB[i0]);
        if ((i1 & 31) == lane) v.xB[i1] = fma(q1, delta, v.xB[i1]);
        if ((i2 & 31) == lane) v.xB[i2] = fma(q2, delta, v.xB[i2]);
        if ((i3 & 31) == lane) v.xB[i3] = fma(q3, delta, v.xB[i3]);
      }
      for (; t < touched; ++t) {
        const int i = rows[t];
        const double tiq = cq[i], f = tiq * inv_p;
        double* Ti = v.T + (int64_t)i * n;
        for (int j = lane; j < n; j += 32) {
          if (j == q) {
            Ti[j] = f;
          } else {
            const double a = v.rval[j];
            if (a != 0.0) Ti[j] = fma(-f, a, Ti[j]);
          }
        }
        if ((i & 31) == lane) v.xB[i] = fma(tiq, delta, v.xB[i]);
      }
    LPW_END
    // ---- row r and the reduced costs
    LPW_LANES
      for (int j = lane; j < n; j += 32) {
        const double a = v.rval[j];
        if (j == q) {
          Tr[j] = inv_p;
          v.d[j] = fd;
        } else if (a != 0.0) {
          Tr[j] = -a * inv_p;
          double dj = fma(-fd, a, v.d[j]);
          if (v.nl[j] < v.nu[j]) {  // Harris: a sign lost within the tolerance is a zero
            const bool atu = v.atup[j] != 0;
            if ((!atu && dj < 0.0 && dj > -16.0 * kDualTol) || (atu && dj > 0.0 && dj < 16.0 * kDualTol)) dj = 0.0;
          }
          v.d[j] = dj;
        }
      }
    LPW_END
    LPW_LANES
      if (lane == 0) {
        v.bid[r] = ev;
        v.nid[q] = lv;
        v.nl[q] = blr, v.nu[q] = bur;
        v.bl[r] = el, v.bu[r] = eu;
        v.atup[q] = sgn < 0.0 ? 1 : 0;
        v.xB[r] = x_enter;
      }
    LPW_END
    ++it;
  }
  // ---- the columns' values and the objective
  LPW_LANES
    for (int j = lane; j < n; j += 32) {
      const int64_t var = v.nid[j];
      if (var < n) x[var * S + e] = nonbasic_value(v.nl[j], v.nu[j], v.atup[j] != 0);
    }
    for (int i = lane; i < m; i += 32) {
      const int64_t var = v.bid[i];
      if (var < n) x[var * S + e] = v.xB[i];
    }
  LPW_END
  LPW_LANES
    double acc = 0.0;
    for (int j = lane; j < n; j += 32) acc = fma(b.c[j], x[(int64_t)j * S + e], acc);
    s.rv[lane] = acc;
  LPW_END
  double z = 0.0;
  for (int l = 0; l < 32; ++l) z += s.rv[l];
  LPW_SYNC();
  *obj = z;
  *iters = it;
  return status;
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(32) lp_solve_warp_kernel(WarpBatch b, int64_t batch, const double* __restrict__ lo,
                                                           const double* __restrict__ up,
                                                           const uint8_t* __restrict__ restart, int restart_all,
                                                           double* x, double* obj, int32_t* status, int32_t* iters) {
  __shared__ double tile[32 * 33];
  __shared__ double rv[32];
  __shared__ int32_t ri[32];
  __shared__ double col[kColScratch];
  __shared__ int32_t rows[kColScratch];
  const int64_t e = blockIdx.x;  // one warp = one block = one program
  if (e >= batch) return;
  WarpScratch s{tile, rv, ri, col, rows, false};
  double z;
  int32_t it;
  const bool rs = restart_all || (restart != nullptr && restart[e] != 0);
  const int st = solve_one_warp(b, e, lo, up, rs, x, &z, &it, s);
  if (threadIdx.x == 0) {
    if (obj) obj[e] = z;
    if (status) status[e] = st;
    if (iters) iters[e] = it;
  }
}
#endif

}  // namespace anm_lp
