/*
 * anm_oracle.c -- CPU restatement of the reference's per-timestep path, plain C, fp64.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (gym_anm_b200/) links, loads or calls
 * this file; it is the checker used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg.  One environment per loop iteration, scalar code, optional OpenMP over
 * the batch.  It is pinned (tests/test_oracle_golden.py) against the fixtures in
 * tests/golden/, which were produced by the UNMODIFIED reference (oracle/gen_golden.py).
 *
 * Each function cites the reference lines it restates (paths relative to
 * /root/reference/gym_anm/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/anm_b200.h"

#define MAXBUS 64
#define MAXDEV 128
#define MAXBR 160
#define MAXUNK (2 * (MAXBUS - 1))
#define NR_TOL 1e-5 /* simulator.py:529 xtol=1e-5 */
#define NR_MAXIT 100 /* solve_load_flow.py:176 lim_iter=100 */
#define FEAS_TOL 1e-12

typedef struct {
  double re, im;
} cplx;
static inline cplx cmul(cplx a, cplx b) { return (cplx){a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
static inline cplx cconj(cplx a) { return (cplx){a.re, -a.im}; }
static inline cplx cadd(cplx a, cplx b) { return (cplx){a.re + b.re, a.im + b.im}; }
static inline cplx csub(cplx a, cplx b) { return (cplx){a.re - b.re, a.im - b.im}; }
static inline double cabs_(cplx a) { return hypot(a.re, a.im); }
static cplx cdiv(cplx a, cplx b) {
  double d = b.re * b.re + b.im * b.im;
  return (cplx){(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}
static inline double clipd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
/* np.maximum(0, x): NaN propagates */
static inline double relu_nan(double x) { return isnan(x) ? x : (x > 0.0 ? x : 0.0); }
static inline double sign_nan(double x) { return isnan(x) ? x : (x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : 0.0)); }

/* ---------------------------------------------------------------------------------------
 * Exact Euclidean projection of (p,q) on {x : a_i p + b_i q <= h_i}; the stand-in for the
 * CVXPY QP of components/devices.py:299-304 and :517-522 (see oracle/shims/cvxpy).
 * ------------------------------------------------------------------------------------- */
static int poly_feasible(int n, const double (*row)[3], double x, double y, int s1, int s2) {
  for (int k = 0; k < n; ++k) {
    if (k == s1 || k == s2 || !isfinite(row[k][2])) continue;
    if (row[k][0] * x + row[k][1] * y - row[k][2] > FEAS_TOL) return 0;
  }
  return 1;
}

static void project_polygon(int n, const double (*row)[3], double p, double q, double* po, double* qo) {
  if (poly_feasible(n, row, p, q, -1, -1)) {
    *po = p;
    *qo = q;
    return;
  }
  double best = INFINITY, bx = NAN, by = NAN;
  for (int i = 0; i < n; ++i) {
    double a = row[i][0], b = row[i][1], h = row[i][2], x, y;
    if (!isfinite(h)) continue;
    if (b == 0.0) {
      x = h / a;
      y = q;
    } else if (a == 0.0) {
      x = p;
      y = h / b;
    } else {
      double t = (a * p + b * q - h) / (a * a + b * b);
      x = p - t * a;
      y = q - t * b;
    }
    if (poly_feasible(n, row, x, y, i, -1)) {
      double d = (x - p) * (x - p) + (y - q) * (y - q);
      if (d < best) best = d, bx = x, by = y;
    }
  }
  for (int i = 0; i < n; ++i) {
    if (!isfinite(row[i][2])) continue;
    for (int j = i + 1; j < n; ++j) {
      if (!isfinite(row[j][2])) continue;
      double det = row[i][0] * row[j][1] - row[j][0] * row[i][1];
      if (det == 0.0) continue;
      double x = (row[i][2] * row[j][1] - row[j][2] * row[i][1]) / det;
      double y = (row[i][0] * row[j][2] - row[j][0] * row[i][2]) / det;
      if (poly_feasible(n, row, x, y, i, j)) {
        double d = (x - p) * (x - p) + (y - q) * (y - q);
        if (d < best) best = d, bx = x, by = y;
      }
    }
  }
  *po = bx;
  *qo = by;
}

/* Generator.map_pq, devices.py:280-304 (7 rows) */
static void gen_map_pq(const double* P, double p_pot, double p, double q, double* po, double* qo) {
  double row[7][3] = {
      {-1, 0, -P[ANM_DP_PMIN]},       {1, 0, P[ANM_DP_PMAX]},         {1, 0, p_pot},
      {0, -1, -P[ANM_DP_QMIN]},       {0, 1, P[ANM_DP_QMAX]},         {-P[ANM_DP_TAU1], 1, P[ANM_DP_RHO1]},
      {P[ANM_DP_TAU2], -1, -P[ANM_DP_RHO2]},
  };
  project_polygon(7, row, p, q, po, qo);
}

/* StorageUnit.map_pq, devices.py:472-522 (10 rows) */
static void des_map_pq(const double* P, double soc, double dt, double p, double q, double* po, double* qo) {
  double eff = P[ANM_DP_EFF];
  double row[10][3] = {
      {-1, 0, -P[ANM_DP_PMIN]},
      {1, 0, P[ANM_DP_PMAX]},
      {0, -1, -P[ANM_DP_QMIN]},
      {0, 1, P[ANM_DP_QMAX]},
      {-P[ANM_DP_TAU1], 1, P[ANM_DP_RHO1]},
      {P[ANM_DP_TAU2], -1, -P[ANM_DP_RHO2]},
      {P[ANM_DP_TAU3], -1, -P[ANM_DP_RHO3]},
      {-P[ANM_DP_TAU4], 1, P[ANM_DP_RHO4]},
      {-1, 0, -(soc - P[ANM_DP_SOCMAX]) / (dt * eff)},
      {1, 0, eff * (soc - P[ANM_DP_SOCMIN]) / dt},
  };
  project_polygon(10, row, p, q, po, qo);
}

/* ---------------------------------------------------------------------------------------
 * Dense LU with partial pivoting, in place, solves A x = b (n <= MAXUNK).  Stands in for
 * scipy.sparse.linalg.spsolve (SuperLU) at solve_load_flow.py:220.
 * ------------------------------------------------------------------------------------- */
static void lu_solve(int n, double* A /* n x n row-major */, double* b) {
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = fabs(A[k * n + k]);
    for (int r = k + 1; r < n; ++r) {
      double v = fabs(A[r * n + k]);
      if (v > best) best = v, piv = r;
    }
    if (piv != k) {
      for (int c = 0; c < n; ++c) {
        double t = A[k * n + c];
        A[k * n + c] = A[piv * n + c];
        A[piv * n + c] = t;
      }
      double t = b[k];
      b[k] = b[piv];
      b[piv] = t;
    }
    double d = A[k * n + k];
    for (int r = k + 1; r < n; ++r) {
      double f = A[r * n + k] / d;
      for (int c = k + 1; c < n; ++c) A[r * n + c] -= f * A[k * n + c];
      b[r] -= f * b[k];
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    double s = b[k];
    for (int c = k + 1; c < n; ++c) s -= A[k * n + c] * b[c];
    b[k] = s / A[k * n + k];
  }
}

typedef struct {
  int N, D, L;
  double dev_p[MAXDEV], dev_q[MAXDEV], p_pot[MAXDEV];
  double bus_p[MAXBUS], bus_q[MAXBUS];
  cplx V[MAXBUS], I[MAXBUS];
  cplx i_from[MAXBR], i_to[MAXBR];
  double br_p[MAXBR], br_q[MAXBR], br_s[MAXBR];
  int n_iter;
} work_t;

static inline cplx Yij(const anm_network_desc* net, int i, int j) {
  const double* y = net->ybus + 2 * ((size_t)i * net->n_bus + j);
  return (cplx){y[0], y[1]};
}

/* _construct_v_from_guess, solve_load_flow.py:167-173 */
static void v_from_x(int N, const double* x, cplx* V) {
  int n = N - 1;
  V[0] = (cplx){1.0, 0.0};
  for (int j = 0; j < n; ++j) V[j + 1] = (cplx){x[n + j] * cos(x[j]), x[n + j] * sin(x[j])};
}

static void y_times_v(const anm_network_desc* net, const cplx* V, cplx* I) {
  int N = net->n_bus;
  for (int i = 0; i < N; ++i) {
    cplx s = {0, 0};
    for (int j = 0; j < N; ++j) {
      cplx y = Yij(net, i, j);
      if (y.re == 0.0 && y.im == 0.0) continue; /* sparse Y: structural zeros are not stored */
      s = cadd(s, cmul(y, V[j]));
    }
    I[i] = s;
  }
}

/* _f, solve_load_flow.py:84-120; returns the inf-norm (NaN if any entry is NaN, like
 * numpy.linalg.norm(F, inf)). */
static double nr_residual(const anm_network_desc* net, const double* x, const double* p, const double* q, double* F) {
  int N = net->n_bus, n = N - 1;
  cplx V[MAXBUS], I[MAXBUS];
  v_from_x(N, x, V);
  y_times_v(net, V, I);
  double nrm = 0.0;
  int has_nan = 0;
  for (int b = 1; b < N; ++b) {
    cplx s = cmul(V[b], cconj(I[b]));
    F[b - 1] = s.re - p[b];
    F[n + b - 1] = s.im - q[b];
  }
  for (int k = 0; k < 2 * n; ++k) {
    if (isnan(F[k])) has_nan = 1;
    if (fabs(F[k]) > nrm) nrm = fabs(F[k]);
  }
  return has_nan ? NAN : nrm;
}

/* _dfdx, solve_load_flow.py:123-164 (dense here) */
static void nr_jacobian(const anm_network_desc* net, const double* x, double* J) {
  int N = net->n_bus, n = N - 1, M = 2 * n;
  cplx V[MAXBUS], I[MAXBUS], E[MAXBUS];
  v_from_x(N, x, V);
  y_times_v(net, V, I);
  for (int j = 0; j < N; ++j) {
    double a = cabs_(V[j]);
    E[j] = (cplx){V[j].re / a, V[j].im / a};
  }
  memset(J, 0, sizeof(double) * M * M);
  for (int b = 1; b < N; ++b)
    for (int j = 1; j < N; ++j) {
      cplx y = Yij(net, b, j);
      int diag = (b == j);
      if (!diag && y.re == 0.0 && y.im == 0.0) continue;
      /* dS/dtheta = j V_b conj(delta I_b - Y_bj V_j) */
      cplx t = cmul(y, V[j]);
      cplx inner = diag ? csub(I[b], t) : (cplx){-t.re, -t.im};
      cplx w = cmul((cplx){-V[b].im, V[b].re}, cconj(inner));
      /* dS/d|V| = delta E_b conj(I_b) + V_b conj(Y_bj E_j) */
      cplx u = cmul(V[b], cconj(cmul(y, E[j])));
      if (diag) u = cadd(cmul(E[b], cconj(I[b])), u);
      J[(b - 1) * M + (j - 1)] = w.re;
      J[(b - 1) * M + (n + j - 1)] = u.re;
      J[(n + b - 1) * M + (j - 1)] = w.im;
      J[(n + b - 1) * M + (n + j - 1)] = u.im;
    }
}

/* solve_pfe_newton_raphson + _newton_raphson_sparse, solve_load_flow.py:7-81, 176-226;
 * branch flows: components/branch.py:153-198.  Returns `stable`. */
static int power_flow(const anm_network_desc* net, work_t* w) {
  int N = net->n_bus, n = N - 1, M = 2 * n;
  double x[MAXUNK], F[MAXUNK];
  static _Thread_local double J[MAXUNK * MAXUNK];
  for (int j = 0; j < n; ++j) x[j] = 0.0, x[n + j] = 1.0; /* flat start, :42 */
  int n_iter = 0;
  double diff = nr_residual(net, x, w->bus_p, w->bus_q, F);
  while (diff > NR_TOL && n_iter < NR_MAXIT) {
    ++n_iter;
    nr_jacobian(net, x, J);
    lu_solve(M, J, F);
    for (int k = 0; k < M; ++k) x[k] -= F[k];
    diff = nr_residual(net, x, w->bus_p, w->bus_q, F);
  }
  w->n_iter = n_iter;
  int converged = !isnan(diff);
  int stable = converged && diff <= NR_TOL; /* :49 */
  v_from_x(N, x, w->V);
  y_times_v(net, w->V, w->I); /* :55 */
  cplx s = cmul(w->V[0], cconj(w->I[0]));
  w->bus_p[0] = isnan(s.re) ? INFINITY : s.re; /* :63-66 */
  w->bus_q[0] = isnan(s.im) ? INFINITY : s.im;
  for (int d = 0; d < net->n_dev; ++d)
    if (net->dev_type[d] == ANM_DEV_SLACK) w->dev_p[d] = w->bus_p[0], w->dev_q[d] = w->bus_q[0]; /* :69-72 */
  for (int l = 0; l < net->n_branch; ++l) {
    const double* B = net->br_param + (size_t)l * ANM_BR_NPARAM;
    cplx ys = {B[ANM_BP_SERIES_RE], B[ANM_BP_SERIES_IM]}, ysh = {B[ANM_BP_SHUNT_RE], B[ANM_BP_SHUNT_IM]};
    cplx tap = {B[ANM_BP_TAP_RE], B[ANM_BP_TAP_IM]};
    cplx vf = w->V[net->br_from[l]], vt = w->V[net->br_to[l]];
    double t2 = cabs_(tap);
    t2 = t2 * t2;
    cplx a = cmul(cadd(ys, ysh), vf);
    a = (cplx){a.re / t2, a.im / t2};
    cplx b = cdiv(cmul((cplx){-ys.re, -ys.im}, vt), cconj(tap));
    w->i_from[l] = cadd(a, b);
    a = cmul(cadd(ys, ysh), vt);
    b = cdiv(cmul((cplx){-ys.re, -ys.im}, vf), tap);
    w->i_to[l] = cadd(a, b);
    cplx sf = cmul(vf, cconj(w->i_from[l])), st = cmul(vt, cconj(w->i_to[l]));
    w->br_p[l] = sf.re;
    w->br_q[l] = sf.im;
    double af = cabs_(sf), at = cabs_(st);
    double mx = (isnan(af) || isnan(at)) ? NAN : (af > at ? af : at);
    w->br_s[l] = sign_nan(sf.re) * mx; /* branch.py:198 */
  }
  return stable;
}

/* Simulator.transition, simulator.py:464-537 (+ _compute_reward :638-683).
 * soc: [n_des] in/out (p.u.).  Set-points ordered generators first then storage units. */
static int transition_one(const anm_network_desc* net, double* soc, const double* p_load, const double* p_potential,
                          const double* p_set, const double* q_set, work_t* w, double* reward, double* e_loss,
                          double* penalty) {
  int N = net->n_bus, D = net->n_dev;
  double m = net->base_mva, dt = net->delta_t;
  int n_gen = 0;
  for (int d = 0; d < D; ++d) n_gen += (net->dev_type[d] == ANM_DEV_GEN || net->dev_type[d] == ANM_DEV_RENEWABLE);
  int il = 0, ig = 0, is = 0;
  for (int d = 0; d < D; ++d) {
    const double* P = net->dev_param + (size_t)d * ANM_DEV_NPARAM;
    int t = net->dev_type[d];
    w->p_pot[d] = 0.0;
    if (t == ANM_DEV_LOAD) { /* Load.map_pq, devices.py:156-167 */
      w->dev_p[d] = clipd(p_load[il++] / m, P[ANM_DP_PMIN], P[ANM_DP_PMAX]);
      w->dev_q[d] = w->dev_p[d] * P[ANM_DP_QP_RATIO];
    } else if (t == ANM_DEV_GEN || t == ANM_DEV_RENEWABLE) {
      w->p_pot[d] = clipd(p_potential[ig] / m, P[ANM_DP_PMIN], P[ANM_DP_PMAX]); /* simulator.py:511 */
      gen_map_pq(P, w->p_pot[d], p_set[ig] / m, q_set[ig] / m, &w->dev_p[d], &w->dev_q[d]);
      ++ig;
    } else if (t == ANM_DEV_STORAGE) {
      des_map_pq(P, soc[is], dt, p_set[n_gen + is] / m, q_set[n_gen + is] / m, &w->dev_p[d], &w->dev_q[d]);
      /* update_soc, devices.py:524-545 */
      if (w->dev_p[d] <= 0)
        soc[is] -= dt * P[ANM_DP_EFF] * w->dev_p[d];
      else
        soc[is] -= dt * w->dev_p[d] / P[ANM_DP_EFF];
      soc[is] = clipd(soc[is], P[ANM_DP_SOCMIN], P[ANM_DP_SOCMAX]);
      ++is;
    } else { /* slack: simulator.py:521-523 */
      w->dev_p[d] = w->dev_q[d] = 0.0;
    }
  }
  for (int b = 0; b < N; ++b) w->bus_p[b] = w->bus_q[b] = 0.0; /* :539-549 */
  for (int d = 0; d < D; ++d) w->bus_p[net->dev_bus[d]] += w->dev_p[d], w->bus_q[net->dev_bus[d]] += w->dev_q[d];
  int stable = power_flow(net, w);
  /* _compute_reward */
  double e = 0.0;
  for (int d = 0; d < D; ++d) {
    int t = net->dev_type[d];
    if (t != ANM_DEV_STORAGE) e += w->dev_p[d];
    if (t == ANM_DEV_RENEWABLE) e += relu_nan(w->p_pot[d] - w->dev_p[d]);
  }
  e *= dt;
  double pen = 0.0;
  for (int b = 0; b < N; ++b) {
    double vm = cabs_(w->V[b]);
    pen += relu_nan(vm - net->bus_vmax[b]) + relu_nan(net->bus_vmin[b] - vm);
  }
  for (int l = 0; l < net->n_branch; ++l)
    pen += relu_nan(fabs(w->br_s[l]) - net->br_param[(size_t)l * ANM_BR_NPARAM + ANM_BP_RATE]);
  pen *= dt * net->lamb;
  *reward = -(e + pen);
  *e_loss = e;
  *penalty = pen;
  return stable;
}

/* Simulator._gather_state, simulator.py:551-636, p.u. / rad, order of anm_b200.h */
static void gather_full_state(const anm_network_desc* net, const work_t* w, const double* soc, const double* aux, int K,
                              double* out) {
  int N = net->n_bus, D = net->n_dev, L = net->n_branch;
  double* o = out;
  for (int b = 0; b < N; ++b) o[b] = w->bus_p[b];
  o += N;
  for (int b = 0; b < N; ++b) o[b] = w->bus_q[b];
  o += N;
  for (int b = 0; b < N; ++b) o[b] = cabs_(w->V[b]);
  o += N;
  for (int b = 0; b < N; ++b) o[b] = atan2(w->V[b].im, w->V[b].re);
  o += N;
  for (int b = 0; b < N; ++b) o[b] = cabs_(w->I[b]);
  o += N;
  for (int b = 0; b < N; ++b) o[b] = atan2(w->I[b].im, w->I[b].re);
  o += N;
  for (int d = 0; d < D; ++d) o[d] = w->dev_p[d];
  o += D;
  for (int d = 0; d < D; ++d) o[d] = w->dev_q[d];
  o += D;
  int is = 0;
  for (int d = 0; d < D; ++d) o[d] = (net->dev_type[d] == ANM_DEV_STORAGE) ? soc[is++] : 0.0;
  o += D;
  for (int d = 0; d < D; ++d)
    o[d] = (net->dev_type[d] == ANM_DEV_GEN || net->dev_type[d] == ANM_DEV_RENEWABLE) ? w->p_pot[d] : 0.0;
  o += D;
  for (int l = 0; l < L; ++l) o[l] = w->br_p[l];
  o += L;
  for (int l = 0; l < L; ++l) o[l] = w->br_q[l];
  o += L;
  for (int l = 0; l < L; ++l) o[l] = w->br_s[l];
  o += L;
  for (int l = 0; l < L; ++l) { /* simulator.py:613 with NumPy>=2 sign(z)=z/|z| */
    double a = cabs_(w->i_from[l]);
    o[l] = (a == 0.0) ? 0.0 : (w->i_from[l].re / a) * a;
  }
  o += L;
  for (int l = 0; l < L; ++l) o[l] = atan2(w->i_from[l].im, w->i_from[l].re);
  o += L;
  for (int k = 0; k < K; ++k) o[k] = aux[k];
}

static int full_offset(const anm_network_desc* net, int quantity) {
  int N = net->n_bus, D = net->n_dev, L = net->n_branch;
  if (quantity <= ANM_Q_BUS_I_ANG) return quantity * N;
  if (quantity <= ANM_Q_GEN_P_MAX) return 6 * N + (quantity - ANM_Q_DEV_P) * D;
  if (quantity <= ANM_Q_BRANCH_I_ANG) return 6 * N + 4 * D + (quantity - ANM_Q_BRANCH_P) * L;
  return 6 * N + 4 * D + 5 * L;
}

/* ANMEnv._extract_state_variables / observation, anm_env.py:562-592, 313-331 */
static void extract_vars(const anm_network_desc* net, const double* full, int n, const anm_var_spec* vars, int clip,
                         double* out) {
  for (int k = 0; k < n; ++k) {
    double v = full[full_offset(net, vars[k].quantity) + vars[k].index] * vars[k].mul;
    if (vars[k].div != 1.0) v /= vars[k].div;
    if (clip) v = clipd(v, vars[k].low, vars[k].high);
    out[k] = v;
  }
}

static void count_devices(const anm_network_desc* net, int* n_load, int* n_gen, int* n_des) {
  *n_load = *n_gen = *n_des = 0;
  for (int d = 0; d < net->n_dev; ++d) {
    int t = net->dev_type[d];
    *n_load += (t == ANM_DEV_LOAD);
    *n_gen += (t == ANM_DEV_GEN || t == ANM_DEV_RENEWABLE);
    *n_des += (t == ANM_DEV_STORAGE);
  }
}

static int check_limits(const anm_network_desc* net) {
  return net->n_bus >= 2 && net->n_bus <= MAXBUS && net->n_dev <= MAXDEV && net->n_branch <= MAXBR;
}

/* ======================================================================================
 * exported entry points (batch loops)
 * ==================================================================================== */

int anm_oracle_project(int n_rows, const double* rows /* [n][3] */, double p, double q, double* out2) {
  project_polygon(n_rows, (const double (*)[3])rows, p, q, &out2[0], &out2[1]);
  return 0;
}

/* Simulator.transition over a batch.  soc [B, n_des] in/out. */
int anm_oracle_transition(const anm_network_desc* net, int64_t B, double* soc, const double* p_load,
                          const double* p_pot, const double* p_set, const double* q_set, double* full_state,
                          double* reward, double* e_loss, double* penalty, uint8_t* converged, int32_t* n_iter) {
  if (!check_limits(net)) return ANM_E_UNSUPPORTED;
  int nl, ng, ns;
  count_devices(net, &nl, &ng, &ns);
  int F = 6 * net->n_bus + 4 * net->n_dev + 5 * net->n_branch;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < B; ++e) {
    work_t w;
    double r, el, pe;
    int ok = transition_one(net, soc + e * ns, p_load + e * nl, p_pot + e * ng, p_set + e * (ng + ns),
                            q_set + e * (ng + ns), &w, &r, &el, &pe);
    if (full_state) gather_full_state(net, &w, soc + e * ns, NULL, 0, full_state + e * F);
    if (reward) reward[e] = r;
    if (e_loss) e_loss[e] = el;
    if (penalty) penalty[e] = pe;
    if (converged) converged[e] = (uint8_t)ok;
    if (n_iter) n_iter[e] = w.n_iter;
  }
  return 0;
}

/* ANMEnv.reset inner loop body (anm_env.py:271-295) + Simulator.reset (simulator.py:225-293). */
int anm_oracle_reset(const anm_network_desc* net, const anm_env_desc* env, int64_t B, const double* s0,
                     const uint8_t* mask, double* soc, double* aux, uint8_t* terminated, double* obs, double* state,
                     uint8_t* converged) {
  if (!check_limits(net)) return ANM_E_UNSUPPORTED;
  int nl, ng, ns, D = net->n_dev, K = env->K;
  count_devices(net, &nl, &ng, &ns);
  int S = 2 * D + ns + ng + K;
  if (S != env->n_state) return ANM_E_INVALID;
  int F = 6 * net->n_bus + 4 * D + 5 * net->n_branch + K;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < B; ++e) {
    if (mask && !mask[e]) continue;
    const double* s = s0 + e * S;
    double p_load[MAXDEV], p_pot[MAXDEV], p_set[MAXDEV], q_set[MAXDEV], full[6 * MAXBUS + 4 * MAXDEV + 5 * MAXBR + 16];
    int il = 0, ig = 0, is = 0;
    for (int d = 0; d < D; ++d) {
      int t = net->dev_type[d];
      const double* P = net->dev_param + (size_t)d * ANM_DEV_NPARAM;
      if (t == ANM_DEV_LOAD)
        p_load[il++] = s[d];
      else if (t == ANM_DEV_GEN || t == ANM_DEV_RENEWABLE) {
        p_set[ig] = s[d], q_set[ig] = s[D + d];
        p_pot[ig] = s[2 * D + ns + ig];
        ++ig;
      } else if (t == ANM_DEV_STORAGE) {
        p_set[ng + is] = s[d], q_set[ng + is] = s[D + d];
        soc[e * ns + is] = (s[d] <= 0) ? P[ANM_DP_SOCMIN] : P[ANM_DP_SOCMAX]; /* simulator.py:273-278 */
        ++is;
      }
    }
    work_t w;
    double r, el, pe;
    int ok = transition_one(net, soc + e * ns, p_load, p_pot, p_set, q_set, &w, &r, &el, &pe);
    for (int k = 0; k < ns; ++k) soc[e * ns + k] = s[2 * D + k] / net->base_mva; /* :284-288 */
    for (int k = 0; k < K; ++k) aux[e * K + k] = s[S - K + k];
    gather_full_state(net, &w, soc + e * ns, aux + e * K, K, full);
    (void)F;
    if (state) extract_vars(net, full, env->n_state, env->state_vars, 0, state + e * env->n_state);
    extract_vars(net, full, env->n_obs, env->obs_vars, 1, obs + e * env->n_obs);
    terminated[e] = (uint8_t)!ok;
    converged[e] = (uint8_t)ok;
  }
  return 0;
}

/* ANMEnv.step, anm_env.py:333-453 (+ ANM6Easy.next_vars, anm6_easy.py:54-65 when the table is used). */
int anm_oracle_step(const anm_network_desc* net, const anm_env_desc* env, int64_t B, double* soc, double* aux,
                    uint8_t* terminated, const double* action, const double* next_vars, double* obs, double* reward,
                    uint8_t* terminated_out, double* state, double* e_loss, double* penalty, int32_t* n_iter,
                    double* full_state) {
  if (!check_limits(net)) return ANM_E_UNSUPPORTED;
  int nl, ng, ns, D = net->n_dev, K = env->K;
  count_devices(net, &nl, &ng, &ns);
  int A = 2 * ng + 2 * ns, NV = nl + ng + K;
  int F = 6 * net->n_bus + 4 * D + 5 * net->n_branch + K;
  if (!next_vars && (env->table_len <= 0 || K < 1)) return ANM_E_INVALID;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < B; ++e) {
    double* ob = obs + e * env->n_obs;
    if (terminated[e]) { /* anm_env.py:365-367 */
      for (int k = 0; k < env->n_obs; ++k) ob[k] = 0.0;
      reward[e] = 0.0;
      terminated_out[e] = 1;
      if (state) memset(state + e * env->n_state, 0, sizeof(double) * env->n_state);
      if (e_loss) e_loss[e] = env->clip_e_loss;
      if (penalty) penalty[e] = env->clip_penalty;
      if (n_iter) n_iter[e] = 0;
      if (full_state) memset(full_state + e * F, 0, sizeof(double) * F);
      continue;
    }
    double vars[MAXDEV + 16], new_aux[16], p_set[MAXDEV], q_set[MAXDEV];
    double full[6 * MAXBUS + 4 * MAXDEV + 5 * MAXBR + 16];
    if (next_vars) {
      memcpy(vars, next_vars + e * NV, sizeof(double) * NV);
    } else {
      int a = (int)fmod(aux[e * K + K - 1] + 1.0, (double)env->table_len);
      memcpy(vars, env->table + (size_t)a * (nl + ng), sizeof(double) * (nl + ng));
      for (int k = 0; k < K; ++k) vars[nl + ng + k] = (k == K - 1) ? (double)a : aux[e * K + k];
    }
    for (int k = 0; k < K; ++k) new_aux[k] = vars[nl + ng + k];
    const double* a = action + e * A; /* anm_env.py:394-410 */
    for (int g = 0; g < ng; ++g) p_set[g] = a[g], q_set[g] = a[ng + g];
    for (int s = 0; s < ns; ++s) p_set[ng + s] = a[2 * ng + s], q_set[ng + s] = a[2 * ng + ns + s];
    work_t w;
    double r, el, pe;
    int ok = transition_one(net, soc + e * ns, vars, vars + nl, p_set, q_set, &w, &r, &el, &pe);
    int term = !ok;
    if (!term) { /* anm_env.py:424-427 */
      el = sign_nan(el) * clipd(fabs(el), 0.0, env->clip_e_loss);
      pe = clipd(pe, 0.0, env->clip_penalty);
      r = -(el + pe);
      for (int k = 0; k < K; ++k) aux[e * K + k] = new_aux[k];
      gather_full_state(net, &w, soc + e * ns, aux + e * K, K, full);
      if (state) extract_vars(net, full, env->n_state, env->state_vars, 0, state + e * env->n_state);
      extract_vars(net, full, env->n_obs, env->obs_vars, 1, ob);
      if (full_state) memcpy(full_state + e * F, full, sizeof(double) * F);
    } else { /* :428-432, 446-448 */
      r = -env->clip_penalty / (1.0 - env->gamma);
      el = env->clip_e_loss;
      pe = env->clip_penalty;
      for (int k = 0; k < env->n_obs; ++k) ob[k] = 0.0;
      if (state) memset(state + e * env->n_state, 0, sizeof(double) * env->n_state);
      if (full_state) memset(full_state + e * F, 0, sizeof(double) * F);
    }
    terminated[e] = (uint8_t)term;
    terminated_out[e] = (uint8_t)term;
    reward[e] = r;
    if (e_loss) e_loss[e] = el;
    if (penalty) penalty[e] = pe;
    if (n_iter) n_iter[e] = w.n_iter;
  }
  return 0;
}
