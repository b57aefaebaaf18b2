/*
 * anm_lp.h -- C ABI of the batched linear-program solver behind the MPC action source
 * (BASELINE config 5: "16384 envs, MPC-constant agent driving actions").
 *
 * Reference interface replaced: MPCAgent._create_optimization_problem (gym_anm/agents/mpc.py:161-319: ONE
 * parametrised convex program per agent, built once) and MPCAgent._solve / _update_parameters
 * (mpc.py:372-393, 395-420: new parameter values, `self.dc_opf.solve()`, read `P_dev.value`).  The
 * reference hands the program to CVXPY; the program is linear, and a batch of B environment instances
 * gives B programs that share the constraint matrix and the cost vector and differ only in their bounds
 * (load / generation forecasts, initial state of charge).  That is what this library solves:
 *
 *     minimise  c . x     subject to   lo[j] <= x[j] <= up[j]            (j <  n: columns)
 *                                       lo[n+i] <= (A x)[i] <= up[n+i]    (i <  m: rows)
 *
 * for every instance of the batch, A [m, n] and c [n] common to the batch, lo / up per instance.
 * Method: bounded dual simplex on the condensed tableau (one warp per program -- or one thread, see
 * ANM_LP_KERNEL_* -- the tableau of every instance resident in HBM between calls: consecutive MPC steps change
 * bounds only, so the previous optimal basis stays dual feasible and a step costs a handful of pivots).  Every
 * column needs a finite bound on the side its cost pulls towards (c[j] > 0: lo[j], c[j] < 0: up[j]); the
 * caller adds box bounds where the program has none.
 *
 * Batch layout: "interleaved" -- element k of instance e lives at [k * stride + e] (stride >= B, the value
 * passed to anm_lp_create), so that neighbouring threads touch neighbouring addresses.  All pointers of the
 * device entry points are CUDA device pointers on the handle's device; `stream` is a cudaStream_t passed as
 * void*.  Calls are asynchronous and never synchronise.  Return value 0 or a negative ANM_LP_E_* code,
 * message from anm_lp_last_error() (thread-local).
 */
#ifndef ANM_LP_H
#define ANM_LP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ANM_LP_OK 0
#define ANM_LP_E_INVALID (-1)
#define ANM_LP_E_CUDA (-2)
#define ANM_LP_E_NOMEM (-3)

/* per-instance status written by a solve */
#define ANM_LP_OPTIMAL 0
#define ANM_LP_INFEASIBLE 1     /* a violated row / basic variable without an entering candidate       */
#define ANM_LP_ITER_LIMIT 2
#define ANM_LP_DUAL_INFEASIBLE 3 /* a column whose cost pulls towards an infinite bound                  */

/* the two device kernels (same algorithm, same pivoting rules): one thread or one warp per program.  The
 * environment variable ANM_LP_KERNEL = "thread" | "warp" read by anm_lp_create overrides the default. */
#define ANM_LP_KERNEL_THREAD 0
#define ANM_LP_KERNEL_WARP 1
#define ANM_LP_KERNEL_DEFAULT ANM_LP_KERNEL_WARP

typedef struct anm_lp_batch* anm_lp_handle;

/* A [m, n] row-major and c [n] are HOST pointers (copied).  `batch` programs, interleaved stride `stride`
 * (>= batch).  max_iter <= 0: 4 (m + n) + 50 pivots per solve. */
int anm_lp_create(int32_t n, int32_t m, const double* a_host, const double* c_host, int64_t batch, int64_t stride,
                  int32_t max_iter, int device, anm_lp_handle* out);
int anm_lp_destroy(anm_lp_handle h);

/* One solve of every program.  lo / up: [(n + m), stride] interleaved.  restart_or_null: [batch] u8, non-zero =
 * forget this instance's basis and start from the all-slack basis (NULL: every instance restarts on the handle's
 * first solve, none afterwards).  Outputs: x [n, stride] interleaved, obj [batch], status [batch], iters [batch]
 * (pivots of this call); any may be NULL except x. */
int anm_lp_solve(anm_lp_handle h, const double* lo_dev, const double* up_dev, const uint8_t* restart_dev_or_null,
                 double* x_dev, double* obj_dev_or_null, int32_t* status_dev_or_null, int32_t* iters_dev_or_null,
                 void* stream);

/* Bytes of device memory the handle holds (tableaux + bases + scratch); the kernel it launches (ANM_LP_KERNEL_*). */
int64_t anm_lp_bytes(anm_lp_handle h);
int anm_lp_kernel(anm_lp_handle h);

/* Host build of the very same solver code (test hook: the CPU suite checks the algorithm against HiGHS without a
 * GPU).  All pointers are host pointers, same layouts; `state_host` is an opaque buffer of
 * anm_debug_lp_state_bytes(n, m, stride) bytes that carries the tableaux between calls (zero-initialised by the
 * caller; restart semantics as above, with `first` non-zero on the first call). */
int64_t anm_debug_lp_state_bytes(int32_t n, int32_t m, int64_t stride);
int anm_debug_lp_solve_host(int32_t n, int32_t m, const double* a_host, const double* c_host, int64_t batch,
                            int64_t stride, int32_t max_iter, void* state_host, int32_t first, const double* lo,
                            const double* up, const uint8_t* restart_or_null, double* x, double* obj_or_null,
                            int32_t* status_or_null, int32_t* iters_or_null);

/* The warp kernel's code on the host: its lane phases run as loops over the 32 lanes, forwards or (reverse_lanes)
 * backwards -- a phase whose lanes depended on each other would differ between the two orders.  `state_host`:
 * anm_debug_lp_warp_state_bytes(n, m, batch) bytes (one contiguous block per program). */
int64_t anm_debug_lp_warp_state_bytes(int32_t n, int32_t m, int64_t batch);
int anm_debug_lp_solve_host_warp(int32_t n, int32_t m, const double* a_host, const double* c_host, int64_t batch,
                                 int64_t stride, int32_t max_iter, void* state_host, int32_t first, const double* lo,
                                 const double* up, const uint8_t* restart_or_null, double* x, double* obj_or_null,
                                 int32_t* status_or_null, int32_t* iters_or_null, int32_t reverse_lanes);

const char* anm_lp_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* ANM_LP_H */
