/*
 * anm_layout.h -- layout of the per-network constant blob and of the per-environment
 * shared-memory workspace.  Shared by the host builder (anm_capi.cu) and the kernels
 * (anm_kernels.cuh).
 *
 * The blob lives once in global memory; every CTA stages it into shared memory with one
 * TMA bulk copy (cp.async.bulk.shared::cluster.global) at kernel entry, so all per-step
 * constant reads (Y-bus, polygon rows, 96-slot tables, observation map) are shared-memory
 * broadcasts.  All offsets are in BYTES from the start of the blob and 16-byte aligned.
 */
#pragma once
#include <stdint.h>

#define ANM_MAX_ROWS 10  /* half-planes per device polygon (devices.py:486-514) */
#define ANM_NPAIRS 45    /* ANM_MAX_ROWS choose 2 */
#define ANM_NCAND (1 + ANM_MAX_ROWS + ANM_NPAIRS)

enum { ANM_MODE_STEP = 0, ANM_MODE_RESET = 1, ANM_MODE_TRANSITION = 2 };

struct AnmConstHeader {
  /* sizes */
  int32_t n_bus, n_dev, n_branch, n_load, n_gen, n_des, n_ctrl, K;
  int32_t n_unk;       /* 2 (n_bus - 1) Newton-Raphson unknowns */
  int32_t n_action, n_state, n_obs, n_next_vars, n_full;
  int32_t table_len, y_nnz, n_jac; /* n_jac: Y entries with row, col >= 1 */
  int32_t need_angles;  /* some state/obs entry is an angle */
  int32_t need_mask;    /* ANM_NEED_* (anm_kernels.cuh): derived quantities some state/obs entry reads */
  int32_t blob_bytes;
  int32_t ws_doubles;   /* per-env workspace size (doubles) */
  int32_t nr_maxit;     /* Newton-Raphson iteration cap: 100 (solve_load_flow.py:176); ANM_DEBUG_NR_MAXIT (environment,
                           read by anm_create) lowers it for single-iteration tests */
  int32_t pad0;
  double base_mva, delta_t, lamb, gamma, clip_e, clip_pen, term_reward;
  /* blob offsets (bytes) */
  int32_t o_vmin, o_vmax;                    /* double[n_bus]                                  */
  int32_t o_dev_bus, o_dev_type, o_dev_slot; /* int[n_dev]  (slot = index inside its category) */
  int32_t o_dev_param;                       /* double[n_dev][16]                              */
  int32_t o_bus_dev_ptr, o_bus_dev_idx;      /* int[n_bus+1], int[n_dev]  devices of each bus  */
  int32_t o_br_from, o_br_to;                /* int[n_branch]                                  */
  int32_t o_br_coef;                         /* double[n_branch][10]: a_ff a_ft a_tf a_tt (re,im) rate pad */
  int32_t o_y_ptr, o_y_col, o_y_val;         /* CSR of Y: int[n_bus+1], int[nnz], double[nnz][2] */
  int32_t o_y_dense;                         /* double[n_bus][n_bus][2] dense Y (register-resident NR) */
  int32_t o_jac_row, o_jac_col, o_jac_y;     /* int[n_jac] each: bus b>=1, bus j>=1, index into y_val */
  int32_t o_ctrl_dev;                        /* int[n_ctrl] device position (gens then storage) */
  int32_t o_ctrl_rows;                       /* double[n_ctrl][3][ANM_MAX_ROWS]: a[], b[], h[]  */
  int32_t o_ctrl_fin;                        /* int[n_ctrl]: bit k = static row k has a finite h (dynamic rows: 0) */
  int32_t o_sv_off, o_sv_mul, o_sv_div;      /* state vars: int[], double[], double[]          */
  int32_t o_ov_off, o_ov_mul, o_ov_div, o_ov_low, o_ov_high; /* obs vars                       */
  int32_t o_table;                           /* double[table_len][n_load+n_gen]                */
  int32_t o_pair_i, o_pair_j;                /* int[ANM_NPAIRS]                                */
  /* projection candidates of every controllable device (anm_capi.cu: candidate_table) */
  int32_t o_cand_ptr;                        /* int[n_ctrl+1]: first candidate of each device               */
  int32_t o_cand_info;                       /* int[ncand]: s1 | s2 << 8 | need << 16                       */
  int32_t o_cand_coef;                       /* double[ncand][8]: kx[4], ky[4] on (p, q, h[s1], h[s2])      */
  int32_t o_cand_short;                      /* int[n_ctrl][2]: (n_short, bit): only the first n_short candidates
                                                count when the finite-row mask has `bit` (anm_capi.cu)       */
  /* radial networks (bus graph = tree rooted at the slack): per non-slack bus b (lane b-1) */
  int32_t is_radial, rad_maxc, rad_maxdepth;
  int32_t o_rad_parent, o_rad_depth;         /* int[n_bus-1]: parent's lane (-1 = slack), depth >= 1 */
  int32_t o_rad_child;                       /* int[n_bus-1][4]: children's lanes, -1 padded         */
  int32_t o_rad_y;                           /* double[n_bus-1][6]: Y_bb, Y_b,parent, Y_parent,b     */
  /* block-sparse LU of the Newton system (any network; default above 9 buses): 2x2 blocks on the filled
   * Y-bus pattern, elimination order and fill computed once on the host (anm_capi.cu: sparse_symbolic) */
  int32_t solver;                            /* 0 dense smem, 1 dense register rows, 2 radial tree, 4 block-sparse */
  int32_t sp_nblk, sp_nsteps;
  int32_t o_sp_blk_i, o_sp_blk_j, o_sp_blk_y; /* int[sp_nblk]: bus i, bus j, index into y_val (-1: fill)       */
  int32_t o_sp_step;                         /* int[sp_nsteps][5]: pivot bus, pivot block, degree, list offset, target offset */
  int32_t o_sp_row, o_sp_col, o_sp_nbr;      /* int lists: blocks (b,j), blocks (i,b), neighbour buses          */
  int32_t o_sp_tgt;                          /* int list: target blocks (i,j), degree^2 per step                */
  int32_t o_sp_diag;                         /* int[n_bus]: diagonal block of each bus (-1 for the slack)       */
  /* per-env workspace offsets (in doubles) */
  int32_t w_in_pl, w_in_pp, w_in_ps, w_in_qs, w_soc, w_aux, w_devp, w_devq, w_ppot, w_busp, w_busq;
  int32_t w_x, w_vre, w_vim, w_ere, w_eim, w_ire, w_iim, w_J, w_rowh;
  int32_t w_brp, w_brq, w_brs, w_brire, w_briim, w_full, w_s0;
  int32_t w_vx; /* (Vre, Vim, Ere, Eim) x n_bus exchange buffer of the register-resident solver */
  int32_t w_dx; /* Newton step written by the shared-memory solver */
  int32_t w_blk; /* block-sparse solver: 4 doubles per block */
};
