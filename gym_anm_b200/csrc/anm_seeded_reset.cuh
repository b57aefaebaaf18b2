/*
 * anm_seeded_reset.cuh -- the random part of ANMEnv.reset on the device.
 *
 * One thread per environment instance, its own PCG64 stream (anm_rng.h).  Kept out of the step kernel on purpose: the
 * hot kernel's registers and instruction footprint stay what they are; a reset is a handful of tiny launches.
 */
#pragma once
#include "anm_kernels.cuh"
#include "anm_rng.h"

namespace anm {

/* ANM6Easy.init_state (anm6_easy.py:25-52) for any network with a built-in next_vars table, the draws in the
 * reference's order: t0 = integers(0, table_len); loads and maximum generation from table[t0]; per generator
 * (ascending id) q = uniform(q_min, q_max) -- per-unit values written where the state vector holds MVAr, the
 * reference's quirk --; per storage unit soc = uniform(soc_min, soc_max) (per-unit, same quirk); aux = t0. */
__device__ inline void draw_init_state(const Cst& C, double* __restrict__ s0, AnmPcg64& rng) {
  const AnmConstHeader& H = *C.H;
  const int D = H.n_dev, nl = H.n_load, ng = H.n_gen, ns = H.n_des, S = H.n_state;
  for (int k = 0; k < S; ++k) s0[k] = 0.0;
  const int t0 = (int)anm_rng_integers(rng, 0, H.table_len);
  const double* trow = C.table + t0 * (nl + ng);
  s0[S - 1] = (double)t0;
  for (int d = 0; d < D; ++d) {
    const int t = C.dev_type[d], slot = C.dev_slot[d];
    const double* P = C.dev_param + d * ANM_DEV_NPARAM;
    if (t == ANM_DEV_LOAD) {
      s0[d] = trow[slot];
      s0[D + d] = trow[slot] * P[ANM_DP_QP_RATIO];
    } else if (t == ANM_DEV_GEN || t == ANM_DEV_RENEWABLE) {
      s0[d] = trow[nl + slot];
      s0[2 * D + ns + slot] = trow[nl + slot];
    }
  }
  for (int c = 0; c < ng; ++c) { /* generators in ascending id, then the storage units: the reference's draw order */
    const int d = C.ctrl_dev[c];
    const double* P = C.dev_param + d * ANM_DEV_NPARAM;
    s0[D + d] = anm_rng_uniform(rng, P[ANM_DP_QMIN], P[ANM_DP_QMAX]);
  }
  for (int c = 0; c < ns; ++c) {
    const double* P = C.dev_param + C.ctrl_dev[ng + c] * ANM_DEV_NPARAM;
    s0[2 * D + c] = anm_rng_uniform(rng, P[ANM_DP_SOCMIN], P[ANM_DP_SOCMAX]);
  }
}

/* PCG64(SeedSequence(seed_first + e)) for every instance, expanded on the device (anm_seed): no host loop over the
 * batch, no blocking copy */
__global__ void seed_streams_kernel(AnmPcg64* __restrict__ rng, uint64_t seed_first, int64_t B) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B) return;
  AnmPcg64 r;
  anm_pcg_seed(r, seed_first + (uint64_t)e);
  rng[e] = r;
}

/* One round of the reset loop (anm_env.py:266-289) for every instance that is still looking for an initial state:
 * if its previous attempt converged it is done (and draws ANM6.reset's date, anm6.py:138), otherwise it draws the next
 * initial state -- unless the attempts are used up (`last`). */
__global__ void seeded_draw_kernel(const unsigned char* __restrict__ blob, int64_t B, AnmPcg64* __restrict__ rng,
                                   uint8_t* __restrict__ need, const uint8_t* __restrict__ mask, bool first, bool last,
                                   const uint8_t* __restrict__ conv_prev, uint8_t* __restrict__ converged, int date_draw,
                                   double* __restrict__ s0_buf, int* __restrict__ n_drawn) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B) return;
  if (first) {
    const bool sel = !mask || mask[e];
    need[e] = sel ? 1 : 0;
    if (sel) converged[e] = 0;
  }
  if (!need[e]) return;
  const Cst C(blob);
  AnmPcg64 r = rng[e];
  if (!first && conv_prev[e]) {
    need[e] = 0;
    converged[e] = 1;
    if (date_draw) (void)anm_rng_integers(r, 1, 365); /* random_date, anm6_env/utils.py:22 */
    rng[e] = r;
    return;
  }
  if (last) return; /* no convergent initial state within the attempts allowed: converged[e] stays 0 */
  draw_init_state(C, s0_buf + e * C.H->n_state, r);
  rng[e] = r;
  atomicAdd(n_drawn, 1);
}

}  // namespace anm
