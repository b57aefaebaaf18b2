/*
 * anm_rng.h -- the random streams behind ANMEnv.reset(seed=...), restated for host and device.
 *
 * The reference seeds every environment with Gymnasium's `np.random.Generator(PCG64(SeedSequence(seed)))`
 * (gymnasium Env.reset -> seeding.np_random; gym_anm/envs/anm_env.py:116, 257) and draws, per
 * ANM6Easy.init_state (anm6_easy.py:25-52): integers(0, 96), uniform(q_min, q_max) per generator,
 * uniform(soc_min, soc_max) per storage unit, and after a successful reset one integers(1, 365)
 * (ANM6.reset -> random_date, anm6.py:138, anm6_env/utils.py:22).  This header follows NumPy's published algorithms
 * (the dependency itself is not part of the reference tree; numpy 2.1.3 in its lock file):
 *   SeedSequence      numpy/random/bit_generator.pyx  (hashmix / mix, pool of four 32-bit words)
 *   PCG64             numpy/random/src/pcg64/pcg64.h  (128-bit LCG, XSL-RR 64-bit output, buffered 32-bit halves)
 *   integers(lo, hi)  numpy/random/src/distributions/distributions.c  random_bounded_uint64_fill -> Lemire, 32-bit
 *   uniform(lo, hi)   lo + (hi - lo) * (next_uint64 >> 11) * 2^-53
 * Checked bit for bit against NumPy on the CPU (tests/test_host_logic.py::test_rng_streams_match_numpy).
 */
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ANM_HD __host__ __device__ __forceinline__
#else
#define ANM_HD inline
#endif

struct AnmPcg64 {          /* one stream: 48 bytes */
  uint64_t state_hi, state_lo, inc_hi, inc_lo;
  uint32_t has_u32, u32;   /* the unused upper half of the last 64-bit draw (pcg64_next32) */
  uint64_t pad;
};

typedef unsigned __int128 anm_u128;

ANM_HD void anm_pcg_step(AnmPcg64& r) {
  const anm_u128 mult = ((anm_u128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull; /* PCG_DEFAULT_MULTIPLIER_128 */
  anm_u128 s = ((anm_u128)r.state_hi << 64) | r.state_lo;
  const anm_u128 inc = ((anm_u128)r.inc_hi << 64) | r.inc_lo;
  s = s * mult + inc;
  r.state_hi = (uint64_t)(s >> 64);
  r.state_lo = (uint64_t)s;
}

ANM_HD uint64_t anm_pcg_next64(AnmPcg64& r) { /* step, then XSL-RR of the new state */
  anm_pcg_step(r);
  const uint64_t x = r.state_hi ^ r.state_lo;
  const unsigned rot = (unsigned)(r.state_hi >> 58);
  return (x >> rot) | (x << ((64u - rot) & 63u));
}

ANM_HD uint32_t anm_pcg_next32(AnmPcg64& r) {
  if (r.has_u32) {
    r.has_u32 = 0;
    return r.u32;
  }
  const uint64_t n = anm_pcg_next64(r);
  r.has_u32 = 1;
  r.u32 = (uint32_t)(n >> 32);
  return (uint32_t)n;
}

ANM_HD double anm_pcg_double(AnmPcg64& r) { return (double)(anm_pcg_next64(r) >> 11) * (1.0 / 9007199254740992.0); }

/* Generator.uniform(lo, hi) */
ANM_HD double anm_rng_uniform(AnmPcg64& r, double lo, double hi) {
  const double u = anm_pcg_double(r);
#if defined(__CUDA_ARCH__)
  return __dadd_rn(lo, __dmul_rn(hi - lo, u)); /* NumPy rounds the product and the sum separately: no FMA contraction */
#else
  return lo + (hi - lo) * u;
#endif
}

/* Generator.integers(lo, hi) for hi - lo <= 2^32 (dtype int64, endpoint=False): Lemire's method on 32-bit draws */
ANM_HD int64_t anm_rng_integers(AnmPcg64& r, int64_t lo, int64_t hi) {
  const uint64_t rng = (uint64_t)(hi - 1 - lo);
  if (rng == 0) return lo;
  if (rng == 0xFFFFFFFFull) return lo + (int64_t)anm_pcg_next32(r);
  const uint32_t rng_excl = (uint32_t)rng + 1u;
  uint64_t m = (uint64_t)anm_pcg_next32(r) * rng_excl;
  uint32_t leftover = (uint32_t)m;
  if (leftover < rng_excl) {
    const uint32_t threshold = (0xFFFFFFFFu - (uint32_t)rng) % rng_excl;
    while (leftover < threshold) {
      m = (uint64_t)anm_pcg_next32(r) * rng_excl;
      leftover = (uint32_t)m;
    }
  }
  return lo + (int64_t)(m >> 32);
}

/* PCG64(SeedSequence(seed)) for a non-negative integer seed < 2^64 */
ANM_HD uint32_t anm_ss_hashmix(uint32_t value, uint32_t& hash_const) {
  value ^= hash_const;
  hash_const *= 0x931e8875u; /* MULT_A */
  value *= hash_const;
  value ^= value >> 16;
  return value;
}
ANM_HD uint32_t anm_ss_mix(uint32_t x, uint32_t y) {
  uint32_t r = 0xca01f9ddu * x - 0x4973f715u * y; /* MIX_MULT_L, MIX_MULT_R */
  r ^= r >> 16;
  return r;
}
ANM_HD void anm_pcg_seed(AnmPcg64& r, uint64_t seed) {
  /* SeedSequence.__init__ / mix_entropy: entropy = little-endian 32-bit words of the seed, spawn_key = () */
  uint32_t ent[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  const int n_ent = (ent[1] != 0u) ? 2 : 1;
  uint32_t pool[4];
  uint32_t hc = 0x43b0d7e5u; /* INIT_A */
  for (int i = 0; i < 4; ++i) pool[i] = anm_ss_hashmix(i < n_ent ? ent[i] : 0u, hc);
  for (int s = 0; s < 4; ++s)
    for (int d = 0; d < 4; ++d)
      if (s != d) pool[d] = anm_ss_mix(pool[d], anm_ss_hashmix(pool[s], hc));
  /* generate_state(4, uint64) = eight 32-bit words */
  uint32_t w[8];
  uint32_t hb = 0x8b51f9ddu; /* INIT_B */
  for (int i = 0; i < 8; ++i) {
    uint32_t v = pool[i & 3];
    v ^= hb;
    hb *= 0x58f38dedu; /* MULT_B */
    v *= hb;
    v ^= v >> 16;
    w[i] = v;
  }
  const uint64_t s0 = (uint64_t)w[0] | ((uint64_t)w[1] << 32), s1 = (uint64_t)w[2] | ((uint64_t)w[3] << 32);
  const uint64_t i0 = (uint64_t)w[4] | ((uint64_t)w[5] << 32), i1 = (uint64_t)w[6] | ((uint64_t)w[7] << 32);
  /* pcg64_set_seed: initstate = (s0 << 64) | s1, initseq = (i0 << 64) | i1; pcg_setseq_128_srandom_r */
  const anm_u128 initstate = ((anm_u128)s0 << 64) | s1;
  const anm_u128 inc = ((((anm_u128)i0 << 64) | i1) << 1) | 1u;
  r.inc_hi = (uint64_t)(inc >> 64);
  r.inc_lo = (uint64_t)inc;
  r.state_hi = r.state_lo = 0;
  r.has_u32 = 0;
  r.u32 = 0;
  r.pad = 0;
  anm_pcg_step(r);
  anm_u128 s = (((anm_u128)r.state_hi << 64) | r.state_lo) + initstate;
  r.state_hi = (uint64_t)(s >> 64);
  r.state_lo = (uint64_t)s;
  anm_pcg_step(r);
}
