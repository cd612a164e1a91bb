/* oracle/kgw_rng.c -- TEST INFRASTRUCTURE ONLY.  See kgw_rng.h. */
#include "kgw_rng.h"

/* One linear-feedback shift engine with word size 32
 * (linear_feedback_shift.hpp:90-98 for seeding, :117-123 for the step). */
static uint32_t lfs_seed(uint32_t s0, int k) {
  uint32_t v = s0;
  if (v < (1u << (32 - k))) v += 1u << (32 - k);
  return v;
}

static uint32_t lfs_step(uint32_t *value, int k, int q, int s) {
  const uint32_t v = *value;
  const uint32_t b = ((v << q) ^ v) >> (k - s);
  const uint32_t mask = 0xFFFFFFFFu << (32 - k);
  *value = ((v & mask) << s) ^ b;
  return *value;
}

void kgw_taus88_seed(kgw_taus88 *g, uint32_t seed) {
  /* taus88.hpp:36-40: (k,q,s) = (31,13,12), (29,2,4), (28,3,17); all seeded with the same value */
  g->v1 = lfs_seed(seed, 31);
  g->v2 = lfs_seed(seed, 29);
  g->v3 = lfs_seed(seed, 28);
}

uint32_t kgw_taus88_next(kgw_taus88 *g) {
  const uint32_t a = lfs_step(&g->v1, 31, 13, 12);
  const uint32_t b = lfs_step(&g->v2, 29, 2, 4);
  const uint32_t c = lfs_step(&g->v3, 28, 3, 17);
  return a ^ b ^ c;
}

double kgw_taus88_uniform(kgw_taus88 *g) {
  return (double)kgw_taus88_next(g) / 4294967296.0;
}

void kgw_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

double kgw_oracle_philox_uniform(uint64_t seed, uint32_t hap, uint32_t stream, uint32_t locus) {
  const uint32_t ctr[4] = {locus >> 2, hap, stream, 0u};
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t out[4];
  kgw_oracle_philox4x32_10(ctr, key, out);
  return (double)out[locus & 3u] / 4294967296.0;
}
