/* oracle/kgw_rng.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Uniform sources for the CPU oracle:
 *   - taus88 replay of the reference's own generator
 *     (boost-1.55.0/boost/random/taus88.hpp:36-40,
 *      linear_feedback_shift.hpp:90-98 seed, :117-123 step,
 *      xor_combine.hpp:119-122, uniform_real_distribution.hpp:50-66: u = x / 2^32)
 *   - an independent restatement of the product's counter-based generator
 *     (Philox4x32-10, Salmon et al. SC'11; keyed (seed, hap, stream, locus) as
 *     specified in include/kgw.h) so that device and host replay can be
 *     cross-checked against a second implementation.
 */
#ifndef KGW_ORACLE_RNG_H
#define KGW_ORACLE_RNG_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint32_t v1, v2, v3; } kgw_taus88;

void kgw_taus88_seed(kgw_taus88 *g, uint32_t seed);     /* boost: default seed is 341 */
uint32_t kgw_taus88_next(kgw_taus88 *g);
double kgw_taus88_uniform(kgw_taus88 *g);               /* x / 4294967296.0 */

void kgw_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* u in [0,1): word (locus & 3) of Philox(key = seed, ctr = (locus >> 2, hap, stream, 0)) * 2^-32 */
double kgw_oracle_philox_uniform(uint64_t seed, uint32_t hap, uint32_t stream, uint32_t locus);

#ifdef __cplusplus
}
#endif
#endif
