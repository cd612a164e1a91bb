// knockoffgwas_b200/csrc/kgw_philox.h
// Counter-based generator shared by device kernels and the host replay (kgw_uniform).
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  Keying is specified in include/kgw.h:
//   u(seed, hap, stream, locus) = Philox(key={seed lo, seed hi}, ctr={locus>>2, hap, stream, 0})[locus&3] / 2^32
// It replaces the reference's per-thread sequential boost::random::taus88
// (snpknock2/src/knockoffs.cpp:596-612) by a stream that does not depend on the thread count.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KGW_HD __host__ __device__ __forceinline__
#else
#define KGW_HD inline
#endif

struct kgw_u32x4 { uint32_t x, y, z, w; };

KGW_HD void kgw_mulhilo32(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
#if defined(__CUDA_ARCH__)
  lo = a * b;
  hi = __umulhi(a, b);
#else
  const uint64_t p = (uint64_t)a * (uint64_t)b;
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
#endif
}

KGW_HD kgw_u32x4 kgw_philox4x32_10(kgw_u32x4 c, uint32_t k0, uint32_t k1) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; r++) {
    uint32_t hi0, lo0, hi1, lo1;
    kgw_mulhilo32(0xD2511F53u, c.x, hi0, lo0);
    kgw_mulhilo32(0xCD9E8D57u, c.z, hi1, lo1);
    kgw_u32x4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// the four uniforms' raw words for loci 4*blk .. 4*blk+3
KGW_HD kgw_u32x4 kgw_philox_block(uint64_t seed, uint32_t hap, uint32_t stream, uint32_t blk) {
  kgw_u32x4 c;
  c.x = blk; c.y = hap; c.z = stream; c.w = 0u;
  return kgw_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
}

KGW_HD double kgw_word_to_uniform(uint32_t x) { return (double)x * (1.0 / 4294967296.0); }
