// knockoffgwas_b200/csrc/kgw_device.cuh
//
// sm_100a kernel for the unrelated-haplotype branch of snpknock2's knockoff sampler
// (msesia/knockoffgwas, snpknock2/src/knockoffs.cpp:636-700 and the four functions it calls:
//  compute_hmm_backward :1041, sample_posterior_MC :1215, sample_knockoff_MC :1270,
//  sample_emission_copula :1469; categorical draws = weighted_choice, utils.cpp:690).
//
// Mapping.  L lanes (4/8/16/32) of a warp cooperate on one haplotype; lane l owns the S
// consecutive HMM states k = l*S .. l*S+S-1 in registers (K <= L*S).  All haplotypes of a warp
// walk the chromosome in lock step, so every shuffle is a full-mask, width-L segmented shuffle.
// Per haplotype, per window pass:
//   sweep 1  backward messages beta_j (unnormalised, exactly rescaled by powers of two),
//            checkpointed every B loci to HBM;
//   sweep 2  per block of B loci: recompute beta from the checkpoint into shared memory, then draw
//            the latent path z_j forward through the block (prefix-scan categorical draw);
//   sweep 3  group knockoff path zk_j (partition-function recursion N_k carried in registers),
//            then knockoff alleles for each 32-locus word with lanes mapped to loci, packed and
//            stored as one 32-bit word.
// The O(K) per-locus recurrence is FP64 / issue bound; the backward table is never spilled
// (HBM sees only checkpoints: 16*K/B bytes per haplotype-SNP instead of 16*K).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kgw_philox.h"

struct __align__(16) KgwLocusHMM {
  double omm;  // 1.0 - mu_j
  double mu;   // mu_j
  double mup;  // 1.0 - (1.0 - mu_j)   (knockoffs.cpp:1251-1256 quirk: emission for h=0, ref=1)
  double a;    // (1.0 - b_j) * alpha,  alpha = 1.0/K (knockoffs.cpp:75)
  double b;    // b_j
  double pad;
};

enum { KGW_KO_FIRST = 1, KGW_KO_LASTGRP = 2, KGW_KO_GZERO = 4, KGW_KO_GSTART = 8, KGW_KO_LASTINGRP = 16 };

struct __align__(16) KgwLocusKO {  // per (window pass, locus of the padded range)
  double vs;    // vstar[t]                       knockoffs.cpp:1298-1313
  double vb;    // vbar[t][.] (same for every state because alpha is uniform)  :1315-1339
  double a;     // (1-b_j)*alpha
  double apb;   // (1-b_j)*alpha + b_j*1.0
  double aa;    // a*a                 T_k(z0)*T_k(z0k), k notin {z0,z0k}     :1352-1354
  double ab;    // (a+b)*a             k equals exactly one of z0, z0k
  double bb;    // (a+b)*(a+b)         k == z0 == z0k
  double sa;    // sum_k (1-b_j)*alpha (K sequential terms): N_sum when g == g_start  :1348-1349
  double g0;    // (vstar[0]*(1-b_j))*alpha: N update when g == 0            :1363-1365
  int32_t flags;
  int32_t z1_idx;   // locus holding z1 for this locus's group (first locus of group g+1), -1 if none
  int32_t z1_next;  // same for group g+1 (prefetched one group ahead), -1 if none
  int32_t pad;
  double pad2;
};

struct __align__(16) KgwEmit {  // sample_emission_copula (knockoffs.cpp:1469-1502) tabulated per locus:
  double w0[8];                 // weights_binary[0] for combo c = h | r_z<<1 | r_zk<<2
  double sum[8];                // 0.0 + weights_binary[0] + weights_binary[1]
};

struct KgwPass {
  int32_t js, je;   // padded range [js, je)           knockoffs.cpp:1222-1226
  int32_t cs, ce;   // core range   [cs, ce)           knockoffs.cpp:677-680
  int32_t ko_off;   // offset of this pass's KgwLocusKO rows
  int32_t w;        // window index (selects references_local[i][w])
  int32_t pad0, pad1;
};

struct KgwParams {
  int32_t N, p, K, W;
  int32_t n_haps, n_passes;
  int32_t stride_w;  // 32-bit words per packed row
  int32_t B;         // checkpoint spacing (4, 8, 16 or 32)
  const uint32_t *Hw;
  const int32_t *ref_local;  // [N][W][K]
  const int32_t *haps;       // [n_haps]
  const KgwLocusHMM *hmm;    // [p]
  const KgwLocusKO *ko;
  const KgwEmit *emit;       // [p]
  const KgwPass *passes;
  uint32_t *Hk;              // [N][stride_w]
  double *ckpt;              // [slots][ckpt_stride]
  unsigned long long ckpt_stride;
  uint8_t *zbuf, *zkbuf;     // [slots][z_stride]
  unsigned long long z_stride;
  unsigned long long seed;
  const double *injected_u;  // [N][3][p] or null
  int32_t *dbg_z, *dbg_zk;   // [n_haps][p] or null
  double *dbg_beta;          // [n_haps][nblk][K] or null (last pass)
  int32_t dbg_nblk;
  int32_t pad;
};

namespace kgw {

constexpr unsigned FULL = 0xffffffffu;
constexpr int STREAM_Z = 0, STREAM_ZK = 1, STREAM_HK = 2;  // include/kgw.h KGW_STREAM_*
constexpr double N_MIN = 1.0e-10;  // knockoffs.cpp:1295

template <int L>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o, L);
  return v;
}

template <int L>
__device__ __forceinline__ int group_sum_int(int v) {
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o, L);
  return v;
}

template <int L>
__device__ __forceinline__ unsigned group_or(unsigned v) {
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) v |= __shfl_xor_sync(FULL, v, o, L);
  return v;
}

// exclusive prefix of `v` over the L lanes of a group (lane order), plus the group total
template <int L>
__device__ __forceinline__ double group_excl_scan(double v, int lig, double &total) {
  double inc = v;
#pragma unroll
  for (int o = 1; o < L; o <<= 1) {
    const double t = __shfl_up_sync(FULL, inc, o, L);
    if (lig >= o) inc += t;
  }
  total = __shfl_sync(FULL, inc, L - 1, L);
  const double prev = __shfl_up_sync(FULL, inc, 1, L);
  return lig == 0 ? 0.0 : prev;
}

template <int L, int S>
struct Smem {
  // per haplotype group: beta block [B][S][L] doubles, uniforms [32] doubles, z chunk, zk chunk
  __host__ __device__ static constexpr size_t bytes(int B) {
    // + 64 B so that consecutive groups start 16 banks apart (8-byte accesses, half-warp phases)
    return (size_t)B * S * L * 8 + 32 * 8 + 64 + 64;
  }
};

// one backward step: beta <- beta_{m-1} from beta_m using locus m (knockoffs.cpp:1047-1074,
// unnormalised; the reference's per-locus normalisation is a positive scale that cancels in every draw)
template <int L, int S>
__device__ __forceinline__ double backward_step(double (&beta)[S], const uint32_t (&mw)[S], int bit,
                                                const KgwLocusHMM &h, int k0, int K) {
  double e[S];
  double acc = 0.0;
#pragma unroll
  for (int s = 0; s < S; s++) {
    const bool match = (mw[s] >> bit) & 1u;
    e[s] = beta[s] * (match ? h.omm : h.mu);
    acc += e[s];
  }
  const double Se = group_sum<L>(acc);
  const double c = h.a * Se;
#pragma unroll
  for (int s = 0; s < S; s++)
    if (k0 + s < K) beta[s] = fma(h.b, e[s], c);
  return Se;
}

template <int L, int S>
__device__ __forceinline__ void load_match_words(uint32_t (&mw)[S], const uint32_t *__restrict__ Hw,
                                                 const size_t (&roff)[S], size_t hoff, int wi, uint32_t &hw) {
  hw = __ldg(Hw + hoff + wi);
#pragma unroll
  for (int s = 0; s < S; s++) mw[s] = ~(hw ^ __ldg(Hw + roff[s] + wi));
}

// uniforms of one 32-locus chunk of one stream into shared memory (one double per locus)
template <int L>
__device__ __forceinline__ void fill_uniforms(double *su, const KgwParams &P, int hap, int stream, int wi, int lig) {
  constexpr int LPL = 32 / L;  // loci per lane
  if (P.injected_u != nullptr) {
    const double *U = P.injected_u + ((size_t)hap * 3 + (stream & 3)) * (size_t)P.p;
#pragma unroll
    for (int c = 0; c < LPL; c++) {
      const int idx = lig * LPL + c;
      const int j = (wi << 5) + idx;
      su[idx] = (j < P.p) ? U[j] : 0.0;
    }
  } else if (LPL >= 4) {
#pragma unroll
    for (int c = 0; c < (LPL >= 4 ? LPL / 4 : 1); c++) {
      const int idx = lig * LPL + 4 * c;
      const kgw_u32x4 r = kgw_philox_block(P.seed, (uint32_t)hap, (uint32_t)stream, (uint32_t)(((wi << 5) + idx) >> 2));
      su[idx + 0] = kgw_word_to_uniform(r.x);
      su[idx + 1] = kgw_word_to_uniform(r.y);
      su[idx + 2] = kgw_word_to_uniform(r.z);
      su[idx + 3] = kgw_word_to_uniform(r.w);
    }
  } else {
    const int idx = lig * LPL;
    const kgw_u32x4 r = kgw_philox_block(P.seed, (uint32_t)hap, (uint32_t)stream, (uint32_t)(((wi << 5) + idx) >> 2));
    const uint32_t w4[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int c = 0; c < LPL; c++) su[idx + c] = kgw_word_to_uniform(w4[(idx + c) & 3]);
  }
}

template <int L, int S, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) unrelated_kernel(const __grid_constant__ KgwParams P) {
  constexpr int GPC = THREADS / L;  // haplotype groups per CTA
  constexpr int GPW = 32 / L;       // haplotype groups per warp
  constexpr int LPL = 32 / L;       // loci per lane in lane<->locus phases
  const int lig = threadIdx.x % L;
  const int gic = threadIdx.x / L;
  const int slot = blockIdx.x * GPC + gic;
  const int nslots = gridDim.x * GPC;
  const int K = P.K;
  const int B = P.B;
  const int Bm = B - 1;
  const int k0 = lig * S;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char *gs = smem_raw + (size_t)gic * Smem<L, S>::bytes(B);
  double *sbeta = reinterpret_cast<double *>(gs);           // [B][S][L]
  double *su = sbeta + (size_t)B * S * L;                   // [32]
  uint8_t *sz = reinterpret_cast<uint8_t *>(su + 32);       // [32]
  uint8_t *szk = sz + 32;                                   // [32]

  double *ck = P.ckpt + (size_t)slot * P.ckpt_stride;       // [nblk][S][L]
  uint8_t *zb = P.zbuf + (size_t)slot * P.z_stride;
  uint8_t *zkb = P.zkbuf + (size_t)slot * P.z_stride;

  const int warp_first = blockIdx.x * GPC + (threadIdx.x >> 5) * GPW;
  for (int t0 = warp_first, t = slot; t0 < P.n_haps; t0 += nslots, t += nslots) {
    const bool active = t < P.n_haps;
    const int tt = active ? t : P.n_haps - 1;
    const int hap = __ldg(P.haps + tt);
    const size_t hoff = (size_t)hap * P.stride_w;

    for (int ip = 0; ip < P.n_passes; ip++) {
      const KgwPass ps = P.passes[ip];
      const int js = ps.js, je = ps.je;
      const int32_t *ref = P.ref_local + ((size_t)hap * P.W + ps.w) * K;
      size_t roff[S];
#pragma unroll
      for (int s = 0; s < S; s++) roff[s] = (size_t)((k0 + s < K) ? __ldg(ref + k0 + s) : hap) * P.stride_w;

      // ------------------------------------------------------------------ sweep 1: backward + checkpoints
      {
        double beta[S];
#pragma unroll
        for (int s = 0; s < S; s++) beta[s] = (k0 + s < K) ? 1.0 / K : 0.0;  // knockoffs.cpp:1046
        {
          const int n = (je - 1) / B;
#pragma unroll
          for (int s = 0; s < S; s++) ck[((size_t)n * S + s) * L + lig] = beta[s];
        }
        int wi = (je - 1) >> 5;
        const int wlo = js >> 5;
        uint32_t mw[S], nx[S], hw;
        load_match_words<L, S>(mw, P.Hw, roff, hoff, wi, hw);
        if (wi > wlo) load_match_words<L, S>(nx, P.Hw, roff, hoff, wi - 1, hw);
        for (int m = je - 1; m > js; m--) {
          if ((m >> 5) != wi) {
            wi = m >> 5;
#pragma unroll
            for (int s = 0; s < S; s++) mw[s] = nx[s];
            if (wi > wlo) load_match_words<L, S>(nx, P.Hw, roff, hoff, wi - 1, hw);
          }
          const KgwLocusHMM h = P.hmm[m];
          const double Se = backward_step<L, S>(beta, mw, m & 31, h, k0, K);
          const int j = m - 1;
          if ((j & Bm) == Bm) {
            // exact power-of-two rescale so that sum_k beta_j(k) lands in [1,2)
            const int ex = (__double2hiint(Se) >> 20) & 0x7ff;
            if (ex > 0 && ex < 2046) {
              const double sc = __hiloint2double((2046 - ex) << 20, 0);
#pragma unroll
              for (int s = 0; s < S; s++) beta[s] *= sc;
            }
            const int n = j / B;
#pragma unroll
            for (int s = 0; s < S; s++) ck[((size_t)n * S + s) * L + lig] = beta[s];
          }
        }
      }

      // debug: normalised backward messages at the checkpoint loci of the last pass
      if (P.dbg_beta != nullptr && ip == P.n_passes - 1) {
        __syncwarp();
        for (int n = js / B; n <= (je - 1) / B; n++) {
          double v[S];
          double acc = 0.0;
#pragma unroll
          for (int s = 0; s < S; s++) { v[s] = ck[((size_t)n * S + s) * L + lig]; acc += v[s]; }
          const double tot = group_sum<L>(acc);
          if (active && n < P.dbg_nblk) {
#pragma unroll
            for (int s = 0; s < S; s++)
              if (k0 + s < K) P.dbg_beta[((size_t)tt * P.dbg_nblk + n) * K + k0 + s] = v[s] / tot;
          }
        }
      }

      // ------------------------------------------------------------------ sweep 2: recompute + draw z
      {
        int zprev = -1;
        int wi = -1;
        const int whi = (je - 1) >> 5;
        uint32_t mw[S], nx[S], hw = 0, hwn = 0;
        load_match_words<L, S>(nx, P.Hw, roff, hoff, js >> 5, hwn);
        for (int n = js / B; n <= (je - 1) / B; n++) {
          const int lo = max(n * B, js);
          const int top = min(n * B + B, je) - 1;
          if ((lo >> 5) != wi) {
            wi = lo >> 5;
#pragma unroll
            for (int s = 0; s < S; s++) mw[s] = nx[s];
            hw = hwn;
            if (wi < whi) load_match_words<L, S>(nx, P.Hw, roff, hoff, wi + 1, hwn);
            __syncwarp();
            fill_uniforms<L>(su, P, hap, STREAM_Z | (ps.w << 2), wi, lig);
            __syncwarp();
          }
          double beta[S];
#pragma unroll
          for (int s = 0; s < S; s++) beta[s] = ck[((size_t)n * S + s) * L + lig];
          const int nb = n * B;
          if (top > lo) {
#pragma unroll
            for (int s = 0; s < S; s++) sbeta[((size_t)(top - nb) * S + s) * L + lig] = beta[s];
          }
          for (int m = top; m > lo; m--) {
            const KgwLocusHMM h = P.hmm[m];
            backward_step<L, S>(beta, mw, m & 31, h, k0, K);
            if (m - 1 > lo) {
#pragma unroll
              for (int s = 0; s < S; s++) sbeta[((size_t)(m - 1 - nb) * S + s) * L + lig] = beta[s];
            }
          }
          // forward draws through the block (knockoffs.cpp:1232-1260)
          for (int j = lo; j <= top; j++) {
            if (j > lo) {
#pragma unroll
              for (int s = 0; s < S; s++) beta[s] = sbeta[((size_t)(j - nb) * S + s) * L + lig];
            }
            const KgwLocusHMM h = P.hmm[j];
            const int bit = j & 31;
            const bool hbit = (hw >> bit) & 1u;
            const double mis = hbit ? h.mu : h.mup;
            double pre[S];
            double run = 0.0, ez = 0.0;
#pragma unroll
            for (int s = 0; s < S; s++) {
              const bool match = (mw[s] >> bit) & 1u;
              const double e = (match ? h.omm : mis) * beta[s];
              if (k0 + s == zprev) ez = e;
              run += e;
              pre[s] = run;
            }
            double total;
            const double base = group_excl_scan<L>(run, lig, total);
            const bool first = (j == js);
            ez = __shfl_sync(FULL, ez, first ? 0 : zprev / S, L);
            const double X = first ? 0.0 : h.b * ez;
            const double R = su[bit] * fma(h.a, total, X);
            int cnt = 0;
#pragma unroll
            for (int s = 0; s < S; s++) {
              const double C = fma(h.a, base + pre[s], (!first && k0 + s >= zprev) ? X : 0.0);
              cnt += (R >= C) ? 1 : 0;
            }
            cnt = group_sum_int<L>(cnt);
            const int z = min(cnt, K - 1);
            zprev = z;
            if (lig == 0) sz[bit] = (uint8_t)z;
          }
          if ((top & 31) == 31 || top == je - 1) {
            __syncwarp();
#pragma unroll
            for (int c = 0; c < LPL; c++) {
              const int idx = lig * LPL + c;
              const int j = (wi << 5) + idx;
              if (j >= js && j < je) zb[j] = sz[idx];
            }
            __syncwarp();
          }
        }
      }
      __syncwarp();

      // ------------------------------------------------------------------ sweep 3: knockoff path + alleles
      {
        double Nold[S], Iv[S];
#pragma unroll
        for (int s = 0; s < S; s++) { Nold[s] = (k0 + s < K) ? 1.0 : 0.0; Iv[s] = Nold[s]; }  // :1293-1294
        int z0 = -1, z0k = -1, zkprev = -1;
        if (js > 0) { z0 = zb[js - 1]; z0k = zkb[js - 1]; }
        const KgwLocusKO *ko_tab = P.ko + ps.ko_off - js;
        int z1 = 0, z1_pref = 0;
        {
          const KgwLocusKO &k = ko_tab[js];
          if (k.z1_idx >= 0) z1_pref = zb[k.z1_idx];
        }
        const int wlo = js >> 5, whi = (je - 1) >> 5;
        for (int wi = wlo; wi <= whi; wi++) {
          const int lo = max(wi << 5, js), hi = min((wi << 5) + 32, je);
          __syncwarp();
          fill_uniforms<L>(su, P, hap, STREAM_ZK | (ps.w << 2), wi, lig);
#pragma unroll
          for (int c = 0; c < LPL; c++) {
            const int idx = lig * LPL + c;
            const int j = (wi << 5) + idx;
            sz[idx] = (j >= js && j < je) ? zb[j] : (uint8_t)0;
          }
          __syncwarp();
          for (int j = lo; j < hi; j++) {
            const KgwLocusKO k = ko_tab[j];
            const int bit = j & 31;
            const int zj = sz[bit];
            double q[S], pre[S];
            double run = 0.0, ez1 = 0.0;
            if (k.flags & KGW_KO_FIRST) {
              z1 = z1_pref;
              if (k.z1_next >= 0) z1_pref = zb[k.z1_next];
              const bool gzero = k.flags & KGW_KO_GZERO;
#pragma unroll
              for (int s = 0; s < S; s++) {
                const int ks = k0 + s;
                const double tt = gzero ? k.a
                                        : ((ks == z0) ? ((ks == z0k) ? k.bb : k.ab) : ((ks == z0k) ? k.ab : k.aa));
                q[s] = tt * Iv[s];
                if (ks == z1) ez1 = q[s];
                run += q[s];
                pre[s] = run;
              }
            } else {
#pragma unroll
              for (int s = 0; s < S; s++) {
                const int ks = k0 + s;
                q[s] = (ks < K) ? ((ks == zkprev) ? k.apb : k.a) : 0.0;
                if (ks == z1) ez1 = q[s];
                run += q[s];
                pre[s] = run;
              }
            }
            double Qtot;
            const double base = group_excl_scan<L>(run, lig, Qtot);
            if (k.flags & KGW_KO_FIRST) {
              // partition function update (knockoffs.cpp:1345-1390)
              const double Nsum = (k.flags & KGW_KO_GSTART) ? k.sa : Qtot;
              double part = 0.0;
              const bool upd = !(k.flags & KGW_KO_LASTGRP);
              const bool gzero = k.flags & KGW_KO_GZERO;
              const double c2 = k.vb * Nsum;
#pragma unroll
              for (int s = 0; s < S; s++) {
                if (k0 + s < K) {
                  double Nv = Nold[s];
                  if (upd) {
                    Nv += c2;
                    Nv += gzero ? k.g0 : k.vs * q[s];
                  }
                  Nv = fmax(Nv, N_MIN);
                  Nold[s] = Nv;
                  part += Nv;
                }
              }
              const double rN = 1.0 / group_sum<L>(part);
#pragma unroll
              for (int s = 0; s < S; s++) {
                if (k0 + s < K) {
                  const double Nv = fmax(Nold[s] * rN, N_MIN);
                  Nold[s] = Nv;
                  Iv[s] = __drcp_rn(Nv);
                }
              }
            }
            // categorical draw (knockoffs.cpp:1392-1421)
            const bool lastgrp = k.flags & KGW_KO_LASTGRP;
            ez1 = __shfl_sync(FULL, ez1, lastgrp ? 0 : z1 / S, L);
            const double X = lastgrp ? 0.0 : k.vs * ez1;
            const double R = su[bit] * fma(k.vb, Qtot, X);
            int cnt = 0;
#pragma unroll
            for (int s = 0; s < S; s++) {
              const double C = fma(k.vb, base + pre[s], (!lastgrp && k0 + s >= z1) ? X : 0.0);
              cnt += (R >= C) ? 1 : 0;
            }
            cnt = group_sum_int<L>(cnt);
            const int zk = min(cnt, K - 1);
            zkprev = zk;
            if (k.flags & KGW_KO_LASTINGRP) { z0 = zj; z0k = zk; }
            if (lig == 0) szk[bit] = (uint8_t)zk;
          }
          __syncwarp();
          // flush zk chunk, then knockoff alleles for the core loci of this word (lanes <-> loci)
          unsigned bits = 0u, mask = 0u;
          const uint32_t hw = __ldg(P.Hw + hoff + wi);
#pragma unroll
          for (int c = 0; c < LPL; c++) {
            const int idx = lig * LPL + c;
            const int j = (wi << 5) + idx;
            if (j >= lo && j < hi) {
              const int zj = sz[idx], zkj = szk[idx];
              zkb[j] = (uint8_t)zkj;
              if (j >= ps.cs && j < ps.ce) {
                const uint32_t w1 = __ldg(P.Hw + (size_t)__ldg(ref + zj) * P.stride_w + wi);
                const uint32_t w2 = __ldg(P.Hw + (size_t)__ldg(ref + zkj) * P.stride_w + wi);
                const int combo = ((hw >> idx) & 1u) | (((w1 >> idx) & 1u) << 1) | (((w2 >> idx) & 1u) << 2);
                double u;
                if (P.injected_u != nullptr) {
                  u = P.injected_u[((size_t)hap * 3 + STREAM_HK) * (size_t)P.p + j];
                } else {
                  const kgw_u32x4 r = kgw_philox_block(P.seed, (uint32_t)hap, (uint32_t)(STREAM_HK | (ps.w << 2)), (uint32_t)(j >> 2));
                  const uint32_t w4[4] = {r.x, r.y, r.z, r.w};
                  u = kgw_word_to_uniform(w4[j & 3]);
                }
                const KgwEmit &em = P.emit[j];
                const unsigned hk = (u * em.sum[combo] < em.w0[combo]) ? 0u : 1u;  // weighted_choice, 2 weights
                bits |= hk << idx;
                mask |= 1u << idx;
              }
            }
          }
          bits = group_or<L>(bits);
          mask = group_or<L>(mask);
          if (lig == 0 && active && mask != 0u) {
            uint32_t *dst = P.Hk + hoff + wi;
            if (mask == 0xffffffffu) *dst = bits;
            else *dst = (*dst & ~mask) | bits;
          }
        }
      }
      __syncwarp();
    }  // passes

    if (active && (P.dbg_z != nullptr || P.dbg_zk != nullptr)) {
      __syncwarp();
      for (int j = lig; j < P.p; j += L) {
        if (P.dbg_z) P.dbg_z[(size_t)tt * P.p + j] = zb[j];
        if (P.dbg_zk) P.dbg_zk[(size_t)tt * P.p + j] = zkb[j];
      }
    }
    __syncwarp();
  }
}

}  // namespace kgw
