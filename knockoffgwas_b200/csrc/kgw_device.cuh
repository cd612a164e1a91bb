// knockoffgwas_b200/csrc/kgw_device.cuh
//
// sm_100a kernel for the unrelated-haplotype branch of snpknock2's knockoff sampler
// (msesia/knockoffgwas, snpknock2/src/knockoffs.cpp:636-700 and the four functions it calls:
//  compute_hmm_backward :1041, sample_posterior_MC :1215, sample_knockoff_MC :1270,
//  sample_emission_copula :1469; categorical draws = weighted_choice, utils.cpp:690).
//
// Mapping.  A warp carries G = 32/L haplotypes; L lanes (1, 2, 4 or 8) cooperate on one haplotype
// and lane `lig` owns the S consecutive HMM states k = lig*S .. lig*S+S-1 in registers (K <= L*S).
// Lane index = lig*G + g, so that for a fixed state the G haplotypes of a warp touch consecutive
// 8-byte words in shared memory and in the HBM scratch (conflict-free / coalesced).  Warps are
// independent (no CTA barrier); each resident warp owns one scratch slot and walks a strided list
// of haplotype batches.  Per batch and window pass:
//   sweep 1  backward messages beta_j (unnormalised, power-of-two rescaled every 8 loci), a
//            checkpoint every B loci and the per-lane emission-weighted sum of every locus to HBM;
//   sweep 2  per block of B loci: checkpoint -> shared-memory rows (B-1 recomputed steps), then the
//            latent path z_j left to right, one pass per locus (the locus total is known from sweep 1);
//   sweep 3  group knockoff path zk_j: at the first locus of a group the partition function N_k
//            (registers) and its reciprocals (one shared-memory row, patched at the <= 3 special
//            states) are updated; inside a group the draw has a closed form; knockoff alleles are
//            drawn per 32-locus word and stored packed.
// The O(K) per-locus recurrences are FP64-issue bound; HBM sees 16*K/B + 8*L bytes per haplotype-SNP.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kgw_philox.h"

struct __align__(16) KgwLocusHMM {  // per locus, built on the host in the reference's evaluation order
  double omm;   // 1.0 - mu_j
  double mu;    // mu_j
  double mup;   // 1.0 - (1.0 - mu_j)   (knockoffs.cpp:1251-1256 quirk: emission for h=0, ref=1)
  double a;     // (1.0 - b_j) * alpha,  alpha = 1.0/K (knockoffs.cpp:75)
  double b;     // b_j
  double aomm;  // a * omm
  double amu;   // a * mu
  double amup;  // a * mup
};

enum { KGW_KO_FIRST = 1, KGW_KO_LASTGRP = 2, KGW_KO_GZERO = 4, KGW_KO_GSTART = 8, KGW_KO_LASTINGRP = 16 };

struct __align__(16) KgwLocusKO {  // per (window pass, locus of the padded range)
  double vs;     // vstar[t]                                          knockoffs.cpp:1298-1313
  double vb;     // vbar[t][.] (same for every state: alpha uniform)   :1315-1339
  double a;      // (1-b_j)*alpha
  double apb;    // (1-b_j)*alpha + b_j*1.0
  double tg;     // first locus of a group: T_k(z0)*T_k(z0k) for k notin {z0,z0k} (= a*a), or a when g == 0
  double f1;     // ((a+b)*a) / (a*a)       weight ratio of a state equal to exactly one of z0, z0k
  double f2;     // ((a+b)*(a+b)) / (a*a)   weight ratio of k == z0 == z0k
  double fz1;    // 1 + vs/vb               weight ratio of k == z1 in the draw (0 in the last group)
  double sa;     // sum_k (1-b_j)*alpha, K sequential terms: N_sum when g == g_start   :1348-1349
  double vstg;   // vs * tg: coefficient of 1/N_old in the partition-function update   :1361-1375
  double gw;     // a * vb: weight of a generic state inside a group (t > 0)           :1397-1419
  double rgw;    // 1 / gw
  int32_t flags;
  int32_t z1_idx;   // locus holding z1 for this locus's group (first locus of group g+1), -1 if none
  int32_t z1_next;  // same for group g+1 (prefetched one group ahead), -1 if none
  int32_t pad;
  double pad2;
};

struct __align__(16) KgwEmit {  // sample_emission_copula (knockoffs.cpp:1469-1502) tabulated per locus:
  double2 ws[8];                // combo c = h | r_z<<1 | r_zk<<2:  x = weights_binary[0],
};                              //                                   y = 0.0 + weights_binary[0] + weights_binary[1]

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
  int32_t n_tasks;   // haplotype batches (G haplotypes each)
  const uint32_t *Hw;
  const int32_t *ref_local;  // [N][W][K]
  const int32_t *haps;       // [n_haps]
  const KgwLocusHMM *hmm;    // [p]
  const KgwLocusKO *ko;
  const KgwEmit *emit;       // [p]
  const KgwPass *passes;
  uint32_t *Hk;              // [N][stride_w]
  double *ckpt;              // [slots][nblk][L*S][G]
  unsigned long long ckpt_stride;
  double *sebuf;             // [slots][p][32]   per-lane emission-weighted sums of sweep 1
  unsigned long long se_stride;
  uint8_t *zbuf, *zkbuf;     // [slots][p][G]
  unsigned long long z_stride;
  unsigned long long seed;
  const double *injected_u;  // [N][3][p] or null
  int32_t *dbg_z, *dbg_zk;   // [n_haps][p] or null
  double *dbg_beta;          // [n_haps][nblk][K] or null (last pass)
  int32_t dbg_nblk;
  int32_t skip_sweeps;       // development only: bit s set = skip sweep s+1 (timing experiments; results invalid)
};

namespace kgw {

constexpr unsigned FULL = 0xffffffffu;
constexpr int STREAM_Z = 0, STREAM_ZK = 1, STREAM_HK = 2;  // include/kgw.h KGW_STREAM_*
constexpr double N_MIN = 1.0e-10;  // knockoffs.cpp:1295
constexpr int RESCALE = 8;         // loci between power-of-two rescales of the backward messages

// ---- reductions over the L lanes of one haplotype (lane = lig*G + g, G = 32/L)
template <int L>
__device__ __forceinline__ double group_sum(double v) {
  constexpr int G = 32 / L;
#pragma unroll
  for (int o = 16; o >= G; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
template <int L>
__device__ __forceinline__ int group_sum_int(int v) {
  constexpr int G = 32 / L;
#pragma unroll
  for (int o = 16; o >= G; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
template <int L>
__device__ __forceinline__ unsigned group_or(unsigned v) {
  constexpr int G = 32 / L;
#pragma unroll
  for (int o = 16; o >= G; o >>= 1) v |= __shfl_xor_sync(FULL, v, o);
  return v;
}
// exclusive prefix over lig (ascending) and the group total
template <int L>
__device__ __forceinline__ double group_excl(double v, int lane, double &total) {
  constexpr int G = 32 / L;
  if (L == 1) { total = v; return 0.0; }
  double inc = v;
#pragma unroll
  for (int o = G; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += t;
  }
  total = __shfl_sync(FULL, inc, (L - 1) * G + (lane % G));
  const double prev = __shfl_up_sync(FULL, inc, G);
  return lane < G ? 0.0 : prev;
}

__device__ __forceinline__ double fast_rcp(double x) {
  // x in [1e-10, ~1e3]: hardware seed + two Newton steps (relative error ~1e-16)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

// cnt += (v <= t): one DSETP and one predicated integer add (the compiler otherwise builds bit masks)
__device__ __forceinline__ void count_le(int &cnt, double v, double t) {
  asm("{\n\t.reg .pred p;\n\tsetp.le.f64 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(cnt) : "d"(v), "d"(t));
}

__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int L, int S, int B>
struct Cfg {
  static constexpr int G = 32 / L;
  static constexpr int KP = L * S;
  // per-warp shared memory: B-1 recomputed message rows + 2 checkpoint staging rows (sweep 2) / 1 reciprocal row
  // (sweep 3), two reference-word tiles, two staging tiles of per-lane sums, two latent-path tiles
  static constexpr size_t rows_bytes = (size_t)(B + 1) * KP * G * 8;
  static constexpr size_t refw_bytes = (size_t)2 * KP * G * 4;
  static constexpr size_t se_bytes = (size_t)2 * B * 32 * 8;
  static constexpr size_t tile_bytes = (size_t)2 * 32 * G;
  __host__ __device__ static constexpr size_t warp_bytes() { return rows_bytes + refw_bytes + se_bytes + tile_bytes; }
};

// uniform of (hap, stream, locus j): Philox block cached per 4 loci
struct UniformSrc {
  const double *inj;  // injected row or null
  unsigned long long seed;
  uint32_t hap, stream;
  kgw_u32x4 blk;
  int cur;
  __device__ __forceinline__ void init(const KgwParams &P, int hap_, int stream_, int w) {
    inj = P.injected_u ? P.injected_u + ((size_t)hap_ * 3 + stream_) * (size_t)P.p : nullptr;
    seed = P.seed;
    hap = (uint32_t)hap_;
    stream = (uint32_t)(stream_ | (w << 2));
    cur = -1;
  }
  __device__ __forceinline__ double get(int j) {
    if (inj) return inj[j];
    if ((j >> 2) != cur) {
      cur = j >> 2;
      blk = kgw_philox_block(seed, hap, stream, (uint32_t)cur);
    }
    const int q = j & 3;
    const uint32_t w = q == 0 ? blk.x : (q == 1 ? blk.y : (q == 2 ? blk.z : blk.w));
    return kgw_word_to_uniform(w);
  }
};

// emission-weighted messages e = beta * E and their in-lane sum (knockoffs.cpp:1052-1060 without the
// constant factor a); beta is overwritten by e.
template <int S, bool PAD>
__device__ __forceinline__ double emit_and_sum(double (&beta)[S], const uint32_t (&mw)[S], uint32_t bm,
                                               double omm, double mu, int ns) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
  for (int s = 0; s < S; s++) {
    double e = beta[s] * mu;
    if (mw[s] & bm) e = beta[s] * omm;
    beta[s] = (!PAD || s < ns) ? e : 0.0;
    if ((s & 3) == 0) a0 += beta[s];
    else if ((s & 3) == 1) a1 += beta[s];
    else if ((s & 3) == 2) a2 += beta[s];
    else a3 += beta[s];
  }
  return (a0 + a1) + (a2 + a3);
}

// PAD = false asserts K == L*S (no padded states: the per-state guards compile away);
// WPC = maximum warps per CTA, fewer may be launched
template <int L, int S, int B, bool PAD, int WPC, int MINB>
__global__ void __launch_bounds__(32 * WPC, MINB) sampler_kernel(const __grid_constant__ KgwParams P) {
  using C = Cfg<L, S, B>;
  constexpr int G = C::G;
  constexpr int KP = C::KP;
  const int lane = threadIdx.x & 31;
  const int wic = threadIdx.x >> 5;
  const int g = lane % G;
  const int lig = lane / G;
  const int k0 = lig * S;
  const int K = P.K;
  const int ns = PAD ? min(S, max(0, K - k0)) : S;  // states owned by this lane
  const int wpc = blockDim.x >> 5;
  const int slot = blockIdx.x * wpc + wic;
  const int nslots = gridDim.x * wpc;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char *ws = smem_raw + (size_t)wic * C::warp_bytes();
  double *rows = reinterpret_cast<double *>(ws);                                  // [B+1][KP][G]
  double *ckst = rows + (size_t)(B - 1) * KP * G;                                 // [2][KP][G] checkpoint staging
  uint32_t *refw = reinterpret_cast<uint32_t *>(ws + C::rows_bytes);              // [2][KP][G]
  double *sest = reinterpret_cast<double *>(ws + C::rows_bytes + C::refw_bytes);  // [2][B][32]
  uint8_t *ztile = ws + C::rows_bytes + C::refw_bytes + C::se_bytes;              // [32][G]
  uint8_t *zktile = ztile + 32 * G;                                               // [32][G]

  double *ck = P.ckpt + (size_t)slot * P.ckpt_stride;    // [nblk][KP][G]
  double *se = P.sebuf + (size_t)slot * P.se_stride;     // [p][32]
  uint8_t *zb = P.zbuf + (size_t)slot * P.z_stride;      // [p][G]
  uint8_t *zkb = P.zkbuf + (size_t)slot * P.z_stride;

  for (int task = slot; task < P.n_tasks; task += nslots) {
    const int tidx = task * G + g;
    const bool active = tidx < P.n_haps;
    const int tt = active ? tidx : P.n_haps - 1;
    const int hap = __ldg(P.haps + tt);
    const size_t hoff = (size_t)hap * P.stride_w;

    for (int ip = 0; ip < P.n_passes; ip++) {
      const KgwPass ps = P.passes[ip];
      const int js = ps.js, je = ps.je;
      const int wlo = js >> 5, whi = (je - 1) >> 5;
      const int32_t *ref = P.ref_local + ((size_t)hap * P.W + ps.w) * K;
      __syncwarp();

      // asynchronous prefetch of the reference words of 32-locus word `w` into tile `buf`
      auto prefetch_words = [&](int w, int buf) {
#pragma unroll
        for (int s = 0; s < S; s++) {
          const int id = (!PAD || s < ns) ? __ldg(ref + k0 + s) : hap;
          cp_async4(refw + buf * KP * G + (k0 + s) * G + g, P.Hw + (size_t)id * P.stride_w + w);
        }
      };

      // ------------------------------------------------------------------ sweep 1: backward + checkpoints
      if (!(P.skip_sweeps & 1)) {
        double beta[S];
        uint32_t mw[S];
#pragma unroll
        for (int s = 0; s < S; s++) beta[s] = (!PAD || s < ns) ? 1.0 / K : 0.0;  // knockoffs.cpp:1046
        {
          const int n = (je - 1) / B;
#pragma unroll
          for (int s = 0; s < S; s++) ck[((size_t)n * KP + k0 + s) * G + g] = beta[s];
        }
        int wi = whi;
        prefetch_words(wi, wi & 1);
        cp_async_wait_all();
        __syncwarp();
        uint32_t hw = __ldg(P.Hw + hoff + wi);
#pragma unroll
        for (int s = 0; s < S; s++) mw[s] = ~(hw ^ refw[(wi & 1) * KP * G + (k0 + s) * G + g]);
        if (wi > wlo) prefetch_words(wi - 1, (wi - 1) & 1);
        KgwLocusHMM h = P.hmm[je - 1];
        for (int m = je - 1; m >= js; m--) {
          if ((m >> 5) != wi) {
            wi = m >> 5;
            cp_async_wait_all();
            __syncwarp();
            hw = __ldg(P.Hw + hoff + wi);
#pragma unroll
            for (int s = 0; s < S; s++) mw[s] = ~(hw ^ refw[(wi & 1) * KP * G + (k0 + s) * G + g]);
            if (wi > wlo) prefetch_words(wi - 1, (wi - 1) & 1);
          }
          const KgwLocusHMM hn = P.hmm[m > js ? m - 1 : js];
          const uint32_t bm = 1u << (m & 31);
          const double pl = emit_and_sum<S, PAD>(beta, mw, bm, h.omm, h.mu, ns);  // beta := e
          se[(size_t)m * 32 + lane] = pl;
          if (m == js) {
            break;
          }
          const double Se = group_sum<L>(pl);
          const double c = h.a * Se;
#pragma unroll
          for (int s = 0; s < S; s++) beta[s] = (!PAD || s < ns) ? fma(h.b, beta[s], c) : 0.0;  // :1062-1068
          const int j = m - 1;
          if ((j % B) == B - 1) {
            if ((j % RESCALE) == RESCALE - 1) {
              // exact power-of-two rescale so that sum_k beta_j(k) stays near 1
              const int ex = (__double2hiint(Se) >> 20) & 0x7ff;
              if (ex > 0 && ex < 2046) {
                const double sc = __hiloint2double((2046 - ex) << 20, 0);
#pragma unroll
                for (int s = 0; s < S; s++) beta[s] *= sc;
              }
            }
            const int n = j / B;
#pragma unroll
            for (int s = 0; s < S; s++) ck[((size_t)n * KP + k0 + s) * G + g] = beta[s];
          }
          h = hn;
        }
      }
      __syncwarp();

      // debug: normalised backward messages at the checkpoint loci of the last pass
      if (P.dbg_beta != nullptr && ip == P.n_passes - 1) {
        for (int n = js / B; n <= (je - 1) / B; n++) {
          double v[S];
          double acc = 0.0;
#pragma unroll
          for (int s = 0; s < S; s++) { v[s] = ck[((size_t)n * KP + k0 + s) * G + g]; acc += v[s]; }
          const double tot = group_sum<L>(acc);
          if (active && n < P.dbg_nblk) {
#pragma unroll
            for (int s = 0; s < S; s++)
              if (!PAD || s < ns) P.dbg_beta[((size_t)tt * P.dbg_nblk + n) * K + k0 + s] = v[s] / tot;
          }
        }
      }

      // ------------------------------------------------------------------ sweep 2: recompute + draw z
      if (!(P.skip_sweeps & 2)) {
        double beta[S];
        uint32_t mw[S];
        UniformSrc us;
        us.init(P, hap, STREAM_Z, ps.w);
        int zprev = 0;
        int wi = wlo;
        prefetch_words(wi, wi & 1);
        cp_async_wait_all();
        __syncwarp();
        uint32_t hw = __ldg(P.Hw + hoff + wi);
#pragma unroll
        for (int s = 0; s < S; s++) mw[s] = ~(hw ^ refw[(wi & 1) * KP * G + (k0 + s) * G + g]);
        if (wi < whi) prefetch_words(wi + 1, (wi + 1) & 1);
        // loci of this word outside [js, je) keep the values of earlier window passes
        for (int i = lane; i < 8 * G; i += 32)
          reinterpret_cast<uint32_t *>(ztile)[i] = reinterpret_cast<const uint32_t *>(zb + (size_t)(wi << 5) * G)[i];
        __syncwarp();
        // checkpoint (= messages of the block's top locus) and per-lane sums of block n -> staging
        // (asynchronous, one block ahead)
        auto prefetch_block = [&](int n) {
          double *dst = ckst + (size_t)(n & 1) * KP * G;
#pragma unroll
          for (int s = 0; s < S; s++)
            cp_async8(dst + (k0 + s) * G + g, ck + ((size_t)n * KP + k0 + s) * G + g);
          const int lo_n = max(n * B, js), top_n = min(n * B + B, je) - 1;
          for (int j = lo_n; j <= top_n; j++)
            cp_async8(sest + ((n & 1) * B + (j - n * B)) * 32 + lane, se + (size_t)j * 32 + lane);
        };
        const int n_last = (je - 1) / B;
        prefetch_block(js / B);
        KgwLocusHMM hj = P.hmm[js];
        KgwLocusHMM hrec = P.hmm[min((js / B) * B + B, je) - 1];
        for (int n = js / B; n <= n_last; n++) {
          const int nb = n * B;
          const int lo = max(nb, js);
          const int top = min(nb + B, je) - 1;
          if ((lo >> 5) != wi) {
            // flush the finished z tile, then switch to the next 32-locus word
            __syncwarp();
            for (int i = lane; i < 8 * G; i += 32)
              reinterpret_cast<uint32_t *>(zb + (size_t)(wi << 5) * G)[i] = reinterpret_cast<uint32_t *>(ztile)[i];
            wi = lo >> 5;
            cp_async_wait_all();
            __syncwarp();
            for (int i = lane; i < 8 * G; i += 32)
              reinterpret_cast<uint32_t *>(ztile)[i] = reinterpret_cast<const uint32_t *>(zb + (size_t)(wi << 5) * G)[i];
            hw = __ldg(P.Hw + hoff + wi);
#pragma unroll
            for (int s = 0; s < S; s++) mw[s] = ~(hw ^ refw[(wi & 1) * KP * G + (k0 + s) * G + g]);
            if (wi < whi) prefetch_words(wi + 1, (wi + 1) & 1);
          }
          cp_async_wait_all();
          __syncwarp();
          if (n < n_last) prefetch_block(n + 1);
          const double *toprow = ckst + (size_t)(n & 1) * KP * G;
          const KgwLocusHMM hrec_next = P.hmm[min(nb + 2 * B, je) - 1];  // top locus of the next block
          if (B > 1) {
            // B-1 recomputed steps below the staged top row
#pragma unroll
            for (int s = 0; s < S; s++) beta[s] = toprow[(k0 + s) * G + g];
            KgwLocusHMM hr = hrec;
            for (int m = top; m > lo; m--) {
              const KgwLocusHMM h = hr;
              hr = P.hmm[m - 1];
              const double pl = emit_and_sum<S, PAD>(beta, mw, 1u << (m & 31), h.omm, h.mu, ns);
              const double Se = group_sum<L>(pl);
              const double c = h.a * Se;
#pragma unroll
              for (int s = 0; s < S; s++) {
                beta[s] = (!PAD || s < ns) ? fma(h.b, beta[s], c) : 0.0;
                rows[((m - 1 - nb) * KP + k0 + s) * G + g] = beta[s];
              }
            }
            __syncwarp();
          }
          hrec = hrec_next;
          // forward draws through the block (knockoffs.cpp:1232-1260)
          for (int j = lo; j <= top; j++) {
            const KgwLocusHMM h = hj;
            hj = P.hmm[j + 1 < je ? j + 1 : j];
            const uint32_t bm = 1u << (j & 31);
            const bool hbit = hw & bm;
            const double amis = hbit ? h.amu : h.amup;
            const double *row = (j == top) ? toprow : rows + (size_t)(j - nb) * KP * G;
            const double pl = sest[((n & 1) * B + (j - nb)) * 32 + lane];
            double Se;
            const double base = group_excl<L>(pl, lane, Se);
            // the one state whose transition weight carries b_j: k == z_{j-1}
            double X = 0.0;
            if (j > js) {
              const double bz = row[zprev * G + g];
              const uint32_t wz = ~(hw ^ refw[(wi & 1) * KP * G + zprev * G + g]);
              X = h.b * (bz * ((wz & bm) ? h.omm : (hbit ? h.mu : h.mup)));
            }
            const double R = us.get(j) * fma(h.a, Se, X);
            const double tA = R, tB = R - X;
            double run = h.a * base;
            int cA = 0, cB = 0;
#pragma unroll
            for (int s = 0; s < S; s++) {
              const double e = row[(k0 + s) * G + g] * ((mw[s] & bm) ? h.aomm : amis);
              run += e;
              count_le(cA, run, tA);
              count_le(cB, run, tB);
            }
            const int cc = group_sum_int<L>(cA | (cB << 16));
            const int k1 = cc & 0xffff, k2 = cc >> 16;
            int z = (k1 < zprev) ? k1 : max(k2, zprev);
            z = min(z, K - 1);
            zprev = z;
            if (lig == 0) ztile[(j & 31) * G + g] = (uint8_t)z;
          }
        }
        __syncwarp();
        for (int i = lane; i < 8 * G; i += 32)
          reinterpret_cast<uint32_t *>(zb + (size_t)(wi << 5) * G)[i] = reinterpret_cast<uint32_t *>(ztile)[i];
      }
      __syncwarp();

      // ------------------------------------------------------------------ sweep 3: knockoff path + alleles
      if (!(P.skip_sweeps & 4)) {
        double Nv[S];
#pragma unroll
        for (int s = 0; s < S; s++) Nv[s] = 1.0;  // knockoffs.cpp:1293-1294 (N == N_old between groups)
        double *rrow = rows;                      // [KP][G] reciprocals 1/N_old, patched per group
        UniformSrc us, uh;
        us.init(P, hap, STREAM_ZK, ps.w);
        uh.init(P, hap, STREAM_HK, ps.w);
        int z0 = 0, z0k = 0, zkprev = 0;
        if (js > 0) { z0 = zb[(size_t)(js - 1) * G + g]; z0k = zkb[(size_t)(js - 1) * G + g]; }
        const KgwLocusKO *ko_tab = P.ko + ps.ko_off - js;
        int z1 = 0, z1_pref = 0;
        {
          const int zi = ko_tab[js].z1_idx;
          if (zi >= 0) z1_pref = zb[(size_t)zi * G + g];
        }
        prefetch_words(wlo, wlo & 1);
        constexpr int TW = (8 * G + 31) / 32;  // 32-bit words of a latent-path tile per lane
        uint32_t zreg[TW], zkreg[TW];
#pragma unroll
        for (int i = 0; i < TW; i++) {
          const bool in = (G >= 4) || (lane + 32 * i < 8 * G);
          zreg[i] = in ? reinterpret_cast<const uint32_t *>(zb + (size_t)(wlo << 5) * G)[lane + 32 * i] : 0u;
          zkreg[i] = in ? reinterpret_cast<const uint32_t *>(zkb + (size_t)(wlo << 5) * G)[lane + 32 * i] : 0u;
        }
        for (int wi = wlo; wi <= whi; wi++) {
          const int lo = max(wi << 5, js), hi = min((wi << 5) + 32, je);
          __syncwarp();
          // latent-path tiles of this word (loaded into registers one word ahead), then prefetch the next
#pragma unroll
          for (int i = 0; i < TW; i++) {
            reinterpret_cast<uint32_t *>(ztile)[lane + 32 * i] = zreg[i];
            reinterpret_cast<uint32_t *>(zktile)[lane + 32 * i] = zkreg[i];
          }
          if (wi < whi) {
#pragma unroll
            for (int i = 0; i < TW; i++) {
              zreg[i] = reinterpret_cast<const uint32_t *>(zb + (size_t)((wi + 1) << 5) * G)[lane + 32 * i];
              zkreg[i] = reinterpret_cast<const uint32_t *>(zkb + (size_t)((wi + 1) << 5) * G)[lane + 32 * i];
            }
          }
          cp_async_wait_all();
          __syncwarp();
          if (wi < whi) prefetch_words(wi + 1, (wi + 1) & 1);
          KgwLocusKO k = ko_tab[lo];
          for (int j = lo; j < hi; j++) {
            const KgwLocusKO kn = ko_tab[j + 1 < je ? j + 1 : j];
            const int bit = j & 31;
            const int zj = ztile[bit * G + g];
            const bool lastg = k.flags & KGW_KO_LASTGRP;
            const double u = us.get(j);
            int zk;
            if (k.flags & KGW_KO_FIRST) {
              z1 = z1_pref;
              if (k.z1_next >= 0) z1_pref = zb[(size_t)k.z1_next * G + g];
              const bool gz = k.flags & KGW_KO_GZERO;
              // pass A: reciprocals of N_old and their sum
              double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
              for (int s = 0; s < S; s++) {
                const double r = (!PAD || s < ns) ? fast_rcp(Nv[s]) : 0.0;
                rrow[(k0 + s) * G + g] = r;
                if ((s & 3) == 0) a0 += r;
                else if ((s & 3) == 1) a1 += r;
                else if ((s & 3) == 2) a2 += r;
                else a3 += r;
              }
              double pls = (a0 + a1) + (a2 + a3);
              __syncwarp();
              // the (at most two) states equal to z0 / z0k carry T_k(z0)*T_k(z0k) != tg  (:1350-1354)
              if (!gz) {
                const double r0 = rrow[z0 * G + g];
                double n0, n0k = 0.0, r0k = 0.0;
                if (z0 == z0k) {
                  n0 = r0 * k.f2;
                } else {
                  n0 = r0 * k.f1;
                  r0k = rrow[z0k * G + g];
                  n0k = r0k * k.f1;
                }
                __syncwarp();
                if (lig == 0) {
                  rrow[z0 * G + g] = n0;
                  if (z0 != z0k) rrow[z0k * G + g] = n0k;
                }
                if (z0 / S == lig) pls += n0 - r0;
                if (z0 != z0k && z0k / S == lig) pls += n0k - r0k;
              }
              __syncwarp();
              const double Sr = group_sum<L>(pls);
              const double Nsum = (k.flags & KGW_KO_GSTART) ? k.sa : k.tg * Sr;  // :1345-1358
              const double c2 = k.vb * Nsum;
              // pass B: partition function update and normaliser (:1360-1390).  The last group keeps N
              // (zero coefficients make the update an exact identity) but still renormalises it.
              const double c2u = lastg ? 0.0 : c2;
              const double vsu = lastg ? 0.0 : k.vstg;
              double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
#pragma unroll
              for (int s = 0; s < S; s++) {
                double nv = fma(vsu, rrow[(k0 + s) * G + g], Nv[s] + c2u);
                nv = (!PAD || s < ns) ? nv : 0.0;
                Nv[s] = nv;
                if ((s & 3) == 0) b0 += nv;
                else if ((s & 3) == 1) b1 += nv;
                else if ((s & 3) == 2) b2 += nv;
                else b3 += nv;
              }
              const double rN = 1.0 / group_sum<L>((b0 + b1) + (b2 + b3));
              // pass C: normalise; clamp to N_min (:1386-1389; the pre-normalisation clamp of :1380 cannot
              // fire because N only grows).  The clamp is exact but rarely needed: a cheap integer minimum of
              // the high words decides whether any state can be below N_min.
              int hmin = 0x7fffffff;
#pragma unroll
              for (int s = 0; s < S; s++) {
                Nv[s] *= rN;
                if (!PAD || s < ns) hmin = min(hmin, __double2hiint(Nv[s]));
              }
              if (hmin <= __double2hiint(N_MIN)) {
#pragma unroll
                for (int s = 0; s < S; s++) Nv[s] = (Nv[s] < N_MIN) ? N_MIN : Nv[s];
              }
              // the state equal to z1 carries the extra vstar term in the draw (:1413-1419)
              __syncwarp();
              if (!lastg) {
                const double rz = rrow[z1 * G + g];
                const double nz = rz * k.fz1;
                __syncwarp();
                if (lig == 0) rrow[z1 * G + g] = nz;
                if (z1 / S == lig) pls += nz - rz;
              }
              __syncwarp();
              double St;
              const double base = group_excl<L>(pls, lane, St);
              const double t = u * St;
              // pass D: inverse-CDF search (weighted_choice, utils.cpp:690-698) in units of vb*tg
              double run = base;
              int cnt = 0;
#pragma unroll
              for (int s = 0; s < S; s++) {
                run += rrow[(k0 + s) * G + g];
                if (!PAD || s < ns) count_le(cnt, run, t);
              }
              zk = min(group_sum_int<L>(cnt), K - 1);
            } else {
              // inside a group every state has weight a*vb except k == zk_{j-1} and k == z1 (:1397-1419)
              const double wp = k.apb * ((!lastg && z1 == zkprev) ? (k.vb + k.vs) : k.vb);
              int pa = zkprev, pb = K;
              double ea = wp - k.gw, eb = 0.0;
              if (!lastg && z1 != zkprev) {
                const double e1 = k.a * (k.vb + k.vs) - k.gw;
                if (z1 < zkprev) { pa = z1; pb = zkprev; eb = ea; ea = e1; }
                else { pb = z1; eb = e1; }
              }
              const double total = fma((double)K, k.gw, ea + eb);
              const double R = u * total;
              auto F = [&](double tval) -> int {  // #{k in [0,K): (k+1)*gw <= tval}
                if (tval < 0.0) return 0;
                if (k.gw <= 0.0) return K;
                double q = tval * k.rgw;
                int kf = q >= (double)K ? K : (int)q;
                if (kf < K && (double)(kf + 1) * k.gw <= tval) kf++;
                if (kf > 0 && (double)kf * k.gw > tval) kf--;
                return kf;
              };
              int ans = F(R);
              if (ans >= pa) {
                ans = max(F(R - ea), pa);
                if (ans >= pb) ans = max(F(R - ea - eb), pb);
              }
              zk = min(ans, K - 1);
            }
            zkprev = zk;
            if (k.flags & KGW_KO_LASTINGRP) { z0 = zj; z0k = zk; }
            if (lig == 0) zktile[bit * G + g] = (uint8_t)zk;
            k = kn;
          }
          __syncwarp();
          // flush the zk tile, then knockoff alleles of the core loci of this word (:1469-1502)
          for (int i = lane; i < 8 * G; i += 32)
            reinterpret_cast<uint32_t *>(zkb + (size_t)(wi << 5) * G)[i] = reinterpret_cast<uint32_t *>(zktile)[i];
          const uint32_t hw = __ldg(P.Hw + hoff + wi);
          const uint32_t *rw = refw + (wi & 1) * KP * G;
          unsigned bits = 0u, mask = 0u;
          const int elo = max(lo, ps.cs), ehi = min(hi, ps.ce);
#pragma unroll 4
          for (int j = elo + lig; j < ehi; j += L) {
            const int bit = j & 31;
            const uint32_t w1 = rw[ztile[bit * G + g] * G + g];
            const uint32_t w2 = rw[zktile[bit * G + g] * G + g];
            const int combo = ((hw >> bit) & 1u) | (((w1 >> bit) & 1u) << 1) | (((w2 >> bit) & 1u) << 2);
            const double2 em = P.emit[j].ws[combo];
            const unsigned hk = (uh.get(j) * em.y < em.x) ? 0u : 1u;  // weighted_choice with 2 weights
            bits |= hk << bit;
            mask |= 1u << bit;
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
        if (P.dbg_z) P.dbg_z[(size_t)tt * P.p + j] = zb[(size_t)j * G + g];
        if (P.dbg_zk) P.dbg_zk[(size_t)tt * P.p + j] = zkb[(size_t)j * G + g];
      }
    }
    __syncwarp();
  }
}

}  // namespace kgw
