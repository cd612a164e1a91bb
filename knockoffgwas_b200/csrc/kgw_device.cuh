// knockoffgwas_b200/csrc/kgw_device.cuh
//
// sm_100a kernel for the unrelated-haplotype branch of snpknock2's knockoff sampler
// (msesia/knockoffgwas, snpknock2/src/knockoffs.cpp:636-700 and the four functions it calls:
//  compute_hmm_backward :1041, sample_posterior_MC :1215, sample_knockoff_MC :1270,
//  sample_emission_copula :1469; categorical draws = weighted_choice, utils.cpp:690).
//
// Mapping.  A warp carries G = 32/L haplotypes; L lanes (1, 2, 4 or 8) cooperate on one haplotype and
// lane `lig` owns the S consecutive HMM states k = lig*S .. lig*S+S-1 in registers (K <= L*S).
// Lane index = lig*G + g.  Warps are independent (no CTA barrier anywhere); each resident warp
// owns one HBM scratch slot and walks a strided list of haplotype batches.  Per batch and window pass:
//   gather   match words ~(H[i] ^ H[ref_k]) of every (state, 32-locus word) -> slot scratch, once; the
//            sweeps stream them back tile by tile with 16-byte cp.async (double-buffered in shared memory);
//   sweep 0  (only when the chromosome is cut into segments because the scratch would not fit) backward
//            messages over the whole range, keeping the message row at the top locus of every segment;
//   then per segment, left to right:
//   sweep 1  backward messages beta_j (unnormalised, exact power-of-two rescale every 8 loci): a message
//            row every B loci, four partial sums of the emission-weighted messages per lane and locus,
//            and the locus total, to HBM;
//   sweep 2  latent path z_j left to right.  The categorical draw is a two-level inverse-CDF search: the
//            partial sums of sweep 1 locate the chunk of <= 8 states that holds the target, and only those
//            states' messages are rebuilt, from the staged message row of the block's top locus by the
//            O(1)-per-state recurrence beta_{m-1}(k) = c_m + b_m e_m(k) (c_m from the stored totals);
//   and finally
//   sweep 3  group knockoff path zk_j: at the first locus of a group the partition function N_k (registers)
//            and its reciprocals (one shared-memory row, patched at the <= 3 special states) are updated,
//            the draw is again two-level (chunk sums of the reciprocals, then one chunk); inside a group
//            the draw has a closed form; knockoff alleles are drawn per 32-locus word and stored packed.
// Shared-memory rows keep two states per 16-byte unit ([state pair][haplotype][2]), so that the owner
// lane moves its S states with ceil(S/2) 128-bit accesses and a chunk is 2 or 4 such units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kgw_philox.h"

struct __align__(16) KgwLocusHMM {  // per locus.  The derived constants of the reference, 1.0 - mu_j and
  double b;     // b_j                   // (1.0 - b_j) * alpha with alpha = 1.0/K (knockoffs.cpp:75), are single
  double mu;    // mu_j                  // IEEE operations and are formed on the device (hmm_consts)
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
  int32_t stride_w;  // 32-bit words per packed row (multiple of 4)
  int32_t n_tasks;   // haplotype batches (G haplotypes each)
  int32_t seg_loci;  // loci per segment (multiple of 32); >= p means one segment
  int32_t pad0;
  double alpha;      // 1.0 / K  (knockoffs.cpp:75), formed on the host
  const uint32_t *Hw;
  const int32_t *ref_local;  // [N][W][K]
  const int32_t *haps;       // [n_haps]
  const KgwLocusHMM *hmm;    // [p]
  const KgwLocusKO *ko;
  const KgwEmit *emit;       // [p]
  const KgwPass *passes;
  uint32_t *Hk;              // [N][stride_w]
  // per-slot scratch
  uint32_t *mwbuf;           // [slots][stride_w][TILE]         match words
  unsigned long long mw_stride;
  double *ckpt;              // [slots][seg_loci/B][ROW]        message rows of the current segment
  unsigned long long ckpt_stride;
  double *segck;             // [slots][nseg][ROW]              message rows at segment tops (sweep 0)
  unsigned long long segck_stride;
  double *csbuf;             // [slots][seg_loci][NC/2][32][2]  partial sums of sweep 1 (NC chunks per lane)
  unsigned long long cs_stride;
  double *sebuf;             // [slots][seg_loci][G]            locus totals of sweep 1
  unsigned long long se_stride;
  uint8_t *zbuf, *zkbuf;     // [slots][p32][G]
  unsigned long long z_stride;
  unsigned long long seed;
  const double *injected_u;  // [N][W][3][p] (indexed by absolute locus) or null
  int32_t *dbg_z, *dbg_zk;   // [n_haps][p] or null
  double *dbg_beta;          // [n_haps][nblk][K] or null (last pass)
  int32_t dbg_nblk;
  int32_t skip_sweeps;       // development only: bit s set = skip sweep s+1 (timing experiments; results invalid)
  long long *phase_cycles;   // development only (-DKGW_PHASES): [16] clock64 sums of slot 0, or null
};

// development only: per-phase cycle counters of the warp in slot 0
#ifdef KGW_PHASES
#define KGW_PH_DECL long long ph_t0 = clock64();
#define KGW_PH(i) do { if (slot == 0 && P.phase_cycles) { const long long t_ = clock64(); if (lane == 0) P.phase_cycles[i] += t_ - ph_t0; ph_t0 = clock64(); } } while (0)
#else
#define KGW_PH_DECL
#define KGW_PH(i) do { } while (0)
#endif

#ifndef KGW_WPC
#define KGW_WPC 2   // maximum warps per CTA (two halve the per-CTA shared-memory reservation per warp)
#endif

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

// group total and exclusive prefix over lig in one exchange (the total is the same in every lane:
// floating-point addition is commutative and every lane adds the same two partial results per step)
template <int L>
__device__ __forceinline__ double group_total_excl(double v, int lane, double &total) {
  constexpr int G = 32 / L;
  if (L == 1) { total = v; return 0.0; }
  if (L == 2) {
    const double o = __shfl_xor_sync(FULL, v, 16);
    total = v + o;
    return lane < G ? 0.0 : o;
  }
  return group_excl<L>(v, lane, total);
}

__device__ __forceinline__ double fast_rcp(double x) {
  // x in [1e-10, ~1e3]: hardware seed (upper 32 bits of x, >= 20 good bits) and one third-order step
  // y (1 + e + e^2), e = 1 - x y: relative error e^3 plus rounding
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);
}

// cnt += (v <= t): one DSETP and one predicated integer add (the compiler otherwise builds bit masks)
__device__ __forceinline__ void count_le(int &cnt, double v, double t) {
  asm("{\n\t.reg .pred p;\n\tsetp.le.f64 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(cnt) : "d"(v), "d"(t));
}

// cnt += (ok && v <= t)
__device__ __forceinline__ void count_le_if(int &cnt, double v, double t, bool ok) {
  asm("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 q, %3, 0;\n\tsetp.le.and.f64 p, %1, %2, q;\n\t@p add.s32 %0, %0, 1;\n\t}"
      : "+r"(cnt) : "d"(v), "d"(t), "r"((int)ok));
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// acc += v when a == b / a < b: one ISETP and one predicated DADD
__device__ __forceinline__ void add_if_eq(double &acc, int a, int b, double v) {
  asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, %2;\n\t@p add.f64 %0, %0, %3;\n\t}" : "+d"(acc) : "r"(a), "r"(b), "d"(v));
}
__device__ __forceinline__ void add_if_lt(double &acc, int a, int b, double v) {
  asm("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %1, %2;\n\t@p add.f64 %0, %0, %3;\n\t}" : "+d"(acc) : "r"(a), "r"(b), "d"(v));
}

__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
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

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier.  Every staged tile of this kernel is
// one contiguous, 16-byte aligned run of the slot scratch or of a per-locus table, so one elected lane moves it with
// ONE instruction and the other lanes keep their issue slots for arithmetic; each buffer has its own barrier, so a
// wait covers exactly the bytes it needs (cp.async.wait_all also waited for the far-ahead prefetches).
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile("{\n\t.reg .pred p;\n\tKGW_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra KGW_DONE;\n\t"
               "bra KGW_WAIT;\n\tKGW_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// Warp barrier that every lane passes only once the shared-memory loads feeding `x` have RETURNED.  A bulk copy that
// refills a buffer is issued by lane 0 through the async proxy and does not queue behind the other lanes' LDS in the
// load/store unit: when that unit is backed up (every warp of the SM leaves its gather at the same time) a refill from
// L2 can land before loads issued ahead of it have read the old contents.  __syncwarp() alone does not wait for
// outstanding loads; the vote needs `x` and therefore does.
__device__ __forceinline__ void syncwarp_after_loads(unsigned x) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %0, 0x80000001;\n\tvote.sync.any.pred p, p, 0xffffffff;\n\t}" ::"r"(x) : "memory");
}

// generic-proxy writes (global scratch, shared rows) before async-proxy accesses of the same bytes.  The bulk copies
// are issued by lane 0 and read what every lane of the warp stored a moment ago (sweep 1 starts at the word the gather
// wrote last): each lane first makes its own stores visible device-wide, then orders them against the async proxy; the
// caller follows with __syncwarp() before lane 0 issues.
__device__ __forceinline__ void fence_async_proxy() {
  __threadfence();
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncwarp();
  asm volatile("fence.proxy.async;" ::: "memory");   // ... and the issuing lane once more after it has synchronised with the writers
}

// streaming stores / loads of the scratch (written once, read once: keep it out of L1)
__device__ __forceinline__ void st_cg(double2 *p, double x, double y) {
  asm volatile("st.global.cg.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void st_cg(double *p, double x) {
  asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(p), "d"(x) : "memory");
}
__device__ __forceinline__ void st_cg(uint4 *p, uint4 v) {
  asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ double2 ld_cg(const double2 *p) {
  double2 v;
  asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_cg(const double *p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// SLIM: the shared-memory diet that lets the one-lane layout keep two warps per scheduler (8 per SM) on long runs of
// batches: ONE message-row buffer in sweep 2 (the row of the next block is fetched when the block starts instead of a
// block ahead), and in sweep 3 the knockoff-table rows in the ring, ONE pair of latent-path tiles in the block buffers.
template <int L, int S, int B, int D, bool SLIM = false>
struct Cfg {
  static constexpr int G = 32 / L;
  static constexpr int KP = L * S;
  static constexpr int SPR = (S + 1) / 2 * 2;   // state slots per lane in a message row
  static constexpr int SPW = (S + 3) / 4 * 4;   // state slots per lane in a match-word tile
  static constexpr int NQ = SPW / 4;            // 16-byte units of a lane in a match-word tile
  static constexpr int NP = SPR / 2;            // 16-byte units of a lane in a message row
  static constexpr int CS = S <= 16 ? 4 : 8;    // states per chunk
  static constexpr int NC = S <= 32 ? 4 : 8;    // chunks per lane (a lane holds at most 64 states)
  static_assert(S <= 64, "at most 64 states per lane");
  static constexpr int CSW = NC * 32;           // partial sums per locus and warp: [NC/2][32 lanes][2]
  static constexpr int ROW = L * SPR * G;       // doubles per message row
  static constexpr int TILE = L * SPW * G;      // 32-bit words per match-word tile
  // per-warp shared memory: two message rows (sweep 2: staged top rows of two blocks; sweep 3: the
  // reciprocal row, the two latent-path tiles and the 32 knockoff-table rows of the current word; sweeps 0/1: the
  // table rows of two words), one match-word tile, a ring of D loci of partial sums (sweep 2; sweep 3 keeps the
  // emission rows of the word there when it holds 4 KB), totals + table rows of two blocks, 16 mbarriers
  static constexpr size_t s3_bytes = (size_t)ROW * 8 + 4 * 32 * G + 32 * sizeof(KgwLocusKO);
  static constexpr size_t rows_bytes = SLIM ? (size_t)ROW * 8 : ((size_t)2 * ROW * 8 > s3_bytes ? (size_t)2 * ROW * 8 : s3_bytes);
  static constexpr size_t tiles_bytes = (size_t)TILE * 4;
  static constexpr size_t ring_bytes = (size_t)D * CSW * 8;
  static constexpr size_t blk_bytes = (size_t)2 * ((size_t)B * G * 8 + (size_t)B * 16);
  static constexpr size_t bar_bytes = 128;
  static_assert(!SLIM || (ring_bytes >= 32 * sizeof(KgwLocusKO) && blk_bytes >= 128 + 2 * 32 * G && ROW * 8 >= 1024),
                "slim layout: knockoff-table rows in the ring, one pair of latent-path tiles in the block buffers");
  __host__ __device__ static constexpr size_t warp_bytes() { return rows_bytes + tiles_bytes + ring_bytes + blk_bytes + bar_bytes; }
};
enum { BAR_S1 = 0, BAR_TILE2 = 1, BAR_BLK = 2, BAR_CS = 4, BAR_W3 = 12, BAR_ZT = 13, BAR_EM = 14, BAR_ROW = 15, KGW_NBARS = 16 };

// uniform of (hap, stream, locus j): Philox block cached per 4 loci
struct UniformSrc {
  const double *inj;  // injected row or null
  unsigned long long seed;
  uint32_t hap, stream;
  kgw_u32x4 blk;
  int cur;
  __device__ __forceinline__ void init(const KgwParams &P, int hap_, int stream_, int w) {
    inj = P.injected_u ? P.injected_u + (((size_t)hap_ * P.W + w) * 3 + stream_) * (size_t)P.p : nullptr;
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

struct HmmConsts { double b, mu, omm, a; };
__device__ __forceinline__ HmmConsts hmm_consts(const KgwLocusHMM h, double alpha) {
  HmmConsts c;
  c.b = h.b;
  c.mu = h.mu;
  c.omm = __dsub_rn(1.0, h.mu);
  c.a = __dmul_rn(__dsub_rn(1.0, h.b), alpha);
  return c;
}

// PAD = false asserts K == L*S (no padded states: the per-state guards compile away)
template <int L, int S, int B, bool PAD, int MINB, int D, bool SLIM = false>
__global__ void __launch_bounds__(32 * KGW_WPC, MINB) sampler_kernel(const __grid_constant__ KgwParams P) {
  using C = Cfg<L, S, B, D, SLIM>;
  static_assert(D >= 2 && D <= 8, "ring depth");
  constexpr int G = C::G;
  constexpr int SPR = C::SPR, SPW = C::SPW, NQ = C::NQ, NP = C::NP, CS = C::CS, NC = C::NC, ROW = C::ROW, TILE = C::TILE;
  constexpr int NCT = L * NC;  // chunks per haplotype
  constexpr int CSW = C::CSW;  // partial sums per locus and warp
  auto chunk_total = [](const double (&cs)[NC]) {   // balanced tree over the chunk sums of a lane
    double t[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) t[c] = cs[c];
#pragma unroll
    for (int w = NC / 2; w > 0; w >>= 1)
#pragma unroll
      for (int c = 0; c < w; c++) t[c] = t[2 * c] + t[2 * c + 1];
    return t[0];
  };
  static_assert(RESCALE % B == 0 && 32 % B == 0, "blocks must not straddle words or rescale points");
  const int lane = threadIdx.x & 31;
  const int g = lane % G;
  const int lig = lane / G;
  const int k0 = lig * S;
  const int K = P.K;
  const int ns = PAD ? min(S, max(0, K - k0)) : S;  // states owned by this lane
  // warps are independent (no CTA barrier anywhere); the host launches CTAs of 1 or KGW_WPC warps
  const int wpc = blockDim.x >> 5;
  const int slot = blockIdx.x * wpc + (threadIdx.x >> 5);
  const int nslots = gridDim.x * wpc;

  extern __shared__ __align__(16) unsigned char smem_all[];
  unsigned char *smem_raw = smem_all + (size_t)(threadIdx.x >> 5) * C::warp_bytes();
  double *rowbuf = reinterpret_cast<double *>(smem_raw);                         // [2][ROW]
  uint32_t *mwt = reinterpret_cast<uint32_t *>(smem_raw + C::rows_bytes);        // [TILE]
  double *ring = reinterpret_cast<double *>(smem_raw + C::rows_bytes + C::tiles_bytes);     // [D][2][32][2] (sweep 2)
  double *blk_se = reinterpret_cast<double *>(smem_raw + C::rows_bytes + C::tiles_bytes + C::ring_bytes);   // [2][B][G]
  KgwLocusHMM *blk_h = reinterpret_cast<KgwLocusHMM *>(blk_se + 2 * B * G);       // [2][B]
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw + C::rows_bytes + C::tiles_bytes + C::ring_bytes + C::blk_bytes);
  constexpr int ZB = SLIM ? 1 : 2;   // buffers of the latent-path tiles in sweep 3
  uint8_t *ztile3 = SLIM ? reinterpret_cast<uint8_t *>(blk_se) + 128 : reinterpret_cast<uint8_t *>(rowbuf + ROW);   // [ZB][32][G] (sweep 3)
  uint8_t *zktile3 = ztile3 + ZB * 32 * G;                                       // [ZB][32][G]
  KgwLocusKO *kot = SLIM ? reinterpret_cast<KgwLocusKO *>(ring) : reinterpret_cast<KgwLocusKO *>(zktile3 + 2 * 32 * G);   // [32] (sweep 3)
  uint32_t *z1slot = reinterpret_cast<uint32_t *>(blk_se);                       // [32]       (sweep 3)
  static_assert(C::blk_bytes >= 128, "z1 slots");
  // one mbarrier per staged buffer (one arrival: the elected lane's expect_tx); `ph` holds the phase parity every
  // lane expects next on each of them.  Every copy that is issued is waited for exactly once.
  unsigned ph = 0;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < KGW_NBARS; i++) mbar_init(bars + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_proxy();
  __syncwarp();
  // barrier between lanes that share a shared-memory row of one haplotype; with one lane per haplotype every lane reads
  // and writes its own entries only and the compiler is free to overlap the draw with the partition-function passes
  auto sync_lanes = [&]() { if (L > 1) __syncwarp(); };
  auto wait_bar = [&](int i) {
    mbar_wait(bars + i, (ph >> i) & 1u);
    ph ^= 1u << i;
  };

  uint32_t *mw = P.mwbuf + (size_t)slot * P.mw_stride;   // [stride_w][TILE]
  double *ck = P.ckpt + (size_t)slot * P.ckpt_stride;    // [seg_loci/B][ROW]
  double *sgk = P.segck + (size_t)slot * P.segck_stride; // [nseg][ROW]
  double *csb = P.csbuf + (size_t)slot * P.cs_stride;    // [seg_loci][128]
  double *seb = P.sebuf + (size_t)slot * P.se_stride;    // [seg_loci][G]
  uint8_t *zb = P.zbuf + (size_t)slot * P.z_stride;      // [p32][G]
  uint8_t *zkb = P.zkbuf + (size_t)slot * P.z_stride;

  // 16-byte units owned by this lane: pair q of a row at rowoff + q*G, quad q of a tile at mwoff + q*G
  const int rowoff = lig * NP * G + g;
  const int mwoff = lig * NQ * G + g;
  // dynamic addressing of state k of this lane's haplotype (slot index = k + (k/S) * padding of a lane)
  auto lane_of = [&](int k) { return L == 1 ? 0 : (L == 2 ? (int)(k >= S) : k / S); };
  auto row_idx = [&](int k) {
    const int kl = (SPR == S) ? k : k + lane_of(k) * (SPR - S);
    return ((kl >> 1) * G + g) * 2 + (kl & 1);
  };
  auto mw_idx = [&](int k) {
    const int kl = (SPW == S) ? k : k + lane_of(k) * (SPW - S);
    return ((kl >> 2) * G + g) * 4 + (kl & 3);
  };

  const int PS = P.seg_loci;

  for (int task = slot; task < P.n_tasks; task += nslots) {
    const int tidx = task * G + g;
    const bool active = tidx < P.n_haps;
    const int tt = active ? tidx : P.n_haps - 1;
    const int hap = __ldg(P.haps + tt);
    const size_t hoff = (size_t)hap * P.stride_w;

    KGW_PH_DECL
    for (int ip = 0; ip < P.n_passes; ip++) {
      const KgwPass ps = P.passes[ip];
      const int js = ps.js, je = ps.je;
      const int wlo = js >> 5, whi = (je - 1) >> 5;
      const int32_t *ref = P.ref_local + ((size_t)hap * P.W + ps.w) * K;
      __syncwarp();

      // ------------------------------------------------------------------ gather: match words -> scratch
      for (int wq = wlo >> 2; wq <= (whi >> 2); wq++) {
        const uint4 hw4 = __ldg(reinterpret_cast<const uint4 *>(P.Hw + hoff) + wq);
#pragma unroll
        for (int q = 0; q < NQ; q++) {
          uint4 v[4];
          bool ok[4];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int s = 4 * q + i;
            ok[i] = (s < S) && (!PAD || s < ns);
            const int id = ok[i] ? __ldg(ref + k0 + s) : hap;
            v[i] = __ldg(reinterpret_cast<const uint4 *>(P.Hw + (size_t)id * P.stride_w) + wq);
          }
          uint4 o0, o1, o2, o3;  // one per word of the quad; components = the four states
          o0.x = ok[0] ? ~(hw4.x ^ v[0].x) : 0u; o0.y = ok[1] ? ~(hw4.x ^ v[1].x) : 0u;
          o0.z = ok[2] ? ~(hw4.x ^ v[2].x) : 0u; o0.w = ok[3] ? ~(hw4.x ^ v[3].x) : 0u;
          o1.x = ok[0] ? ~(hw4.y ^ v[0].y) : 0u; o1.y = ok[1] ? ~(hw4.y ^ v[1].y) : 0u;
          o1.z = ok[2] ? ~(hw4.y ^ v[2].y) : 0u; o1.w = ok[3] ? ~(hw4.y ^ v[3].y) : 0u;
          o2.x = ok[0] ? ~(hw4.z ^ v[0].z) : 0u; o2.y = ok[1] ? ~(hw4.z ^ v[1].z) : 0u;
          o2.z = ok[2] ? ~(hw4.z ^ v[2].z) : 0u; o2.w = ok[3] ? ~(hw4.z ^ v[3].z) : 0u;
          o3.x = ok[0] ? ~(hw4.w ^ v[0].w) : 0u; o3.y = ok[1] ? ~(hw4.w ^ v[1].w) : 0u;
          o3.z = ok[2] ? ~(hw4.w ^ v[2].w) : 0u; o3.w = ok[3] ? ~(hw4.w ^ v[3].w) : 0u;
          uint4 *dst = reinterpret_cast<uint4 *>(mw + (size_t)(4 * wq) * TILE) + mwoff + q * G;
          st_cg(dst, o0);
          st_cg(dst + TILE / 4, o1);
          st_cg(dst + 2 * (TILE / 4), o2);
          st_cg(dst + 3 * (TILE / 4), o3);
        }
      }
      __syncwarp();

      KGW_PH(0);   // gather
      // match-word tile of word `w` -> shared memory: one bulk copy on barrier `bar` (elected lane; the caller adds
      // the tile's bytes to that barrier's expect_tx)
      auto bulk_tile = [&](int w, int bar) { bulk_g2s(mwt, mw + (size_t)w * TILE, TILE * 4, bars + bar); };
      // this lane's match words of the staged tile -> registers
      auto load_mw = [&](uint32_t (&mwr)[S]) {
        const uint4 *src = reinterpret_cast<const uint4 *>(mwt) + mwoff;
#pragma unroll
        for (int q = 0; q < NQ; q++) {
          const uint4 t = src[q * G];
          if (4 * q + 0 < S) mwr[4 * q + 0] = t.x;
          if (4 * q + 1 < S) mwr[(4 * q + 1 < S) ? 4 * q + 1 : 0] = t.y;
          if (4 * q + 2 < S) mwr[(4 * q + 2 < S) ? 4 * q + 2 : 0] = t.z;
          if (4 * q + 3 < S) mwr[(4 * q + 3 < S) ? 4 * q + 3 : 0] = t.w;
        }
      };
      // a value that depends on every 16-byte load of load_mw
      auto mw_any = [&](const uint32_t (&mwr)[S]) {
        unsigned x = 0;
#pragma unroll
        for (int q = 0; q < NQ; q++) x |= mwr[4 * q];
        return x;
      };
      // message row registers -> global row (this lane's pairs; padded slots are written as zero)
      auto store_row = [&](const double (&beta)[S], double *row) {
        double2 *dst = reinterpret_cast<double2 *>(row) + rowoff;
#pragma unroll
        for (int q = 0; q < NP; q++) st_cg(dst + q * G, beta[2 * q], (2 * q + 1 < S) ? beta[(2 * q + 1 < S) ? 2 * q + 1 : 0] : 0.0);
      };

      // backward recursion from locus `hi_m` (messages of hi_m in `beta`) down to `lo_m`; FULLSTORE:
      // rows every B loci + partial sums + totals (sweep 1, scratch relative to seg0); otherwise only
      // the rows at segment tops (sweep 0).
      auto backward = [&](double (&beta)[S], int lo_m, int hi_m, int seg0, bool fullstore) {
        uint32_t mwr[S];
        int wi = hi_m >> 5;
        const int wl = lo_m >> 5;
        // match-word tile and table rows of the 32 loci of word w -> shared memory (table rows: hmt[w & 1]; the row
        // buffers are idle during sweeps 0 and 1; the table is padded to whole words)
        KgwLocusHMM *hmt = reinterpret_cast<KgwLocusHMM *>(rowbuf);   // [2][32]
        auto stage_word = [&](int w) {
          if (lane == 0) {
            mbar_expect_tx(bars + BAR_S1, TILE * 4 + 32 * (unsigned)sizeof(KgwLocusHMM));
            bulk_tile(w, BAR_S1);
            bulk_g2s(hmt + (w & 1) * 32, P.hmm + (w << 5), 32 * (unsigned)sizeof(KgwLocusHMM), bars + BAR_S1);
          }
        };
        fence_async_proxy();   // the scratch and the row buffers were last written through the generic proxy
        __syncwarp();
        stage_word(wi);
        wait_bar(BAR_S1);
        load_mw(mwr);
        syncwarp_after_loads(mw_any(mwr));   // every lane HOLDS its quads in registers: the tile buffer is free
        if (wi > wl) stage_word(wi - 1);
        for (int m = hi_m; m >= lo_m; m--) {
          if ((m >> 5) != wi) {
            wi = m >> 5;
            wait_bar(BAR_S1);
            load_mw(mwr);
            syncwarp_after_loads(mw_any(mwr));
            if (wi > wl) stage_word(wi - 1);
          }
          const HmmConsts h = hmm_consts(hmt[(wi & 1) * 32 + (m & 31)], P.alpha);
          const uint32_t bm = 1u << (m & 31);
          // emission-weighted messages e = beta * E (knockoffs.cpp:1052-1060 without the constant factor a)
          double cs[NC];
#pragma unroll
          for (int c = 0; c < NC; c++) cs[c] = 0.0;
#pragma unroll
          for (int s = 0; s < S; s++) {
            double e = beta[s] * h.mu;
            if (mwr[s] & bm) e = beta[s] * h.omm;
            if (PAD && s >= ns) e = 0.0;
            beta[s] = e;
            cs[s / CS] += e;
          }
          const double pl = chunk_total(cs);
          const double Se = group_sum<L>(pl);
          if (fullstore) {
            double2 *d = reinterpret_cast<double2 *>(csb + (size_t)(m - seg0) * CSW);
#pragma unroll
            for (int h = 0; h < NC / 2; h++) st_cg(d + h * 32 + lane, cs[2 * h], cs[2 * h + 1]);
            if (lig == 0) st_cg(seb + (size_t)(m - seg0) * G + g, Se);
          }
          if (m == lo_m) break;
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
            if (fullstore) store_row(beta, ck + (size_t)((j - seg0) / B) * ROW);
            else if ((j % PS) == PS - 1) store_row(beta, sgk + (size_t)(j / PS) * ROW);
          }
        }
      };

      const int seg_first = js / PS, seg_last = (je - 1) / PS;
      // ------------------------------------------------------------------ sweep 0: segment-top rows
      if (seg_last > seg_first && !(P.skip_sweeps & 1)) {
        double beta[S];
#pragma unroll
        for (int s = 0; s < S; s++) beta[s] = (!PAD || s < ns) ? 1.0 / K : 0.0;  // knockoffs.cpp:1046
        backward(beta, seg_first * PS + PS - 1, je - 1, 0, false);
        __syncwarp();
      }

      UniformSrc usz;
      usz.init(P, hap, STREAM_Z, ps.w);
      int zprev = 0;

      for (int sg = seg_first; sg <= seg_last; sg++) {
        const int seg0 = sg * PS;                 // scratch origin of this segment
        const int slo = max(seg0, js), shi = min(seg0 + PS, je);   // [slo, shi)
        __syncwarp();
        // ---------------------------------------------------------------- sweep 1: backward + stores
        if (!(P.skip_sweeps & 1)) {
          double beta[S];
          if (sg == seg_last) {
#pragma unroll
            for (int s = 0; s < S; s++) beta[s] = (!PAD || s < ns) ? 1.0 / K : 0.0;  // knockoffs.cpp:1046
          } else {
            const double2 *src = reinterpret_cast<const double2 *>(sgk + (size_t)sg * ROW) + rowoff;
#pragma unroll
            for (int q = 0; q < NP; q++) {
              const double2 t = ld_cg(src + q * G);
              beta[2 * q] = t.x;
              if (2 * q + 1 < S) beta[(2 * q + 1 < S) ? 2 * q + 1 : 0] = t.y;
            }
          }
          store_row(beta, ck + (size_t)((shi - 1 - seg0) / B) * ROW);
          backward(beta, slo, shi - 1, seg0, true);
        }
        __syncwarp();
        KGW_PH(1);   // sweep 1

        // debug: normalised backward messages at the block-top loci of the last pass
        if (P.dbg_beta != nullptr && ip == P.n_passes - 1) {
          for (int n = slo / B; n <= (shi - 1) / B; n++) {
            double v[S];
            double acc = 0.0;
            const double2 *src = reinterpret_cast<const double2 *>(ck + (size_t)(n - seg0 / B) * ROW) + rowoff;
#pragma unroll
            for (int q = 0; q < NP; q++) {
              const double2 t = ld_cg(src + q * G);
              v[2 * q] = t.x;
              if (2 * q + 1 < S) v[(2 * q + 1 < S) ? 2 * q + 1 : 0] = t.y;
            }
#pragma unroll
            for (int s = 0; s < S; s++) acc += v[s];
            const double tot = group_sum<L>(acc);
            if (active && n < P.dbg_nblk) {
#pragma unroll
              for (int s = 0; s < S; s++)
                if (!PAD || s < ns) P.dbg_beta[((size_t)tt * P.dbg_nblk + n) * K + k0 + s] = v[s] / tot;
            }
          }
        }

        // ---------------------------------------------------------------- sweep 2: latent path z
        // No register-destination global load sits in this loop: everything it reads from HBM arrives by
        // cp.async (message rows a block ahead, block constants a block ahead, partial sums a locus ahead),
        // so that no scoreboard wait of the draw's arithmetic is tied to a load in flight.
        if (!(P.skip_sweeps & 2)) {
          int wi = slo >> 5;
          // row, locus totals and table rows of block n -> buffers n & 1, one barrier per buffer pair
          // (slim layout: the row has one buffer and its own barrier; it is fetched when its block starts)
          auto stage_block = [&](int n) {
            if (lane == 0) {
              unsigned long long *bar = bars + BAR_BLK + (n & 1);
              mbar_expect_tx(bar, (SLIM ? 0 : ROW * 8) + B * G * 8 + B * (unsigned)sizeof(KgwLocusHMM));
              if (!SLIM) bulk_g2s(rowbuf + (size_t)(n & 1) * ROW, ck + (size_t)(n - seg0 / B) * ROW, ROW * 8, bar);
              bulk_g2s(blk_se + (n & 1) * B * G, seb + (size_t)(n * B - seg0) * G, B * G * 8, bar);
              bulk_g2s(blk_h + (n & 1) * B, P.hmm + n * B, B * (unsigned)sizeof(KgwLocusHMM), bar);
            }
          };
          auto stage_row = [&](int n) {
            if (lane == 0) {
              mbar_expect_tx(bars + BAR_ROW, ROW * 8);
              bulk_g2s(rowbuf, ck + (size_t)(n - seg0 / B) * ROW, ROW * 8, bars + BAR_ROW);
            }
          };
          // partial sums of locus j -> ring slot j % D
          auto stage_locus = [&](int j) {
            if (lane == 0) {
              unsigned long long *bar = bars + BAR_CS + (j % D);
              mbar_expect_tx(bar, CSW * 8);
              bulk_g2s(ring + (j % D) * CSW, csb + (size_t)(j - seg0) * CSW, CSW * 8, bar);
            }
          };
          auto stage_tile2 = [&](int w) {
            if (lane == 0) {
              mbar_expect_tx(bars + BAR_TILE2, TILE * 4);
              bulk_tile(w, BAR_TILE2);
            }
          };
          fence_async_proxy();   // sweep 1 wrote the scratch through the generic proxy
          __syncwarp();
          stage_tile2(wi);
          bool tile_pending = true;
          stage_block(slo / B);
#pragma unroll
          for (int d = 0; d < D; d++)
            if (slo + d < shi) stage_locus(slo + d);
          uint32_t hw = __ldg(P.Hw + hoff + wi);
          const int n_last = (shi - 1) / B;
          for (int n = slo / B; n <= n_last; n++) {
            const int nb = n * B;
            const int lo = max(nb, slo);
            const int top = min(nb + B, shi) - 1;
            __syncwarp();   // every lane is done with block n - 1 (its row buffer is refilled below) and with the tile
            if ((lo >> 5) != wi) {
              wi = lo >> 5;
              hw = __ldg(P.Hw + hoff + wi);
              stage_tile2(wi);
              tile_pending = true;
            }
            if (SLIM) stage_row(n);   // every lane is past block n - 1 (the barrier above)
            if (n < n_last) stage_block(n + 1);
            wait_bar(BAR_BLK + (n & 1));
            if (SLIM) wait_bar(BAR_ROW);
            if (tile_pending) { wait_bar(BAR_TILE2); tile_pending = false; }
            const double *bse = blk_se + (n & 1) * B * G;
            const KgwLocusHMM *bh = blk_h + (n & 1) * B;
            // constants of the recurrence beta_{m-1}(k) = (b_m E_m(k)) beta_m(k) + a_m Se_m for the loci of the
            // block (loci above `top` get the identity step); the staged table rows and totals stay in shared
            // memory until the block's last locus
            double tbo[B], tbm[B], tcm[B];
#pragma unroll
            for (int mm = 1; mm < B; mm++) {
              const int m = nb + mm;
              if (m >= lo && m <= top) {
                const HmmConsts h = hmm_consts(bh[mm], P.alpha);
                tbo[mm] = h.b * h.omm; tbm[mm] = h.b * h.mu;
                tcm[mm] = h.a * bse[mm * G + g];
              } else {
                tbo[mm] = 1.0; tbm[mm] = 1.0; tcm[mm] = 0.0;
              }
            }
            const double *row = rowbuf + (size_t)(SLIM ? 0 : (n & 1)) * ROW;
            const uint32_t *tile = mwt;
            const int bit0 = nb & 31;

#pragma unroll
            for (int jj = 0; jj < B; jj++) {
              const int j = nb + jj;
              if (j < lo || j > top) continue;
              wait_bar(BAR_CS + (j % D));
              // partial sums of this locus: the L lanes of the haplotype, NC chunks each
              double csv[NCT];
              {
                const double2 *slot = reinterpret_cast<const double2 *>(ring + (j % D) * CSW);
#pragma unroll
                for (int lg = 0; lg < L; lg++) {
#pragma unroll
                  for (int h = 0; h < NC / 2; h++) {
                    const double2 t = slot[h * 32 + lg * G + g];
                    csv[lg * NC + 2 * h] = t.x; csv[lg * NC + 2 * h + 1] = t.y;
                  }
                }
              }
              const double u = usz.get(j);
              const HmmConsts hj = hmm_consts(bh[jj], P.alpha);
              const double sej = bse[jj * G + g];
              KGW_PH(2);   // sweep 2: waits + staging reads
              const int bit = bit0 + jj;
              const bool hbit = (hw >> bit) & 1u;
              const double aj = hj.a, bj = hj.b, ommj = hj.omm, muj = hj.mu;
              const double mupj = __dsub_rn(1.0, ommj);              // knockoffs.cpp:1251-1256 quirk
              const double aomm = __dmul_rn(aj, ommj);
              const double amis = hbit ? __dmul_rn(aj, muj) : __dmul_rn(aj, mupj);
              // message of locus j at one state: staged top row, then the steps top -> j
              auto chain = [&](double v, uint32_t mword) {
#pragma unroll
                for (int mm = B - 1; mm > jj; mm--)
                  v = ((mword >> (bit0 + mm)) & 1u) ? fma(v, tbo[mm], tcm[mm]) : fma(v, tbm[mm], tcm[mm]);
                return v;
              };
              // the one state whose transition weight carries b_j: k == z_{j-1}
              double X = 0.0;
              if (j > js) {
                const uint32_t mwz = tile[mw_idx(zprev)];
                const double bz = chain(row[row_idx(zprev)], mwz);
                X = bj * (bz * (((mwz >> bit) & 1u) ? ommj : (hbit ? muj : mupj)));
              }
              const double R = u * fma(aj, sej, X);
              const double tA = R, tB = R - X;
              KGW_PH(3);   // sweep 2: X term
              // chunk level: #chunks whose end prefix is <= threshold
              int CA = 0, CB = 0;
              {
                double T = 0.0;
#pragma unroll
                for (int c = 0; c < NCT; c++) {
                  T += csv[c];
                  const double PT = aj * T;
                  count_le(CA, PT, tA);
                  count_le(CB, PT, tB);
                }
              }
              const int lgz = lane_of(zprev);
              const int cz = lgz * NC + (zprev - lgz * S) / CS;
              const int csel = (CA > cz) ? CB : CA;
              const int cst = min(csel, NCT - 1);
              // `cst` depends on every partial sum of the slot: all lanes have their values, refill it D loci ahead
              syncwarp_after_loads((unsigned)cst);
              if (j + D < shi) stage_locus(j + D);
              double Tb = 0.0;
#pragma unroll
              for (int c = 0; c < NCT - 1; c++) add_if_lt(Tb, c, cst, csv[c]);
              const int lgs = cst / NC, ci = cst - lgs * NC;
              const int nss = PAD ? min(S, max(0, K - lgs * S)) : S;
              const int n_in = min(CS, nss - ci * CS);       // valid states of the chunk (may be <= 0)
              const int kb = lgs * S + ci * CS;
              KGW_PH(4);   // sweep 2: chunk level
              const int klr = lgs * SPR + ci * CS, klw = lgs * SPW + ci * CS;
              double run = aj * Tb;
              int cA = 0, cB = 0;
              {
                // every lane walks all CS slots of its chunk (invalid ones contribute zero weight and are not
                // counted; their reads are redirected to the chunk's first units): no divergent branch
                const double2 *r2 = reinterpret_cast<const double2 *>(row) + (klr >> 1) * G + g;
                const uint4 *t4 = reinterpret_cast<const uint4 *>(tile) + (klw >> 2) * G + g;
#pragma unroll
                for (int qq = 0; qq < CS / 4; qq++) {
                  const int qa = (4 * qq < n_in) ? qq : 0;
                  const int pb = (4 * qq + 2 < n_in) ? 2 * qq + 1 : 0;
                  const uint4 mq = t4[qa * G];
                  const double2 va = r2[(2 * qa) * G];
                  const double2 vb = r2[pb * G];
                  const double v4[4] = {va.x, va.y, vb.x, vb.y};
                  const uint32_t m4[4] = {mq.x, mq.y, mq.z, mq.w};
#pragma unroll
                  for (int i = 0; i < 4; i++) {
                    const bool ok = 4 * qq + i < n_in;
                    const double bv = chain(v4[i], m4[i]);
                    const double e = bv * (((m4[i] >> bit) & 1u) ? aomm : amis);
                    run += ok ? e : 0.0;
                    count_le_if(cA, run, tA, ok);
                    count_le_if(cB, run, tB, ok);
                  }
                }
              }
              int z;
              if (CA > cz) {
                z = max(kb + cB, zprev);
              } else {
                const int k1 = kb + cA;
                const int k2 = (CB == CA) ? kb + cB : 0;
                z = (k1 < zprev) ? k1 : max(k2, zprev);
              }
              z = min(z, K - 1);
              zprev = z;
              if (lig == 0) zb[(size_t)j * G + g] = (uint8_t)z;
              KGW_PH(5);   // sweep 2: refinement
            }
          }
        }
        __syncwarp();
      }  // segments

      // ------------------------------------------------------------------ sweep 3: knockoff path + alleles
      if (!(P.skip_sweeps & 4)) {
        double Nv[S];
#pragma unroll
        for (int s = 0; s < S; s++) Nv[s] = (!PAD || s < ns) ? 1.0 : 0.0;  // knockoffs.cpp:1293-1294 (N == N_old between groups)
        double sumN = (double)K;                                          // sum_k N_k of this haplotype
        double *rrow = rowbuf;                                            // [ROW] reciprocals 1/N_old, patched per group
        double2 *rrow2 = reinterpret_cast<double2 *>(rrow);
        UniformSrc us, uh;
        us.init(P, hap, STREAM_ZK, ps.w);
        uh.init(P, hap, STREAM_HK, ps.w);
        int z0 = 0, z0k = 0, zkprev = 0;
        if (js > 0) { z0 = zb[(size_t)(js - 1) * G + g]; z0k = zkb[(size_t)(js - 1) * G + g]; }
        const KgwLocusKO *ko_tab = P.ko + ps.ko_off - js;
        // As in sweep 2, everything the per-locus loop reads from global memory is staged in shared memory: bulk
        // copies per word (latent-path tiles a word ahead; match-word tile, knockoff-table rows and emission rows of
        // the word -- the tables are padded to whole words on either side), and the byte z1 of the next group by a
        // per-lane cp.async one group ahead.
        constexpr bool EMIT_IN_RING = !SLIM && C::ring_bytes >= 32 * sizeof(KgwEmit);
        static_assert(sizeof(KgwEmit) == sizeof(KgwLocusKO), "the two per-locus tables may share a shared-memory tile");
        KgwEmit *emt = EMIT_IN_RING ? reinterpret_cast<KgwEmit *>(ring) : reinterpret_cast<KgwEmit *>(kot);   // [32], indexed by bit
        auto stage_ztiles = [&](int w) {
          if (lane == 0) {
            mbar_expect_tx(bars + BAR_ZT, 2 * 32 * G);
            bulk_g2s(ztile3 + (size_t)(w & (ZB - 1)) * 32 * G, zb + (size_t)(w << 5) * G, 32 * G, bars + BAR_ZT);
            bulk_g2s(zktile3 + (size_t)(w & (ZB - 1)) * 32 * G, zkb + (size_t)(w << 5) * G, 32 * G, bars + BAR_ZT);
          }
        };
        int z1 = 0, z1_byte = 0;
        auto stage_z1 = [&](int locus) {   // the aligned 4 bytes holding z[locus] of this lane's haplotype
          const size_t off = (size_t)locus * G + g;
          z1_byte = (int)(off & 3);
          cp_async4(z1slot + lane, zb + (off & ~(size_t)3));
        };
        fence_async_proxy();   // sweep 2 wrote the latent path, earlier sweeps the row buffers, through the generic proxy
        __syncwarp();
        {
          const int zi = ko_tab[js].z1_idx;
          if (zi >= 0) stage_z1(zi);
        }
        if (!SLIM) stage_ztiles(wlo);
        for (int wi = wlo; wi <= whi; wi++) {
          const int lo = max(wi << 5, js), hi = min((wi << 5) + 32, je);
          __syncwarp();   // the previous word is finished in every lane (tile reads of its allele draws)
          if (SLIM) stage_ztiles(wi);   // one pair of latent-path tiles: fetched with the word
          if (lane == 0) {
            // match words (for the allele draws at the end of the word), knockoff-table rows and, when the ring can
            // hold them, the emission rows of the word's 32 loci
            mbar_expect_tx(bars + BAR_W3, TILE * 4 + 32 * (unsigned)sizeof(KgwLocusKO) + (EMIT_IN_RING ? 32 * (unsigned)sizeof(KgwEmit) : 0u));
            bulk_tile(wi, BAR_W3);
            bulk_g2s(kot, ko_tab + (wi << 5), 32 * (unsigned)sizeof(KgwLocusKO), bars + BAR_W3);
            if (EMIT_IN_RING) bulk_g2s(emt, P.emit + (wi << 5), 32 * (unsigned)sizeof(KgwEmit), bars + BAR_W3);
          }
          wait_bar(BAR_ZT);
          wait_bar(BAR_W3);
          if (!SLIM && wi < whi) stage_ztiles(wi + 1);
          uint8_t *zt = ztile3 + (size_t)(wi & (ZB - 1)) * 32 * G, *zkt = zktile3 + (size_t)(wi & (ZB - 1)) * 32 * G;
          for (int j = lo; j < hi; j++) {
            const KgwLocusKO &k = kot[j & 31];
            const int bit = j & 31;
            const int zj = zt[bit * G + g];
            const bool lastg = k.flags & KGW_KO_LASTGRP;
            const double u = us.get(j);
            int zk;
            if (k.flags & KGW_KO_FIRST) {
              cp_async_wait_all();   // z1 of this group was requested at the first locus of the previous one
              z1 = (k.z1_idx >= 0) ? (int)reinterpret_cast<const uint8_t *>(z1slot + lane)[z1_byte] : 0;
              if (k.z1_next >= 0) stage_z1(k.z1_next);
              const bool gz = k.flags & KGW_KO_GZERO;
              // pass A: reciprocals of N_old -> shared row, chunk sums
              double cs[NC];
#pragma unroll
              for (int c = 0; c < NC; c++) cs[c] = 0.0;
#pragma unroll
              for (int q = 0; q < NP; q++) {
                const int s0 = 2 * q, s1 = 2 * q + 1;
                const double r0 = (!PAD || s0 < ns) ? fast_rcp(Nv[s0]) : 0.0;
                const double r1 = (s1 < S && (!PAD || s1 < ns)) ? fast_rcp(Nv[(s1 < S) ? s1 : 0]) : 0.0;
                rrow2[rowoff + q * G] = make_double2(r0, r1);
                cs[s0 / CS] += r0;
                if (s1 < S) cs[((s1 < S) ? s1 : 0) / CS] += r1;
              }
              sync_lanes();
              KGW_PH(6);   // sweep 3: loop head + pass A
              // Special states, read by every lane of the haplotype.  k == z0 / z0k carry T_k(z0)*T_k(z0k) != tg
              // (:1350-1354) in the N update and in the draw; k == z1 carries the extra vstar term in the draw
              // only (:1413-1419), on top of the former if the states coincide.
              const bool same = (z0 == z0k);
              const double r0 = rrow[row_idx(z0)], r0k = rrow[row_idx(z0k)], rz1 = rrow[row_idx(z1)];
              const double n0 = r0 * (same ? k.f2 : k.f1), n0k = r0k * k.f1;
              const double d0 = gz ? 0.0 : n0 - r0;
              const double d0k = (gz || same) ? 0.0 : n0k - r0k;
              const double rz = (!gz && z1 == z0) ? n0 : ((!gz && z1 == z0k) ? n0k : rz1);
              const double nz = rz * k.fz1;
              const double d1 = lastg ? 0.0 : nz - rz;
              sync_lanes();   // every lane has read the unpatched row
              const int lg0 = lane_of(z0), lg0k = lane_of(z0k), lg1 = lane_of(z1);
              {
                // the owner lane patches its entries (it re-reads them in pass B) and its chunk sums
                if (!gz && lg0 == lig) rrow[row_idx(z0)] = n0;
                if (!gz && !same && lg0k == lig) rrow[row_idx(z0k)] = n0k;
                const int c0 = (!gz && lg0 == lig) ? (z0 - lg0 * S) / CS : -1;
                const int c0k = (!gz && !same && lg0k == lig) ? (z0k - lg0k * S) / CS : -1;
#pragma unroll
                for (int c = 0; c < NC; c++) {
                  add_if_eq(cs[c], c0, c, d0);
                  add_if_eq(cs[c], c0k, c, d0k);
                }
              }
              KGW_PH(7);   // sweep 3: specials
              double Sr;
              const double base0 = group_total_excl<L>(chunk_total(cs), lane, Sr);
              const double Nsum = (k.flags & KGW_KO_GSTART) ? k.sa : k.tg * Sr;  // :1345-1358
              const double c2 = k.vb * Nsum;
              // pass B: partition function update, normalised (:1360-1390).  The last group keeps N (zero
              // coefficients make the update an exact identity) but still renormalises it.  The normaliser
              // sum_k (N_k + c2 + vstg r_k) is formed from the haplotype's running sum_k N_k.
              const double c2u = lastg ? 0.0 : c2;
              const double vsu = lastg ? 0.0 : k.vstg;
              const double norm = fma(vsu, Sr, fma((double)K, c2u, sumN));
              const double rN = fast_rcp(norm);
              const double cA_ = vsu * rN, cC_ = c2u * rN;
              KGW_PH(8);   // sweep 3: reduction + scalars
#pragma unroll
              for (int q = 0; q < NP; q++) {
                const int s0 = 2 * q, s1 = 2 * q + 1;
                const double2 rr = rrow2[rowoff + q * G];
                Nv[s0] = (!PAD || s0 < ns) ? fma(cA_, rr.x, fma(Nv[s0], rN, cC_)) : 0.0;
                if (s1 < S) Nv[(s1 < S) ? s1 : 0] = (!PAD || s1 < ns) ? fma(cA_, rr.y, fma(Nv[(s1 < S) ? s1 : 0], rN, cC_)) : 0.0;
              }
              // clamp to N_min (:1386-1389; the pre-normalisation clamp of :1380 cannot fire because N only
              // grows).  Every N_k is at least cC_, so the clamp pass runs only when cC_ < N_min.
              sumN = 1.0;
              if (__any_sync(FULL, cC_ < N_MIN)) {
                double acc = 0.0;
#pragma unroll
                for (int s = 0; s < S; s++) {
                  if (!PAD || s < ns) {
                    Nv[s] = (Nv[s] < N_MIN) ? N_MIN : Nv[s];
                    acc += Nv[s];
                  }
                }
                acc = group_sum<L>(acc);
                if (cC_ < N_MIN) sumN = acc;
              }
              // the draw: weights r_k (patched), in units of vb*tg
              {
                if (!lastg && lg1 == lig) rrow[row_idx(z1)] = nz;   // after this lane's pass B reads
                const int c1 = (!lastg && lg1 == lig) ? (z1 - lg1 * S) / CS : -1;
#pragma unroll
                for (int c = 0; c < NC; c++) add_if_eq(cs[c], c1, c, d1);
              }
              sync_lanes();
              KGW_PH(9);   // sweep 3: pass B + clamp + z1 patch
              // two-level inverse-CDF search (weighted_choice, utils.cpp:690-698)
              const double St = Sr + d1;
              const double base = base0 + ((!lastg && lg1 < lig) ? d1 : 0.0);
              const double t = u * St;
              int cnt = 0;
              {
                double run = base;
#pragma unroll
                for (int c = 0; c < NC; c++) {
                  run += cs[c];
                  count_le(cnt, run, t);
                }
              }
              const int cst = min(group_sum_int<L>(cnt), NCT - 1);
              const int lgs = cst / NC, ci = cst - lgs * NC;
              double tb_own = base;
#pragma unroll
              for (int c = 0; c < NC - 1; c++) add_if_lt(tb_own, c, ci, cs[c]);
              double run = (L == 1) ? tb_own : __shfl_sync(FULL, tb_own, lgs * G + g);
              const int nss = PAD ? min(S, max(0, K - lgs * S)) : S;
              const int n_in = min(CS, nss - ci * CS);
              const int klr = lgs * SPR + ci * CS;
              int c2n = 0;
              {
                const double2 *r2 = rrow2 + (klr >> 1) * G + g;
#pragma unroll
                for (int q = 0; q < CS / 2; q++) {
                  const double2 rr = r2[((2 * q < n_in) ? q : 0) * G];
                  run += (2 * q < n_in) ? rr.x : 0.0;
                  count_le_if(c2n, run, t, 2 * q < n_in);
                  run += (2 * q + 1 < n_in) ? rr.y : 0.0;
                  count_le_if(c2n, run, t, 2 * q + 1 < n_in);
                }
              }
              zk = min(lgs * S + ci * CS + c2n, K - 1);
              sync_lanes();   // the row is rewritten at the next group
              KGW_PH(10);  // sweep 3: draw
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
            if (lig == 0) zkt[bit * G + g] = (uint8_t)zk;
          }
          fence_async_proxy();   // this word's zk tile was written through the generic proxy; a bulk copy refills it two words on
          __syncwarp();
          // flush the zk tile, then knockoff alleles of the core loci of this word (:1469-1502)
          for (int i = lane; i < 8 * G; i += 32)
            reinterpret_cast<uint32_t *>(zkb + (size_t)(wi << 5) * G)[i] = reinterpret_cast<uint32_t *>(zkt)[i];
          const uint32_t hw = __ldg(P.Hw + hoff + wi);
          const uint32_t *tile = mwt;
          unsigned bits = 0u, mask = 0u;
          const int elo = max(lo, ps.cs), ehi = min(hi, ps.ce);
          if (!EMIT_IN_RING) {
            // emission-table rows of the word -> over the knockoff-table rows (done for this word)
            if (lane == 0) {
              mbar_expect_tx(bars + BAR_EM, 32 * (unsigned)sizeof(KgwEmit));
              bulk_g2s(emt, P.emit + (wi << 5), 32 * (unsigned)sizeof(KgwEmit), bars + BAR_EM);
            }
            wait_bar(BAR_EM);
          }
          // a lane draws the alleles of whole groups of four loci (one Philox block each)
          for (int q4 = lig; q4 < 8; q4 += L) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const int bit = 4 * q4 + i;
              const int j = (wi << 5) + bit;
              if (j >= elo && j < ehi) {
                // reference bit = target bit where the match word says "equal", its complement elsewhere
                const uint32_t w1 = ~(tile[mw_idx(zt[bit * G + g])] ^ hw);
                const uint32_t w2 = ~(tile[mw_idx(zkt[bit * G + g])] ^ hw);
                const int combo = ((hw >> bit) & 1u) | (((w1 >> bit) & 1u) << 1) | (((w2 >> bit) & 1u) << 2);
                const double2 em = emt[bit].ws[combo];
                const unsigned hk = (uh.get(j) * em.y < em.x) ? 0u : 1u;  // weighted_choice with 2 weights
                bits |= hk << bit;
                mask |= 1u << bit;
              }
            }
          }
          bits = group_or<L>(bits);
          mask = group_or<L>(mask);
          KGW_PH(11);  // sweep 3: word tail (flush + alleles)
          if (lig == 0 && active && mask != 0u) {
            uint32_t *dst = P.Hk + hoff + wi;
            if (mask == 0xffffffffu) *dst = bits;
            else *dst = (*dst & ~mask) | bits;
          }
        }
        cp_async_wait_all();
      }
      __syncwarp();
    }  // passes

    __syncwarp();
    if (active && (P.dbg_z != nullptr || P.dbg_zk != nullptr)) {
      for (int j = lig; j < P.p; j += L) {
        if (P.dbg_z) P.dbg_z[(size_t)tt * P.p + j] = zb[(size_t)j * G + g];
        if (P.dbg_zk) P.dbg_zk[(size_t)tt * P.p + j] = zkb[(size_t)j * G + g];
      }
    }
    __syncwarp();
  }
}

}  // namespace kgw
