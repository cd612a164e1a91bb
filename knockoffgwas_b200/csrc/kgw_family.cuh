// knockoffgwas_b200/csrc/kgw_family.cuh
//
// sm_100a kernel for the IBD-conditioned (related) branch of snpknock2's knockoff sampler:
//   Knockoffs::generate_related_cluster   snpknock2/src/knockoffs.cpp:465-571
//   FamilyBP (loopy belief propagation)   snpknock2/src/family_bp.cpp:79-601
//   FamilyKnockoffs                       snpknock2/src/family_knockoffs.cpp:77-454
//   Knockoffs::sample_emission            snpknock2/src/knockoffs.cpp:1430-1441
//
// Mapping: one warp per IBD family (families are independent, knockoffs.cpp:440-445), families handed out
// through an atomic counter; lane l owns the HMM states k = l, l+32, ... (SF = ceil(K/32) per lane), so
// no message entry is ever touched by two lanes (the final draws and the family knockoffs, which keep their
// per-state values in registers, use lane l = states l*SF .. l*SF+SF-1 instead: one scan per draw).  The left
// BP messages of a family, double [members][p][K], the right messages at the loci where something reads them
// (see below) and the per-(member, locus, state) emission pattern bytes live in an HBM scratch slot owned by the
// warp; neighbour sets are piecewise constant along the chromosome and come from the host as intervals
// (with the junction differences of family_bp.cpp:352-378 precomputed).  Node emissions are table
// look-ups: the reference's pow(prod, 1/(n+1)) (family_bp.cpp:441-454) takes one of 2^(n+1) values,
// tabulated on the host with the reference's own operation order.
//
// Exact work skipping in run_bp (family_bp.cpp:318-419).  A message is a deterministic function of its
// inputs: the neighbouring message on its own chain and, only at a junction, messages of other members.
// Recomputing it from bit-identical inputs returns the stored bits and adds exactly 0 to the
// convergence sum.  Interval boundaries b (a member's neighbour list changes between b-1 and b) define
// "special" loci [b-2, b+1]: every junction node, every node pinned by condition_on (:236-255) and every
// consumer of a pinned or cross-chain message lies there.  A sweep recomputes (1) everything in the
// very first iteration, (2) the special loci, (3) a node whose own-chain input changed bits earlier in
// the same sweep; all other nodes are skipped.  The perturbation from a junction dies out bit for bit
// after a few hundred loci, so iterations after the first cost a few per cent of a full sweep (85 % of
// the reference's message updates are such no-ops on the synthetic pairs; the rule is audited on the CPU
// by tests/test_family_skip_rule.py: 0 violations).
//
// Right messages are not stored in plain stretches.  Nothing reads them there except the sweep itself (the left
// sweep, the junction draws and the final draws use right messages only at special loci or analytically), and
// the sweep needs the old value of a node only to measure its change: when a stretch is re-entered with a changed
// input, the OLD chain is recomputed beside the new one from the old input (bit-identical to what a stored copy
// would hold) until the two agree bit for bit.  A slot therefore holds 8 K + K bytes per (member, locus) instead
// of 16 K + K, which is what bounds the number of resident families on long chromosomes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "kgw_device.cuh"

// Two compiled classes of the kernel (template parameter BIG).  The SMALL class keeps the per-member bookkeeping of a
// sweep in registers (bit masks, 3-bit neighbour counts) and the running messages in shared memory, and looks node
// emissions up by an 8-bit match pattern: families of at most KGW_FAM_MAXN haplotypes whose nodes have at most
// KGW_FAM_MAXNB neighbours -- pairs, trios, sibships.  Everything else goes to the BIG class, which has no limit (the
// reference has none, family_bp.cpp:41-61): bookkeeping and running messages live in the slot's HBM scratch and a node
// with more than KGW_FAM_MAXNB neighbours forms its emission from the packed rows directly.
#define KGW_FAM_MAXN 8   // members per family, SMALL class
#define KGW_FAM_MAXNB 7  // neighbours of a node whose emission is a table look-up / pattern byte
#ifndef KGW_FAM_MINB
#define KGW_FAM_MINB 4   // resident CTAs of 4 warps per SM the register allocation aims at
#endif

struct KgwFamInterval {  // maximal run of loci over which neighbors[i][j] is constant (family_bp.cpp:41-61)
  int32_t j0, j1;        // inclusive
  int32_t nn;            // number of neighbours
  int32_t ndl, ndr;      // sizes of neigh_diff_left at j0 / neigh_diff_right at j1 (family_bp.cpp:365-378)
  int32_t nb_off;        // into KgwFamParams::ivlist: nn family-local indices, reference push_back order
  int32_t dl_off;        // ... ndl indices
  int32_t dr_off;        // ... ndr indices
};

struct KgwFamRange {   // special loci s0 .. s1, inclusive; sorted, disjoint
  int32_t s0, s1;
  int32_t off;         // row of locus max(0, s0 - 1) in the family's stored right messages
  int32_t pad;
};

// phases of a family (KgwFamParams::phase_mask); KOCOPY: only the copy operations of FamilyKnockoffs::run (the pair
// class draws in a kernel of its own, kgw_pairs.cuh)
enum { KGW_FPH_INIT = 1, KGW_FPH_BP = 2, KGW_FPH_DRAW = 4, KGW_FPH_KO = 8, KGW_FPH_OUT = 16, KGW_FPH_ALL = 31, KGW_FPH_KOCOPY = 32 };
enum { KGW_FOP_COPYZ = 0, KGW_FOP_DRAW = 1, KGW_FOP_COPYZK = 2 };
struct KgwFamOp {  // FamilyKnockoffs::run (family_knockoffs.cpp:155-270) flattened in execution order
  int32_t kind;    // COPYZ: Zk[i][a..b] = Z[i][a..b];  DRAW: sample_knockoff_MC(i, groups a..b);  COPYZK: Zk[i][a..b] = Zk[src][a..b]
  int32_t i, a, b, src;
  int32_t pad[3];
};

struct KgwPairOp {   // one DRAW operation of a pair (kgw_pairs.cuh)
  int32_t f;         // the pair: index into fam_list
  int32_t i;         // member
  int32_t j0, j1;    // loci grp_first[g_begin] .. grp_last[g_end]
};

struct KgwFamDesc {
  int32_t n;          // members
  int32_t mem_off;    // into members[]
  int32_t op_off, n_ops;
  int32_t out_off;    // row offset of this family in the debug outputs
  int32_t rng_off, n_rng;   // special ranges of this family
  int32_t ivs_off;    // into iv_start[]: n + 1 entries for this family
};

struct KgwFamLocus {  // per locus
  double b;      // b_j
  double a_bp;   // (1.0 - b_j) / (double)K        family_bp.cpp:458,530
  double a_al;   // (1.0 - b_j) * alpha            family_bp.cpp:243,249; family_knockoffs.cpp
  double mu;     // mut_rates[j]
  double w;      // K * a_bp + b  (= 1 up to rounding): sum over k of the unnormalised transition weights
  double r1;     // a_bp / w
};

struct KgwFamParams {
  int32_t N, p, K, n_families, n_groups, stride_w, n_max, uniform_mu;
  const uint32_t *Hw;
  const int32_t *ref_global;   // [N][K]
  const KgwFamDesc *fams;
  const int32_t *fam_list;     // [n_families] descriptors of this launch's class
  const int32_t *members;      // absolute ids
  const int32_t *iv_start;     // per (family, member): first interval; one sentinel per family
  const KgwFamInterval *ivs;
  const int32_t *ivlist;       // neighbour / junction-difference lists of the intervals
  const KgwFamRange *ranges;
  const KgwFamOp *ops;
  const KgwFamLocus *loc;      // [p]
  const KgwLocusKO *ko;        // [p] single-window table (vstar / vbar / flags of every locus)
  const int32_t *groups;       // [p]
  const int32_t *grp_first;    // [n_groups]
  const int32_t *grp_last;     // [n_groups]
  const double *emtab;         // [8][256] node emission by (neighbours, match pattern); uniform mu only
  uint32_t *Hk;                // [N][stride_w]
  double *msg;                 // [slots][n_max][p][K]      messages_left
  double *mrc;                 // [slots][n_max][rtot][K]   messages_right at the stored loci (special ranges + one locus before)
  int32_t rtot, W;             // stored loci per member (maximum over the families); windows of the problem (injected_u layout)
  uint8_t *pat;                // [slots][n_max][p][K]
  uint8_t *bytes;              // [slots][2][n_max][p]   Z (255 = node still active), Zk
  double *pmg;                 // BIG class: [slots][2][n_max][KP] running messages of a sweep (SMALL: shared memory)
  int32_t *flg;                // BIG class: [slots][5][n_max] per-member bookkeeping of a sweep (SMALL: registers)
  int32_t *work;               // next family to hand out (zeroed before every launch)
  int32_t phase_mask;          // KGW_FPH_* phases this launch runs
  int32_t slot_by_family;      // scratch slot = family's index in the batch (phases split over launches) instead of the warp's
  int32_t f_base;              // first family (index into fam_list) of this launch's batch; n_families = one past its last
  int32_t pad_;
  unsigned long long seed;
  const double *injected_u;    // [N][W][3][p] or null (the related branch reads window block 0)
  int32_t *dbg_z, *dbg_zk;     // [members total][p] or null
  int32_t *status;             // 0 ok; -10 missing knockoff variants; -11 inconsistency on Zk
  long long *phase_cycles;     // -DKGW_PHASES: [8] summed lane-0 clock64 deltas per phase, else null
};

#ifdef KGW_PHASES
#define KGW_FAM_PHASE(idx) do { if (P.phase_cycles) { const long long t_ = clock64(); if (lane == 0) atomicAdd((unsigned long long *)&P.phase_cycles[idx], (unsigned long long)(t_ - fam_t0)); fam_t0 = clock64(); } } while (0)
#else
#define KGW_FAM_PHASE(idx) do { } while (0)
#endif

namespace kgw {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ void warp_sum2(double &a, double &b) {   // two sums in one shuffle tree
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ta = __shfl_xor_sync(FULL, a, o);
    const double tb = __shfl_xor_sync(FULL, b, o);
    a += ta;
    b += tb;
  }
}
__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// L2 prefetch of the `bytes` starting at `ptr` (one 128-byte line per lane)
__device__ __forceinline__ void fam_prefetch(const void *ptr, int bytes, int lane) {
  if (lane * 128 < bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(ptr) + lane * 128));
}

__device__ __forceinline__ double fam_uniform(const KgwFamParams &P, int hap, int stream, int j) {
  if (P.injected_u) return P.injected_u[((size_t)hap * P.W * 3 + stream) * (size_t)P.p + j];   // window block 0 of [N][W][3][p]
  const kgw_u32x4 r = kgw_philox_block(P.seed, (uint32_t)hap, (uint32_t)stream, (uint32_t)(j >> 2));
  const int q = j & 3;
  return kgw_word_to_uniform(q == 0 ? r.x : (q == 1 ? r.y : (q == 2 ? r.z : r.w)));
}

// weighted_choice (utils.cpp:690-698) over states k = s*32 + lane: first k with u*S < cumulative weight.
// The cumulative sums are non-decreasing in k, so the count of prefixes <= R is one ballot per s.
template <int SF>
__device__ __forceinline__ int warp_choice(const double (&w)[SF], double u, int K, int lane) {
  double tot = 0.0;
  double inc[SF];
#pragma unroll
  for (int s = 0; s < SF; s++) inc[s] = warp_incl_scan(w[s], lane);
#pragma unroll
  for (int s = 0; s < SF; s++) {
    const double c = inc[s];
    inc[s] = tot + c;
    tot += __shfl_sync(FULL, c, 31);
  }
  const double R = u * tot;
  int cnt = 0;
#pragma unroll
  for (int s = 0; s < SF; s++) cnt += __popc(__ballot_sync(FULL, s * 32 + lane < K && inc[s] <= R));
  return min(cnt, K - 1);
}

// The same draw when lane l owns the CONTIGUOUS states l*SF .. l*SF+SF-1 (the final draws and the family knockoffs
// keep their per-state values in registers only, so they are free to choose the ownership): one scan over the
// lane totals instead of SF scans, then the state inside the lane that holds u*S.
template <int SF>
__device__ __forceinline__ int warp_choice_contig(const double (&w)[SF], double u, int K, int lane) {
  double T = 0.0;
#pragma unroll
  for (int s = 0; s < SF; s++) T += w[s];
  const double inc = warp_incl_scan(T, lane);
  const double R = u * __shfl_sync(FULL, inc, 31);
  const int L = __popc(__ballot_sync(FULL, inc <= R));   // first lane whose inclusive sum exceeds R (32: none)
  const double r = R - (inc - T);
  double acc = 0.0;
  int c = 0;
#pragma unroll
  for (int s = 0; s < SF; s++) {
    acc += w[s];
    c += (acc <= r) ? 1 : 0;
  }
  const int k = __shfl_sync(FULL, lane * SF + min(c, SF - 1), min(L, 31));
  return (L >= 32) ? K - 1 : min(k, K - 1);
}

// SF: states per lane; UMU: mu is the same at every locus (init_hmm, knockoffs.cpp:1520) and node emissions
// come from the shared-memory table, else they are formed per state with pow() as the reference does
// P.phase_mask selects the phases a launch runs (KGW_FPH_*).  The PAIR class (kgw_pairs.cuh) runs belief propagation
// here, one warp per pair, and the chain-sequential phases (final draws, family knockoffs) in kernels of their own with
// several chains per warp; the state between the launches lives in the scratch slot, which is then the family's index
// in the batch (P.slot_by_family) instead of the resident warp's.
template <int SF, bool UMU, bool BIG>
__global__ void __launch_bounds__(128, KGW_FAM_MINB) family_kernel(const __grid_constant__ KgwFamParams P) {
  constexpr int KP = SF * 32;
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int wpc = blockDim.x >> 5;
  const int slot = blockIdx.x * wpc + wid;
  const int p = P.p, K = P.K;
  const size_t pK = (size_t)p * K;

  extern __shared__ __align__(16) unsigned char fam_smem[];
  double *tab = reinterpret_cast<double *>(fam_smem);                       // [8][256], shared by the CTA
  // [n_max][KP] last message of each chain, and its value before this sweep touched it
  double *pm = BIG ? P.pmg + (size_t)slot * 2 * P.n_max * KP : tab + 8 * 256 + (size_t)wid * 2 * P.n_max * KP;
  double *pm_old = pm + (size_t)P.n_max * KP;
  // BIG class: per-member bookkeeping of a sweep, [5][n_max]: neighbour count at the sweep position, message one locus
  // back computed / changed in this sweep, and the same two for the stretch or locus being processed
  int32_t *fl_nn = BIG ? P.flg + (size_t)slot * 5 * P.n_max : nullptr;
  int32_t *fl_cp = fl_nn + P.n_max, *fl_gp = fl_cp + P.n_max, *fl_cn = fl_gp + P.n_max, *fl_gn = fl_cn + P.n_max;
  if (UMU) {
    for (int t = threadIdx.x; t < 8 * 256; t += blockDim.x) tab[t] = P.emtab[t];
  }
  __syncthreads();
  const double p0 = 1.0 / ((double)K);   // initial value of every message (family_bp.cpp:421-427)

  // the phases `mask` of family `f` (index into this launch's list) in scratch slot `sub` of the warp
  auto family_phases = [&](const int f, const size_t fslot, const unsigned mask) {
    double *ml_base = P.msg + fslot * P.n_max * pK;                       // messages_left [n][p][K]
    double *mrc = P.mrc + fslot * P.n_max * (size_t)P.rtot * K;          // messages_right, stored loci only
    uint8_t *pat_base = P.pat + fslot * P.n_max * pK;
    uint8_t *Zb = P.bytes + fslot * 2 * P.n_max * p;     // 255 = not drawn yet = node still active
    uint8_t *Zkb = Zb + (size_t)P.n_max * p;
    const KgwFamDesc fd = P.fams[P.fam_list[f]];
    const int n = fd.n;
    const int32_t *mem = P.members + fd.mem_off;
    const int32_t *ivs0 = P.iv_start + fd.ivs_off;   // n+1 entries for this family
    auto nbl = [&](const KgwFamInterval *iv, int t) -> int { return __ldg(P.ivlist + iv->nb_off + t); };
    auto dll = [&](const KgwFamInterval *iv, int t) -> int { return __ldg(P.ivlist + iv->dl_off + t); };
    auto drl = [&](const KgwFamInterval *iv, int t) -> int { return __ldg(P.ivlist + iv->dr_off + t); };
    const KgwFamRange *rngs = P.ranges + fd.rng_off;
    const int nrng = fd.n_rng;
#ifdef KGW_PHASES
    long long fam_t0 = clock64();
#endif

    auto find_iv = [&](int i, int j) -> const KgwFamInterval * {
      const KgwFamInterval *iv = P.ivs + ivs0[i];
      while (iv->j1 < j) iv++;
      return iv;
    };
    // right message row of (i, j) inside the stored span of a range (ext0 = max(0, s0 - 1) .. s1)
    auto mr_row = [&](int i, int j, int ext0, int off) -> double * {
      return mrc + ((size_t)i * P.rtot + (size_t)(off + (j - ext0))) * K;
    };
    auto mr_find = [&](int i, int j) -> double * {   // outside the sweeps: look the range up
      for (int r = 0; r < nrng; r++) {
        const KgwFamRange rg = rngs[r];
        const int ext0 = max(0, rg.s0 - 1);
        if (j >= ext0 && j <= rg.s1) return mr_row(i, j, ext0, rg.off);
      }
      return mrc;   // not reached: every caller asks for a stored locus
    };
    // emission of node (i, j) for state k (family_bp.cpp:437-454)
    auto emission_of = [&](unsigned pt, int j, int nn) -> double {
      if (UMU) return tab[nn * 256 + pt];
      const double mu = P.loc[j].mu;
      double v = (pt & 1u) ? 1.0 - mu : mu;
      for (int t = 0; t < nn; t++) v *= ((pt >> (t + 1)) & 1u) ? 1.0 - mu : mu;
      if (nn > 0) v = pow(v, 1.0 / (double)(nn + 1));
      return v;
    };

    // emission of a node with more than KGW_FAM_MAXNB neighbours (BIG class): the same product in the same order
    // (family_bp.cpp:441-454), match bits read from the packed rows
    auto emission_big = [&](int i, const KgwFamInterval *iv, int j, int k) -> double {
      const double mu = P.loc[j].mu;
      const uint32_t rw = __ldg(P.Hw + (size_t)__ldg(P.ref_global + (size_t)mem[i] * K + k) * P.stride_w + (j >> 5));
      auto same = [&](int m) { return ((~(rw ^ __ldg(P.Hw + (size_t)mem[m] * P.stride_w + (j >> 5)))) >> (j & 31)) & 1u; };
      double v = same(i) ? 1.0 - mu : mu;
      for (int t = 0; t < iv->nn; t++) v *= same(nbl(iv, t)) ? 1.0 - mu : mu;
      return pow(v, 1.0 / (double)(iv->nn + 1));
    };
    // emission of node (i, j) at state k: pattern byte `pt` when the node has few neighbours
    auto emission_at = [&](unsigned pt, int i, const KgwFamInterval *iv, int nn, int j, int k) -> double {
      if (BIG && nn > KGW_FAM_MAXNB) return emission_big(i, iv, j, k);
      return emission_of(pt, j, nn);
    };

    // ---- initialise: nothing drawn, match patterns.  The messages are NOT written: the first BP
    //      iteration substitutes their initial value 1/K wherever the reference would read it.
    if (mask & KGW_FPH_INIT)
    for (int i = 0; i < n; i++) {
      const int i_abs = mem[i];
      const int32_t *ref = P.ref_global + (size_t)i_abs * K;
      for (int j = lane; j < p; j += 32) { Zb[(size_t)i * p + j] = 255; Zkb[(size_t)i * p + j] = 255; }
      // row 0 of a member's left messages is never touched by belief propagation (the left sweep ends at locus 1);
      // it is row p of the member before it: ones, the empty product a draw at the last locus sees (kgw_pairs.cuh)
      for (int k = lane; k < K; k += 32) {
        ml_base[(size_t)i * pK + k] = 1.0;
        if (i == n - 1) ml_base[(size_t)n * pK + k] = 1.0;   // (row 0 of the next slot, or the padding behind the last one)
      }
      for (int k = lane; k < K; k += 32) {
        const size_t rrow = (size_t)__ldg(ref + k) * P.stride_w;
        const KgwFamInterval *iv = P.ivs + ivs0[i];
        int ivj1 = iv->j1, ivnn = iv->nn;
        for (int w = 0; w <= (p - 1) >> 5; w++) {
          const uint32_t rw = __ldg(P.Hw + rrow + w);
          const uint32_t m0 = ~(rw ^ __ldg(P.Hw + (size_t)i_abs * P.stride_w + w));
          const int jlo = w << 5, jhi = min(jlo + 32, p);
          uint8_t *dst = pat_base + ((size_t)i * p + jlo) * K + k;
          while (ivj1 < jlo) { iv++; ivj1 = iv->j1; ivnn = iv->nn; }
          const int nnp = min(ivnn, KGW_FAM_MAXNB);   // nodes with more neighbours do not use their pattern byte
          if (ivj1 >= jhi - 1 && ivnn <= 1) {   // one neighbour list for the whole word, pairs
            const uint32_t m1 = (ivnn == 1) ? ~(rw ^ __ldg(P.Hw + (size_t)mem[nbl(iv, 0)] * P.stride_w + w)) : 0u;
            for (int j = jlo; j < jhi; j++) {
              const int q = j & 31;
              *dst = (uint8_t)(((m0 >> q) & 1u) | (((m1 >> q) & 1u) << 1));
              dst += K;
            }
          } else if (ivj1 >= jhi - 1) {   // one neighbour list for the whole word
            uint32_t mt[KGW_FAM_MAXNB];
#pragma unroll
            for (int t = 0; t < KGW_FAM_MAXNB; t++)
              mt[t] = (t < nnp) ? ~(rw ^ __ldg(P.Hw + (size_t)mem[nbl(iv, t)] * P.stride_w + w)) : 0u;
            for (int j = jlo; j < jhi; j++) {
              const int q = j & 31;
              unsigned pt = (m0 >> q) & 1u;
#pragma unroll
              for (int t = 0; t < KGW_FAM_MAXNB; t++) pt |= ((mt[t] >> q) & 1u) << (t + 1);
              *dst = (uint8_t)pt;
              dst += K;
            }
          } else {
            for (int j = jlo; j < jhi; j++) {
              while (ivj1 < j) { iv++; ivj1 = iv->j1; ivnn = iv->nn; }
              unsigned pt = (m0 >> (j & 31)) & 1u;
              for (int t = 0; t < min(ivnn, KGW_FAM_MAXNB); t++) {
                const uint32_t hw2 = __ldg(P.Hw + (size_t)mem[nbl(iv, t)] * P.stride_w + w);
                pt |= ((~(rw ^ hw2) >> (j & 31)) & 1u) << (t + 1);
              }
              *dst = (uint8_t)pt;
              dst += K;
            }
          }
        }
      }
    }
    __syncwarp();
    KGW_FAM_PHASE(0);

    // One direction of one BP iteration (family_bp.cpp:341-405 and the message formulas :429-559):
    // dir 0 = right (writes messages_right[i][j], j = 0 .. p-2), 1 = left (messages_left[i][j], j = p-1 .. 1).
    // `full`: first iteration of the first run_bp -- every node is computed and the "old" value of every
    // message (also of the left messages read at a junction by the first right sweep) is 1/K.
    // The sweep alternates between the special ranges (locus by locus, members in order, junction terms,
    // pinned nodes) and the plain stretches between them.  In a plain stretch every node depends on its own
    // chain only, so the members are taken one after the other with the running message in registers; a
    // chain is followed only while its message changes bits (or everywhere when `full`).
    // Message formula (:456-497): with q_k = f_k * (product of the incoming messages at k) and S = sum_k q_k the
    // reference forms const = sum_k a q_k, m_k = const + b q_k and divides by sum_k m_k = S (K a + b); here
    // new_k = a / (K a + b) + (b / (S (K a + b))) q_k: one multiply and one fused multiply-add per state and
    // one warp reduction per message (the same number up to rounding; sums over k are tree-ordered anyway).
    // Returns this lane's share of sum |new - old|.
    auto sweep = [&](auto dir_tag, const bool full) -> double {
      constexpr int dir = decltype(dir_tag)::value;
      double diff = 0.0;
      unsigned comp_prev = 0, chg_prev = 0;   // members whose message one locus back was computed (value in pm) / changed bits in this sweep
      constexpr int step = (dir == 0) ? 1 : -1;
      const int jlast = (dir == 0) ? p - 2 : 1;   // last locus of the sweep
      int j = (dir == 0) ? 0 : p - 1;
      unsigned nnpack = 0;                    // neighbour count of every member at the current locus, 3 bits each
      // (BIG class: the same bookkeeping in the slot's scratch, one int per member; every lane writes the same values)
      auto nn_get = [&](int i) -> int { return BIG ? fl_nn[i] : (int)((nnpack >> (3 * i)) & 7u); };
      auto nn_set = [&](int i, int v) {
        if (BIG) fl_nn[i] = v;
        else nnpack = (nnpack & ~(7u << (3 * i))) | ((unsigned)v << (3 * i));
      };
      auto cp_get = [&](int i) -> bool { return BIG ? (fl_cp[i] != 0) : (((comp_prev >> i) & 1u) != 0); };
      auto gp_get = [&](int i) -> bool { return BIG ? (fl_gp[i] != 0) : (((chg_prev >> i) & 1u) != 0); };
      for (int i = 0; i < n; i++) {
        nn_set(i, find_iv(i, j)->nn);
        if (BIG) { fl_cp[i] = 0; fl_gp[i] = 0; }
      }
      bool valid[SF];
#pragma unroll
      for (int s = 0; s < SF; s++) valid[s] = (s * 32 + lane < K);
      int prev_ext0 = 0, prev_off = 0;        // stored span of the range behind the sweep position

      for (int rr = 0; rr <= nrng; rr++) {
        int s_first = 0, s_last = -1, cur_ext0 = 0, cur_off = 0;
        if (rr < nrng) {
          const KgwFamRange rg = rngs[dir == 0 ? rr : nrng - 1 - rr];
          s_first = (dir == 0) ? rg.s0 : rg.s1;
          s_last = (dir == 0) ? rg.s1 : rg.s0;
          cur_ext0 = max(0, rg.s0 - 1);
          cur_off = rg.off;
        }
        // ---- plain stretch j .. e (sweep order)
        int e = (rr < nrng) ? s_first - step : jlast;
        if (dir == 0 ? (e > jlast) : (e < jlast)) e = jlast;
        const int len = (e - j) * step + 1;
        if (len > 0) {
          unsigned comp_new = 0, chg_new = 0;
          if (BIG) for (int i = 0; i < n; i++) { fl_cn[i] = 0; fl_gn[i] = 0; }
          const bool has_in = (dir == 0) ? (j > 0) : (j < p - 1);
          if (!BIG && full && n == 2 && (!has_in || comp_prev == 3u)) {
            // First iteration of a pair (the common family): the two chains of the stretch are independent, so they
            // run in lockstep -- one loop, one set of locus constants, one shuffle tree for both sums, and twice
            // the instruction-level parallelism.  Nothing is compared (the old value of every message is 1/K).
            const int nnA = (int)(nnpack & 7u), nnB = (int)((nnpack >> 3) & 7u);
            constexpr ptrdiff_t sgn = step;
            const ptrdiff_t sK = sgn * K;
            const uint8_t *ppA = pat_base + (size_t)j * K + lane, *ppB = ppA + pK;
            double *dpA = ml_base + (size_t)j * K + lane, *dpB = dpA + pK;   // left sweep: every row is stored
            const KgwFamLocus *lp = P.loc + (dir == 0 ? j + 1 : j);
            double vA[SF], vB[SF];
            unsigned pcA[SF], pcB[SF];
#pragma unroll
            for (int s = 0; s < SF; s++) {
              vA[s] = 1.0;
              vB[s] = 1.0;
              if (valid[s] && has_in) { vA[s] = pm[s * 32 + lane]; vB[s] = pm[KP + s * 32 + lane]; }
              pcA[s] = valid[s] ? ppA[s * 32] : 0u;
              pcB[s] = valid[s] ? ppB[s * 32] : 0u;
            }
            double lb = lp->b, lw = lp->w, lr1 = lp->r1;
            for (int t = 0; t < len; t++) {
              unsigned pnA[SF], pnB[SF];
              const bool more = (t + 1 < len);
#pragma unroll
              for (int s = 0; s < SF; s++) {
                pnA[s] = (valid[s] && more) ? ppA[sK + s * 32] : 0u;
                pnB[s] = (valid[s] && more) ? ppB[sK + s * 32] : 0u;
              }
              const double lbn = more ? lp[step].b : 0.0, lwn = more ? lp[step].w : 1.0, lr1n = more ? lp[step].r1 : 0.0;
              if ((t & 3) == 0 && t + 12 < len) {
                fam_prefetch(ppA + (dir == 0 ? 8 : 11) * sK - lane, 4 * K, lane);
                fam_prefetch(ppB + (dir == 0 ? 8 : 11) * sK - lane, 4 * K, lane);
              }
              double qA[SF], qB[SF];
              double sA = 0.0, sB = 0.0;
#pragma unroll
              for (int s = 0; s < SF; s++) {
                qA[s] = 0.0;
                qB[s] = 0.0;
                if (valid[s]) {
                  qA[s] = emission_of(pcA[s], j + t * step, nnA) * vA[s];
                  qB[s] = emission_of(pcB[s], j + t * step, nnB) * vB[s];
                  sA += qA[s];
                  sB += qB[s];
                }
              }
              warp_sum2(sA, sB);
              const double cA = lb * fast_rcp(sA * lw), cB = lb * fast_rcp(sB * lw);
#pragma unroll
              for (int s = 0; s < SF; s++) {
                if (valid[s]) {
                  vA[s] = lr1 + cA * qA[s];
                  vB[s] = lr1 + cB * qB[s];
                  diff += fabs(vA[s] - p0);
                  diff += fabs(vB[s] - p0);
                  if (dir == 1) { dpA[s * 32] = vA[s]; dpB[s * 32] = vB[s]; }
                }
              }
              if (dir == 0 && !more && rr < nrng) {   // the next range reads the last plain row
                double *rowA = mr_row(0, e, cur_ext0, cur_off), *rowB = mr_row(1, e, cur_ext0, cur_off);
#pragma unroll
                for (int s = 0; s < SF; s++) if (valid[s]) { rowA[s * 32 + lane] = vA[s]; rowB[s * 32 + lane] = vB[s]; }
              }
              ppA += sK;
              ppB += sK;
              dpA += sK;
              dpB += sK;
              lp += step;
              lb = lbn;
              lw = lwn;
              lr1 = lr1n;
#pragma unroll
              for (int s = 0; s < SF; s++) { pcA[s] = pnA[s]; pcB[s] = pnB[s]; }
            }
#pragma unroll
            for (int s = 0; s < SF; s++) if (valid[s]) { pm[s * 32 + lane] = vA[s]; pm[KP + s * 32 + lane] = vB[s]; }
            comp_new = 3u;
            chg_new = 3u;
          } else
          for (int i = 0; i < n; i++) {
            if (!(full || gp_get(i))) continue;
            const int nn = nn_get(i);
            const KgwFamInterval *ivp = (BIG && nn > KGW_FAM_MAXNB) ? find_iv(i, j) : nullptr;   // constant over a plain stretch
            const size_t rowo = ((size_t)i * p + j) * K + lane;
            constexpr ptrdiff_t sgn = step;
            const ptrdiff_t sK = sgn * K;
            const uint8_t *pp = pat_base + rowo;
            const KgwFamLocus *lp = P.loc + (dir == 0 ? j + 1 : j);
            const bool hasv = (dir == 0) ? (j > 0) : (j < p - 1);
            unsigned ptc[SF];
#pragma unroll
            for (int s = 0; s < SF; s++) ptc[s] = valid[s] ? pp[s * 32] : 0u;
            double lb = lp->b, lw = lp->w, lr1 = lp->r1;
            bool ch = true;
            int t = 0;
            double v[SF];
            if constexpr (dir == 0) {
              // right sweep: nothing of the stretch is stored except its last locus (the next range reads it);
              // when not `full`, vo[] follows the chain as it was before this sweep
              double vo[SF];
              const double *prow = (hasv && !cp_get(i)) ? mr_row(i, j - 1, prev_ext0, prev_off) : nullptr;
#pragma unroll
              for (int s = 0; s < SF; s++) {
                v[s] = 1.0;   // no incoming message at the first node of the sweep
                vo[s] = 1.0;
                if (valid[s] && hasv) {
                  v[s] = prow ? prow[s * 32 + lane] : pm[i * KP + s * 32 + lane];
                  vo[s] = full ? v[s] : pm_old[i * KP + s * 32 + lane];
                }
              }
              double *lastrow = (rr < nrng) ? mr_row(i, e, cur_ext0, cur_off) : nullptr;
              for (; t < len; t++) {
                unsigned ptn[SF];
                const bool more = (t + 1 < len);
#pragma unroll
                for (int s = 0; s < SF; s++) ptn[s] = (valid[s] && more) ? pp[sK + s * 32] : 0u;
                const double lbn = more ? lp[step].b : 0.0, lwn = more ? lp[step].w : 1.0, lr1n = more ? lp[step].r1 : 0.0;
                if ((t & 3) == 0 && t + 12 < len) fam_prefetch(pp + 8 * sK - lane, 4 * K, lane);
                double q[SF], qo[SF];
                double qsum = 0.0, qosum = 0.0;
#pragma unroll
                for (int s = 0; s < SF; s++) {
                  q[s] = 0.0;
                  qo[s] = 0.0;
                  if (valid[s]) {
                    const double fk = emission_at(ptc[s], i, ivp, nn, j + t, s * 32 + lane);
                    q[s] = fk * v[s];
                    qsum += q[s];
                    if (!full) { qo[s] = fk * vo[s]; qosum += qo[s]; }
                  }
                }
                bool ne = false;
                if (full) {
                  const double c2 = lb * fast_rcp(warp_sum(qsum) * lw);
#pragma unroll
                  for (int s = 0; s < SF; s++) {
                    if (valid[s]) {
                      v[s] = lr1 + c2 * q[s];
                      diff += fabs(v[s] - p0);
                    }
                  }
                  ch = true;
                } else {
                  warp_sum2(qsum, qosum);
                  const double c2 = lb * fast_rcp(qsum * lw), c2o = lb * fast_rcp(qosum * lw);
#pragma unroll
                  for (int s = 0; s < SF; s++) {
                    if (valid[s]) {
                      v[s] = lr1 + c2 * q[s];
                      vo[s] = lr1 + c2o * qo[s];
                      diff += fabs(v[s] - vo[s]);
                      ne |= (v[s] != vo[s]);
                    }
                  }
                  ch = __any_sync(FULL, ne);
                }
                if (!more && lastrow && ch) {
#pragma unroll
                  for (int s = 0; s < SF; s++) if (valid[s]) lastrow[s * 32 + lane] = v[s];
                }
                pp += sK;
                lp += step;
                lb = lbn;
                lw = lwn;
                lr1 = lr1n;
#pragma unroll
                for (int s = 0; s < SF; s++) ptc[s] = ptn[s];
                if (!full && !ch) { t++; break; }   // the two chains agree bit for bit: everything further on is clean
              }
            } else {
              double *dp = ml_base + rowo;
#pragma unroll
              for (int s = 0; s < SF; s++) {
                v[s] = 1.0;
                if (valid[s] && hasv) v[s] = cp_get(i) ? pm[i * KP + s * 32 + lane] : dp[s * 32 - sK];
              }
              for (; t < len; t++) {
                // fetches for the next locus of this chain, and L2 prefetches further ahead
                unsigned ptn[SF];
                double oldv[SF];
                const bool more = (t + 1 < len);
#pragma unroll
                for (int s = 0; s < SF; s++) {
                  ptn[s] = (valid[s] && more) ? pp[sK + s * 32] : 0u;
                  oldv[s] = (valid[s] && !full) ? dp[s * 32] : (valid[s] ? p0 : 0.0);
                }
                const double lbn = more ? lp[step].b : 0.0, lwn = more ? lp[step].w : 1.0, lr1n = more ? lp[step].r1 : 0.0;
                if ((t & 3) == 0 && t + 12 < len) {   // rows t+8 .. t+11 of this chain (contiguous in memory)
                  fam_prefetch(pp + 11 * sK - lane, 4 * K, lane);
                  if (!full) fam_prefetch(dp + 11 * sK - lane, 32 * K, lane);
                }
                double q[SF];
                double qsum = 0.0;
#pragma unroll
                for (int s = 0; s < SF; s++) {
                  q[s] = 0.0;
                  if (valid[s]) {
                    q[s] = emission_at(ptc[s], i, ivp, nn, j - t, s * 32 + lane) * v[s];
                    qsum += q[s];
                  }
                }
                const double S = warp_sum(qsum);
                const double c2 = lb * fast_rcp(S * lw);
                bool ne = false;
#pragma unroll
                for (int s = 0; s < SF; s++) {
                  if (valid[s]) {
                    const double nv = lr1 + c2 * q[s];
                    diff += fabs(nv - oldv[s]);
                    ne |= (nv != oldv[s]);
                    v[s] = nv;
                  }
                }
                ch = __any_sync(FULL, ne);
                if (full || ch) {
#pragma unroll
                  for (int s = 0; s < SF; s++) if (valid[s]) dp[s * 32] = v[s];
                }
                pp += sK;
                dp += sK;
                lp += step;
                lb = lbn;
                lw = lwn;
                lr1 = lr1n;
#pragma unroll
                for (int s = 0; s < SF; s++) ptc[s] = ptn[s];
                if (!full && !ch) { t++; break; }   // bit-identical message: everything further along this chain is clean
              }
            }
            if (t == len) {   // the chain reached the end of the stretch: hand its last message to the next range
#pragma unroll
              for (int s = 0; s < SF; s++) if (valid[s]) pm[i * KP + s * 32 + lane] = v[s];
              if (BIG) { fl_cn[i] = 1; fl_gn[i] = ch ? 1 : 0; }
              else { comp_new |= 1u << i; if (ch) chg_new |= 1u << i; }
            }
          }
          comp_prev = comp_new;
          chg_prev = chg_new;
          if (BIG) {
            __syncwarp();
            for (int i = 0; i < n; i++) { fl_cp[i] = fl_cn[i]; fl_gp[i] = fl_gn[i]; }
            __syncwarp();
          }
          j = e + step;
        }
        if (rr == nrng) break;
        // ---- special range s_first .. s_last, locus by locus
        for (; (dir == 0) ? (j <= s_last && j <= jlast) : (j >= s_last && j >= jlast); j += step) {
          const KgwFamLocus lc = P.loc[dir == 0 ? j + 1 : j];
          const bool has_prev = (dir == 0) ? (j > 0) : (j < p - 1);
          unsigned comp_cur = 0, chg_cur = 0;
          for (int i = 0; i < n; i++) {
            const KgwFamInterval *iv = find_iv(i, j);
            const int nn = iv->nn;
            nn_set(i, nn);
            if (BIG) { fl_cn[i] = 0; fl_gn[i] = 0; }
            if (Zb[(size_t)i * p + j] != 255) continue;   // pinned by condition_on: inactive node
            const int ndl = (j == iv->j0 && j > 0) ? iv->ndl : 0;
            const int ndr = (j == iv->j1 && j < p - 1) ? iv->ndr : 0;
            const size_t rowo = ((size_t)i * p + j) * K;
            double *dst = (dir == 0) ? mr_row(i, j, cur_ext0, cur_off) : ml_base + rowo;
            // own chain, one locus back in sweep order (for the right sweep: inside the stored span of this range)
            const double *prevg = (dir == 0) ? mr_row(i, j - 1, cur_ext0, cur_off) : ml_base + rowo + K;
            const bool prev_sm = cp_get(i);
            double msg[SF], oldv[SF];
            double qsum = 0.0;
#pragma unroll
            for (int s = 0; s < SF; s++) {
              const int k = s * 32 + lane;
              msg[s] = 0.0;
              oldv[s] = 0.0;
              if (k < K) {
                double q = emission_at(pat_base[rowo + k], i, iv, nn, j, k);
                oldv[s] = full ? p0 : dst[k];
                if (has_prev) q *= prev_sm ? pm[i * KP + k] : prevg[k];
                if (dir == 0) {
                  for (int t = 0; t < ndl; t++) q *= mr_row(dll(iv, t), j - 1, cur_ext0, cur_off)[k];
                  for (int t = 0; t < ndr; t++) q *= full ? p0 : ml_base[((size_t)drl(iv, t) * p + j + 1) * K + k];
                } else {
                  for (int t = 0; t < ndr; t++) q *= ml_base[((size_t)drl(iv, t) * p + j + 1) * K + k];
                  for (int t = 0; t < ndl; t++) q *= mr_row(dll(iv, t), j - 1, cur_ext0, cur_off)[k];
                }
                msg[s] = q;
                qsum += q;
              }
            }
            const double S = warp_sum(qsum);
            const double c2 = lc.b * fast_rcp(S * lc.w);
            bool ne = false;
#pragma unroll
            for (int s = 0; s < SF; s++) {
              if (s * 32 + lane < K) {
                const double nv = lc.r1 + c2 * msg[s];
                diff += fabs(nv - oldv[s]);
                ne |= (nv != oldv[s]);
                msg[s] = nv;
              }
            }
            const bool ch = __any_sync(FULL, ne);
#pragma unroll
            for (int s = 0; s < SF; s++) {
              const int k = s * 32 + lane;
              if (k < K) {
                if (full || ch) dst[k] = msg[s];
                pm[i * KP + k] = msg[s];
                if (dir == 0) pm_old[i * KP + k] = oldv[s];
              }
            }
            if (BIG) { fl_cn[i] = 1; fl_gn[i] = ch ? 1 : 0; }
            else { comp_cur |= 1u << i; if (ch) chg_cur |= 1u << i; }
          }
          comp_prev = comp_cur;
          chg_prev = chg_cur;
          if (BIG) {
            __syncwarp();
            for (int i = 0; i < n; i++) { fl_cp[i] = fl_cn[i]; fl_gp[i] = fl_gn[i]; }
            __syncwarp();
          }
        }
        prev_ext0 = cur_ext0;
        prev_off = cur_off;
      }
      return diff;
    };

    // run_bp (family_bp.cpp:318-419)
    auto run_bp = [&](const bool first) {
      for (int it = 0; it < 20; it++) {
        double diff = 0.0;
        for (int dir = 0; dir < 2; dir++) {
          double dt = (dir == 0) ? sweep(std::integral_constant<int, 0>{}, first && it == 0)
                                 : sweep(std::integral_constant<int, 1>{}, first && it == 0);
          KGW_FAM_PHASE((first && it == 0) ? 1 : 2);
          dt = warp_sum(dt) / (double)((size_t)p * (size_t)n);
          diff += dt;
        }
        if (diff <= 1e-3) break;
      }
    };

    // condition_on (family_bp.cpp:236-255), forward = false (the final pass keeps its pins in registers)
    auto condition_on = [&](int i, int j, int kz) {
      for (int l = lane; l < K; l += 32) {
        if (j < p - 1) {
          const KgwFamLocus lc = P.loc[j + 1];
          mr_find(i, j)[l] = lc.a_al + lc.b * (double)(l == kz);
        }
        if (j > 0) {
          const KgwFamLocus lc = P.loc[j];
          ml_base[((size_t)i * p + j) * K + l] = lc.a_al + lc.b * (double)(l == kz);
        }
      }
      if (lane == 0) Zb[(size_t)i * p + j] = (uint8_t)kz;
    };

    // sample_posterior_joint at a junction node (family_bp.cpp:569-601)
    auto draw_junction = [&](int i, int j, const KgwFamInterval *iv) -> int {
      const int ndl = (j == iv->j0 && j > 0) ? iv->ndl : 0, ndr = (j == iv->j1 && j < p - 1) ? iv->ndr : 0;
      double w[SF];
#pragma unroll
      for (int s = 0; s < SF; s++) {
        const int k = s * 32 + lane;
        w[s] = 0.0;
        if (k < K) {
          double v = emission_at(pat_base[((size_t)i * p + j) * K + k], i, iv, iv->nn, j, k);
          if (j > 0) {
            v *= mr_find(i, j - 1)[k];
            for (int t = 0; t < ndl; t++) v *= mr_find(dll(iv, t), j - 1)[k];
          }
          if (j < p - 1) {
            v *= ml_base[((size_t)i * p + j + 1) * K + k];
            for (int t = 0; t < ndr; t++) v *= ml_base[((size_t)drl(iv, t) * p + j + 1) * K + k];
          }
          w[s] = v;
        }
      }
      return warp_choice<SF>(w, fam_uniform(P, mem[i], STREAM_Z, j), K, lane);
    };

    // ---- FamilyBP::run (family_bp.cpp:79-100) and sample_posterior (:160-230): BP, then the junction nodes in
    //      (member, locus) order -- they sit at interval ends (:168-193) -- each followed by a BP re-run (:194-211).
    //      One loop with a cursor over (member, interval, end), so that run_bp is instantiated once.
    if (mask & KGW_FPH_BP) {
      int ci = 0, ce = 0;
      const KgwFamInterval *civ = P.ivs + ivs0[0];
      bool first = true;
      for (;;) {
        run_bp(first);
        first = false;
        int fj = -1;
        while (ci < n) {
          if (civ >= P.ivs + ivs0[ci + 1]) {
            ci++;
            if (ci < n) civ = P.ivs + ivs0[ci];
            ce = 0;
            continue;
          }
          if (ce >= 2) { civ++; ce = 0; continue; }
          const int e = ce++;
          if (e == 1 && civ->j1 == civ->j0) continue;
          const int j = (e == 0) ? civ->j0 : civ->j1;
          const bool jl = (j == civ->j0) && j > 0 && civ->ndl > 0, jr = (j == civ->j1) && j < p - 1 && civ->ndr > 0;
          if (!(jl || jr)) continue;
          if (Zb[(size_t)ci * p + j] != 255) continue;
          fj = j;
          break;
        }
        if (fj < 0) break;
        const int kz = draw_junction(ci, fj, civ);
        __syncwarp();
        condition_on(ci, fj, kz);
        for (int t = 0; t < civ->nn; t++) condition_on(nbl(civ, t), fj, kz);
        __syncwarp();
        KGW_FAM_PHASE(3);
      }
    }
    // every remaining node, member by member, loci ascending (:217-229).  By then every node to the left on
    // the same chain has been drawn, so messages_right[i][j-1] is the point-mass transition condition_on
    // would have stored (:243-246) and is formed from the previous draw; nothing reads the right messages
    // afterwards, so they are not written.  messages_left[i][j+1] and the pattern bytes are fetched one
    // locus ahead of the (serial) draw.  Lane l owns the contiguous states l*SF .. here (warp_choice_contig).
    if (mask & KGW_FPH_DRAW)
    for (int i = 0; i < n; i++) {
      const KgwFamInterval *iv = P.ivs + ivs0[i];
      int ivj1 = iv->j1, ivnn = iv->nn;
      unsigned nbpack = 0;   // SMALL class: the neighbour list of the current interval, 3 bits each
      if (!BIG) for (int t = 0; t < ivnn; t++) nbpack |= (unsigned)nbl(iv, t) << (3 * t);
      const uint8_t *zrow = Zb + (size_t)i * p;
      const double *mlrow = ml_base + (size_t)i * pK;
      const uint8_t *ptrow = pat_base + (size_t)i * pK;
      int zprev = 0;
      double la_c = P.loc[0].a_al, lb_c = P.loc[0].b;
      double ml_c[SF];
      unsigned pt_c[SF];
#pragma unroll
      for (int s = 0; s < SF; s++) {
        const int k = lane * SF + s;
        pt_c[s] = (k < K) ? ptrow[k] : 0u;
        ml_c[s] = (k < K && p > 1) ? mlrow[(size_t)K + k] : 1.0;
      }
      for (int w = 0; w <= (p - 1) >> 5; w++) {
        const int jw = (w << 5) + lane;
        const int zw = (jw < p) ? zrow[jw] : 0;
        const double uw = (jw < p) ? fam_uniform(P, mem[i], STREAM_Z, jw) : 0.0;   // lane q holds the uniform of locus 32 w + q
        const int jhi = min(32, p - (w << 5));
        for (int q = 0; q < jhi; q++) {
          const int j = (w << 5) + q;
          double ml_n[SF];
          unsigned pt_n[SF];
#pragma unroll
          for (int s = 0; s < SF; s++) {   // next locus: pattern j+1, left message j+2
            const int k = lane * SF + s;
            pt_n[s] = (k < K && j + 1 < p) ? ptrow[(size_t)(j + 1) * K + k] : 0u;
            ml_n[s] = (k < K && j + 2 < p) ? mlrow[(size_t)(j + 2) * K + k] : 1.0;
          }
          const double la_n = (j + 1 < p) ? P.loc[j + 1].a_al : 0.0, lb_n = (j + 1 < p) ? P.loc[j + 1].b : 0.0;
          if ((q & 3) == 0 && j + 14 < p) {
            fam_prefetch(ptrow + (size_t)(j + 9) * K, 4 * K, lane);
            fam_prefetch(mlrow + (size_t)(j + 10) * K, 32 * K, lane);
          }
          const int zj = __shfl_sync(FULL, zw, q);
          if (zj != 255) {
            zprev = zj;
          } else {
            while (ivj1 < j) {
              iv++;
              ivj1 = iv->j1;
              ivnn = iv->nn;
              nbpack = 0;
              if (!BIG) for (int t = 0; t < ivnn; t++) nbpack |= (unsigned)nbl(iv, t) << (3 * t);
            }
            double wv[SF];
#pragma unroll
            for (int s = 0; s < SF; s++) {
              const int k = lane * SF + s;
              wv[s] = 0.0;
              if (k < K) {
                double v = emission_at(pt_c[s], i, iv, ivnn, j, k);
                if (j > 0) v *= la_c + lb_c * (double)(k == zprev);
                if (j < p - 1) v *= ml_c[s];
                wv[s] = v;
              }
            }
            const int kz = warp_choice_contig<SF>(wv, __shfl_sync(FULL, uw, q), K, lane);
            if (lane == 0) {
              Zb[(size_t)i * p + j] = (uint8_t)kz;
              for (int t = 0; t < ivnn; t++) Zb[(size_t)(BIG ? nbl(iv, t) : (int)((nbpack >> (3 * t)) & 7u)) * p + j] = (uint8_t)kz;
            }
            zprev = kz;
          }
#pragma unroll
          for (int s = 0; s < SF; s++) { pt_c[s] = pt_n[s]; ml_c[s] = ml_n[s]; }
          la_c = la_n;
          lb_c = lb_n;
        }
      }
      __syncwarp();
    }

    KGW_FAM_PHASE(4);
    // ---- FamilyKnockoffs::run (family_knockoffs.cpp:77-293) as a flat op list
    const KgwFamOp *ops = P.ops + fd.op_off;
    if (mask & (KGW_FPH_KO | KGW_FPH_KOCOPY))
    for (int q = 0; q < fd.n_ops; q++) {
      const KgwFamOp op = ops[q];
      if (op.kind == KGW_FOP_DRAW && !(mask & KGW_FPH_KO)) continue;
      if (op.kind == KGW_FOP_COPYZ) {
        for (int j = op.a + lane; j <= op.b; j += 32) Zkb[(size_t)op.i * p + j] = Zb[(size_t)op.i * p + j];
      } else if (op.kind == KGW_FOP_COPYZK) {
        for (int j = op.a + lane; j <= op.b; j += 32) Zkb[(size_t)op.i * p + j] = Zkb[(size_t)op.src * p + j];
      } else {
        // FamilyKnockoffs::sample_knockoff_MC (family_knockoffs.cpp:295-454), groups op.a .. op.b
        const int i = op.i, g_begin = op.a, g_end = op.b;
        const uint8_t *z = Zb + (size_t)i * p;
        uint8_t *zk = Zkb + (size_t)i * p;
        // N_old enters only through 1/N_old: its reciprocal is formed once per group and multiplied in
        // (x * (1/y) instead of x / y moves a weight by an ulp, see DESIGN.md "Arithmetic vs the reference")
        double Nv[SF], rNo[SF];
#pragma unroll
        for (int s = 0; s < SF; s++) { Nv[s] = 1.0; rNo[s] = 1.0; }
        int zk_last = 0;
        int uw_word = -1;
        double uw = 0.0;   // lane q holds the uniform of locus 32 uw_word + q
        for (int g = g_begin; g <= g_end; g++) {
          const int e0 = P.grp_first[g], e1 = P.grp_last[g];
          const KgwLocusKO k0v = P.ko[e0];
          const bool lastg = (g == P.n_groups - 1);
          const int z0 = (g > 0) ? z[e0 - 1] : 0;
          const int z0k = zk_last;
          const int z1 = lastg ? 0 : z[e1 + 1];
          // T_k for the first locus of the group (:358-375, :378-398, :423-436)
          double tq[SF];
          double part = 0.0;
#pragma unroll
          for (int s = 0; s < SF; s++) {
            const int k = lane * SF + s;
            double tmp = k0v.a;
            if (g > 0) {
              tmp = (k == z0) ? k0v.apb : k0v.a;
              if (g != g_begin) tmp = tmp * ((k == z0k) ? k0v.apb : k0v.a);
            }
            tq[s] = (g == 0) ? tmp : tmp * rNo[s];   // T_k / N_old[k]
            if (k < K) part += tq[s];
          }
          const double N_sum = warp_sum(part);
          double nn_part = 0.0;
#pragma unroll
          for (int s = 0; s < SF; s++) {
            const int k = lane * SF + s;
            if (!lastg) {
              Nv[s] += k0v.vb * N_sum;
              Nv[s] += k0v.vs * tq[s];   // vstar[0] == vs of the first locus
            }
            if (Nv[s] < N_MIN) Nv[s] = N_MIN;
            if (k < K) nn_part += Nv[s];
          }
          const double r_norm = fast_rcp(warp_sum(nn_part));
#pragma unroll
          for (int s = 0; s < SF; s++) {
            Nv[s] *= r_norm;
            if (Nv[s] < N_MIN) Nv[s] = N_MIN;
          }
          int zkprev = 0;
          for (int j = e0; j <= e1; j++) {
            const KgwLocusKO kv = P.ko[j];
            if ((j >> 5) != uw_word) {
              uw_word = j >> 5;
              const int jl = (uw_word << 5) + lane;
              uw = (jl < p) ? fam_uniform(P, mem[i], STREAM_ZK, jl) : 0.0;
            }
            double w[SF];
#pragma unroll
            for (int s = 0; s < SF; s++) {
              const int k = lane * SF + s;
              double wk;
              if (j > e0) wk = (k == zkprev) ? kv.apb : kv.a;
              else wk = (g == 0) ? tq[s] * rNo[s] : tq[s];
              wk *= lastg ? kv.vb : (kv.vb + kv.vs * (double)(k == z1));
              w[s] = (k < K) ? wk : 0.0;
            }
            zkprev = warp_choice_contig<SF>(w, __shfl_sync(FULL, uw, j & 31), K, lane);
            if (lane == 0) zk[j] = (uint8_t)zkprev;
          }
          zk_last = zkprev;
#pragma unroll
          for (int s = 0; s < SF; s++) rNo[s] = fast_rcp(Nv[s]);
        }
      }
      __syncwarp();
    }

    KGW_FAM_PHASE(5);
    if (!(mask & KGW_FPH_OUT)) return;
    // ---- the reference's final checks (family_knockoffs.cpp:272-291)
    int bad = 0;
    for (int i = 0; i < n; i++) {
      const KgwFamInterval *iv = P.ivs + ivs0[i];
      for (int j = lane; j < p; j += 32) {
        if (Zkb[(size_t)i * p + j] == 255) bad |= 1;
      }
      for (; iv < P.ivs + ivs0[i + 1]; iv++) {
        for (int t = 0; t < iv->nn; t++)
          for (int j = iv->j0 + lane; j <= iv->j1; j += 32)
            if (Zkb[(size_t)i * p + j] != Zkb[(size_t)nbl(iv, t) * p + j]) bad |= 2;
      }
    }
    bad = __reduce_or_sync(FULL, bad);
    if (bad && lane == 0) atomicMin(P.status, (bad & 1) ? -10 : -11);

    // ---- plain emission with the global references (knockoffs.cpp:505-510, :1430-1441), packed rows
    for (int i = 0; i < n; i++) {
      const int i_abs = mem[i];
      const int32_t *ref = P.ref_global + (size_t)i_abs * K;
      for (int w = 0; w <= (p - 1) >> 5; w++) {
        const int j = (w << 5) + lane;
        unsigned hk = 0;
        if (j < p) {
          const int zkj = Zkb[(size_t)i * p + j];
          const uint32_t rw = __ldg(P.Hw + (size_t)__ldg(ref + min(zkj, K - 1)) * P.stride_w + w);
          const double mu = P.loc[j].mu;
          const double w1 = ((rw >> lane) & 1u) ? 1.0 - mu : mu;   // mutate(), knockoffs.cpp:1626
          const double w0 = 1.0 - w1;
          double sum = 0.0;
          sum += w0;
          sum += w1;
          hk = (fam_uniform(P, i_abs, STREAM_HK, j) * sum < w0) ? 0u : 1u;
        }
        const unsigned bits = __ballot_sync(FULL, hk);
        if (lane == 0) P.Hk[(size_t)i_abs * P.stride_w + w] = bits;
      }
      if (P.dbg_z || P.dbg_zk) {
        for (int j = lane; j < p; j += 32) {
          if (P.dbg_z) P.dbg_z[(size_t)(fd.out_off + i) * p + j] = Zb[(size_t)i * p + j];
          if (P.dbg_zk) P.dbg_zk[(size_t)(fd.out_off + i) * p + j] = Zkb[(size_t)i * p + j];
        }
      }
    }
    __syncwarp();
    KGW_FAM_PHASE(6);
  };

  for (;;) {
    int f = 0;
    if (lane == 0) f = atomicAdd(P.work, 1);
    f = __shfl_sync(FULL, f, 0) + P.f_base;
    if (f >= P.n_families) break;
    family_phases(f, P.slot_by_family ? (size_t)(f - P.f_base) : (size_t)slot, (unsigned)P.phase_mask);
  }
}

}  // namespace kgw
