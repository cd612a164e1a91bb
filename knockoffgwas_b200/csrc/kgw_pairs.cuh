// knockoffgwas_b200/csrc/kgw_pairs.cuh
//
// The PAIR class of the related branch: IBD families of two haplotypes (what RaPID + union-find mostly produces,
// metadata.cpp:242-293).  kgw::family_kernel (kgw_family.cuh) gives a family one warp, lane = HMM state: every node
// then pays a warp-wide reduction or scan, and the two chain-sequential phases -- the final draws of every remaining
// node (FamilyBP::sample_posterior, family_bp.cpp:217-229) and the family knockoffs (FamilyKnockoffs::sample_knockoff_MC,
// family_knockoffs.cpp:295-454) -- are bound by that chain of shuffles, not by arithmetic.  Here they run as kernels of
// their own in the unrelated sampler's layout: L lanes per chain, each owning S consecutive states in registers
// (K <= L*S), 32/L chains per warp in lock step, so a scan is log2(L) shuffle steps shared by 32/L chains and the
// weights are formed S per lane without idle lanes.  The state between the launches (left messages, pattern bytes,
// latent paths) lives in the pair's scratch slot (KgwFamParams::slot_by_family).
//
//   pair_draw_kernel   the two chains of a pair side by side at the same locus.  Legal because member 1's node at a locus
//                      depends on member 0 only through the neighbour copy at that locus (family_bp.cpp:224-227), which
//                      is a shuffle here; the left messages are final after belief propagation (condition_on with
//                      forward = true touches right messages only, :236-255).
//   pair_ko_kernel     the DRAW operations of FamilyKnockoffs::run (:172-193, :196-270) are independent of each other (each
//                      reads Z and its own Zk only), so they are handed out to lane groups through an atomic counter,
//                      longest first.  The copy operations run afterwards (family_kernel, KGW_FPH_KOCOPY); the host
//                      rewrites a pair's list so that this order gives the reference's result (kgw_api.cu).
#pragma once
#include "kgw_family.cuh"
#include <type_traits>


namespace kgw {

// sum over the L consecutive lanes of a chain
template <int L>
__device__ __forceinline__ double lg_sum(double v) {
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// weighted_choice (utils.cpp:690-698) over K weights held S per lane by the L consecutive lanes of a chain (lane m of
// the chain owns states m*S .. m*S+S-1): one scan over the lane totals, then the state inside the lane that holds u*sum.
// Same tie rules as warp_choice_contig (kgw_family.cuh): first k with u*S < cumulative weight, last state otherwise.
template <int L, int S>
__device__ __forceinline__ int lg_choice(const double (&w)[S], double u, int K, int lane) {
  const int m = lane & (L - 1);
  double T = 0.0;
#pragma unroll
  for (int s = 0; s < S; s++) T += w[s];
  double inc = T;
#pragma unroll
  for (int o = 1; o < L; o <<= 1) {
    const double t = __shfl_up_sync(FULL, inc, o, L);
    if (m >= o) inc += t;
  }
  const double R = u * __shfl_sync(FULL, inc, L - 1, L);
  const unsigned bal = __ballot_sync(FULL, inc <= R);
  const int Lc = __popc((bal >> (lane & ~(L - 1))) & (unsigned)((1ull << L) - 1ull));   // first lane of the chain whose inclusive sum exceeds R
  const double r = R - (inc - T);
  double acc = 0.0;
  int c = 0;
#pragma unroll
  for (int s = 0; s < S; s++) {
    acc += w[s];
    c += (acc <= r) ? 1 : 0;
  }
  const int k = __shfl_sync(FULL, m * S + min(c, S - 1), min(Lc, L - 1), L);
  return (Lc >= L) ? K - 1 : min(k, K - 1);
}

// Uniforms of one (haplotype, stream) for the L lanes of a chain: lane m keeps the four uniforms of loci
// base + 4m .. base + 4m + 3 (one Philox block), base = j & ~(4L - 1)
template <int L>
struct LgUniform {
  double u0, u1, u2, u3;
  __device__ __forceinline__ void refill(const KgwFamParams &P, int hap, int stream, int j, int m) {
    const int jl = (j & ~(4 * L - 1)) + 4 * m;
    if (P.injected_u) {
      const double *src = P.injected_u + ((size_t)hap * P.W * 3 + stream) * (size_t)P.p;   // window block 0 of [N][W][3][p]
      u0 = (jl < P.p) ? src[jl] : 0.0;
      u1 = (jl + 1 < P.p) ? src[jl + 1] : 0.0;
      u2 = (jl + 2 < P.p) ? src[jl + 2] : 0.0;
      u3 = (jl + 3 < P.p) ? src[jl + 3] : 0.0;
    } else {
      const kgw_u32x4 r = kgw_philox_block(P.seed, (uint32_t)hap, (uint32_t)stream, (uint32_t)(jl >> 2));
      u0 = kgw_word_to_uniform(r.x);
      u1 = kgw_word_to_uniform(r.y);
      u2 = kgw_word_to_uniform(r.z);
      u3 = kgw_word_to_uniform(r.w);
    }
  }
  __device__ __forceinline__ double get(int j) const {
    const int q = j & 3;
    const double v = (q & 2) ? ((q & 1) ? u3 : u2) : ((q & 1) ? u1 : u0);
    return __shfl_sync(FULL, v, (j >> 2) & (L - 1), L);
  }
};

// node emission (family_bp.cpp:437-454) from the pattern byte: table when mu is the same at every locus (UMU)
template <bool UMU>
__device__ __forceinline__ double pair_emission(const double *tab, unsigned pt, double mu, int nn) {
  if (UMU) return tab[nn * 256 + pt];
  double v = (pt & 1u) ? 1.0 - mu : mu;
  for (int t = 0; t < nn; t++) v *= ((pt >> (t + 1)) & 1u) ? 1.0 - mu : mu;
  if (nn > 0) v = pow(v, 1.0 / (double)(nn + 1));
  return v;
}

// Staging of a chain's rows through shared memory (both pair kernels that walk a chain locus by locus): a block of RB
// consecutive loci of the chain's pattern rows (and left-message rows) is ONE contiguous run of the pair's scratch slot,
// so lane 0 moves it with one cp.async.bulk per chain and array, completing on the stage's mbarrier.  The runs start at
// any byte (K bytes per pattern row): the copy starts at the 16-byte boundary below and the readers add the offset.
struct PairStage {
  // bytes a buffer needs for `bytes` payload bytes starting at any offset within a 16-byte unit
  __host__ __device__ static constexpr unsigned buf(unsigned bytes) { return (bytes + 15u + 15u) / 16u * 16u; }
};
// issue: copy [src, src + bytes) -- widened to 16-byte boundaries -- to dst; returns the bytes the barrier will see
__device__ __forceinline__ unsigned pair_stage_copy(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(src);
  const unsigned off = (unsigned)(a & 15u), n = (off + bytes + 15u) & ~15u;
  bulk_g2s(dst, reinterpret_cast<const void *>(a - off), n, bar);
  return n;
}
// ---- every remaining node of a pair (family_bp.cpp:217-229), both members at the same locus.
// Straight-line code per state: a state beyond K carries a zero left message and looks its emission up in a zero block
// of the table, so its weight is 0; the first locus carries a unit
// transition factor; row p of a member's left messages -- row 0 of the next member, which belief propagation never
// touches -- holds ones (written by the INIT phase), so the last locus needs no special case either.
// Inputs: blocks of RB loci (pattern rows, left-message rows, per-locus constants) travel by bulk copies into a
// two-stage shared-memory ring a block ahead of their use (register-destination loads a locus or two ahead left the
// kernel waiting for L2: 2.9 long-scoreboard stalls per issue); the Z byte is a plain load two loci ahead.
template <int L, int S, bool UMU>
__global__ void __launch_bounds__(128) pair_draw_kernel(const __grid_constant__ KgwFamParams P, const int RB) {
  constexpr int PW = 32 / (2 * L), G = 2 * PW;   // pairs, chains per warp
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int m = lane & (L - 1), c = (lane / L) & 1, q = lane / (2 * L), g = lane / L;
  const int p = P.p, K = P.K;
  const size_t pK = (size_t)p * K;
  __shared__ double tab[2 * 512];   // [2][256] emissions by (neighbours, pattern), then 512 zeros
  extern __shared__ __align__(16) unsigned char pair_smem[];
  for (int t = threadIdx.x; t < 2 * 512; t += blockDim.x) tab[t] = (UMU && t < 512) ? P.emtab[t] : 0.0;
  // per warp: two stages of { G left-message buffers, G pattern buffers, RB per-locus constants }, two mbarriers
  const unsigned MLB = PairStage::buf((unsigned)RB * K * 8), PTB = PairStage::buf((unsigned)RB * K), LCB = (unsigned)RB * (unsigned)sizeof(KgwFamLocus);
  const unsigned STB = G * (MLB + PTB) + LCB;
  unsigned char *wbase = pair_smem + (size_t)wid * (2 * STB + 16);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(wbase + 2 * STB);
  if (lane == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int f_first = P.f_base + (blockIdx.x * wpc + wid) * PW;
  if (f_first >= P.n_families) return;
  const bool act = f_first + q < P.n_families;
  const int f = act ? f_first + q : P.n_families - 1;   // idle lane groups shadow the batch's last pair and store nothing
  const size_t fslot = (size_t)(f - P.f_base);
  const KgwFamDesc fd = P.fams[P.fam_list[f]];
  const int i_abs = P.members[fd.mem_off + c];
  const KgwFamInterval *iv = P.ivs + P.iv_start[fd.ivs_off + c];
  int ivj1 = iv->j1, ivnn = iv->nn;
  uint8_t *zrow = P.bytes + fslot * 2 * P.n_max * p + (size_t)c * p;
  int toff[S];   // table offset of the state: 0, or 512 (the zero block) beyond K
  bool valid[S];
#pragma unroll
  for (int s = 0; s < S; s++) { valid[s] = (m * S + s < K); toff[s] = valid[s] ? 0 : 512; }
  const double *ml_chain = P.msg + fslot * P.n_max * pK + (size_t)c * pK;   // row 0 of this chain's left messages
  const uint8_t *pt_chain = P.pat + fslot * P.n_max * pK + (size_t)c * pK;
  // lane 0 issues every chain's copies: the chains' base pointers, by shuffle
  unsigned long long mlb[G], ptb[G];
#pragma unroll
  for (int t = 0; t < G; t++) {
    mlb[t] = __shfl_sync(FULL, (unsigned long long)reinterpret_cast<uintptr_t>(ml_chain), t * L);
    ptb[t] = __shfl_sync(FULL, (unsigned long long)reinterpret_cast<uintptr_t>(pt_chain), t * L);
  }
  // block b = loci b*RB .. b*RB+RB-1: pattern rows and constants of those loci, left-message rows of the loci + 1.
  // Lane 0 keeps the source addresses of the next block to request and steps them by one block per call.
  const unsigned ml_blk = (unsigned)RB * K * 8, pt_blk = (unsigned)RB * K;
#pragma unroll
  for (int t = 0; t < G; t++) mlb[t] += (unsigned long long)K * 8;   // left messages: one locus on
  unsigned long long lcsrc = (unsigned long long)reinterpret_cast<uintptr_t>(P.loc);
  int fill_b = 0;
  auto fill = [&]() {   // lane 0: request block fill_b
    const int st = fill_b & 1;
    unsigned char *sb = wbase + (size_t)st * STB;
    unsigned tot = LCB;
#pragma unroll
    for (int t = 0; t < G; t++) {
      tot += (((unsigned)mlb[t] & 15u) + ml_blk + 15u) & ~15u;
      tot += (((unsigned)ptb[t] & 15u) + pt_blk + 15u) & ~15u;
    }
    mbar_expect_tx(bars + st, tot);
#pragma unroll
    for (int t = 0; t < G; t++) {
      pair_stage_copy(sb + (size_t)t * MLB, reinterpret_cast<const void *>(mlb[t]), ml_blk, bars + st);
      pair_stage_copy(sb + (size_t)G * MLB + (size_t)t * PTB, reinterpret_cast<const void *>(ptb[t]), pt_blk, bars + st);
      mlb[t] += ml_blk;
      ptb[t] += pt_blk;
    }
    bulk_g2s(sb + (size_t)G * (MLB + PTB), reinterpret_cast<const void *>(lcsrc), LCB, bars + st);
    lcsrc += LCB;
    fill_b++;
  };
  const int nblocks = (p + RB - 1) / RB;
  if (lane == 0) {
    fill();
    if (nblocks > 1) fill();
  }
  // this lane's view of block b: pointers to its states of the block's first locus
  const double *mls = nullptr;
  const uint8_t *pts = nullptr;
  const KgwFamLocus *lcs = nullptr;
  unsigned long long am = (unsigned long long)reinterpret_cast<uintptr_t>(ml_chain + K), ap = (unsigned long long)reinterpret_cast<uintptr_t>(pt_chain);
  auto enter = [&](int b) {   // every lane: wait for block b, point at its first locus, step the addresses to block b + 1
    const int st = b & 1;
    mbar_wait(bars + st, (unsigned)((b >> 1) & 1));
    unsigned char *sb = wbase + (size_t)st * STB;
    mls = reinterpret_cast<const double *>(sb + (size_t)g * MLB + ((unsigned)am & 15u)) + m * S;
    pts = sb + (size_t)G * MLB + (size_t)g * PTB + ((unsigned)ap & 15u) + m * S;
    lcs = reinterpret_cast<const KgwFamLocus *>(sb + (size_t)G * (MLB + PTB));
    am += ml_blk;
    ap += pt_blk;
  };
  enter(0);
  double ml_c[S], la_c = 1.0, lb_c = 0.0, mu_c = lcs->mu;   // j == 0: no right message (factor 1)
  unsigned pt_c[S];
#pragma unroll
  for (int s = 0; s < S; s++) { pt_c[s] = valid[s] ? (unsigned)pts[s] : 0u; ml_c[s] = valid[s] ? mls[s] : 0.0; }   // (beyond K: stale shared memory)
  int z_c = zrow[0], z_n = zrow[1];
  int zprev = -1;
  LgUniform<L> U;
  int jj = 0, b = 0;            // position inside the block, block
  for (int j = 0; j < p; j++) {
    if ((j & (4 * L - 1)) == 0) U.refill(P, i_abs, STREAM_Z, j, m);
    const double u = U.get(j);
    while (ivj1 < j) { iv++; ivj1 = iv->j1; ivnn = iv->nn; }
    const bool active = (z_c == 255);
    const double lab = la_c + lb_c * 1.0;   // the point-mass right message of the previous draw (:243-246) at k == zprev
    const int zrel = zprev - m * S;
    const int nnoff = ivnn * 256;
    double wv[S];
#pragma unroll
    for (int s = 0; s < S; s++) {
      double v;
      if (UMU) v = tab[(int)pt_c[s] + nnoff + toff[s]];
      else v = valid[s] ? pair_emission<false>(tab, pt_c[s], mu_c, ivnn) : 0.0;
      v *= (s == zrel) ? lab : la_c;
      wv[s] = v * ml_c[s];
    }
    // the next locus's inputs from the ring (a new block: its stage has landed; the stage left behind is refilled)
    const int z_nn = zrow[j + 2];
    if (++jj == RB) {
      jj = 0;
      b++;
      __syncwarp();   // every lane has used the last row of the block left behind
      if (lane == 0 && b + 1 < nblocks) fill();
      if (b < nblocks) enter(b);
    } else {
      mls += K;
      pts += K;
      lcs++;
    }
#pragma unroll
    for (int s = 0; s < S; s++) { pt_c[s] = valid[s] ? (unsigned)pts[s] : 0u; ml_c[s] = valid[s] ? mls[s] : 0.0; }   // (beyond K: stale shared memory)
    la_c = lcs->a_al;
    lb_c = lcs->b;
    mu_c = lcs->mu;
    const int kz = lg_choice<L, S>(wv, u, K, lane);
    const int kz_o = __shfl_xor_sync(FULL, kz, L);
    const bool act_o = __shfl_xor_sync(FULL, (int)active, L) != 0;
    // member 0 is visited first (:217-229): its draw at a shared locus pins member 1's node as well; a draw of member 1
    // there (member 0 pinned earlier, which belief propagation's junction draws never leave behind) overwrites Z[0][j]
    int zst = -1;   // value to store
    int znext = z_c;
    if (active) { zst = kz; znext = kz; }
    if (ivnn == 1 && act_o) {
      if (c == 1) { zst = kz_o; znext = kz_o; }
      else if (!active) zst = kz_o;
    }
    if (act && m == 0 && zst >= 0) zrow[j] = (uint8_t)zst;
    zprev = znext;
    z_c = z_n;
    z_n = z_nn;
  }
}

// ---- the DRAW operations of the pairs of a batch (FamilyKnockoffs::sample_knockoff_MC, family_knockoffs.cpp:295-454).
// Per state the code is straight-line: the three special states (z0, z0k, z1) are selected by compile-time register
// index against a per-lane relative index, states beyond K are masked by a 0/1 factor (their 1/N_old is 0, so T_k/N_old
// vanishes), the clamp N_k >= 1e-10 before the normalisation can never bind (N_k only grows from a clamped value,
// :378-398) and the one after it runs only when the lower bound vbar * N_sum / norm of every N_k is near it.
template <int L, int S>
__global__ void __launch_bounds__(128) pair_ko_kernel(const __grid_constant__ KgwFamParams P, const KgwPairOp *ops, const int n_ops) {
  const int lane = threadIdx.x & 31;
  const int m = lane & (L - 1);
  const int p = P.p, K = P.K;
  bool done = false;
  int j = 0, j0 = 0, j1 = -1, i_abs = 0;
  const uint8_t *z = nullptr;
  uint8_t *zk = nullptr;
  double Nv[S], rNo[S], tq[S], vmask[S];
#pragma unroll
  for (int s = 0; s < S; s++) { vmask[s] = (m * S + s < K) ? 1.0 : 0.0; Nv[s] = 1.0; rNo[s] = vmask[s]; tq[s] = 0.0; }
  int zk_last = 0, zkprev = 0, zm1 = 0, z1 = 0, z1n = 0;
  double kv_a = 0.0, kv_apb = 0.0, kv_vb = 0.0, kv_vs = 0.0;
  int kv_flags = 0, kv_z1next = -1;
  LgUniform<L> U;
  bool fresh = false;
  for (;;) {
    const bool need = !done && j > j1;
    if (__any_sync(FULL, need)) {
      int idx = n_ops;
      if (need && m == 0) idx = atomicAdd(P.work, 1);
      idx = __shfl_sync(FULL, idx, 0, L);
      if (need) {
        if (idx >= n_ops) {
          done = true;
        } else {
          const KgwPairOp op = ops[idx];
          const size_t fslot = (size_t)(op.f - P.f_base);
          const KgwFamDesc fd = P.fams[P.fam_list[op.f]];
          i_abs = P.members[fd.mem_off + op.i];
          uint8_t *zb = P.bytes + fslot * 2 * P.n_max * p;
          z = zb + (size_t)op.i * p;
          zk = zb + (size_t)P.n_max * p + (size_t)op.i * p;
          j0 = op.j0;
          j1 = op.j1;
          j = j0;
#pragma unroll
          for (int s = 0; s < S; s++) { Nv[s] = 1.0; rNo[s] = vmask[s]; }
          zk_last = 0;
          zkprev = 0;
          const KgwLocusKO *kp = P.ko + j0;
          kv_a = kp->a; kv_apb = kp->apb; kv_vb = kp->vb; kv_vs = kp->vs; kv_flags = kp->flags; kv_z1next = kp->z1_next;
          const int zi = kp->z1_idx;
          zm1 = (j0 > 0) ? z[j0 - 1] : 0;
          z1n = (zi >= 0) ? z[zi] : 0;   // becomes z1 at the first locus of the group
          fresh = true;
        }
      }
    }
    if (__all_sync(FULL, done)) break;
    const bool run = !done;
    const int jc = run ? j : 0;   // idle lane groups go through the motions at locus 0 and store nothing
    // next locus: table row and Z[j] (the z0 of a group starting at j+1)
    const bool nx = run && (jc + 1 <= j1);
    const KgwLocusKO *kn = P.ko + (nx ? jc + 1 : jc);
    const double kn_a = kn->a, kn_apb = kn->apb, kn_vb = kn->vb, kn_vs = kn->vs;
    const int kn_flags = kn->flags, kn_z1next = kn->z1_next;
    const int z_at_j = run ? z[jc] : 0;
    if (fresh || (jc & (4 * L - 1)) == 0) {
      U.refill(P, i_abs, STREAM_ZK, jc, m);
      fresh = false;
    }
    const double u = U.get(jc);
    const bool first = (kv_flags & KGW_KO_FIRST) != 0, lastg = (kv_flags & KGW_KO_LASTGRP) != 0, gzero = (kv_flags & KGW_KO_GZERO) != 0;

    // first locus of a group: T_k / N_old[k] and the partition function N (:358-398, :423-436).  ALL: every lane group of
    // the warp is at such a locus (singleton groups: always), results are assigned instead of selected.
    auto group_head = [&](auto all_tag) {
      constexpr bool ALL = decltype(all_tag)::value;
      int z1_pref = 0;
      if (run && first && kv_z1next >= 0) z1_pref = z[kv_z1next];   // z1 of the next group, a group ahead
      const bool gb = (jc == j0);
      // T_k = T(z0) * T(z0k): a * a in general, (a+b) at a state equal to z0 or z0k; first group of the operation: no z0k
      // factor; group 0: a (:358-375)
      const double cg = (gzero || gb) ? kv_a : kv_a * kv_a;
      const double c0 = gzero ? kv_a : (gb ? kv_apb : kv_apb * kv_a);
      const double c1 = (zm1 == zk_last) ? kv_apb * kv_apb : kv_a * kv_apb;
      const int z0rel = gzero ? -1000 : zm1 - m * S;
      const int z0krel = (gzero || gb) ? -1000 : zk_last - m * S;
      double tqs[S], nv[S];
      double part = 0.0;
#pragma unroll
      for (int s = 0; s < S; s++) {
        double cf = (s == z0rel) ? c0 : cg;
        cf = (s == z0krel) ? c1 : cf;
        tqs[s] = cf * rNo[s];
        part += tqs[s];
      }
      const double N_sum = lg_sum<L>(part);
      const double vb_u = lastg ? 0.0 : kv_vb, vs_u = lastg ? 0.0 : kv_vs;
      double nn_part = 0.0;
#pragma unroll
      for (int s = 0; s < S; s++) {
        double x = fma(vb_u, N_sum, Nv[s]);
        x = fma(vs_u, tqs[s], x);
        nv[s] = x;
        nn_part = fma(x, vmask[s], nn_part);
      }
      const double r_norm = fast_rcp(lg_sum<L>(nn_part));
      const bool tight = !(vb_u * N_sum * r_norm > 4.0 * N_MIN);
#pragma unroll
      for (int s = 0; s < S; s++) nv[s] *= r_norm;
      if (__any_sync(FULL, tight)) {
#pragma unroll
        for (int s = 0; s < S; s++) if (nv[s] < N_MIN) nv[s] = N_MIN;
      }
      if (ALL || first) {
#pragma unroll
        for (int s = 0; s < S; s++) { Nv[s] = nv[s]; tq[s] = tqs[s]; }
        z1 = z1n;
        z1n = z1_pref;
      }
    };
    const bool allfirst = __all_sync(FULL, first || !run);
    const double vbs = fma(kv_vs, 1.0, kv_vb);   // vbar + vstar at k == z1 (vstar is 0 in the last group)
    double w[S];
    if (allfirst) {
      // every lane group of the warp is at the first locus of a group (singleton groups: always): the same steps as
      // group_head, in place, and the draw's weights right behind them
      int z1_pref = 0;
      if (run && kv_z1next >= 0) z1_pref = z[kv_z1next];
      const bool gb = (jc == j0);
      const double cg = (gzero || gb) ? kv_a : kv_a * kv_a;
      const double c0 = gzero ? kv_a : (gb ? kv_apb : kv_apb * kv_a);
      const double c1 = (zm1 == zk_last) ? kv_apb * kv_apb : kv_a * kv_apb;
      const int z0rel = gzero ? -1000 : zm1 - m * S;
      const int z0krel = (gzero || gb) ? -1000 : zk_last - m * S;
      double part = 0.0;
#pragma unroll
      for (int s = 0; s < S; s++) {
        double cf = (s == z0rel) ? c0 : cg;
        cf = (s == z0krel) ? c1 : cf;
        tq[s] = cf * rNo[s];
        part += tq[s];
      }
      const double N_sum = lg_sum<L>(part);
      const double vb_u = lastg ? 0.0 : kv_vb, vs_u = lastg ? 0.0 : kv_vs;
      double nn_part = 0.0;
#pragma unroll
      for (int s = 0; s < S; s++) {
        Nv[s] = fma(vs_u, tq[s], fma(vb_u, N_sum, Nv[s]));
        nn_part = fma(Nv[s], vmask[s], nn_part);
      }
      z1 = z1n;
      z1n = z1_pref;
      const int z1rel = z1 - m * S;
#pragma unroll
      for (int s = 0; s < S; s++) w[s] = tq[s] * ((s == z1rel) ? vbs : kv_vb);
      const double r_norm = fast_rcp(lg_sum<L>(nn_part));
      const bool tight = !(vb_u * N_sum * r_norm > 4.0 * N_MIN);
#pragma unroll
      for (int s = 0; s < S; s++) Nv[s] *= r_norm;
      if (__any_sync(FULL, tight)) {
#pragma unroll
        for (int s = 0; s < S; s++) if (Nv[s] < N_MIN) Nv[s] = N_MIN;
      }
    } else {
      if (__any_sync(FULL, run && first)) group_head(std::false_type{});
      const int z1rel = z1 - m * S;
      const int zkrel = zkprev - m * S;
#pragma unroll
      for (int s = 0; s < S; s++) {
        const double wk = first ? tq[s] : ((s == zkrel) ? kv_apb : kv_a) * vmask[s];
        w[s] = wk * ((s == z1rel) ? vbs : kv_vb);
      }
    }
    zkprev = lg_choice<L, S>(w, u, K, lane);
    if (run && m == 0) zk[jc] = (uint8_t)zkprev;
    if (kv_flags & KGW_KO_LASTINGRP) {
      zk_last = zkprev;
#pragma unroll
      for (int s = 0; s < S; s++) rNo[s] = fast_rcp(Nv[s]) * vmask[s];
    }
    if (run) {
      zm1 = z_at_j;
      kv_a = kn_a; kv_apb = kn_apb; kv_vb = kn_vb; kv_vs = kn_vs; kv_flags = kn_flags; kv_z1next = kn_z1next;
      j++;
    }
  }
}

}  // namespace kgw
