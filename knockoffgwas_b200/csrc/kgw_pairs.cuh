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

#ifndef KGW_DRAW_RING
#define KGW_DRAW_RING 2   // loci between the request of a draw's inputs and the draw (measured: 1 -> 40.6 ms, 2 -> 38.0 ms, 4 -> 50 ms
#endif                    // on 2368 pairs x 10 000 SNPs, K=50: deeper rings put far-ahead and near-term loads on one scoreboard)

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

// ---- every remaining node of a pair (family_bp.cpp:217-229), both members at the same locus.
// Straight-line code per state, constant offsets from one running pointer per array: a state beyond K reads past the
// row (the next row, or the padding behind the last slot) and looks its emission up in a zero block of the table, so
// its weight is 0; the first locus carries a unit transition factor; row p of a member's left messages -- row 0 of the
// next member, which belief propagation never touches -- holds ones (written by the INIT phase), so the last locus
// needs no special case either.
template <int L, int S, bool UMU>
__global__ void __launch_bounds__(128) pair_draw_kernel(const __grid_constant__ KgwFamParams P) {
  constexpr int PW = 32 / (2 * L);   // pairs per warp
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int m = lane & (L - 1), c = (lane / L) & 1, q = lane / (2 * L);
  const int p = P.p, K = P.K;
  const size_t pK = (size_t)p * K;
  __shared__ double tab[2 * 512];   // [2][256] emissions by (neighbours, pattern), then 512 zeros
  for (int t = threadIdx.x; t < 2 * 512; t += blockDim.x) tab[t] = (UMU && t < 512) ? P.emtab[t] : 0.0;
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
  // this lane's states m*S .. m*S+S-1 of the current locus's rows
  const double *mlp = P.msg + fslot * P.n_max * pK + (size_t)c * pK + m * S;
  const uint8_t *ptp = P.pat + fslot * P.n_max * pK + (size_t)c * pK + m * S;
  const KgwFamLocus *lp = P.loc;   // [p + 1]

  // The per-locus inputs (pattern bytes, left-message entries, transition constants, the Z byte) sit in a register
  // ring D loci deep: a row is requested D loci before its draw, which covers an L2 hit; L2 prefetches run 12 + D ahead.
  // Rows behind the last locus (the ring runs past it) are the next member's rows or the padding: never used.
  constexpr int D = KGW_DRAW_RING;
  double mlr[D][S], lar[D], lbr[D], mur[D];
  unsigned ptr_[D][S];
  int zr[D];
#pragma unroll
  for (int d = 0; d < D; d++) {
#pragma unroll
    for (int s = 0; s < S; s++) { ptr_[d][s] = ptp[(size_t)d * K + s]; mlr[d][s] = mlp[(size_t)(d + 1) * K + s]; }
    lar[d] = (d == 0) ? 1.0 : lp[min(d, p)].a_al;   // j == 0: no right message (factor 1)
    lbr[d] = (d == 0) ? 0.0 : lp[min(d, p)].b;
    mur[d] = lp[min(d, p)].mu;
    zr[d] = zrow[d];
  }
  int zprev = -1;
  LgUniform<L> U;
  const int pf_off = m * 128;   // this lane's line of a row being prefetched
  for (int jb = 0; jb < p; jb += D) {
#pragma unroll
    for (int d = 0; d < D; d++) {
      const int j = jb + d;
      if (j >= p) break;
      if (j + 12 + D < p) {
        const char *row = reinterpret_cast<const char *>(mlp - m * S + (size_t)(12 + D) * K);
        if (pf_off < K * 8) prefetch_l2(row + pf_off);
        if (L * 128 < 256 * 8 && pf_off + L * 128 < K * 8) prefetch_l2(row + pf_off + L * 128);
        if (m == 0) prefetch_l2(ptp + (size_t)(12 + D) * K);
      }
      if ((j & (4 * L - 1)) == 0) U.refill(P, i_abs, STREAM_Z, j, m);
      const double u = U.get(j);
      while (ivj1 < j) { iv++; ivj1 = iv->j1; ivnn = iv->nn; }
      const int z_c = zr[d];
      const bool active = (z_c == 255);
      const double la_c = lar[d];
      const double lab = la_c + lbr[d] * 1.0;   // the point-mass right message of the previous draw (:243-246) at k == zprev
      const int zrel = zprev - m * S;
      const int nnoff = ivnn * 256;
      double wv[S];
#pragma unroll
      for (int s = 0; s < S; s++) {
        double v;
        if (UMU) v = tab[(int)ptr_[d][s] + nnoff + toff[s]];
        else v = valid[s] ? pair_emission<false>(tab, ptr_[d][s], mur[d], ivnn) : 0.0;
        v *= (s == zrel) ? lab : la_c;
        wv[s] = v * mlr[d][s];
      }
      // refill the ring entry: locus j + D (pattern, Z, constants) and left message j + D + 1
      {
        const uint8_t *pn = ptp + (size_t)D * K;
        const double *mn = mlp + (size_t)(D + 1) * K;
#pragma unroll
        for (int s = 0; s < S; s++) { ptr_[d][s] = pn[s]; mlr[d][s] = mn[s]; }
        const KgwFamLocus *ln = lp + min(D, p - j);
        lar[d] = ln->a_al;
        lbr[d] = ln->b;
        mur[d] = ln->mu;
        zr[d] = zrow[j + D];
      }
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
      mlp += K;
      ptp += K;
      lp++;
    }
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
    const int z1rel_old = z1 - m * S;

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
    if (allfirst) group_head(std::true_type{});
    else if (__any_sync(FULL, run && first)) group_head(std::false_type{});

    const int z1rel = z1 - m * S;
    (void)z1rel_old;
    const double vbs = fma(kv_vs, 1.0, kv_vb);   // vbar + vstar at k == z1 (vstar is 0 in the last group)
    double w[S];
    if (allfirst) {
#pragma unroll
      for (int s = 0; s < S; s++) w[s] = tq[s] * ((s == z1rel) ? vbs : kv_vb);
    } else {
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
