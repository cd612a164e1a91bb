// knockoffgwas_b200/csrc/kgw_family.cuh
//
// sm_100a kernel for the IBD-conditioned (related) branch of snpknock2's knockoff sampler:
//   Knockoffs::generate_related_cluster   snpknock2/src/knockoffs.cpp:465-571
//   FamilyBP (loopy belief propagation)   snpknock2/src/family_bp.cpp:79-601
//   FamilyKnockoffs                       snpknock2/src/family_knockoffs.cpp:77-454
//   Knockoffs::sample_emission            snpknock2/src/knockoffs.cpp:1430-1441
//
// Mapping: one warp per IBD family (families are independent, knockoffs.cpp:440-445); lane l owns the
// HMM states k = l, l+32, ... (SF = ceil(K/32) per lane).  The BP messages of a family,
// double [2][members][p][K], and the per-(member, locus, state) emission pattern bytes live in an HBM
// scratch slot owned by the warp; neighbour sets are piecewise constant along the chromosome and come
// from the host as intervals (with the junction differences of family_bp.cpp:352-378 precomputed).
// Node emissions are table look-ups: the reference's pow(prod, 1/(n+1)) (family_bp.cpp:441-454) takes
// one of 2^(n+1) values, tabulated on the host with the reference's own operation order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kgw_device.cuh"

#define KGW_FAM_MAXN 8   // members per family
#define KGW_FAM_MAXNB 7  // neighbours of a node

struct KgwFamInterval {  // maximal run of loci over which neighbors[i][j] is constant (family_bp.cpp:41-61)
  int32_t j0, j1;        // inclusive
  int32_t nn;            // number of neighbours
  int32_t nb[KGW_FAM_MAXNB];   // family-local indices, reference push_back order
  int32_t ndl, ndr;      // sizes of neigh_diff_left at j0 / neigh_diff_right at j1 (family_bp.cpp:365-378)
  int32_t dl[KGW_FAM_MAXNB];
  int32_t dr[KGW_FAM_MAXNB];
  int32_t pad;
};

enum { KGW_FOP_COPYZ = 0, KGW_FOP_DRAW = 1, KGW_FOP_COPYZK = 2 };
struct KgwFamOp {  // FamilyKnockoffs::run (family_knockoffs.cpp:155-270) flattened in execution order
  int32_t kind;    // COPYZ: Zk[i][a..b] = Z[i][a..b];  DRAW: sample_knockoff_MC(i, groups a..b);  COPYZK: Zk[i][a..b] = Zk[src][a..b]
  int32_t i, a, b, src;
  int32_t pad[3];
};

struct KgwFamDesc {
  int32_t n;          // members
  int32_t mem_off;    // into members[] and iv_start[] (iv_start has one extra entry per family)
  int32_t op_off, n_ops;
  int32_t out_off;    // row offset of this family in the debug outputs
  int32_t pad[3];
};

struct KgwFamLocus {  // per locus
  double b;      // b_j
  double a_bp;   // (1.0 - b_j) / (double)K        family_bp.cpp:458,530
  double a_al;   // (1.0 - b_j) * alpha            family_bp.cpp:243,249; family_knockoffs.cpp
  double mu;     // mut_rates[j]
};

struct KgwFamParams {
  int32_t N, p, K, n_families, n_groups, stride_w, n_max, uniform_mu;
  const uint32_t *Hw;
  const int32_t *ref_global;   // [N][K]
  const KgwFamDesc *fams;
  const int32_t *members;      // absolute ids
  const int32_t *iv_start;     // per (family, member): first interval; one sentinel per family
  const KgwFamInterval *ivs;
  const KgwFamOp *ops;
  const KgwFamLocus *loc;      // [p]
  const KgwLocusKO *ko;        // [p] single-window table (vstar / vbar / flags of every locus)
  const int32_t *groups;       // [p]
  const int32_t *grp_first;    // [n_groups]
  const int32_t *grp_last;     // [n_groups]
  const double *emtab;         // [8][256] node emission by (neighbours, match pattern); uniform mu only
  uint32_t *Hk;                // [N][stride_w]
  double *msg;                 // [slots][2][n_max][p][K]
  uint8_t *pat;                // [slots][n_max][p][K]
  uint8_t *bytes;              // [slots][3][n_max][p]   inactive, Z, Zk
  unsigned long long seed;
  const double *injected_u;    // [N][3][p] or null
  int32_t *dbg_z, *dbg_zk;     // [members total][p] or null
  int32_t *status;             // 0 ok; -10 missing knockoff variants; -11 inconsistency on Zk
};

namespace kgw {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

__device__ __forceinline__ double fam_uniform(const KgwFamParams &P, int hap, int stream, int j) {
  if (P.injected_u) return P.injected_u[((size_t)hap * 3 + stream) * (size_t)P.p + j];
  const kgw_u32x4 r = kgw_philox_block(P.seed, (uint32_t)hap, (uint32_t)stream, (uint32_t)(j >> 2));
  const int q = j & 3;
  return kgw_word_to_uniform(q == 0 ? r.x : (q == 1 ? r.y : (q == 2 ? r.z : r.w)));
}

// weighted_choice (utils.cpp:690-698) over states k = s*32 + lane: first k with u*S < cumulative weight
template <int SF>
__device__ __forceinline__ int warp_choice(const double (&w)[SF], double u, int K, int lane) {
  double tot = 0.0;
  double inc[SF];
#pragma unroll
  for (int s = 0; s < SF; s++) {
    const double c = warp_incl_scan(w[s], lane);
    inc[s] = tot + c;
    tot += __shfl_sync(FULL, c, 31);
  }
  const double R = u * tot;
  int cnt = 0;
#pragma unroll
  for (int s = 0; s < SF; s++) cnt += (s * 32 + lane < K && inc[s] <= R) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
  return min(cnt, K - 1);
}

template <int SF>
__global__ void __launch_bounds__(128) family_kernel(const __grid_constant__ KgwFamParams P) {
  const int lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const int slot = blockIdx.x * wpc + (threadIdx.x >> 5);
  const int nslots = gridDim.x * wpc;
  const int p = P.p, K = P.K;
  const size_t pK = (size_t)p * K;
  double *mr_base = P.msg + (size_t)slot * 2 * P.n_max * pK;   // messages_right [n][p][K]
  double *ml_base = mr_base + (size_t)P.n_max * pK;            // messages_left
  uint8_t *pat_base = P.pat + (size_t)slot * P.n_max * pK;
  uint8_t *inact = P.bytes + (size_t)slot * 3 * P.n_max * p;
  uint8_t *Zb = inact + (size_t)P.n_max * p;
  uint8_t *Zkb = Zb + (size_t)P.n_max * p;

  extern __shared__ __align__(16) unsigned char fam_smem[];
  double *tab = reinterpret_cast<double *>(fam_smem) + (size_t)(threadIdx.x >> 5) * 8 * 256;
  if (P.uniform_mu) {
    for (int t = lane; t < 8 * 256; t += 32) tab[t] = P.emtab[t];
  }
  __syncwarp();

  for (int f = slot; f < P.n_families; f += nslots) {
    const KgwFamDesc fd = P.fams[f];
    const int n = fd.n;
    const int32_t *mem = P.members + fd.mem_off;
    const int32_t *ivs0 = P.iv_start + fd.mem_off + f;   // n+1 entries for this family

    auto find_iv = [&](int i, int j) -> const KgwFamInterval * {
      const KgwFamInterval *iv = P.ivs + ivs0[i];
      while (iv->j1 < j) iv++;
      return iv;
    };
    // emission of node (i, j) for state k (family_bp.cpp:437-454)
    auto node_emission = [&](int i, int j, int k, int nn) -> double {
      const unsigned pt = pat_base[((size_t)i * p + j) * K + k];
      if (P.uniform_mu) return tab[nn * 256 + pt];
      const double mu = P.loc[j].mu;
      double v = (pt & 1u) ? 1.0 - mu : mu;
      for (int t = 0; t < nn; t++) v *= ((pt >> (t + 1)) & 1u) ? 1.0 - mu : mu;
      if (nn > 0) v = pow(v, 1.0 / (double)(nn + 1));
      return v;
    };

    // ---- initialise: match patterns, messages 1/K (family_bp.cpp:421-427), nothing inactive, Z unset
    const double p0 = 1.0 / ((double)K);
    for (int i = 0; i < n; i++) {
      const int i_abs = mem[i];
      const int32_t *ref = P.ref_global + (size_t)i_abs * K;
      for (int j = lane; j < p; j += 32) { inact[(size_t)i * p + j] = 0; Zb[(size_t)i * p + j] = 255; Zkb[(size_t)i * p + j] = 255; }
      for (int k = lane; k < K; k += 32) {
        const size_t rrow = (size_t)__ldg(ref + k) * P.stride_w;
        const KgwFamInterval *iv = P.ivs + ivs0[i];
        for (int w = 0; w <= (p - 1) >> 5; w++) {
          const uint32_t rw = __ldg(P.Hw + rrow + w);
          const uint32_t m0 = ~(rw ^ __ldg(P.Hw + (size_t)i_abs * P.stride_w + w));
          const int jhi = min((w << 5) + 32, p);
          for (int j = w << 5; j < jhi; j++) {
            while (iv->j1 < j) iv++;
            unsigned pt = (m0 >> (j & 31)) & 1u;
            for (int t = 0; t < iv->nn; t++) {
              const uint32_t hw2 = __ldg(P.Hw + (size_t)mem[iv->nb[t]] * P.stride_w + w);
              pt |= ((~(rw ^ hw2) >> (j & 31)) & 1u) << (t + 1);
            }
            const size_t o = ((size_t)i * p + j) * K + k;
            pat_base[o] = (uint8_t)pt;
            mr_base[o] = p0;
            ml_base[o] = p0;
          }
        }
      }
    }
    __syncwarp();

    // one BP message (family_bp.cpp:429-559): dir 0 = right (writes messages_right[i][j]), 1 = left.
    // Returns this lane's share of sum_k |new - old|.
    auto send_message = [&](int i, int j, int dir, const KgwFamInterval *iv) -> double {
      const bool jl = (j == iv->j0) && j > 0, jr = (j == iv->j1) && j < p - 1;
      const int ndl = jl ? iv->ndl : 0, ndr = jr ? iv->ndr : 0;
      const KgwFamLocus lc = P.loc[dir == 0 ? j + 1 : j];
      double cst[SF], msg[SF];
      double csum = 0.0;
#pragma unroll
      for (int s = 0; s < SF; s++) {
        const int k = s * 32 + lane;
        cst[s] = 0.0;
        msg[s] = 0.0;
        if (k < K) {
          const double fk = node_emission(i, j, k, iv->nn);
          double c = lc.a_bp * fk, m = lc.b * fk;
          if (dir == 0) {
            if (j > 0) { const double v = mr_base[((size_t)i * p + j - 1) * K + k]; c *= v; m *= v; }
            for (int t = 0; t < ndl; t++) { const double v = mr_base[((size_t)iv->dl[t] * p + j - 1) * K + k]; c *= v; m *= v; }
            for (int t = 0; t < ndr; t++) { const double v = ml_base[((size_t)iv->dr[t] * p + j + 1) * K + k]; c *= v; m *= v; }
          } else {
            if (j < p - 1) { const double v = ml_base[((size_t)i * p + j + 1) * K + k]; c *= v; m *= v; }
            for (int t = 0; t < ndr; t++) { const double v = ml_base[((size_t)iv->dr[t] * p + j + 1) * K + k]; c *= v; m *= v; }
            for (int t = 0; t < ndl; t++) { const double v = mr_base[((size_t)iv->dl[t] * p + j - 1) * K + k]; c *= v; m *= v; }
          }
          cst[s] = c;
          msg[s] = m;
          csum += c;
        }
      }
      const double const_msg = warp_sum(csum);
      double msum = 0.0;
#pragma unroll
      for (int s = 0; s < SF; s++) {
        if (s * 32 + lane < K) { msg[s] = const_msg + msg[s]; msum += msg[s]; }
      }
      const double inv = 1.0 / warp_sum(msum);
      double *dst = (dir == 0 ? mr_base : ml_base) + ((size_t)i * p + j) * K;
      double diff = 0.0;
#pragma unroll
      for (int s = 0; s < SF; s++) {
        const int k = s * 32 + lane;
        if (k < K) {
          const double v = msg[s] * inv;
          diff += fabs(v - dst[k]);
          dst[k] = v;
        }
      }
      return diff;
    };

    // run_bp (family_bp.cpp:318-419)
    auto run_bp = [&]() {
      for (int it = 0; it < 20; it++) {
        double diff = 0.0;
        for (int dir = 0; dir < 2; dir++) {
          double dt = 0.0;
          for (int jj = 0; jj < p; jj++) {
            const int j = (dir == 0) ? jj : p - 1 - jj;
            if (dir == 0 ? (j >= p - 1) : (j <= 0)) continue;
            for (int i = 0; i < n; i++) {
              if (inact[(size_t)i * p + j]) continue;
              dt += send_message(i, j, dir, find_iv(i, j));
            }
            __syncwarp();
          }
          dt = warp_sum(dt) / (double)((size_t)p * (size_t)n);
          diff += dt;
        }
        if (diff <= 1e-3) break;
      }
    };

    // condition_on (family_bp.cpp:236-255)
    auto condition_on = [&](int i, int j, int kz, bool forward) {
      for (int l = lane; l < K; l += 32) {
        if (j < p - 1) {
          const KgwFamLocus lc = P.loc[j + 1];
          mr_base[((size_t)i * p + j) * K + l] = lc.a_al + lc.b * (double)(l == kz);
        }
        if (!forward && j > 0) {
          const KgwFamLocus lc = P.loc[j];
          ml_base[((size_t)i * p + j) * K + l] = lc.a_al + lc.b * (double)(l == kz);
        }
      }
      if (lane == 0) inact[(size_t)i * p + j] = 1;
    };

    // sample_posterior_joint (family_bp.cpp:569-601)
    auto draw_node = [&](int i, int j, const KgwFamInterval *iv, bool junction) -> int {
      const bool jl = junction && (j == iv->j0) && j > 0, jr = junction && (j == iv->j1) && j < p - 1;
      const int ndl = jl ? iv->ndl : 0, ndr = jr ? iv->ndr : 0;
      double w[SF];
#pragma unroll
      for (int s = 0; s < SF; s++) {
        const int k = s * 32 + lane;
        w[s] = 0.0;
        if (k < K) {
          double v = node_emission(i, j, k, iv->nn);
          if (j > 0) {
            v *= mr_base[((size_t)i * p + j - 1) * K + k];
            for (int t = 0; t < ndl; t++) v *= mr_base[((size_t)iv->dl[t] * p + j - 1) * K + k];
          }
          if (j < p - 1) {
            v *= ml_base[((size_t)i * p + j + 1) * K + k];
            for (int t = 0; t < ndr; t++) v *= ml_base[((size_t)iv->dr[t] * p + j + 1) * K + k];
          }
          w[s] = v;
        }
      }
      return warp_choice<SF>(w, fam_uniform(P, mem[i], STREAM_Z, j), K, lane);
    };

    // ---- FamilyBP::run (family_bp.cpp:79-100) and sample_posterior (:160-230)
    run_bp();
    for (int i = 0; i < n; i++) {
      for (int j = 0; j < p; j++) {
        if (inact[(size_t)i * p + j]) continue;
        const KgwFamInterval *iv = find_iv(i, j);
        const bool jl = (j == iv->j0) && j > 0, jr = (j == iv->j1) && j < p - 1;
        if (!((jl && iv->ndl > 0) || (jr && iv->ndr > 0))) continue;  // not a junction (:176-193)
        const int kz = draw_node(i, j, iv, true);
        __syncwarp();
        if (lane == 0) Zb[(size_t)i * p + j] = (uint8_t)kz;
        condition_on(i, j, kz, false);
        for (int t = 0; t < iv->nn; t++) {
          const int i2 = iv->nb[t];
          if (lane == 0) Zb[(size_t)i2 * p + j] = (uint8_t)kz;
          condition_on(i2, j, kz, false);
        }
        __syncwarp();
        run_bp();
      }
    }
    for (int i = 0; i < n; i++) {
      for (int j = 0; j < p; j++) {
        if (inact[(size_t)i * p + j]) continue;
        const KgwFamInterval *iv = find_iv(i, j);
        const int kz = draw_node(i, j, iv, false);
        __syncwarp();
        if (lane == 0) Zb[(size_t)i * p + j] = (uint8_t)kz;
        condition_on(i, j, kz, true);
        for (int t = 0; t < iv->nn; t++) {
          const int i2 = iv->nb[t];
          if (lane == 0) Zb[(size_t)i2 * p + j] = (uint8_t)kz;
          condition_on(i2, j, kz, true);
        }
        __syncwarp();
      }
    }

    // ---- FamilyKnockoffs::run (family_knockoffs.cpp:77-293) as a flat op list
    const KgwFamOp *ops = P.ops + fd.op_off;
    for (int q = 0; q < fd.n_ops; q++) {
      const KgwFamOp op = ops[q];
      if (op.kind == KGW_FOP_COPYZ) {
        for (int j = op.a + lane; j <= op.b; j += 32) Zkb[(size_t)op.i * p + j] = Zb[(size_t)op.i * p + j];
      } else if (op.kind == KGW_FOP_COPYZK) {
        for (int j = op.a + lane; j <= op.b; j += 32) Zkb[(size_t)op.i * p + j] = Zkb[(size_t)op.src * p + j];
      } else {
        // FamilyKnockoffs::sample_knockoff_MC (family_knockoffs.cpp:295-454), groups op.a .. op.b
        const int i = op.i, g_begin = op.a, g_end = op.b;
        const uint8_t *z = Zb + (size_t)i * p;
        uint8_t *zk = Zkb + (size_t)i * p;
        double Nv[SF], No[SF];
#pragma unroll
        for (int s = 0; s < SF; s++) { Nv[s] = 1.0; No[s] = 1.0; }
        for (int g = g_begin; g <= g_end; g++) {
          const int e0 = P.grp_first[g], e1 = P.grp_last[g];
          const KgwLocusKO k0v = P.ko[e0];
          const bool lastg = (g == P.n_groups - 1);
          const int z0 = (g > 0) ? z[e0 - 1] : 0;
          const int z0k = (g > 0) ? zk[e0 - 1] : 0;
          const int z1 = lastg ? 0 : z[e1 + 1];
          // T_k for the first locus of the group (:358-375, :378-398, :423-436)
          double tq[SF];
          double part = 0.0;
#pragma unroll
          for (int s = 0; s < SF; s++) {
            const int k = s * 32 + lane;
            double tmp = k0v.a;
            if (g > 0) {
              tmp = (k == z0) ? k0v.apb : k0v.a;
              if (g != g_begin) tmp = tmp * ((k == z0k) ? k0v.apb : k0v.a);
            }
            tq[s] = tmp;
            if (k < K) part += (g == 0) ? tmp : tmp / No[s];
          }
          const double N_sum = warp_sum(part);
          double nn_part = 0.0;
#pragma unroll
          for (int s = 0; s < SF; s++) {
            const int k = s * 32 + lane;
            if (!lastg) {
              Nv[s] += k0v.vb * N_sum;
              Nv[s] += (g == 0) ? (k0v.vs * tq[s]) : (k0v.vs * tq[s] / No[s]);   // vstar[0] == vs of the first locus
            }
            if (Nv[s] < N_MIN) Nv[s] = N_MIN;
            if (k < K) nn_part += Nv[s];
          }
          const double N_norm = warp_sum(nn_part);
#pragma unroll
          for (int s = 0; s < SF; s++) {
            Nv[s] /= N_norm;
            if (Nv[s] < N_MIN) Nv[s] = N_MIN;
          }
          int zkprev = 0;
          for (int j = e0; j <= e1; j++) {
            const KgwLocusKO kv = P.ko[j];
            double w[SF];
#pragma unroll
            for (int s = 0; s < SF; s++) {
              const int k = s * 32 + lane;
              double wk;
              if (j > e0) wk = (k == zkprev) ? kv.apb : kv.a;
              else wk = tq[s] / No[s];
              wk *= lastg ? kv.vb : (kv.vb + kv.vs * (double)(k == z1));
              w[s] = (k < K) ? wk : 0.0;
            }
            zkprev = warp_choice<SF>(w, fam_uniform(P, mem[i], STREAM_ZK, j), K, lane);
            if (lane == 0) zk[j] = (uint8_t)zkprev;
          }
          __syncwarp();
#pragma unroll
          for (int s = 0; s < SF; s++) No[s] = Nv[s];
        }
      }
      __syncwarp();
    }

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
            if (Zkb[(size_t)i * p + j] != Zkb[(size_t)iv->nb[t] * p + j]) bad |= 2;
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
  }
}

}  // namespace kgw
