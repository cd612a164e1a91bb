// knockoffgwas_b200/csrc/kgw_strict.cu -- audit mode of the unrelated sampler (kgw_generate_strict, include/kgw.h).
//
// The product kernels (kgw_device.cuh) split the K states of a haplotype over registers and lanes and sum them as a
// tree, with fused multiply-adds; the reference sums every quantity over k left to right in IEEE double
// (backward_const, backwardSum: knockoffs.cpp:1052-1073; N_sum, N_norm: :1345-1390; weights_sum: utils.cpp:690-698).
// A categorical draw can differ between the two only where u * S falls within a few ulps of a prefix sum.  This file
// is the other side of that statement: the same sampler, one thread per haplotype, every sum sequential over k and
// every product in the reference's left-to-right order, each operation a correctly rounded intrinsic so that nothing
// is contracted.  It is slow by construction (a thread walks p x K tables in HBM) and exists so that the fast path can
// be audited on the device against the reference's arithmetic: tests/test_gpu_strict.py requires it to equal the
// strict-IEEE CPU restatement bit for bit (paths, knockoffs and every backward message), and the fast path to equal it.
//
// Per-thread tables live in HBM, interleaved over the threads of a launch (element idx of thread t at idx * T + t) so
// that a warp's accesses coalesce.  The reference's worker scratch is reproduced as it is: the latent paths persist
// from one window pass to the next (sample_knockoff_MC reads z and zk of the group before the padded window,
// knockoffs.cpp:1360-1376).
#include "kgw_common.h"
#include "kgw_philox.h"
#include "../../include/kgw.h"

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

int kgw_fail_(int code, const std::string &msg);   // kgw_api.cu

namespace {

__device__ __forceinline__ double dA(double x, double y) { return __dadd_rn(x, y); }
__device__ __forceinline__ double dS(double x, double y) { return __dsub_rn(x, y); }
__device__ __forceinline__ double dM(double x, double y) { return __dmul_rn(x, y); }
__device__ __forceinline__ double dD(double x, double y) { return __ddiv_rn(x, y); }

struct StrictParams {
  const uint8_t *H;
  long long row_bytes;
  const int32_t *ref_local;
  const double *b, *mu;
  const int32_t *groups, *grp_first, *grp_last, *win_start, *win_end;
  int N, p, K, W, max_group;
  unsigned long long seed;
  const double *injected_u;
  const int32_t *haps;   // the haplotypes of this launch
  int n, T;              // how many, and the interleave (threads of the launch)
  double *backward, *weights, *Nv, *Nold, *vstar, *vbar;
  int32_t *z, *zk;
  uint8_t *hk;
  uint8_t *hk_rows;      // [n][row_bytes]
  int32_t *z_out, *zk_out;   // [n][p] or null
  double *beta_out;          // [n][p][K] or null
};

struct Thread {
  const StrictParams &P;
  const int t, i;
  const size_t T;
  __device__ Thread(const StrictParams &P_, int t_) : P(P_), t(t_), i(P_.haps[t_]), T((size_t)P_.T) {}
  __device__ size_t ix(size_t idx) const { return idx * T + (size_t)t; }
  // chaplotype::operator[] (chaplotype.h:25)
  __device__ int bit(int row, int j) const { return (P.H[(size_t)row * P.row_bytes + (j >> 3)] >> (j & 7)) & 1; }
  // Knockoffs::mutate (knockoffs.cpp:1626-1629)
  __device__ static double mutate(int h, double mut) { return h == 1 ? dS(1.0, mut) : mut; }
  __device__ double uniform(int stream, int j, int w) const {
    if (P.injected_u) return P.injected_u[(((size_t)i * P.W + w) * 3 + stream) * (size_t)P.p + j];
    const kgw_u32x4 r = kgw_philox_block(P.seed, (uint32_t)i, (uint32_t)stream | ((uint32_t)w << 2), (uint32_t)(j >> 2));
    const int q = j & 3;
    return kgw_word_to_uniform(q == 0 ? r.x : (q == 1 ? r.y : (q == 2 ? r.z : r.w)));
  }
  // weighted_choice (utils.cpp:690-698) on the thread's weights
  __device__ int choice(int n, double u) const {
    double sum = 0.0;
    for (int k = 0; k < n; k++) sum = dA(sum, P.weights[ix(k)]);
    double R = dM(u, sum);
    for (int k = 0; k < n; k++) {
      const double w = P.weights[ix(k)];
      if (R < w) return k;
      R = dS(R, w);
    }
    return n - 1;
  }
  __device__ void window_pad(int w, int &js, int &je) const {   // knockoffs.cpp:1222-1226, :1276-1280
    js = P.win_start[w];
    je = P.win_end[w];
    if (w > 0) js = P.win_start[w - 1];
    if (w < P.W - 1) je = P.win_end[w + 1];
  }
  __device__ double &bw(int j, int k) const { return P.backward[ix((size_t)j * P.K + k)]; }

  // compute_hmm_backward (knockoffs.cpp:1041-1075)
  __device__ void backward(const int32_t *ref, int js, int je) const {
    const int K = P.K;
    const double alpha = dD(1.0, (double)K);   // knockoffs.cpp:75
    for (int k = 0; k < K; k++) bw(je - 1, k) = dD(1.0, (double)K);
    for (int j = je - 2; j >= js; j--) {
      const int h = bit(i, j + 1);
      const double bj = P.b[j + 1], mut = P.mu[j + 1];
      const double a = dM(dS(1.0, bj), alpha), om = dS(1.0, mut);
      double bc = 0.0;
      for (int k = 0; k < K; k++) {   // :1052-1060
        const double e = (h == bit(ref[k], j + 1)) ? om : mut;
        bc = dA(bc, dM(dM(bw(j + 1, k), a), e));
      }
      double sum = 0.0;
      for (int k = 0; k < K; k++) {   // :1062-1068
        const double e = (h == bit(ref[k], j + 1)) ? om : mut;
        const double v = dA(bc, dM(dM(bj, bw(j + 1, k)), e));
        bw(j, k) = v;
        sum = dA(sum, v);
      }
      for (int k = 0; k < K; k++) bw(j, k) = dD(bw(j, k), sum);   // :1069-1073
    }
  }

  // sample_posterior_MC (knockoffs.cpp:1215-1261)
  __device__ void posterior(int w, const int32_t *ref) const {
    const int K = P.K;
    const double alpha = dD(1.0, (double)K);
    int js, je;
    window_pad(w, js, je);
    backward(ref, js, je);
    {
      const double a0 = dM(dS(1.0, P.b[js]), alpha);
      const int h = bit(i, js);
      for (int k = 0; k < K; k++) {   // :1233-1242
        const double th = mutate(bit(ref[k], js), P.mu[js]);
        P.weights[ix(k)] = dM(dM(a0, h == 1 ? th : dS(1.0, th)), bw(js, k));
      }
      P.z[ix(js)] = choice(K, uniform(KGW_STREAM_Z, js, w));
    }
    for (int j = js + 1; j < je; j++) {   // :1245-1260
      const double bj = P.b[j], mut = P.mu[j];
      const double a = dM(dS(1.0, bj), alpha);
      const int h = bit(i, j), zp = P.z[ix(j - 1)];
      for (int k = 0; k < K; k++) {
        const double th = mutate(bit(ref[k], j), mut);
        const double tr = dA(a, dM(bj, k == zp ? 1.0 : 0.0));
        P.weights[ix(k)] = dM(dM(tr, h == 1 ? th : dS(1.0, th)), bw(j, k));
      }
      P.z[ix(j)] = choice(K, uniform(KGW_STREAM_Z, j, w));
    }
  }

  // (1 - b) alpha + b [k == z], squared over the original and the knockoff path: the factor of :1350-1355, :1368-1372, :1404-1406
  __device__ static double pair_factor(double b, double alpha, int k, int z0, int z0k) {
    double tmp = dA(dM(dS(1.0, b), alpha), dM(b, k == z0 ? 1.0 : 0.0));
    tmp = dM(tmp, dA(dM(dS(1.0, b), alpha), dM(b, k == z0k ? 1.0 : 0.0)));
    return tmp;
  }

  // sample_knockoff_MC (knockoffs.cpp:1270-1428); groups are contiguous runs (set_groups, :96-107)
  __device__ void knockoff(int w) const {
    const int K = P.K;
    const double alpha = dD(1.0, (double)K);
    const double *b = P.b;
    int js, je;
    window_pad(w, js, je);
    const int g_start = P.groups[js], g_end = P.groups[je - 1] + 1;
    for (int k = 0; k < K; k++) { P.Nv[ix(k)] = 1.0; P.Nold[ix(k)] = 1.0; }   // :1293-1294
    const double N_min = 1.0e-10;
    for (int g = g_start; g < g_end; g++) {
      const int e0 = P.grp_first[g], size = P.grp_last[g] - e0 + 1;
      for (int s = size - 1; s >= 0; s--) {   // vstar (:1298-1313)
        if (s < size - 1) P.vstar[ix(s)] = dM(P.vstar[ix(s + 1)], b[e0 + s + 1]);
        else P.vstar[ix(s)] = (g < g_end - 1) ? b[e0] : 0.0;
      }
      for (int s = size - 1; s >= 0; s--) {   // vbar (:1315-1339)
        double sum_a = 0.0;
        if (s < size - 1)
          for (int l = 0; l < K; l++) sum_a = dA(sum_a, dM(dS(1.0, b[e0 + s + 1]), alpha));
        for (int x = 0; x < K; x++) {
          double v;
          if (s < size - 1) {
            v = dM(P.vbar[ix((size_t)(s + 1) * K + x)], dA(sum_a, b[e0 + s + 1]));
            v = dA(v, dM(dM(P.vstar[ix(s + 1)], dS(1.0, b[e0 + s + 1])), alpha));
          } else {
            v = (g < g_end - 1) ? dM(dS(1.0, b[P.grp_first[g + 1]]), alpha) : 1.0;
          }
          P.vbar[ix((size_t)s * K + x)] = v;
        }
      }
      const double b0 = b[e0];
      int z0 = 0, z0k = 0;
      if (g > 0) { z0 = P.z[ix(P.grp_last[g - 1])]; z0k = P.zk[ix(P.grp_last[g - 1])]; }
      double N_sum = 0.0;   // :1345-1358
      for (int k = 0; k < K; k++) {
        if (g == g_start) N_sum = dA(N_sum, dM(dS(1.0, b0), alpha));
        else N_sum = dA(N_sum, dD(pair_factor(b0, alpha, k, z0, z0k), P.Nold[ix(k)]));
      }
      if (g < g_end - 1) {   // partition function (:1360-1376); note g == 0 here against g == g_start above, as in the reference
        const double vs0 = P.vstar[ix(0)];
        for (int k = 0; k < K; k++) {
          double n = dA(P.Nv[ix(k)], dM(P.vbar[ix(k)], N_sum));
          if (g == 0) n = dA(n, dM(dM(vs0, dS(1.0, b0)), alpha));
          else n = dA(n, dD(dM(vs0, pair_factor(b0, alpha, k, z0, z0k)), P.Nold[ix(k)]));
          P.Nv[ix(k)] = n;
        }
      }
      double N_norm = 0.0;   // :1378-1390
      for (int k = 0; k < K; k++) {
        double n = P.Nv[ix(k)];
        if (n < N_min) n = N_min;
        P.Nv[ix(k)] = n;
        N_norm = dA(N_norm, n);
      }
      for (int k = 0; k < K; k++) {
        double n = dD(P.Nv[ix(k)], N_norm);
        if (n < N_min) n = N_min;
        P.Nv[ix(k)] = n;
      }
      for (int s = 0; s < size; s++) {   // sampling weights (:1392-1422)
        const int j = e0 + s;
        const int zprev = s > 0 ? P.zk[ix(j - 1)] : 0;
        const int z1 = (g == g_end - 1) ? 0 : P.z[ix(P.grp_first[g + 1])];
        for (int k = 0; k < K; k++) {
          double wk = 1.0;
          if (s > 0) {
            wk = dM(wk, dA(dM(dS(1.0, b[j]), alpha), dM(b[j], k == zprev ? 1.0 : 0.0)));
          } else {
            if (g == 0) wk = dM(wk, dM(dS(1.0, b0), alpha));
            else {
              wk = dM(wk, dA(dM(dS(1.0, b0), alpha), dM(b0, k == z0 ? 1.0 : 0.0)));
              wk = dM(wk, dA(dM(dS(1.0, b0), alpha), dM(b0, k == z0k ? 1.0 : 0.0)));
            }
            wk = dD(wk, P.Nold[ix(k)]);
          }
          if (g == g_end - 1) wk = dM(wk, P.vbar[ix((size_t)s * K + 0)]);
          else wk = dM(wk, dA(P.vbar[ix((size_t)s * K + z1)], dM(P.vstar[ix(s)], k == z1 ? 1.0 : 0.0)));
          P.weights[ix(k)] = wk;
        }
        P.zk[ix(j)] = choice(K, uniform(KGW_STREAM_ZK, j, w));
      }
      for (int k = 0; k < K; k++) P.Nold[ix(k)] = P.Nv[ix(k)];   // :1424-1426
    }
  }

  // sample_emission_copula (knockoffs.cpp:1469-1502), lambda = 0.5 at the call site (:677), loci of the window proper
  __device__ void emission(int w, const int32_t *ref) const {
    const double lambda = 0.5;
    for (int j = P.win_start[w]; j < P.win_end[w]; j++) {
      const double p_orig1 = mutate(bit(ref[P.z[ix(j)]], j), P.mu[j]);
      const double p_ko1 = mutate(bit(ref[P.zk[ix(j)]], j), P.mu[j]);
      double temp = dS(p_orig1, dS(1.0, p_ko1)), temp2 = 0.0;
      if (temp < 0) { temp2 = -temp; temp = 0.0; }
      double w0, w1;
      if (bit(i, j) == 0) {
        w1 = dS(1.0, dD(temp2, dS(1.0, p_orig1)));
        w0 = dD(temp2, dS(1.0, p_orig1));
      } else {
        w1 = dD(temp, p_orig1);
        w0 = dS(1.0, dD(temp, p_orig1));
      }
      w0 = dA(dM(lambda, w0), dM(dS(1.0, lambda), dS(1.0, p_ko1)));
      w1 = dA(dM(lambda, w1), dM(dS(1.0, lambda), p_ko1));
      // weighted_choice over two weights
      const double R = dM(uniform(KGW_STREAM_HK, j, w), dA(dA(0.0, w0), w1));
      P.hk[ix(j)] = (R < w0) ? 0 : 1;
    }
  }
};

__global__ void __launch_bounds__(64) strict_sampler_kernel(StrictParams P) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.n) return;
  Thread th(P, t);
  for (int w = 0; w < P.W; w++) {   // generate_unrelated_worker (knockoffs.cpp:654-681)
    const int32_t *ref = P.ref_local + ((size_t)th.i * P.W + w) * P.K;
    th.posterior(w, ref);
    th.knockoff(w);
    th.emission(w, ref);
  }
  // set_Hk -> chaplotype(vector<int>) (knockoffs.cpp:299, chaplotype.cpp:25-29)
  uint8_t *row = P.hk_rows + (size_t)t * P.row_bytes;
  for (long long q = 0; q < P.row_bytes; q++) {
    unsigned v = 0;
    for (int r = 0; r < 8; r++) {
      const long long j = q * 8 + r;
      if (j < P.p && P.hk[th.ix((size_t)j)]) v |= 1u << r;
    }
    row[q] = (uint8_t)v;
  }
  if (P.z_out) for (int j = 0; j < P.p; j++) P.z_out[(size_t)t * P.p + j] = P.z[th.ix(j)];
  if (P.zk_out) for (int j = 0; j < P.p; j++) P.zk_out[(size_t)t * P.p + j] = P.zk[th.ix(j)];
  if (P.beta_out) {
    const size_t n = (size_t)P.p * P.K;
    for (size_t e = 0; e < n; e++) P.beta_out[(size_t)t * n + e] = P.backward[th.ix(e)];
  }
}

struct DevBufs {
  std::vector<void *> ptrs;
  ~DevBufs() { for (void *q : ptrs) cudaFree(q); }
  template <typename T>
  cudaError_t get(T **out, size_t bytes) {
    const cudaError_t e = kgw_malloc(out, std::max<size_t>(bytes, 16));
    if (e == cudaSuccess) ptrs.push_back(*out);
    return e;
  }
  template <typename T>
  cudaError_t up(T **out, const T *src, size_t count) {
    cudaError_t e = get(out, sizeof(T) * count);
    if (e == cudaSuccess && count) e = cudaMemcpy(*out, src, sizeof(T) * count, cudaMemcpyHostToDevice);
    return e;
  }
};

}   // namespace

#define KGW_CUDA_S(call)                                                                              \
  do {                                                                                                \
    const cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) return kgw_fail_(e_ == cudaErrorMemoryAllocation ? KGW_ERR_NOMEM : KGW_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

extern "C" int32_t kgw_generate_strict(const kgw_problem *prob, int32_t device, uint8_t *hk_out, kgw_debug *dbg) {
  if (!prob || !hk_out) return kgw_fail_(KGW_ERR_INVALID, "kgw_generate_strict: null argument");
  if (prob->abi_version != KGW_ABI_VERSION) return kgw_fail_(KGW_ERR_INVALID, "kgw_problem.abi_version does not match this library");
  const int N = prob->N, p = prob->p, K = prob->K, W = prob->W, n_all = prob->n_unrelated;
  if (N <= 0 || p <= 0 || K <= 0 || W <= 0 || n_all < 0 || prob->row_bytes * 8 < p) return kgw_fail_(KGW_ERR_INVALID, "kgw_generate_strict: bad sizes");
  if (!prob->H || !prob->ref_local || !prob->b || !prob->mu || !prob->groups || !prob->win_start || !prob->win_end || (n_all && !prob->unrelated))
    return kgw_fail_(KGW_ERR_INVALID, "kgw_generate_strict: null array in kgw_problem");
  if (prob->families && prob->families->n_families > 0)
    return kgw_fail_(KGW_ERR_UNSUPPORTED, "kgw_generate_strict audits the unrelated branch only: pass the problem without families");
  for (int c = 0; c < n_all; c++)
    if (prob->unrelated[c] < 0 || prob->unrelated[c] >= N) return kgw_fail_(KGW_ERR_INVALID, "unrelated haplotype index out of range");
  for (size_t e = 0; e < (size_t)N * W * K; e++)
    if (prob->ref_local[e] < 0 || prob->ref_local[e] >= N) return kgw_fail_(KGW_ERR_INVALID, "reference haplotype index out of range");
  // groups: contiguous, non-decreasing, dense (Knockoffs::set_groups, knockoffs.cpp:96-107)
  const int G = prob->groups[p - 1] + 1;
  if (prob->groups[0] != 0 || G <= 0) return kgw_fail_(KGW_ERR_INVALID, "groups must start at 0");
  std::vector<int32_t> gfirst(G, -1), glast(G, -1);
  for (int j = 0; j < p; j++) {
    const int g = prob->groups[j];
    if (j > 0 && (g < prob->groups[j - 1] || g > prob->groups[j - 1] + 1)) return kgw_fail_(KGW_ERR_INVALID, "groups must be contiguous, non-decreasing runs");
    if (gfirst[g] < 0) gfirst[g] = j;
    glast[g] = j;
  }
  int max_group = 1;
  for (int g = 0; g < G; g++) max_group = std::max(max_group, glast[g] - gfirst[g] + 1);
  for (int w = 0; w < W; w++)
    if (prob->win_start[w] < 0 || prob->win_end[w] > p || prob->win_start[w] >= prob->win_end[w]) return kgw_fail_(KGW_ERR_INVALID, "bad window bounds");
  if (n_all == 0) return KGW_OK;

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return kgw_fail_(KGW_ERR_CUDA, "no CUDA device: the audit mode runs on the GPU like the product path (there is no CPU fallback)");
  KGW_CUDA_S(cudaSetDevice(device));
  DevBufs D;
  StrictParams P = {};
  P.N = N; P.p = p; P.K = K; P.W = W; P.max_group = max_group;
  P.row_bytes = prob->row_bytes;
  P.seed = prob->seed;
  uint8_t *dH; int32_t *dref, *dgroups, *dgf, *dgl, *dws, *dwe; double *db, *dmu, *dinj = nullptr;
  KGW_CUDA_S(D.up(&dH, prob->H, (size_t)N * prob->row_bytes));
  KGW_CUDA_S(D.up(&dref, prob->ref_local, (size_t)N * W * K));
  KGW_CUDA_S(D.up(&db, prob->b, (size_t)p));
  KGW_CUDA_S(D.up(&dmu, prob->mu, (size_t)p));
  KGW_CUDA_S(D.up(&dgroups, prob->groups, (size_t)p));
  KGW_CUDA_S(D.up(&dgf, gfirst.data(), (size_t)G));
  KGW_CUDA_S(D.up(&dgl, glast.data(), (size_t)G));
  KGW_CUDA_S(D.up(&dws, prob->win_start, (size_t)W));
  KGW_CUDA_S(D.up(&dwe, prob->win_end, (size_t)W));
  if (prob->injected_u) KGW_CUDA_S(D.up(&dinj, prob->injected_u, (size_t)N * W * 3 * p));
  P.H = dH; P.ref_local = dref; P.b = db; P.mu = dmu; P.groups = dgroups; P.grp_first = dgf; P.grp_last = dgl;
  P.win_start = dws; P.win_end = dwe; P.injected_u = dinj;

  const bool want_z = dbg && dbg->z, want_zk = dbg && dbg->zk, want_beta = dbg && dbg->beta_ckpt;
  if (want_beta && dbg->n_ckpt < p) return kgw_fail_(KGW_ERR_INVALID, "kgw_generate_strict writes every backward row: beta_ckpt needs n_ckpt >= p");
  // haplotypes per launch: what the tables of a thread take, against the free memory
  const size_t per_thread = sizeof(double) * ((size_t)p * K + 3 * (size_t)K + (size_t)max_group * (K + 1)) + 9 * (size_t)p +
                            (size_t)prob->row_bytes + (want_z ? 4 * (size_t)p : 0) + (want_zk ? 4 * (size_t)p : 0) +
                            (want_beta ? sizeof(double) * (size_t)p * K : 0);
  size_t free_b = 0, total_b = 0;
  KGW_CUDA_S(cudaMemGetInfo(&free_b, &total_b));
  size_t cap = (size_t)(0.8 * (double)free_b) / per_thread;
  if (cap == 0) return kgw_fail_(KGW_ERR_NOMEM, "kgw_generate_strict: not even one haplotype's tables fit in device memory");
  if (cap >= 64) cap &= ~(size_t)63;
  const int T = (int)std::min<size_t>(cap, (size_t)n_all);
  P.T = T;
  int32_t *dhaps;
  KGW_CUDA_S(D.get(&dhaps, sizeof(int32_t) * T));
  KGW_CUDA_S(D.get(&P.backward, sizeof(double) * (size_t)p * K * T));
  KGW_CUDA_S(D.get(&P.weights, sizeof(double) * (size_t)K * T));
  KGW_CUDA_S(D.get(&P.Nv, sizeof(double) * (size_t)K * T));
  KGW_CUDA_S(D.get(&P.Nold, sizeof(double) * (size_t)K * T));
  KGW_CUDA_S(D.get(&P.vstar, sizeof(double) * (size_t)max_group * T));
  KGW_CUDA_S(D.get(&P.vbar, sizeof(double) * (size_t)max_group * K * T));
  KGW_CUDA_S(D.get(&P.z, sizeof(int32_t) * (size_t)p * T));
  KGW_CUDA_S(D.get(&P.zk, sizeof(int32_t) * (size_t)p * T));
  KGW_CUDA_S(D.get(&P.hk, (size_t)p * T));
  KGW_CUDA_S(D.get(&P.hk_rows, (size_t)prob->row_bytes * T));
  if (want_z) KGW_CUDA_S(D.get(&P.z_out, sizeof(int32_t) * (size_t)p * T));
  if (want_zk) KGW_CUDA_S(D.get(&P.zk_out, sizeof(int32_t) * (size_t)p * T));
  if (want_beta) KGW_CUDA_S(D.get(&P.beta_out, sizeof(double) * (size_t)p * K * T));
  P.haps = dhaps;
  std::vector<uint8_t> rows((size_t)prob->row_bytes * T);
  for (int base = 0; base < n_all; base += T) {
    const int n = std::min(T, n_all - base);
    P.n = n;
    KGW_CUDA_S(cudaMemcpy(dhaps, prob->unrelated + base, sizeof(int32_t) * n, cudaMemcpyHostToDevice));
    // the reference's worker scratch starts zeroed (z, zk: knockoffs.cpp:645-652)
    KGW_CUDA_S(cudaMemset(P.z, 0, sizeof(int32_t) * (size_t)p * T));
    KGW_CUDA_S(cudaMemset(P.zk, 0, sizeof(int32_t) * (size_t)p * T));
    strict_sampler_kernel<<<(n + 63) / 64, 64>>>(P);
    KGW_CUDA_S(cudaGetLastError());
    KGW_CUDA_S(cudaDeviceSynchronize());
    KGW_CUDA_S(cudaMemcpy(rows.data(), P.hk_rows, (size_t)prob->row_bytes * n, cudaMemcpyDeviceToHost));
    for (int c = 0; c < n; c++)
      memcpy(hk_out + (size_t)prob->unrelated[base + c] * prob->row_bytes, rows.data() + (size_t)c * prob->row_bytes, (size_t)prob->row_bytes);
    if (want_z) KGW_CUDA_S(cudaMemcpy(dbg->z + (size_t)base * p, P.z_out, sizeof(int32_t) * (size_t)p * n, cudaMemcpyDeviceToHost));
    if (want_zk) KGW_CUDA_S(cudaMemcpy(dbg->zk + (size_t)base * p, P.zk_out, sizeof(int32_t) * (size_t)p * n, cudaMemcpyDeviceToHost));
    if (want_beta) KGW_CUDA_S(cudaMemcpy(dbg->beta_ckpt + (size_t)base * p * K, P.beta_out, sizeof(double) * (size_t)p * K * n, cudaMemcpyDeviceToHost));
  }
  if (want_beta) {
    dbg->n_ckpt = p;
    if (dbg->ckpt_locus) for (int j = 0; j < p; j++) dbg->ckpt_locus[j] = j;
  }
  return KGW_OK;
}
