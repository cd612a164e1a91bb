// knockoffgwas_b200/csrc/kgw_em.cu -- `--estimate-hmm` on the device.
//
// Replaces Knockoffs::EM / EM_step / EM_worker (snpknock2/src/knockoffs.cpp:706-960).  As the reference stands one EM
// step only re-estimates the mutation rates: the recombination scale is not updated (`new_rho = rho`, :929-930;
// EM_solve_rho is commented out) and the Xi sums (:803-830) feed nothing.  A step is, for every haplotype (related or
// not, :862-864) and every window pass: the backward messages (compute_hmm_backward :1041-1075), the forward messages
// (compute_hmm_forward :1087-1125), the posterior marginals of the core loci (:777-787) and the expected mismatch
//     gamma_j += sum_{k : H[ref_k][j] != H[i][j]} marginal_j(k)                                   (:789-798)
// then mu_j := clamp(gamma_j / N, 1e-6, 1e-3) (:917-926) and the step's change is the root mean square of
// |mu_new - mu| / max(mu, 1e-5) (:908-926); iterations stop below 5e-2 or after 20 (:716-736).
//
// Kernel: one warp per haplotype (persistent grid, haplotypes handed out by an atomic counter); lane l owns the states
// l, l+32, ...; the backward pass writes its normalised message rows to a per-warp HBM slot (p x K doubles -- the table
// the reference materialises), the forward pass keeps its row in registers, multiplies it with the stored backward row
// and reduces the two sums of the marginal ratio (all states / mismatching states) with one shuffle tree.  The marginal
// is a ratio, so both recursions run on messages normalised per locus exactly where the reference normalises them.
// gamma_j is accumulated with one fp64 atomic add per (haplotype, locus): the sum over haplotypes is order dependent at
// the last bits, which is why this row is checked against the oracle to a relative tolerance (tests/test_gpu_em.py:
// 1e-9; the north star asks 1e-5 for floating-point results) and not bit for bit.
#include "../../include/kgw.h"
#include "kgw_common.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

int kgw_fail_(int code, const std::string &msg);   // kgw_api.cu

namespace {

struct EmParams {
  int32_t N, p, K, W, stride_w, pad;
  const uint32_t *Hw;        // [N][stride_w]
  const int32_t *ref_local;  // [N][W][K]
  const double *b, *mu;      // [p]
  const int32_t *pass;       // [W][4]: padded range [js, je), core range [cs, ce)   (knockoffs.cpp:1222-1226)
  double *gamma;             // [p]
  double *beta;              // [slots][p][K] backward rows of the warp's current haplotype
  int32_t *work;
};

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void warp_sum2(double &a, double &b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ta = __shfl_xor_sync(FULL, a, o), tb = __shfl_xor_sync(FULL, b, o);
    a += ta;
    b += tb;
  }
}
__device__ __forceinline__ double warp_sum(double a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(FULL, a, o);
  return a;
}

template <int SF>
__global__ void __launch_bounds__(128) em_kernel(const EmParams P) {
  const int lane = threadIdx.x & 31;
  const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int K = P.K, p = P.p;
  double *brow = P.beta + (size_t)slot * p * K;
  const double alpha = 1.0 / K;   // knockoffs.cpp:75
  for (;;) {
    int i = 0;
    if (lane == 0) i = atomicAdd(P.work, 1);
    i = __shfl_sync(FULL, i, 0);
    if (i >= P.N) break;
    const uint32_t *hrow = P.Hw + (size_t)i * P.stride_w;
    for (int w = 0; w < P.W; w++) {
      const int js = P.pass[4 * w], je = P.pass[4 * w + 1], cs = P.pass[4 * w + 2], ce = P.pass[4 * w + 3];
      const int32_t *ref = P.ref_local + ((size_t)i * P.W + w) * K;
      const uint32_t *rrow[SF];
      bool valid[SF];
#pragma unroll
      for (int s = 0; s < SF; s++) {
        valid[s] = s * 32 + lane < K;
        rrow[s] = P.Hw + (size_t)(valid[s] ? ref[s * 32 + lane] : i) * P.stride_w;
      }
      // ---- backward messages, normalised per locus (:1041-1075); row j -> brow[j][.]
      double bt[SF];
      uint32_t mw[SF];   // match words of the current 32-locus word: bit q set = reference k equals the target at locus 32 w + q
      int wcur = -1;
      auto load_words = [&](int word) {
        const uint32_t hw = hrow[word];
#pragma unroll
        for (int s = 0; s < SF; s++) mw[s] = ~(hw ^ rrow[s][word]);
        wcur = word;
      };
#pragma unroll
      for (int s = 0; s < SF; s++) {
        bt[s] = valid[s] ? alpha : 0.0;   // 1.0 / K, :1046
        if (valid[s]) brow[(size_t)(je - 1) * K + s * 32 + lane] = bt[s];
      }
      for (int j = je - 2; j >= js; j--) {   // uses locus j + 1
        const int m = j + 1;
        if ((m >> 5) != wcur) load_words(m >> 5);
        const double bm = P.b[m], mum = P.mu[m];
        const double a = (1.0 - bm) * alpha;
        double e[SF], part = 0.0;
#pragma unroll
        for (int s = 0; s < SF; s++) {
          e[s] = valid[s] ? bt[s] * (((mw[s] >> (m & 31)) & 1u) ? 1.0 - mum : mum) : 0.0;
          part += e[s];
        }
        const double c = a * warp_sum(part);           // sum_k beta a E (:1052-1060)
        double tot = 0.0;
#pragma unroll
        for (int s = 0; s < SF; s++) {
          bt[s] = valid[s] ? c + bm * e[s] : 0.0;       // (:1062-1068)
          tot += bt[s];
        }
        const double rn = 1.0 / warp_sum(tot);         // (:1069-1073)
#pragma unroll
        for (int s = 0; s < SF; s++) {
          bt[s] *= rn;
          if (valid[s]) brow[(size_t)j * K + s * 32 + lane] = bt[s];
        }
      }
      __syncwarp();
      // ---- forward messages (:1087-1125) and the marginal ratio of the core loci (:777-798)
      double ft[SF];
#pragma unroll
      for (int s = 0; s < SF; s++) ft[s] = 0.0;
      for (int j = js; j < je; j++) {
        if ((j >> 5) != wcur) load_words(j >> 5);
        const double bj = P.b[j], muj = P.mu[j];
        const double a = (1.0 - bj) * alpha;
        double tot = 0.0;
#pragma unroll
        for (int s = 0; s < SF; s++) {
          const double E = ((mw[s] >> (j & 31)) & 1u) ? 1.0 - muj : muj;
          ft[s] = valid[s] ? ((j == js) ? alpha : a + bj * ft[s]) * E : 0.0;
          tot += ft[s];
        }
        const double rn = 1.0 / warp_sum(tot);
        double sall = 0.0, smis = 0.0;
#pragma unroll
        for (int s = 0; s < SF; s++) {
          ft[s] *= rn;
          if (valid[s] && j >= cs && j < ce) {
            const double mg = brow[(size_t)j * K + s * 32 + lane] * ft[s];
            sall += mg;
            if (!((mw[s] >> (j & 31)) & 1u)) smis += mg;
          }
        }
        if (j >= cs && j < ce) {
          warp_sum2(sall, smis);
          if (lane == 0) atomicAdd(P.gamma + j, smis / sall);
        }
      }
      __syncwarp();
    }
  }
}

#define EM_CUDA(call)                                                                                  \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) { cleanup(); return kgw_fail_(KGW_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } \
  } while (0)

}  // namespace

extern "C" int32_t kgw_estimate_hmm(const kgw_problem *prob, const kgw_options *opt, const double *mu_init, double *mu_out,
                                    int32_t *iterations, double *deltas) {
  if (!prob || !mu_init || !mu_out) return kgw_fail_(KGW_ERR_INVALID, "null argument");
  if (prob->abi_version != KGW_ABI_VERSION) return kgw_fail_(KGW_ERR_INVALID, "abi_version mismatch");
  const int N = prob->N, p = prob->p, K = prob->K, W = prob->W;
  if (N <= 0 || p <= 0 || K <= 0 || W <= 0) return kgw_fail_(KGW_ERR_INVALID, "N, p, K, W must be positive");
  if (!prob->H || !prob->ref_local || !prob->b || !prob->win_start || !prob->win_end) return kgw_fail_(KGW_ERR_INVALID, "null input buffer");
  if (prob->row_bytes < (p + 7) / 8) return kgw_fail_(KGW_ERR_INVALID, "row_bytes < ceil(p/8)");
  if (K > 256) return kgw_fail_(KGW_ERR_UNSUPPORTED, "K > 256 is not built");
  for (size_t t = 0; t < (size_t)N * W * K; t++)
    if (prob->ref_local[t] < 0 || prob->ref_local[t] >= N) return kgw_fail_(KGW_ERR_INVALID, "reference index out of range");
  std::vector<int32_t> pass((size_t)4 * W);
  for (int w = 0; w < W; w++) {
    pass[4 * w] = (w > 0) ? prob->win_start[w - 1] : prob->win_start[w];        // knockoffs.cpp:1222-1226
    pass[4 * w + 1] = (w < W - 1) ? prob->win_end[w + 1] : prob->win_end[w];
    pass[4 * w + 2] = prob->win_start[w];
    pass[4 * w + 3] = prob->win_end[w];
    if (pass[4 * w] < 0 || pass[4 * w + 1] > p || pass[4 * w + 2] < pass[4 * w] || pass[4 * w + 3] > pass[4 * w + 1] || pass[4 * w + 2] >= pass[4 * w + 3])
      return kgw_fail_(KGW_ERR_INVALID, "inconsistent window boundaries");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return kgw_fail_(KGW_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  const int device = opt ? opt->device : 0;
  if (device < 0 || device >= ndev) return kgw_fail_(KGW_ERR_INVALID, "bad device ordinal");

  uint32_t *dH = nullptr;
  int32_t *dref = nullptr, *dpass = nullptr, *dwork = nullptr;
  double *db = nullptr, *dmu = nullptr, *dgamma = nullptr, *dbeta = nullptr;
  auto cleanup = [&]() {
    cudaFree(dH); cudaFree(dref); cudaFree(dpass); cudaFree(dwork); cudaFree(db); cudaFree(dmu); cudaFree(dgamma); cudaFree(dbeta);
  };
  EM_CUDA(cudaSetDevice(device));
  int sms = 0;
  EM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int stride_w = (int)(((size_t)prob->row_bytes + 15) / 16 * 4);
  EM_CUDA(kgw_malloc(&dH, (size_t)N * stride_w * 4));
  EM_CUDA(cudaMemset(dH, 0, (size_t)N * stride_w * 4));
  EM_CUDA(cudaMemcpy2D(dH, (size_t)stride_w * 4, prob->H, (size_t)prob->row_bytes, (size_t)(p + 7) / 8, N, cudaMemcpyHostToDevice));
  EM_CUDA(kgw_malloc(&dref, sizeof(int32_t) * (size_t)N * W * K));
  EM_CUDA(cudaMemcpy(dref, prob->ref_local, sizeof(int32_t) * (size_t)N * W * K, cudaMemcpyHostToDevice));
  EM_CUDA(kgw_malloc(&dpass, sizeof(int32_t) * 4 * W));
  EM_CUDA(cudaMemcpy(dpass, pass.data(), sizeof(int32_t) * 4 * W, cudaMemcpyHostToDevice));
  EM_CUDA(kgw_malloc(&db, sizeof(double) * p));
  EM_CUDA(cudaMemcpy(db, prob->b, sizeof(double) * p, cudaMemcpyHostToDevice));
  EM_CUDA(kgw_malloc(&dmu, sizeof(double) * p));
  EM_CUDA(kgw_malloc(&dgamma, sizeof(double) * p));
  EM_CUDA(kgw_malloc(&dwork, sizeof(int32_t)));
  // one backward table per resident warp: as many warps as the SMs hold (16 per SM) and half the free memory allow
  size_t free_b = 0, total_b = 0;
  EM_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const size_t per_slot = sizeof(double) * (size_t)p * K;
  long long slots = std::min<long long>((long long)sms * 16, (long long)((double)free_b * 0.5 / (double)per_slot));
  slots = std::min<long long>(slots, N);
  if (slots < 1) { cleanup(); return kgw_fail_(KGW_ERR_NOMEM, "not enough device memory for one backward table"); }
  const int wpc = slots >= 4 ? 4 : 1;
  const int grid = (int)(slots / wpc);
  slots = (long long)grid * wpc;
  EM_CUDA(kgw_malloc(&dbeta, per_slot * (size_t)slots));
  const int SF = (K + 31) / 32;
  EmParams P{};
  P.N = N; P.p = p; P.K = K; P.W = W; P.stride_w = stride_w;
  P.Hw = dH; P.ref_local = dref; P.b = db; P.mu = dmu; P.pass = dpass; P.gamma = dgamma; P.beta = dbeta; P.work = dwork;

  std::vector<double> mu(mu_init, mu_init + p), gamma(p);
  const double delta_min = 5e-2;   // knockoffs.cpp:716
  const int it_max = 20;           // :717
  int it = 0;
  bool converged = false;
  for (it = 0; it < it_max; it++) {
    EM_CUDA(cudaMemcpy(dmu, mu.data(), sizeof(double) * p, cudaMemcpyHostToDevice));
    EM_CUDA(cudaMemset(dgamma, 0, sizeof(double) * p));
    EM_CUDA(cudaMemset(dwork, 0, sizeof(int32_t)));
    switch (SF) {
      case 1: em_kernel<1><<<grid, 32 * wpc>>>(P); break;
      case 2: em_kernel<2><<<grid, 32 * wpc>>>(P); break;
      case 3: em_kernel<3><<<grid, 32 * wpc>>>(P); break;
      case 4: em_kernel<4><<<grid, 32 * wpc>>>(P); break;
      case 5: em_kernel<5><<<grid, 32 * wpc>>>(P); break;
      case 6: em_kernel<6><<<grid, 32 * wpc>>>(P); break;
      case 7: em_kernel<7><<<grid, 32 * wpc>>>(P); break;
      default: em_kernel<8><<<grid, 32 * wpc>>>(P); break;
    }
    EM_CUDA(cudaGetLastError());
    EM_CUDA(cudaMemcpy(gamma.data(), dgamma, sizeof(double) * p, cudaMemcpyDeviceToHost));
    // EM_step, host part (:903-930): mean over haplotypes, clamp, relative change
    const double n_inv = 1.0 / (double)N;
    double delta_mutation = 0.0;
    for (int j = 0; j < p; j++) {
      double g = 0.0 + gamma[j] * n_inv;
      g = std::max(1e-6, g);
      g = std::min(1e-3, g);
      const double dj = std::abs(g - mu[j]) / std::max(mu[j], 1e-5);
      delta_mutation += dj * dj;
      gamma[j] = g;
    }
    delta_mutation = std::sqrt(delta_mutation / (double)p);
    const double delta_parameters = std::max(0.0, delta_mutation);   // the recombination scale does not move (:929-930)
    mu = gamma;
    if (deltas) deltas[it] = delta_parameters;
    if (delta_parameters < delta_min) { converged = true; break; }
  }
  std::copy(mu.begin(), mu.end(), mu_out);
  if (iterations) *iterations = converged ? it + 1 : it;
  cleanup();
  return converged ? KGW_OK : 1;   // 1: not converged after 20 steps (the reference prints a warning and goes on, :739-748)
}
