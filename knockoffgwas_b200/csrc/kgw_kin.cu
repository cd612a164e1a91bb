// knockoffgwas_b200/csrc/kgw_kin.cu -- the two data passes of the bifurcating k-means of `--compute-references`
// (SURVEY.md section 8f, rank 2): KinshipComputer::update_assignments and update_centroids
// (snpknock2/src/kinship_computer.cpp:314-405, distance :239-250).
//
// The reference walks, for every member of the cluster being split and for both centroids, every thinned SNP of every
// chromosome with one vector<bool>-style bit read per SNP (chaplotype::operator[]), on one thread:
//     distance(i, mu) = sum over chr != chr_out, j:  (H[chr][i][j] - mu[chr][j])^2        (sequential, one accumulator)
//     dist(member)    = distance(i0, mu) + distance(i1, mu),  i0 = 2*(i/2), i1 = i0 + 1   (both haplotypes of the sample)
// and the centroids are the column means of the members of each side.  Here the thinned rows of all chromosomes sit in
// HBM as one packed bit matrix (kgw_kinship handle); a k-means iteration is two launches:
//   kin_distance_kernel  one thread per (member, centroid): the same additions in the same order (fp64, no fused
//                        multiply-add), so the comparison dist_0 < dist_1 (:331) sees the reference's numbers as a
//                        strict-IEEE build computes them;
//   kin_centroid_kernel  one thread per (32 loci, 64 members): integer counts of set bits over the members of each side
//                        (atomics; exact in any order), divided once per locus as the reference divides its sums (:399-402).
// The control flow around them (initial centroids from the reference's own random stream, the size rebalancing by
// multimap, acceptance of a split, the list of unfinished clusters) is integer bookkeeping and stays on the host.
#include "../../include/kgw.h"
#include "kgw_common.h"

#include <cuda_runtime.h>

#include <string>
#include <vector>

int kgw_fail_(int code, const std::string &msg);   // kgw_api.cu

struct kgw_kinship {
  int device = 0;
  int N = 0, p = 0, stride_w = 0;
  uint32_t *dH = nullptr;        // [N][stride_w] packed rows, LSB first
  uint8_t *duse = nullptr;       // [p] loci that count (all ones, or zeros on the chromosome left out)
  double *dmu = nullptr;         // [2][p]
  double *dd = nullptr;          // [2][cap]
  int32_t *dmem = nullptr, *dasg = nullptr;
  unsigned int *dcnt = nullptr;  // [2][p]
  int cap = 0;
};

namespace {

#define KIN_CUDA(call)                                                                                  \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return kgw_fail_(KGW_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// thread t: member t / 2, centroid t % 2
__global__ void kin_distance_kernel(const uint32_t *__restrict__ H, int stride_w, int p, const uint8_t *__restrict__ use,
                                    const double *__restrict__ mu, const int32_t *__restrict__ members, int n, double *__restrict__ out, int cap) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * n) return;
  const int m = t >> 1, side = t & 1;
  const int i = members[m];
  const double *c = mu + (size_t)side * p;
  double tot[2];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const uint32_t *row = H + (size_t)(2 * (i / 2) + h) * stride_w;
    double acc = 0.0;
    for (int w = 0; w * 32 < p; w++) {
      const uint32_t bits = row[w];
      const int jn = min(32, p - w * 32);
      for (int q = 0; q < jn; q++) {
        const int j = w * 32 + q;
        if (!use[j]) continue;
        const double d = __dsub_rn((double)((bits >> q) & 1u), c[j]);
        acc = __dadd_rn(acc, __dmul_rn(d, d));   // tot_dist += pow(hj - mu, 2), :246
      }
    }
    tot[h] = acc;
  }
  out[(size_t)side * cap + m] = __dadd_rn(tot[0], tot[1]);   // distance(i0, ...) + distance(i1, ...), :327-328
}

// thread (w, chunk): loci 32 w .. 32 w + 31 of KIN_CHUNK members; counts of set bits among the members of each side,
// added to the totals with integer atomics (exact in any order)
constexpr int KIN_CHUNK = 64;
__global__ void kin_centroid_kernel(const uint32_t *__restrict__ H, int stride_w, int p, const int32_t *__restrict__ members,
                                    const int32_t *__restrict__ assign, int n, unsigned int *__restrict__ cnt) {
  const int nw = (p + 31) / 32;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int w = (int)(t % nw);
  const long long chunk = t / nw;
  const long long m0 = chunk * KIN_CHUNK;
  if (m0 >= n) return;
  const int m1 = (int)min((long long)n, m0 + KIN_CHUNK);
  unsigned int c0[32], c1[32];
#pragma unroll
  for (int q = 0; q < 32; q++) { c0[q] = 0u; c1[q] = 0u; }
  for (int m = (int)m0; m < m1; m++) {
    const uint32_t bits = H[(size_t)members[m] * stride_w + w];
    const int a = assign[m];
#pragma unroll
    for (int q = 0; q < 32; q++) {
      const unsigned b = (bits >> q) & 1u;
      c0[q] += (a == 0) ? b : 0u;
      c1[q] += (a == 1) ? b : 0u;
    }
  }
#pragma unroll
  for (int q = 0; q < 32; q++) {
    const int j = w * 32 + q;
    if (j < p) {
      if (c0[q]) atomicAdd(cnt + j, c0[q]);
      if (c1[q]) atomicAdd(cnt + (size_t)p + j, c1[q]);
    }
  }
}

int ensure_capacity(kgw_kinship *k, int n) {
  if (n <= k->cap) return KGW_OK;
  cudaFree(k->dd); cudaFree(k->dmem); cudaFree(k->dasg);
  k->dd = nullptr; k->dmem = nullptr; k->dasg = nullptr;
  k->cap = 0;
  KIN_CUDA(kgw_malloc(&k->dd, sizeof(double) * 2 * (size_t)n));
  KIN_CUDA(kgw_malloc(&k->dmem, sizeof(int32_t) * (size_t)n));
  KIN_CUDA(kgw_malloc(&k->dasg, sizeof(int32_t) * (size_t)n));
  k->cap = n;
  return KGW_OK;
}

}  // namespace

extern "C" {

int32_t kgw_kinship_create(int32_t N, int32_t p, int64_t row_bytes, const uint8_t *H, int32_t device, kgw_kinship **out) {
  if (!out) return kgw_fail_(KGW_ERR_INVALID, "null argument");
  *out = nullptr;
  if (N <= 0 || (N & 1) || p <= 0 || !H) return kgw_fail_(KGW_ERR_INVALID, "bad argument (N must be positive and even)");
  if (row_bytes < (p + 7) / 8) return kgw_fail_(KGW_ERR_INVALID, "row_bytes < ceil(p/8)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return kgw_fail_(KGW_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  if (device < 0 || device >= ndev) return kgw_fail_(KGW_ERR_INVALID, "bad device ordinal");
  KIN_CUDA(cudaSetDevice(device));
  kgw_kinship *k = new kgw_kinship();
  k->device = device; k->N = N; k->p = p; k->stride_w = (p + 31) / 32;
  auto bail = [&](int rc) { kgw_kinship_destroy(k); return rc; };
  if (kgw_malloc(&k->dH, sizeof(uint32_t) * (size_t)N * k->stride_w) != cudaSuccess) return bail(kgw_fail_(KGW_ERR_NOMEM, "device memory for the thinned rows"));
  if (cudaMemset(k->dH, 0, sizeof(uint32_t) * (size_t)N * k->stride_w) != cudaSuccess ||
      cudaMemcpy2D(k->dH, sizeof(uint32_t) * (size_t)k->stride_w, H, (size_t)row_bytes, (size_t)(p + 7) / 8, (size_t)N, cudaMemcpyHostToDevice) != cudaSuccess)
    return bail(kgw_fail_(KGW_ERR_CUDA, "upload of the thinned rows failed"));
  if (kgw_malloc(&k->duse, (size_t)p) != cudaSuccess || kgw_malloc(&k->dmu, sizeof(double) * 2 * (size_t)p) != cudaSuccess ||
      kgw_malloc(&k->dcnt, sizeof(unsigned int) * 2 * (size_t)p) != cudaSuccess)
    return bail(kgw_fail_(KGW_ERR_NOMEM, "device memory for the centroids"));
  *out = k;
  return KGW_OK;
}

void kgw_kinship_destroy(kgw_kinship *k) {
  if (!k) return;
  cudaSetDevice(k->device);
  cudaFree(k->dH); cudaFree(k->duse); cudaFree(k->dmu); cudaFree(k->dd); cudaFree(k->dmem); cudaFree(k->dasg); cudaFree(k->dcnt);
  delete k;
}

int32_t kgw_kinship_distances(kgw_kinship *k, const int32_t *members, int32_t n, const double *mu0, const double *mu1,
                              const uint8_t *use, double *d0, double *d1) {
  if (!k || !members || n <= 0 || !mu0 || !mu1 || !d0 || !d1) return kgw_fail_(KGW_ERR_INVALID, "bad argument");
  for (int t = 0; t < n; t++) if (members[t] < 0 || members[t] >= k->N) return kgw_fail_(KGW_ERR_INVALID, "member out of range");
  KIN_CUDA(cudaSetDevice(k->device));
  int rc = ensure_capacity(k, n);
  if (rc != KGW_OK) return rc;
  const int p = k->p;
  KIN_CUDA(cudaMemcpy(k->dmem, members, sizeof(int32_t) * n, cudaMemcpyHostToDevice));
  KIN_CUDA(cudaMemcpy(k->dmu, mu0, sizeof(double) * p, cudaMemcpyHostToDevice));
  KIN_CUDA(cudaMemcpy(k->dmu + p, mu1, sizeof(double) * p, cudaMemcpyHostToDevice));
  if (use) KIN_CUDA(cudaMemcpy(k->duse, use, (size_t)p, cudaMemcpyHostToDevice));
  else KIN_CUDA(cudaMemset(k->duse, 1, (size_t)p));
  kin_distance_kernel<<<(2 * n + 127) / 128, 128>>>(k->dH, k->stride_w, p, k->duse, k->dmu, k->dmem, n, k->dd, k->cap);
  KIN_CUDA(cudaGetLastError());
  KIN_CUDA(cudaMemcpy(d0, k->dd, sizeof(double) * n, cudaMemcpyDeviceToHost));
  KIN_CUDA(cudaMemcpy(d1, k->dd + k->cap, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return KGW_OK;
}

int32_t kgw_kinship_centroids(kgw_kinship *k, const int32_t *members, const int32_t *assign, int32_t n, double *mu0, double *mu1,
                              int32_t *n0_out, int32_t *n1_out) {
  if (!k || !members || !assign || n <= 0 || !mu0 || !mu1) return kgw_fail_(KGW_ERR_INVALID, "bad argument");
  for (int t = 0; t < n; t++) if (members[t] < 0 || members[t] >= k->N) return kgw_fail_(KGW_ERR_INVALID, "member out of range");
  KIN_CUDA(cudaSetDevice(k->device));
  int rc = ensure_capacity(k, n);
  if (rc != KGW_OK) return rc;
  const int p = k->p;
  KIN_CUDA(cudaMemcpy(k->dmem, members, sizeof(int32_t) * n, cudaMemcpyHostToDevice));
  KIN_CUDA(cudaMemcpy(k->dasg, assign, sizeof(int32_t) * n, cudaMemcpyHostToDevice));
  const int nw = (p + 31) / 32;
  KIN_CUDA(cudaMemset(k->dcnt, 0, sizeof(unsigned int) * 2 * (size_t)p));
  const long long threads = (long long)nw * ((n + KIN_CHUNK - 1) / KIN_CHUNK);
  kin_centroid_kernel<<<(unsigned)((threads + 127) / 128), 128>>>(k->dH, k->stride_w, p, k->dmem, k->dasg, n, k->dcnt);
  KIN_CUDA(cudaGetLastError());
  std::vector<unsigned int> cnt(2 * (size_t)p);
  KIN_CUDA(cudaMemcpy(cnt.data(), k->dcnt, sizeof(unsigned int) * 2 * (size_t)p, cudaMemcpyDeviceToHost));
  int n0 = 0, n1 = 0;
  for (int t = 0; t < n; t++) { n0 += (assign[t] == 0); n1 += (assign[t] == 1); }
  // the reference adds doubles 0.0 / 1.0 (exact integers) and divides by the int count (:399-402)
  for (int j = 0; j < p; j++) {
    double a = (double)cnt[j], b = (double)cnt[(size_t)p + j];
    a /= n0;
    b /= n1;
    mu0[j] = a;
    mu1[j] = b;
  }
  if (n0_out) *n0_out = n0;
  if (n1_out) *n1_out = n1;
  return KGW_OK;
}

}  // extern "C"
