// knockoffgwas_b200/csrc/kgw_refs.cu -- reference selection: the K nearest haplotypes of every haplotype inside
// its kinship cluster (scope table f, rank 2).
//
// Replaces the worker loop of KinshipComputer::assign_references (snpknock2/src/kinship_computer.cpp:856-881 ->
// assign_references_worker :722-762 -> findNeighbors :174-233) for one window, the step that produces the
// reference-index buffer of the sampler:
//   * distances are Hamming distances on the THINNED chromosome (every `compression`-th SNP, Metadata::thin,
//     metadata.cpp:563-580; chaplotype::hamming, chaplotype.cpp:62-66);
//   * members of the same IBD family are pushed to distance 1e6 (:197-199), the haplotype itself is skipped;
//   * the K smallest are kept, ties in cluster order (multimap insertion order + strict `dij < top_key`, :207-218).
// The clusters (bifurcating k-means, :409-691) and the merge of the lists inside IBD families
// (combine_references_families, :765-835) stay on the host.
//
// Three kernels, all integer work:
//   thin_kernel      packed rows -> packed rows of the kept SNPs;
//   distance_kernel  one CTA per 64 x 64 tile of a cluster's pair matrix: row chunks staged (transposed) in shared
//                    memory, 4 x 4 pairs per thread, xor + popc; family / self masks applied on the way out;
//   select_kernel    one warp per haplotype: two-level 10-bit radix histogram finds the K-th smallest distance,
//                    an ordered ballot compaction takes everything below it plus the first ties in cluster order,
//                    a bitonic sort of the <= 256 (distance, position) keys gives the reference's output order.
#include "kgw_common.h"
#include "../../include/kgw.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

int kgw_fail_(int code, const std::string &msg);   // kgw_api.cu

namespace {

constexpr int TILE = 64;        // haplotypes per tile side
constexpr int WCH = 32;         // 32-bit words of a row staged per step
constexpr unsigned D_SELF = 0xffffffffu;
constexpr unsigned D_RELATED = 1000000u;   // kinship_computer.cpp:199

struct TileDesc { int32_t cluster, i0, j0, pad; };

__global__ void thin_kernel(const uint8_t *__restrict__ H, size_t row_bytes, uint32_t *__restrict__ Ht, int stride_t, int N, int p,
                            int compression, int pt) {
  const size_t total = (size_t)N * stride_t;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t i = idx / stride_t;
    const int w = (int)(idx - i * stride_t);
    uint32_t out = 0u;
    for (int b = 0; b < 32; b++) {
      const long long t = 32LL * w + b;            // thinned SNP index
      if (t < pt) {
        const long long j = t * compression;       // original SNP index (j % compression == 0)
        out |= (uint32_t)((H[i * row_bytes + (size_t)(j >> 3)] >> (j & 7)) & 1u) << b;
      }
    }
    Ht[idx] = out;
  }
}

// D[doff[c] + i * n_c + j] = distance between members i and j of cluster c
__global__ void __launch_bounds__(256) distance_kernel(const uint32_t *__restrict__ Ht, int stride_t, int nwords,
                                                       const TileDesc *__restrict__ tiles, const int32_t *__restrict__ clust_off,
                                                       const int32_t *__restrict__ members, const int32_t *__restrict__ family_of,
                                                       const long long *__restrict__ doff, uint32_t *__restrict__ D) {
  __shared__ __align__(16) uint32_t At[WCH][TILE], Bt[WCH][TILE];   // transposed: [word][haplotype]
  const TileDesc td = tiles[blockIdx.x];
  const int base = clust_off[td.cluster], n = clust_off[td.cluster + 1] - base;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  unsigned acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0u;
  // loader mapping: thread t stages row t/4 (of 64), words (t%4)*8 .. +7 of the chunk
  const int lr = threadIdx.x >> 2, lw = (threadIdx.x & 3) * 8;
  const int ia = min(td.i0 + lr, n - 1), jb = min(td.j0 + lr, n - 1);
  const uint32_t *rowA = Ht + (size_t)members[base + ia] * stride_t;
  const uint32_t *rowB = Ht + (size_t)members[base + jb] * stride_t;
  for (int w0 = 0; w0 < nwords; w0 += WCH) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int w = w0 + lw + q;
      At[lw + q][lr] = (w < nwords) ? __ldg(rowA + w) : 0u;
      Bt[lw + q][lr] = (w < nwords) ? __ldg(rowB + w) : 0u;
    }
    __syncthreads();
#pragma unroll 8
    for (int w = 0; w < WCH; w++) {
      const uint4 av = *reinterpret_cast<const uint4 *>(&At[w][ty * 4]);
      const uint4 bv = *reinterpret_cast<const uint4 *>(&Bt[w][tx * 4]);
      const uint32_t a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] += __popc(a4[a] ^ b4[b]);
    }
    __syncthreads();
  }
  uint32_t *Dc = D + doff[td.cluster];
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int i = td.i0 + ty * 4 + a;
    if (i >= n) continue;
    const int hi = members[base + i];
    const int fi = family_of ? family_of[hi] : -1;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int j = td.j0 + tx * 4 + b;
      if (j >= n) continue;
      const int hj = members[base + j];
      unsigned d = acc[a][b];
      if (fi >= 0 && family_of[hj] == fi) d = D_RELATED;   // same IBD family (:191-199)
      if (hi == hj) d = D_SELF;                             // `if(i==j) continue` (:186)
      Dc[(size_t)i * n + j] = d;
    }
  }
}

// one warp per (cluster member); out[hap][0..K) = members in (distance, cluster position) order, -1 padding
__global__ void __launch_bounds__(128) select_kernel(const uint32_t *__restrict__ D, const long long *__restrict__ doff,
                                                     const int32_t *__restrict__ clust_off, const int32_t *__restrict__ members,
                                                     const int32_t *__restrict__ row_cluster, int row0, int row1, int K,
                                                     int32_t *__restrict__ out) {
  __shared__ unsigned hist[4][1024];
  __shared__ unsigned long long keys[4][256];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int row = row0 + blockIdx.x * 4 + wid;   // index into the concatenated member list
  if (row >= row1) return;                       // whole warp (no CTA barrier below)
  const int c = row_cluster[row];
  const int base = clust_off[c], n = clust_off[c + 1] - base;
  const int i = row - base;
  const uint32_t *Dr = D + doff[c] + (size_t)i * n;
  const int Kp = min(K, n - 1);
  unsigned *h = hist[wid];
  unsigned long long *ks = keys[wid];
  const int hap = members[row];
  for (int k = lane; k < K; k += 32) out[(size_t)hap * K + k] = -1;
  if (Kp <= 0) return;
  // level 1: bits 19..10 (distances are < 2^20: at most the thinned SNP count, or 1e6 for relatives)
  for (int b = lane; b < 1024; b += 32) h[b] = 0u;
  __syncwarp();
  for (int l = lane; l < n; l += 32) {
    const unsigned d = Dr[l];
    if (d != D_SELF) atomicAdd(&h[min(d >> 10, 1023u)], 1u);
  }
  __syncwarp();
  auto find_bin = [&](unsigned need, unsigned &before) -> int {   // smallest bin whose cumulative count reaches `need`
    unsigned run = 0u;
    int found = -1;
    for (int b0 = 0; b0 < 1024 && found < 0; b0 += 32) {
      const unsigned v = h[b0 + lane];
      unsigned inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      const unsigned hit = __ballot_sync(0xffffffffu, run + inc >= need);
      if (hit) {
        const int fl = __ffs(hit) - 1;
        found = b0 + fl;
        const unsigned inc_f = __shfl_sync(0xffffffffu, inc, fl), v_f = __shfl_sync(0xffffffffu, v, fl);
        before = run + inc_f - v_f;
      } else {
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
    }
    return found;
  };
  unsigned before1 = 0u;
  const int b1 = find_bin((unsigned)Kp, before1);
  // level 2: bits 9..0 inside bin b1
  __syncwarp();
  for (int b = lane; b < 1024; b += 32) h[b] = 0u;
  __syncwarp();
  for (int l = lane; l < n; l += 32) {
    const unsigned d = Dr[l];
    if (d != D_SELF && (int)min(d >> 10, 1023u) == b1) atomicAdd(&h[d & 1023u], 1u);
  }
  __syncwarp();
  unsigned before2 = 0u;
  const int b2 = find_bin((unsigned)Kp - before1, before2);
  const unsigned thr = ((unsigned)b1 << 10) | (unsigned)b2;     // the Kp-th smallest distance
  const unsigned n_lt = before1 + before2;                      // how many are strictly below it
  const unsigned need_eq = (unsigned)Kp - n_lt;                 // ties taken in cluster order
  // ordered compaction
  unsigned got_lt = 0u, got_eq = 0u;
  for (int l0 = 0; l0 < n; l0 += 32) {
    const int l = l0 + lane;
    const unsigned d = (l < n) ? Dr[l] : D_SELF;
    const bool lt = (d != D_SELF) && d < thr, eq = (d != D_SELF) && d == thr;
    const unsigned m_lt = __ballot_sync(0xffffffffu, lt), m_eq = __ballot_sync(0xffffffffu, eq);
    const unsigned below = (1u << lane) - 1u;
    if (lt) ks[got_lt + __popc(m_lt & below)] = ((unsigned long long)d << 32) | (unsigned)l;
    if (eq) {
      const unsigned r = got_eq + __popc(m_eq & below);
      if (r < need_eq) ks[n_lt + r] = ((unsigned long long)d << 32) | (unsigned)l;
    }
    got_lt += __popc(m_lt);
    got_eq += __popc(m_eq);
  }
  for (int k = Kp + lane; k < 256; k += 32) ks[k] = ~0ull;
  __syncwarp();
  // bitonic sort of 256 keys by one warp
  for (int size = 2; size <= 256; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = lane; t < 128; t += 32) {
        const int lo = 2 * t - (t & (stride - 1));   // index with bit `stride` clear
        const int hi2 = lo + stride;
        const bool up = ((lo & size) == 0);
        const unsigned long long x = ks[lo], y = ks[hi2];
        if ((x > y) == up) { ks[lo] = y; ks[hi2] = x; }
      }
      __syncwarp();
    }
  }
  for (int k = lane; k < Kp; k += 32) out[(size_t)hap * K + k] = members[base + (int)(ks[k] & 0xffffffffu)];
}

#define REF_CUDA(call)                                                                                 \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) return kgw_fail_(KGW_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

}  // namespace

extern "C" int32_t kgw_assign_references(int32_t N, int32_t p, int64_t row_bytes, const uint8_t *H, int32_t K, int32_t compression,
                                         int32_t n_clusters, const int32_t *clust_off, const int32_t *clust_members,
                                         const int32_t *family_of, int32_t device, int32_t *ref_out, float *kernel_ms) {
  if (N <= 0 || p <= 0 || !H || !ref_out || !clust_off || !clust_members || n_clusters <= 0)
    return kgw_fail_(KGW_ERR_INVALID, "bad argument");
  if (K <= 0 || K > 256) return kgw_fail_(KGW_ERR_UNSUPPORTED, "K must lie in 1..256");
  if (compression < 1) return kgw_fail_(KGW_ERR_INVALID, "compression must be >= 1 (metadata.cpp:564)");
  if (row_bytes < (p + 7) / 8) return kgw_fail_(KGW_ERR_INVALID, "row_bytes < ceil(p/8)");
  const int n_rows = clust_off[n_clusters];
  if (clust_off[0] != 0 || n_rows <= 0) return kgw_fail_(KGW_ERR_INVALID, "clust_off must start at 0");
  for (int c = 0; c < n_clusters; c++)
    if (clust_off[c + 1] < clust_off[c]) return kgw_fail_(KGW_ERR_INVALID, "clust_off must be non-decreasing");
  for (int r = 0; r < n_rows; r++)
    if (clust_members[r] < 0 || clust_members[r] >= N) return kgw_fail_(KGW_ERR_INVALID, "cluster member out of range");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return kgw_fail_(KGW_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  REF_CUDA(cudaSetDevice(device));

  const int pt = (p + compression - 1) / compression;          // SNPs kept by Metadata::thin
  const int nwords = (pt + 31) / 32;
  const int stride_t = (nwords + 3) / 4 * 4;
  uint8_t *dH = nullptr;
  uint32_t *dHt = nullptr, *dD = nullptr;
  int32_t *dmem = nullptr, *doffc = nullptr, *dfam = nullptr, *drowc = nullptr, *dout = nullptr;
  long long *ddoff = nullptr;
  TileDesc *dtiles = nullptr;
  auto cleanup = [&]() {
    cudaFree(dH); cudaFree(dHt); cudaFree(dD); cudaFree(dmem); cudaFree(doffc); cudaFree(dfam); cudaFree(drowc);
    cudaFree(dout); cudaFree(ddoff); cudaFree(dtiles);
  };
#define REF_TRY(call) do { cudaError_t e2_ = (call); if (e2_ != cudaSuccess) { cleanup(); return kgw_fail_(KGW_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e2_)); } } while (0)
  REF_TRY(kgw_malloc(&dH, (size_t)N * row_bytes));
  REF_TRY(cudaMemcpy(dH, H, (size_t)N * row_bytes, cudaMemcpyHostToDevice));
  REF_TRY(kgw_malloc(&dHt, sizeof(uint32_t) * (size_t)N * stride_t));
  REF_TRY(kgw_malloc(&dmem, sizeof(int32_t) * n_rows));
  REF_TRY(cudaMemcpy(dmem, clust_members, sizeof(int32_t) * n_rows, cudaMemcpyHostToDevice));
  REF_TRY(kgw_malloc(&doffc, sizeof(int32_t) * (n_clusters + 1)));
  REF_TRY(cudaMemcpy(doffc, clust_off, sizeof(int32_t) * (n_clusters + 1), cudaMemcpyHostToDevice));
  if (family_of) {
    REF_TRY(kgw_malloc(&dfam, sizeof(int32_t) * N));
    REF_TRY(cudaMemcpy(dfam, family_of, sizeof(int32_t) * N, cudaMemcpyHostToDevice));
  }
  std::vector<int32_t> rowc(n_rows);
  for (int c = 0; c < n_clusters; c++)
    for (int r = clust_off[c]; r < clust_off[c + 1]; r++) rowc[r] = c;
  REF_TRY(kgw_malloc(&drowc, sizeof(int32_t) * n_rows));
  REF_TRY(cudaMemcpy(drowc, rowc.data(), sizeof(int32_t) * n_rows, cudaMemcpyHostToDevice));
  REF_TRY(kgw_malloc(&dout, sizeof(int32_t) * (size_t)N * K));
  REF_TRY(cudaMemset(dout, 0xff, sizeof(int32_t) * (size_t)N * K));
  REF_TRY(kgw_malloc(&ddoff, sizeof(long long) * n_clusters));

  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  thin_kernel<<<148 * 8, 256>>>(dH, (size_t)row_bytes, dHt, stride_t, N, p, compression, pt);
  REF_TRY(cudaGetLastError());

  // clusters in batches whose pair matrices fit in a quarter of the free memory
  size_t free_b = 0, total_b = 0;
  REF_TRY(cudaMemGetInfo(&free_b, &total_b));
  const size_t budget = std::max<size_t>(free_b / 4, (size_t)64 << 20);
  int c0 = 0;
  while (c0 < n_clusters) {
    std::vector<long long> doff(n_clusters, 0);
    std::vector<TileDesc> tiles;
    size_t words = 0;
    int c1 = c0;
    while (c1 < n_clusters) {
      const size_t n = (size_t)(clust_off[c1 + 1] - clust_off[c1]);
      if (c1 > c0 && (words + n * n) * 4 > budget) break;
      doff[c1] = (long long)words;
      words += n * n;
      for (int i0 = 0; i0 < (int)n; i0 += TILE)
        for (int j0 = 0; j0 < (int)n; j0 += TILE) tiles.push_back({c1, i0, j0, 0});
      c1++;
    }
    if (words * 4 > free_b / 2) { cleanup(); return kgw_fail_(KGW_ERR_NOMEM, "a kinship cluster is too large for its pair matrix to fit in device memory"); }
    cudaFree(dD); dD = nullptr;
    cudaFree(dtiles); dtiles = nullptr;
    if (!tiles.empty()) {
      REF_TRY(kgw_malloc(&dD, std::max<size_t>(words, 1) * 4));
      REF_TRY(kgw_malloc(&dtiles, sizeof(TileDesc) * tiles.size()));
      REF_TRY(cudaMemcpy(dtiles, tiles.data(), sizeof(TileDesc) * tiles.size(), cudaMemcpyHostToDevice));
      REF_TRY(cudaMemcpy(ddoff, doff.data(), sizeof(long long) * n_clusters, cudaMemcpyHostToDevice));
      distance_kernel<<<(unsigned)tiles.size(), 256>>>(dHt, stride_t, nwords, dtiles, doffc, dmem, dfam, ddoff, dD);
      REF_TRY(cudaGetLastError());
      const int r0 = clust_off[c0], r1 = clust_off[c1];   // the rows (cluster members) of this batch
      if (r1 > r0) {
        select_kernel<<<(unsigned)((r1 - r0 + 3) / 4), 128>>>(dD, ddoff, doffc, dmem, drowc, r0, r1, K, dout);
        REF_TRY(cudaGetLastError());
      }
    }
    REF_TRY(cudaDeviceSynchronize());
    c0 = c1;
  }
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  if (kernel_ms) cudaEventElapsedTime(kernel_ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  REF_TRY(cudaMemcpy(ref_out, dout, sizeof(int32_t) * (size_t)N * K, cudaMemcpyDeviceToHost));
  cleanup();
  return KGW_OK;
}
