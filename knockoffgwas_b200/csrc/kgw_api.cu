// knockoffgwas_b200/csrc/kgw_api.cu -- extern "C" launch layer (include/kgw.h).
//
// Replaces the body of Knockoffs::generate() (snpknock2/src/knockoffs.cpp:311-327) for the
// unrelated branch: flat host buffers in, per-locus tables built on the host in the reference's
// own evaluation order, kernels of kgw_device.cuh, packed Hk out.  There is NO CPU fallback:
// without an sm_100 device every compute entry point returns KGW_ERR_CUDA.
#include "../../include/kgw.h"
#include "kgw_device.cuh"
#include "kgw_variants.h"
#include "kgw_family.cuh"
#include "kgw_common.h"

#include <cmath>
#include <chrono>
#include <cstdlib>

#include <algorithm>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <numeric>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define KGW_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      return fail(KGW_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
  } while (0)

// KGW_TIMING=1: stage timings of the host-side entry points on stderr (development aid)
struct StageTimer {
  bool on;
  cudaStream_t st;
  std::chrono::steady_clock::time_point t0;
  explicit StageTimer(cudaStream_t s) : on(getenv("KGW_TIMING") != nullptr), st(s), t0(std::chrono::steady_clock::now()) {}
  void lap(const char *what) {
    if (!on) return;
    cudaStreamSynchronize(st);
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[kgw] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

// Packed rows travel over PCIe as ONE contiguous copy; these kernels convert between the caller's
// row pitch and the 16-byte-aligned device pitch (a pitched cudaMemcpy2D issues one small DMA per row).
__global__ void rows_to_pitched(const uint8_t *__restrict__ src, size_t src_pitch, uint8_t *__restrict__ dst,
                                size_t dst_pitch, size_t row_len, size_t n_rows) {
  const size_t total = n_rows * dst_pitch;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / dst_pitch, c = i - r * dst_pitch;
    dst[i] = c < row_len ? src[r * src_pitch + c] : (uint8_t)0;
  }
}
__global__ void rows_from_pitched(const uint8_t *__restrict__ src, size_t src_pitch, const int32_t *__restrict__ rows,
                                  int64_t row0, uint8_t *__restrict__ dst, size_t row_len, size_t n_rows) {
  const size_t total = n_rows * row_len;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / row_len, c = i - r * row_len;
    const size_t sr = rows ? (size_t)rows[r] : (size_t)row0 + r;
    dst[i] = src[sr * src_pitch + c];
  }
}

// pick the compiled instantiation: smallest lane count whose states-per-lane fit 32 registers-pairs,
// then the smallest S that holds K, then the requested checkpoint spacing
const KgwVariant *pick_variant(int K, int lanes, int B) {
  int L = lanes;
  if (L <= 0) {
    L = 1;
    while (L < 8 && (K + L - 1) / L > 32) L *= 2;
  }
  int n = 0;
  const KgwVariant *tab = nullptr;
  switch (L) {
    case 1: tab = kgw_variants_L1(&n); break;
    case 2: tab = kgw_variants_L2(&n); break;
    case 4: tab = kgw_variants_L4(&n); break;
    case 8: tab = kgw_variants_L8(&n); break;
    default: return nullptr;
  }
  const KgwVariant *best = nullptr;
  for (int i = 0; i < n; i++) {
    const KgwVariant &v = tab[i];
    if (v.L * v.S < K || v.B != B) continue;
    if (!v.pad && v.L * v.S != K) continue;
    if (!best || v.S < best->S || (v.S == best->S && !v.pad)) best = &v;
  }
  return best;
}

// Device-buffer pool.  Knockoffs::generate() is called once per (chromosome, resolution)
// (main.cpp:159-181) with identical shapes; cudaMalloc + cudaFree of a plan's buffers (the scratch slab
// alone is tens of GB) costs as much as the kernel.  A destroyed plan parks its buffers here, keyed by
// (device, size); the next plan of the same shape takes them back.  kgw_release_workspace() frees them.
struct DevPool {
  struct Item { int dev; size_t bytes; void *ptr; };
  std::mutex mu;
  std::vector<Item> items;
  long long live[64] = {};   // bytes handed out by pool_alloc and not given back, per device (under `mu`)
  void add_live(int dev, long long d) { std::lock_guard<std::mutex> lk(mu); if (dev >= 0 && dev < 64) live[dev] += d; }
  long long live_bytes(int dev) { std::lock_guard<std::mutex> lk(mu); return (dev >= 0 && dev < 64) ? live[dev] : 0; }
  // a buffer of at least `need` bytes and at most 1.25 x that (exact shapes in practice), or nullptr
  void *take(int dev, size_t need, size_t *got) {
    std::lock_guard<std::mutex> lk(mu);
    int best = -1;
    for (int i = 0; i < (int)items.size(); i++)
      if (items[i].dev == dev && items[i].bytes >= need && items[i].bytes <= need + need / 4 + 4096 &&
          (best < 0 || items[i].bytes < items[best].bytes)) best = i;
    if (best < 0) return nullptr;
    void *p = items[best].ptr;
    *got = items[best].bytes;
    items.erase(items.begin() + best);
    return p;
  }
  void give(int dev, void *p, size_t n) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64) live[dev] -= (long long)n;
    if (items.size() >= 64) {   // bounded: drop the oldest entry
      cudaSetDevice(items[0].dev);
      cudaFree(items[0].ptr);
      items.erase(items.begin());
      cudaSetDevice(dev);
    }
    items.push_back({dev, n, p});
  }
  // bytes parked for `dev`: memory a new plan of this device can count on (take() or clear() returns it)
  size_t parked(int dev) {
    std::lock_guard<std::mutex> lk(mu);
    size_t n = 0;
    for (const Item &it : items) if (it.dev == dev) n += it.bytes;
    return n;
  }
  // free everything of at least `min_bytes` parked for `dev` (all devices if dev < 0); returns the bytes released
  size_t clear(int dev = -1, size_t min_bytes = 0) {
    std::lock_guard<std::mutex> lk(mu);
    size_t freed = 0;
    int cur = 0;
    cudaGetDevice(&cur);
    for (size_t i = 0; i < items.size();) {
      if ((dev < 0 || items[i].dev == dev) && items[i].bytes >= min_bytes) {
        cudaSetDevice(items[i].dev);
        cudaFree(items[i].ptr);
        freed += items[i].bytes;
        items.erase(items.begin() + i);
      } else {
        i++;
      }
    }
    cudaSetDevice(cur);
    return freed;
  }
};
DevPool g_pool;

// cudaMalloc through the pool.  A large request that no parked buffer fits belongs to a differently shaped
// problem: the large parked buffers of the device are released first (the plan sized itself counting on
// them); on out-of-memory everything parked is released.
void mem_cache_invalidate(int dev);
cudaError_t pool_alloc(int dev, void **out, size_t bytes, size_t *got) {
  if (bytes == 0) bytes = 16;
  *out = g_pool.take(dev, bytes, got);
  if (*out) { g_pool.add_live(dev, (long long)*got); return cudaSuccess; }
  if (bytes >= ((size_t)64 << 20)) g_pool.clear(dev, (size_t)64 << 20);
  cudaError_t e = cudaMalloc(out, bytes);
  if (e == cudaErrorMemoryAllocation) {
    mem_cache_invalidate(dev);   // the cached figure of mem_available() was too optimistic
    if (g_pool.clear(dev) > 0) {
      cudaGetLastError();
      e = cudaMalloc(out, bytes);
    }
  }
  *got = bytes;
  if (e == cudaSuccess) g_pool.add_live(dev, (long long)bytes);
  return e;
}

// Free device memory a new plan can count on: cudaMemGetInfo's figure plus the parked buffers.  cudaMemGetInfo now and
// then takes 5-70 ms on a device that holds tens of GB (measured: the one-shot kgw_generate of C2, 38.5 ms, then takes
// 50-110 ms), so for small plans the figure is derived from a cached query: free + parked + live (the bytes the pool has
// handed out) does not change under anything the pool does.  Other allocations in the process do change it: a plan that
// needs a large share of the memory, every 64th call, and any out-of-memory answer go back to the driver.
struct MemCache { std::mutex mu; bool valid[64] = {}; long long f0[64] = {}; size_t total[64] = {}; int uses[64] = {}; } g_memcache;
void mem_cache_invalidate(int dev) {
  std::lock_guard<std::mutex> lk(g_memcache.mu);
  if (dev >= 0 && dev < 64) g_memcache.valid[dev] = false;
}
cudaError_t mem_available(int dev, double rough_need, size_t *avail, size_t *total) {
  const long long live = g_pool.live_bytes(dev);
  if (dev >= 0 && dev < 64) {
    std::lock_guard<std::mutex> lk(g_memcache.mu);
    if (g_memcache.valid[dev] && ++g_memcache.uses[dev] < 64) {
      const long long a = g_memcache.f0[dev] - live;
      if (a > 0 && rough_need < 0.4 * (double)a) { *avail = (size_t)a; *total = g_memcache.total[dev]; return cudaSuccess; }
    }
  }
  size_t free_b = 0, total_b = 0;
  cudaError_t e = cudaMemGetInfo(&free_b, &total_b);
  if (e != cudaSuccess) return e;
  const size_t parked = g_pool.parked(dev);
  if (dev >= 0 && dev < 64) {
    std::lock_guard<std::mutex> lk(g_memcache.mu);
    g_memcache.valid[dev] = true; g_memcache.uses[dev] = 0; g_memcache.total[dev] = total_b;
    g_memcache.f0[dev] = (long long)free_b + (long long)parked + g_pool.live_bytes(dev);
  }
  *avail = free_b + parked;
  *total = total_b;
  return cudaSuccess;
}

struct Tables {
  std::vector<KgwLocusHMM> hmm;
  std::vector<KgwLocusKO> ko;
  std::vector<KgwEmit> emit;
  std::vector<KgwPass> passes;
};

// Per-locus tables, every formula in the reference's left-to-right evaluation order
// (compiled with -ffp-contract=off).
int build_tables(int p, int K, int W, const double *b, const double *mu, const int32_t *groups,
                 const int32_t *win_start, const int32_t *win_end, Tables &T, bool hmm_too) {
  const double alpha = 1.0 / K;  // knockoffs.cpp:75
  if (hmm_too) {
    T.hmm.resize(p);
    T.emit.resize(p);
    for (int j = 0; j < p; j++) {
      KgwLocusHMM &h = T.hmm[j];
      h.b = b[j];
      h.mu = mu[j];
      // 1.0 - mu, the draw's 1.0 - (1.0 - mu) (knockoffs.cpp:1251-1256), (1.0 - b) * alpha and the products
      // a * theta (:1254) are single IEEE operations on these numbers and are formed on the device
      // sample_emission_copula, lambda = 0.5 (knockoffs.cpp:1469-1502, call site :677)
      const double lambda = 0.5;
      for (int c = 0; c < 8; c++) {
        const int hbit = c & 1, rz = (c >> 1) & 1, rzk = (c >> 2) & 1;
        const double p_orig1 = rz ? (1.0 - mu[j]) : mu[j];
        const double p_ko1 = rzk ? (1.0 - mu[j]) : mu[j];
        double temp = p_orig1 - (1.0 - p_ko1);
        double temp2 = 0.0;
        if (temp < 0) { temp2 = -temp; temp = 0.0; }
        double w0, w1;
        if (hbit == 0) {
          w1 = 1.0 - temp2 / (1.0 - p_orig1);
          w0 = temp2 / (1.0 - p_orig1);
        } else {
          w1 = temp / (p_orig1);
          w0 = 1.0 - temp / (p_orig1);
        }
        w0 = lambda * w0 + (1 - lambda) * (1 - p_ko1);
        w1 = lambda * w1 + (1 - lambda) * p_ko1;
        double sum = 0.0;
        sum += w0;
        sum += w1;  // std::accumulate(w.begin(), w.end(), 0.0), utils.cpp:691
        T.emit[j].ws[c].x = w0;
        T.emit[j].ws[c].y = sum;
      }
    }
  }
  // groups -> first/last locus (set_groups, knockoffs.cpp:96-107)
  if (groups[0] != 0) return fail(KGW_ERR_INVALID, "groups[0] must be 0");
  for (int j = 0; j < p; j++) {
    if (!(b[j] >= 0.0) || !(1.0 - b[j] > 0.0))
      return fail(KGW_ERR_UNSUPPORTED, "b[j] must lie in [0, 1): a locus with zero recombination weight (1 - b_j == 0) is not supported");
    if (!(mu[j] > 0.0) || !(mu[j] < 1.0)) return fail(KGW_ERR_INVALID, "mut_rates[j] must lie in (0, 1)");
  }
  const int G = groups[p - 1] + 1;
  std::vector<int> gfirst(G, -1), glast(G, -1);
  for (int j = 0; j < p; j++) {
    const int g = groups[j];
    if (g < 0 || g >= G || (j > 0 && (g < groups[j - 1] || g > groups[j - 1] + 1)))
      return fail(KGW_ERR_INVALID, "groups must be contiguous, non-decreasing and dense");
    if (gfirst[g] < 0) gfirst[g] = j;
    glast[g] = j;
  }
  T.passes.clear();
  T.ko.clear();
  // sum_k (1 - b_j) * alpha, K terms added one after the other as the reference does (:1326-1328, :1348-1349).  The
  // same sum is needed for every locus: K passes over the loci instead of p chains of K dependent additions -- each
  // locus still sees its K additions in order, but the inner loop runs over independent loci (vector units).
  std::vector<double> sum_a_of(p, 0.0), term(p);
  for (int j = 0; j < p; j++) term[j] = (1.0 - b[j]) * alpha;
  for (int l = 0; l < K; l++)
    for (int j = 0; j < p; j++) sum_a_of[j] += term[j];
  std::vector<double> vstar, vbar;
  for (int w = 0; w < W; w++) {
    KgwPass ps;
    ps.js = (w > 0) ? win_start[w - 1] : win_start[w];       // knockoffs.cpp:1222-1226
    ps.je = (w < W - 1) ? win_end[w + 1] : win_end[w];
    ps.cs = win_start[w];
    ps.ce = win_end[w];
    ps.w = w;
    ps.ko_off = (int)T.ko.size();
    ps.pad0 = ps.pad1 = 0;
    if (ps.js < 0 || ps.je > p || ps.js >= ps.je || ps.cs < ps.js || ps.ce > ps.je || ps.cs >= ps.ce)
      return fail(KGW_ERR_INVALID, "inconsistent window boundaries");
    const int g_start = groups[ps.js], g_end = groups[ps.je - 1] + 1;  // :1282-1283
    if (gfirst[g_start] != ps.js || glast[g_end - 1] != ps.je - 1)
      return fail(KGW_ERR_UNSUPPORTED,
                  "a variant group straddles a window boundary (the reference defines windows at "
                  "boundaries of the coarsest partition, metadata.cpp:86-108)");
    T.ko.resize(T.ko.size() + (size_t)(ps.je - ps.js));
    for (int g = g_start; g < g_end; g++) {
      const int e0 = gfirst[g];
      const int m = glast[g] - e0 + 1;
      vstar.assign(m, 0.0);
      vbar.assign(m, 0.0);
      for (int t = m - 1; t >= 0; t--) {  // :1298-1313
        if (t < m - 1) vstar[t] = vstar[t + 1] * b[e0 + t + 1];
        else vstar[t] = (g < g_end - 1) ? b[e0] : 0.0;
      }
      for (int t = m - 1; t >= 0; t--) {  // :1315-1339
        if (t < m - 1) {
          const double sum_a = sum_a_of[e0 + t + 1];
          vbar[t] = vbar[t + 1] * (sum_a + b[e0 + t + 1]);
          vbar[t] += vstar[t + 1] * (1.0 - b[e0 + t + 1]) * alpha;
        } else {
          vbar[t] = (g < g_end - 1) ? (1.0 - b[gfirst[g + 1]]) * alpha : 1.0;
        }
      }
      for (int t = 0; t < m; t++) {
        const int j = e0 + t;
        KgwLocusKO &k = T.ko[(size_t)ps.ko_off + (j - ps.js)];
        memset(&k, 0, sizeof(k));
        k.vs = vstar[t];
        k.vb = vbar[t];
        k.a = (1.0 - b[j]) * alpha;
        k.apb = (1.0 - b[j]) * alpha + b[j] * 1.0;
        const double aa = k.a * k.a, ab = k.apb * k.a, bb = k.apb * k.apb;  // T_k(z0)*T_k(z0k), :1352-1354
        k.tg = (g == 0) ? k.a : aa;
        k.f1 = ab / aa;
        k.f2 = bb / aa;
        k.fz1 = (g < g_end - 1) ? 1.0 + k.vs / k.vb : 0.0;
        k.sa = sum_a_of[j];  // :1348-1349
        k.vstg = vstar[0] * k.tg;                                // :1364 / :1373 (only read when t == 0)
        k.gw = k.a * k.vb;
        k.rgw = 1.0 / k.gw;
        k.flags = (t == 0 ? KGW_KO_FIRST : 0) | (g == g_end - 1 ? KGW_KO_LASTGRP : 0) |
                  (g == 0 ? KGW_KO_GZERO : 0) | (g == g_start ? KGW_KO_GSTART : 0) |
                  (t == m - 1 ? KGW_KO_LASTINGRP : 0);
        k.z1_idx = (g < g_end - 1) ? gfirst[g + 1] : -1;
        k.z1_next = (g + 1 < g_end - 1) ? gfirst[g + 2] : -1;
      }
    }
    T.passes.push_back(ps);
  }
  return KGW_OK;
}

}  // namespace


namespace {

// ------------------------------------------------------------------------------------------------
// Related branch (generate_related, knockoffs.cpp:329-571): host pre-processing of the families into
// the flat structures of kgw_family.cuh.
// One of the two compiled classes of the family kernel (kgw_family.cuh): its families, launch shape and slot scratch
struct FamClass {
  int count = 0, n_max = 0, rtot = 1, grid = 0, threads = 128, members = 0;
  size_t nslots = 0;
  size_t smem = 0;
  std::vector<int32_t> list;                      // descriptor indices of the class
  int32_t *dlist = nullptr, *dwork = nullptr, *dflg = nullptr;
  double *dmsg = nullptr, *dmrc = nullptr, *dpmg = nullptr;   // (pooled with the plan's buffers)
  uint8_t *dpat = nullptr, *dbytes = nullptr;
};

struct FamilyState {
  int n_families = 0, members_total = 0, SF = 1;
  FamClass cls[3];                                // [0] small families (registers / shared memory), [1] everything else,
                                                  // [2] pairs: phases split over launches (kgw_pairs.cuh)
  std::vector<int32_t> h_members;                 // concatenated, family order
  std::vector<int32_t> h_fam_off, h_seg_off, h_seg_jmin, h_seg_jmax, h_seg_ids_off, h_seg_ids;
  // device
  KgwFamDesc *dfams = nullptr;
  int32_t *dmembers = nullptr, *divstart = nullptr, *dgroups = nullptr, *dgfirst = nullptr, *dglast = nullptr, *dstatus = nullptr;
  int32_t *divlist = nullptr;
  KgwFamRange *dranges = nullptr;
  long long *dphase = nullptr;
  cudaStream_t side = nullptr;             // the related branch runs beside the unrelated one when both fit on the SMs
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool overlap = false;
  bool pair_kernels = true;                // the pair class runs its final draws / knockoffs in kgw_pairs.cuh's kernels
  bool pair_ops_ok = false;                // ... the knockoffs only when every pair's operation list allows it (fam_upload_partition)
  std::vector<int32_t> pair_op_start;      // per batch of the pair class: first DRAW operation in dpair_ops (+ end)
  KgwPairOp *dpair_ops = nullptr;
  size_t cap_pair_ops = 0;
  int32_t *dref_global = nullptr;
  KgwFamInterval *divs = nullptr;
  KgwFamOp *dops = nullptr;
  KgwFamLocus *dloc = nullptr;
  KgwLocusKO *dko = nullptr;
  double *demtab = nullptr;
  size_t cap_ops = 0, cap_gfirst = 0, cap_glast = 0;   // bytes behind dops / dgfirst / dglast (grow-only)
  int32_t *ddbg_z = nullptr, *ddbg_zk = nullptr;
  std::vector<KgwFamDesc> fams;                   // op_off / n_ops are refreshed by set_groups
  KgwFamParams P{};
  ~FamilyState() {
    cudaFree(dfams); cudaFree(dmembers); cudaFree(divstart); cudaFree(dgroups); cudaFree(dgfirst); cudaFree(dglast);
    cudaFree(dstatus); cudaFree(dref_global); cudaFree(divs); cudaFree(dops); cudaFree(dloc); cudaFree(dko);
    cudaFree(demtab); cudaFree(ddbg_z); cudaFree(ddbg_zk);   // dmsg / dpat / dbytes belong to the plan's pool
    cudaFree(dpair_ops);
#ifdef KGW_PHASES
    if (dphase) {
      long long h[8];
      cudaMemcpy(h, dphase, sizeof(h), cudaMemcpyDeviceToHost);
      static const char *nm[7] = {"init patterns", "BP first iteration", "BP later sweeps", "junction draws", "final draws", "knockoff ops", "checks+alleles"};
      long long tot = 0;
      for (int i = 0; i < 7; i++) tot += h[i];
      for (int i = 0; i < 7; i++) fprintf(stderr, "[kgw family phases] %-20s %14lld cycles  %5.1f %%\n", nm[i], h[i], 100.0 * h[i] / (double)tot);
      cudaFree(dphase);
    }
#endif
    cudaFree(divlist); cudaFree(dranges);
    for (FamClass &c : cls) { cudaFree(c.dlist); cudaFree(c.dwork); }
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (side) cudaStreamDestroy(side);
  }
};

// std::set_difference on the inputs as they are (family_bp.cpp:365-378)
int fam_set_difference(const std::vector<int> &A, const std::vector<int> &B, std::vector<int32_t> &out) {
  size_t ia = 0, ib = 0;
  int no = 0;
  while (ia < A.size() && ib < B.size()) {
    if (A[ia] < B[ib]) { out.push_back(A[ia++]); no++; }
    else if (B[ib] < A[ia]) ib++;
    else { ia++; ib++; }
  }
  while (ia < A.size()) { out.push_back(A[ia++]); no++; }
  return no;
}

// neighbour intervals of every member of family f (family_bp.cpp:41-61 run-length encoded along j)
// (lists of any length: the intervals point into `ivlist`; *max_nn = the largest neighbour count of the family)
int fam_build_intervals(const FamilyState &F, int f, int p, std::vector<std::vector<KgwFamInterval>> &out,
                        std::vector<int32_t> &ivlist, int *max_nn) {
  const int m0 = F.h_fam_off[f], n = F.h_fam_off[f + 1] - m0;
  std::unordered_map<int, int> local_of;
  for (int i = n - 1; i >= 0; i--) local_of[F.h_members[m0 + i]] = i;
  auto local = [&](int id) { auto it = local_of.find(id); return it == local_of.end() ? -1 : it->second; };
  *max_nn = 0;
  const int s0 = F.h_seg_off[f], s1 = F.h_seg_off[f + 1];
  out.assign(n, {});
  for (int i = 0; i < n; i++) {
    std::vector<int> cuts = {0, p};
    for (int s = s0; s < s1; s++) {
      bool member = false;
      for (int a = F.h_seg_ids_off[s]; a < F.h_seg_ids_off[s + 1]; a++) member |= (local(F.h_seg_ids[a]) == i);
      if (member) { cuts.push_back(F.h_seg_jmin[s]); cuts.push_back(F.h_seg_jmax[s] + 1); }
    }
    std::sort(cuts.begin(), cuts.end());
    cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
    std::vector<std::vector<int>> lists;
    std::vector<std::pair<int, int>> ranges;
    for (size_t c = 0; c + 1 < cuts.size(); c++) {
      const int j0 = cuts[c], j1 = cuts[c + 1] - 1;
      if (j0 < 0 || j1 >= p || j0 > j1) return fail(KGW_ERR_INVALID, "IBD segment outside [0, p)");
      std::vector<int> nb;  // neighbors[i][j0], push_back order: segments in order, members in id order
      for (int s = s0; s < s1; s++) {
        if (j0 < F.h_seg_jmin[s] || j0 > F.h_seg_jmax[s]) continue;
        bool member = false;
        for (int a = F.h_seg_ids_off[s]; a < F.h_seg_ids_off[s + 1]; a++) member |= (local(F.h_seg_ids[a]) == i);
        if (!member) continue;
        for (int a = F.h_seg_ids_off[s]; a < F.h_seg_ids_off[s + 1]; a++) {
          const int i2 = local(F.h_seg_ids[a]);
          if (i2 < 0) return fail(KGW_ERR_INVALID, "IBD segment lists a haplotype outside its family");
          if (i2 != i) nb.push_back(i2);
        }
      }
      if (!lists.empty() && lists.back() == nb) ranges.back().second = j1;
      else { lists.push_back(nb); ranges.push_back({j0, j1}); }
    }
    for (size_t c = 0; c < lists.size(); c++) {
      KgwFamInterval iv;
      memset(&iv, 0, sizeof(iv));
      iv.j0 = ranges[c].first; iv.j1 = ranges[c].second; iv.nn = (int)lists[c].size();
      *max_nn = std::max(*max_nn, iv.nn);
      iv.nb_off = (int32_t)ivlist.size();
      ivlist.insert(ivlist.end(), lists[c].begin(), lists[c].end());
      iv.dl_off = (int32_t)ivlist.size();
      if (c > 0) iv.ndl = fam_set_difference(lists[c], lists[c - 1], ivlist);
      iv.dr_off = (int32_t)ivlist.size();
      if (c + 1 < lists.size()) iv.ndr = fam_set_difference(lists[c], lists[c + 1], ivlist);
      out[i].push_back(iv);
    }
  }
  return KGW_OK;
}

// FamilyKnockoffs::run (family_knockoffs.cpp:155-270) as a flat list of operations
void fam_build_ops(const FamilyState &F, int f, const int32_t *groups, const std::vector<int> &gfirst,
                   const std::vector<int> &glast, std::vector<KgwFamOp> &ops) {
  const int num_groups = (int)gfirst.size();
  const int m0 = F.h_fam_off[f], n = F.h_fam_off[f + 1] - m0;
  auto local = [&](int id) { for (int i = 0; i < n; i++) if (F.h_members[m0 + i] == id) return i; return -1; };
  const int s0 = F.h_seg_off[f], s1 = F.h_seg_off[f + 1];
  auto push = [&](int kind, int i, int a, int b, int src) {
    KgwFamOp op;
    memset(&op, 0, sizeof(op));
    op.kind = kind; op.i = i; op.a = a; op.b = b; op.src = src;
    ops.push_back(op);
  };
  for (int s = s0; s < s1; s++) {  // trivial knockoffs in the groups of the segment end points (:155-168)
    const int g_min = groups[F.h_seg_jmin[s]], g_max = groups[F.h_seg_jmax[s]];
    for (int a = F.h_seg_ids_off[s]; a < F.h_seg_ids_off[s + 1]; a++) {
      const int i = local(F.h_seg_ids[a]);
      push(KGW_FOP_COPYZ, i, gfirst[g_min], glast[g_min], 0);
      push(KGW_FOP_COPYZ, i, gfirst[g_max], glast[g_max], 0);
    }
  }
  for (int s = s0; s < s1; s++) {  // one draw inside each segment, copied to the other members (:172-193)
    const int i0 = local(F.h_seg_ids[F.h_seg_ids_off[s]]);
    const int j_min = F.h_seg_jmin[s], j_max = F.h_seg_jmax[s];
    int g_begin = groups[j_min] + 1;
    if (groups[j_min] == 0) g_begin = 0;
    int g_end = groups[j_max] - 1;
    if (groups[j_max] == num_groups - 1) g_end = groups[j_max];
    if (g_end >= g_begin) {
      push(KGW_FOP_DRAW, i0, g_begin, g_end, 0);
      for (int a = F.h_seg_ids_off[s] + 1; a < F.h_seg_ids_off[s + 1]; a++)
        push(KGW_FOP_COPYZK, local(F.h_seg_ids[a]), gfirst[g_begin], glast[g_end], i0);
    }
  }
  for (int i = 0; i < n; i++) {  // complement ranges (:196-270)
    std::vector<int> gl, gr;
    for (int s = s0; s < s1; s++) {
      bool member = false;
      for (int a = F.h_seg_ids_off[s]; a < F.h_seg_ids_off[s + 1]; a++) member |= (local(F.h_seg_ids[a]) == i);
      if (member) { gl.push_back(groups[F.h_seg_jmin[s]]); gr.push_back(groups[F.h_seg_jmax[s]]); }
    }
    std::sort(gl.begin(), gl.end());
    std::sort(gr.begin(), gr.end());
    int g_begin = 0, g_end = num_groups - 1;
    for (size_t q = 0; q < gl.size(); q++) {
      g_end = gl[q] - 1;
      if (g_end >= 0 && g_begin <= g_end) push(KGW_FOP_DRAW, i, g_begin, g_end, 0);
      g_begin = gr[q] + 1;
    }
    if (g_end < num_groups) {
      g_end = num_groups - 1;
      if (g_begin < num_groups && g_begin <= g_end) push(KGW_FOP_DRAW, i, g_begin, g_end, 0);
    }
  }
}

}  // namespace

typedef void (*kgw_fam_kernel_fn)(const KgwFamParams);
kgw_fam_kernel_fn kgw_family_kernel_big_(int SF, bool uniform_mu);   // kgw_family_big.cu
int kgw_pair_draw_launch(const KgwFamParams &P, int sm_count, cudaStream_t st);   // kgw_pairs.cu
int kgw_pair_ko_launch(const KgwFamParams &P, const KgwPairOp *ops, int n_ops, int sm_count, cudaStream_t st);

struct kgw_plan {
  int device = 0, sm_count = 148;
  int N = 0, p = 0, K = 0, W = 0, n_haps = 0;
  int64_t row_bytes = 0;
  int stride_w = 0;
  int B = 4, nblk = 0;
  int seg_loci = 0, nseg = 1;
  const KgwVariant *var = nullptr;
  kgw_kernel_fn fn = nullptr;      // the compiled form of the variant chosen for this problem
  int G = 0, n_tasks = 0;
  int grid = 0, threads = 32;
  size_t smem = 0;
  size_t scratch_bytes = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<double> b, mu;
  std::vector<int32_t> win_start, win_end;
  // device buffers
  uint32_t *dHw = nullptr, *dHk = nullptr;
  int32_t *dref = nullptr, *dhaps = nullptr;
  KgwLocusHMM *dhmm = nullptr;
  KgwLocusKO *dko = nullptr, *dko_base = nullptr;   // dko = dko_base + 32 rows of padding
  KgwEmit *demit = nullptr;
  KgwPass *dpass = nullptr;
  double *dckpt = nullptr, *dinj = nullptr, *dse = nullptr, *dcs = nullptr, *dsegck = nullptr;
  uint32_t *dmw = nullptr;
  uint8_t *dstage = nullptr;       // contiguous staging of packed rows (H in, Hk out)
  size_t stage_bytes = 0;
  int32_t *dout_rows = nullptr;
  unsigned char *slab = nullptr;      // one allocation behind the per-slot scratch arrays
  size_t slab_bytes = 0;
  size_t slab_off[6] = {};            // offsets of dckpt, dcs, dse, dsegck, dz, dzk behind dmw
  size_t z_bytes = 0;                 // bytes of each of dz / dzk
  // The sampler's scratch and the family classes' message slots are never in use at the same time (the two branches run
  // one after the other on one stream): with IBD families in the call the scratch lives INSIDE the largest message
  // buffer, and the families get the memory it would have taken (fam_create).  The latent-path bytes the sampler
  // expects zeroed are cleared before every launch then.
  bool scratch_aliased = false;
  void point_scratch(unsigned char *base) {
    dmw = (uint32_t *)base;
    dckpt = (double *)(base + slab_off[0]); dcs = (double *)(base + slab_off[1]); dse = (double *)(base + slab_off[2]);
    dsegck = (double *)(base + slab_off[3]); dz = base + slab_off[4]; dzk = base + slab_off[5];
    P.mwbuf = dmw; P.ckpt = dckpt; P.csbuf = dcs; P.sebuf = dse; P.segck = dsegck; P.zbuf = dz; P.zkbuf = dzk;
  }
  uint8_t *dz = nullptr, *dzk = nullptr;
  int32_t *ddbg_z = nullptr, *ddbg_zk = nullptr;
  double *ddbg_beta = nullptr;
  size_t ko_capacity = 0;
  KgwParams P{};
  std::vector<int32_t> haps;
  FamilyState *fam = nullptr;      // related branch (null when the problem has no families)
  std::vector<int32_t> out_rows;   // rows of Hk written by this plan: unrelated + family members
  // multi-device plan (kgw_options.n_devices > 1): this object only holds one ordinary plan per device, each with
  // its shard of the work list (every device keeps the whole H and the references; Hk rows are disjoint)
  std::vector<kgw_plan *> subs;
  bool primary_has_all = false;    // the first device's Hk already holds the other devices' rows

  std::vector<std::pair<void *, size_t>> owned;   // pooled device buffers of this plan
  template <typename T>
  cudaError_t alloc(T **out, size_t bytes) {
    void *p = nullptr;
    size_t got = 0;
    cudaError_t e = pool_alloc(device, &p, bytes, &got);
    if (e == cudaSuccess) owned.push_back({p, got});
    *out = (T *)p;
    return e;
  }
  template <typename T>
  void release(T *&ptr) {   // back to the pool before the plan dies (buffers that are re-sized)
    for (size_t i = 0; i < owned.size(); i++)
      if (owned[i].first == (void *)ptr) { g_pool.give(device, owned[i].first, owned[i].second); owned.erase(owned.begin() + i); break; }
    ptr = nullptr;
  }

  ~kgw_plan() {
    if (!subs.empty()) {
      for (kgw_plan *s : subs) delete s;
      return;
    }
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
#ifdef KGW_PHASES
    if (P.phase_cycles) {
      long long h[16];
      cudaMemcpy(h, P.phase_cycles, sizeof(h), cudaMemcpyDeviceToHost);
      static const char *nm[12] = {"gather", "sweep1", "s2 wait+stage", "s2 X term", "s2 chunk level", "s2 refinement",
                                   "s3 head+passA", "s3 specials", "s3 reduction", "s3 passB+z1", "s3 draw", "s3 word tail"};
      long long tot = 0;
      for (int i = 0; i < 12; i++) tot += h[i];
      for (int i = 0; i < 12; i++) fprintf(stderr, "[kgw phases] %-16s %12lld cycles  %5.1f %%\n", nm[i], h[i], 100.0 * h[i] / (double)tot);
      cudaFree(P.phase_cycles);
    }
#endif
    delete fam;
    for (auto &o : owned) g_pool.give(device, o.first, o.second);
    cudaFree(ddbg_z); cudaFree(ddbg_zk); cudaFree(ddbg_beta);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream && own_stream) cudaStreamDestroy(stream);
  }
};

namespace {


int fam_upload_partition(kgw_plan *pl, const int32_t *groups) {
  FamilyState *F = pl->fam;
  const int p = pl->p;
  std::vector<int32_t> one_start(1, 0), one_end(1, p);
  Tables T;
  int rc = build_tables(p, pl->K, 1, pl->b.data(), pl->mu.data(), groups, one_start.data(), one_end.data(), T, false);
  if (rc != KGW_OK) return rc;
  const int G = groups[p - 1] + 1;
  std::vector<int> gfirst(G, -1), glast(G, -1);
  for (int j = 0; j < p; j++) { if (gfirst[groups[j]] < 0) gfirst[groups[j]] = j; glast[groups[j]] = j; }
  std::vector<KgwFamOp> ops;
  for (int f = 0; f < F->n_families; f++) {
    F->fams[f].op_off = (int)ops.size();
    fam_build_ops(*F, f, groups, gfirst, glast, ops);
    F->fams[f].n_ops = (int)ops.size() - F->fams[f].op_off;
  }
  // Pair class (kgw_pairs.cuh): the DRAW operations of a batch run side by side in one kernel, the copies afterwards.
  // That is the reference's result when, per pair, (1) nothing written by a DRAW or copied by a COPYZK is written again
  // later, (2) the source of a COPYZK is not written after the copy; a COPYZ whose groups a later operation overwrites
  // is dropped.  The reference's lists for pairs always qualify (the DRAW ranges avoid the boundary groups unless they
  // start at group 0 / end at the last one, family_knockoffs.cpp:172-193); checked here all the same, and if one does
  // not, the class runs its operation lists in order inside family_kernel.
  std::vector<KgwPairOp> pair_ops;
  F->pair_op_start.assign(1, 0);
  F->pair_ops_ok = true;
  {
    FamClass &C = F->cls[2];
    struct Rg { int i, a, b; };   // member, loci a..b
    auto range_of = [&](const KgwFamOp &op) -> Rg {
      if (op.kind == KGW_FOP_DRAW) return Rg{op.i, gfirst[op.a], glast[op.b]};
      return Rg{op.i, op.a, op.b};
    };
    auto hits = [](const Rg &x, int i, int a, int b) { return x.i == i && x.a <= b && a <= x.b; };
    std::vector<char> drop;
    for (int t = 0; t < C.count && F->pair_ops_ok; t++) {
      KgwFamDesc &d = F->fams[C.list[t]];
      const int o0 = d.op_off, o1 = d.op_off + d.n_ops;
      drop.assign(d.n_ops, 0);
      for (int a = o0; a < o1 && F->pair_ops_ok; a++) {
        const Rg ra = range_of(ops[a]);
        for (int b = a + 1; b < o1; b++) {
          const Rg rb = range_of(ops[b]);
          const bool over = hits(rb, ra.i, ra.a, ra.b);
          if (ops[a].kind == KGW_FOP_COPYZ) { if (over) drop[a - o0] = 1; }
          else if (over) F->pair_ops_ok = false;
          if (ops[a].kind == KGW_FOP_COPYZK && hits(rb, ops[a].src, ra.a, ra.b)) F->pair_ops_ok = false;
        }
      }
      if (!F->pair_ops_ok) break;
      for (int a = o0; a < o1; a++) if (drop[a - o0]) { ops[a].a = 0; ops[a].b = -1; }   // an empty copy
    }
    if (const char *e = getenv("KGW_PAIR_KO_GENERIC")) { if (e[0] == '1') F->pair_ops_ok = false; }   // tests: the in-order path
    if (F->pair_ops_ok && C.count > 0) {
      const int nsl = (int)std::max<size_t>(C.nslots, 1);
      for (int f0 = 0; f0 < C.count; f0 += nsl) {
        const size_t first = pair_ops.size();
        for (int t = f0; t < std::min(C.count, f0 + nsl); t++) {
          const KgwFamDesc &d = F->fams[C.list[t]];
          for (int a = d.op_off; a < d.op_off + d.n_ops; a++)
            if (ops[a].kind == KGW_FOP_DRAW) pair_ops.push_back(KgwPairOp{t, ops[a].i, gfirst[ops[a].a], glast[ops[a].b]});
        }
        std::stable_sort(pair_ops.begin() + first, pair_ops.end(),
                         [](const KgwPairOp &x, const KgwPairOp &y) { return x.j1 - x.j0 > y.j1 - y.j0; });   // longest first
        F->pair_op_start.push_back((int32_t)pair_ops.size());
      }
    }
  }
  std::vector<int32_t> gf(gfirst.begin(), gfirst.end()), gl(glast.begin(), glast.end());
  // grow-only device buffers (a plan sees one partition after the other, main.cpp:159-181) and copies on the
  // plan's stream; one synchronisation at the end because the host vectors go out of scope
  auto grow = [&](auto **ptr, size_t *cap, size_t need) -> cudaError_t {
    if (need <= *cap && *ptr) return cudaSuccess;
    cudaError_t e = cudaStreamSynchronize(pl->stream);   // a launch in flight may still read the old buffer
    if (e != cudaSuccess) return e;
    cudaFree(*ptr);
    *ptr = nullptr;
    *cap = need + need / 4 + 64;
    return kgw_malloc(ptr, *cap);
  };
  KGW_CUDA(grow(&F->dops, &F->cap_ops, sizeof(KgwFamOp) * std::max<size_t>(ops.size(), 1)));
  KGW_CUDA(grow(&F->dpair_ops, &F->cap_pair_ops, sizeof(KgwPairOp) * std::max<size_t>(pair_ops.size(), 1)));
  KGW_CUDA(cudaMemcpyAsync(F->dpair_ops, pair_ops.data(), sizeof(KgwPairOp) * pair_ops.size(), cudaMemcpyHostToDevice, pl->stream));
  KGW_CUDA(grow(&F->dgfirst, &F->cap_gfirst, sizeof(int32_t) * G));
  KGW_CUDA(grow(&F->dglast, &F->cap_glast, sizeof(int32_t) * G));
  if (!F->dko) KGW_CUDA(kgw_malloc(&F->dko, sizeof(KgwLocusKO) * p));
  if (!F->dgroups) KGW_CUDA(kgw_malloc(&F->dgroups, sizeof(int32_t) * p));
  KGW_CUDA(cudaMemcpyAsync(F->dops, ops.data(), sizeof(KgwFamOp) * ops.size(), cudaMemcpyHostToDevice, pl->stream));
  KGW_CUDA(cudaMemcpyAsync(F->dgfirst, gf.data(), sizeof(int32_t) * G, cudaMemcpyHostToDevice, pl->stream));
  KGW_CUDA(cudaMemcpyAsync(F->dglast, gl.data(), sizeof(int32_t) * G, cudaMemcpyHostToDevice, pl->stream));
  KGW_CUDA(cudaMemcpyAsync(F->dko, T.ko.data(), sizeof(KgwLocusKO) * p, cudaMemcpyHostToDevice, pl->stream));
  KGW_CUDA(cudaMemcpyAsync(F->dgroups, groups, sizeof(int32_t) * p, cudaMemcpyHostToDevice, pl->stream));
  KGW_CUDA(cudaMemcpyAsync(F->dfams, F->fams.data(), sizeof(KgwFamDesc) * F->n_families, cudaMemcpyHostToDevice, pl->stream));
  KGW_CUDA(cudaStreamSynchronize(pl->stream));
  F->P.ops = F->dops; F->P.grp_first = F->dgfirst; F->P.grp_last = F->dglast; F->P.ko = F->dko;
  F->P.groups = F->dgroups; F->P.n_groups = G;
  return KGW_OK;
}

struct DevInfo { int major = 0, minor = 0, multiProcessorCount = 0; };

typedef ::kgw_fam_kernel_fn fam_kernel_fn;
fam_kernel_fn fam_kernel_for(int SF, bool uniform_mu, bool big) {
  if (big) return kgw_family_kernel_big_(SF, uniform_mu);
  if (uniform_mu)
    return SF == 1 ? kgw::family_kernel<1, true, false> : (SF == 2 ? kgw::family_kernel<2, true, false> : (SF == 4 ? kgw::family_kernel<4, true, false> : kgw::family_kernel<8, true, false>));
  return SF == 1 ? kgw::family_kernel<1, false, false> : (SF == 2 ? kgw::family_kernel<2, false, false> : (SF == 4 ? kgw::family_kernel<4, false, false> : kgw::family_kernel<8, false, false>));
}

int fam_create(kgw_plan *pl, const kgw_problem *prob, const DevInfo &dp) {
  const kgw_families *fi = prob->families;
  const int p = pl->p, K = pl->K, N = pl->N;
  if (!prob->ref_global) return fail(KGW_ERR_INVALID, "families need ref_global (references_global, knockoffs.h:141)");
  if (K > 255) return fail(KGW_ERR_UNSUPPORTED, "K > 255 with IBD families is not built (the related branch keeps 255 as its 'not drawn yet' state)");
  if (!fi->fam_off || !fi->fam_members || !fi->seg_off || !fi->seg_jmin || !fi->seg_jmax || !fi->seg_ids_off || !fi->seg_ids)
    return fail(KGW_ERR_INVALID, "null array in kgw_families");
  FamilyState *F = new FamilyState();
  pl->fam = F;
  F->n_families = fi->n_families;
  F->h_fam_off.assign(fi->fam_off, fi->fam_off + fi->n_families + 1);
  F->h_seg_off.assign(fi->seg_off, fi->seg_off + fi->n_families + 1);
  const int nseg = F->h_seg_off.back();
  F->h_members.assign(fi->fam_members, fi->fam_members + F->h_fam_off.back());
  F->h_seg_jmin.assign(fi->seg_jmin, fi->seg_jmin + nseg);
  F->h_seg_jmax.assign(fi->seg_jmax, fi->seg_jmax + nseg);
  F->h_seg_ids_off.assign(fi->seg_ids_off, fi->seg_ids_off + nseg + 1);
  F->h_seg_ids.assign(fi->seg_ids, fi->seg_ids + F->h_seg_ids_off.back());
  F->members_total = (int)F->h_members.size();
  for (int id : F->h_members) if (id < 0 || id >= N) return fail(KGW_ERR_INVALID, "family member out of range");
  {
    // metadata.unrelated and metadata.related_families are disjoint (metadata.cpp:242-293) and a haplotype belongs
    // to one family: both branches write the member's row of Hk
    std::vector<uint8_t> seen(N, 0);
    for (int id : pl->haps) seen[id] = 1;
    for (int id : F->h_members) {
      if (seen[id] == 1) return fail(KGW_ERR_INVALID, "a family member is also listed in `unrelated`");
      if (seen[id] == 2) return fail(KGW_ERR_INVALID, "a haplotype is listed in two families (or twice in one)");
      seen[id] = 2;
    }
  }
  {
    int32_t mn = 0, mx = 0;
    for (size_t t = 0; t < (size_t)N * K; t++) { mn = std::min(mn, prob->ref_global[t]); mx = std::max(mx, prob->ref_global[t]); }
    if (mn < 0 || mx >= N) return fail(KGW_ERR_INVALID, "global reference index out of range");
  }

  for (int f = 0; f < F->n_families; f++) {
    for (int sg = F->h_seg_off[f]; sg < F->h_seg_off[f + 1]; sg++) {
      if (F->h_seg_jmin[sg] < 0 || F->h_seg_jmax[sg] >= p || F->h_seg_jmin[sg] > F->h_seg_jmax[sg])
        return fail(KGW_ERR_INVALID, "IBD segment outside [0, p)");
      for (int a = F->h_seg_ids_off[sg]; a < F->h_seg_ids_off[sg + 1]; a++) {
        bool found = false;
        for (int m = F->h_fam_off[f]; m < F->h_fam_off[f + 1]; m++) found |= (F->h_members[m] == F->h_seg_ids[a]);
        if (!found) return fail(KGW_ERR_INVALID, "IBD segment lists a haplotype outside its family");
      }
    }
  }
  std::vector<KgwFamInterval> ivs;
  std::vector<int32_t> iv_start, ivlist;
  std::vector<KgwFamRange> ranges;
  F->fams.resize(F->n_families);
  bool pair_class = true;   // development: KGW_PAIR_CLASS=0 sends pairs through the one-launch family kernel
  { const char *e = getenv("KGW_PAIR_CLASS"); if (e && e[0] == '0') pair_class = false; }
  { const char *e = getenv("KGW_PAIR_KERNELS"); if (e && e[0] == '0') F->pair_kernels = false; }   // development: phases split, old code
  for (int f = 0; f < F->n_families; f++) {
    const int n = F->h_fam_off[f + 1] - F->h_fam_off[f];
    if (n < 1) return fail(KGW_ERR_INVALID, "empty family");
    std::vector<std::vector<KgwFamInterval>> per;
    int max_nn = 0;
    int rc = fam_build_intervals(*F, f, p, per, ivlist, &max_nn);
    if (rc != KGW_OK) return rc;
    KgwFamDesc &d = F->fams[f];
    memset(&d, 0, sizeof(d));
    d.n = n; d.mem_off = F->h_fam_off[f]; d.out_off = F->h_fam_off[f];
    d.ivs_off = (int32_t)iv_start.size();
    std::vector<int> bnd;   // loci b > 0 where some member's neighbour list changes between b-1 and b
    for (int i = 0; i < n; i++) {
      iv_start.push_back((int32_t)ivs.size());
      ivs.insert(ivs.end(), per[i].begin(), per[i].end());
      for (const KgwFamInterval &iv : per[i]) if (iv.j0 > 0) bnd.push_back(iv.j0);
    }
    iv_start.push_back((int32_t)ivs.size());  // sentinel
    // special loci [b-2, b+1] (kgw_family.cuh): junction nodes, pinned nodes and their consumers
    std::sort(bnd.begin(), bnd.end());
    d.rng_off = (int32_t)ranges.size();
    for (int b : bnd) {
      const int s0 = std::max(0, b - 2), s1 = std::min(p - 1, b + 1);
      if ((int)ranges.size() > d.rng_off && s0 <= ranges.back().s1 + 1) ranges.back().s1 = std::max(ranges.back().s1, s1);
      else ranges.push_back(KgwFamRange{s0, s1, 0, 0});
    }
    d.n_rng = (int32_t)ranges.size() - d.rng_off;
    // right messages are stored at the special loci and one locus before each range (kgw_family.cuh)
    int stored = 0;
    for (int r = d.rng_off; r < (int)ranges.size(); r++) {
      ranges[r].off = stored;
      stored += ranges[r].s1 - std::max(0, ranges[r].s0 - 1) + 1;
    }
    // class of the family: registers / shared memory / pattern bytes suffice for small ones (kgw_family.cuh)
    const bool big_family = (n > KGW_FAM_MAXN || max_nn > KGW_FAM_MAXNB);
    FamClass &C = F->cls[big_family ? 1 : ((n == 2 && pair_class) ? 2 : 0)];
    C.list.push_back(f);
    C.count++;
    C.members += n;
    C.n_max = std::max(C.n_max, n);
    C.rtot = std::max(C.rtot, std::max(stored, 1));
  }
  // node emission table (family_bp.cpp:437-454) when mu is the same at every locus (init_hmm, knockoffs.cpp:1520)
  bool uniform = true;
  for (int j = 1; j < p; j++) uniform &= (pl->mu[j] == pl->mu[0]);
  std::vector<double> emtab(8 * 256, 0.0);
  if (uniform) {
    const double mu = pl->mu[0];
    for (int nn = 0; nn < 8; nn++)
      for (int pt = 0; pt < (1 << (nn + 1)); pt++) {
        double v = (pt & 1) ? 1.0 - mu : mu;                                  // emission_prob, family_bp.cpp:628
        for (int t = 0; t < nn; t++) v *= ((pt >> (t + 1)) & 1) ? 1.0 - mu : mu;
        if (nn > 0) { const double num_neighbors = nn + 1; v = std::pow(v, 1.0 / num_neighbors); }
        emtab[nn * 256 + pt] = v;
      }
  }
  std::vector<KgwFamLocus> loc(p + 16);   // (zero entries behind the last locus: kgw_pairs.cuh reads whole blocks of loci)
  memset(&loc[p], 0, 16 * sizeof(KgwFamLocus));
  const double alpha = 1.0 / K;
  for (int j = 0; j < p; j++) {
    loc[j].b = pl->b[j];
    loc[j].a_bp = (1.0 - pl->b[j]) / (double)K;
    loc[j].a_al = (1.0 - pl->b[j]) * alpha;
    loc[j].mu = pl->mu[j];
    loc[j].w = (double)K * loc[j].a_bp + loc[j].b;
    loc[j].r1 = loc[j].a_bp / loc[j].w;
  }
  KGW_CUDA(cudaMalloc(&F->dfams, sizeof(KgwFamDesc) * F->n_families));
  KGW_CUDA(cudaMalloc(&F->dmembers, sizeof(int32_t) * F->members_total));
  KGW_CUDA(cudaMalloc(&F->divstart, sizeof(int32_t) * iv_start.size()));
  KGW_CUDA(cudaMalloc(&F->divs, sizeof(KgwFamInterval) * ivs.size()));
  KGW_CUDA(cudaMalloc(&F->dloc, sizeof(KgwFamLocus) * (p + 16)));
  KGW_CUDA(cudaMalloc(&F->demtab, sizeof(double) * emtab.size()));
  KGW_CUDA(cudaMalloc(&F->dref_global, sizeof(int32_t) * (size_t)N * K));
  KGW_CUDA(cudaMalloc(&F->dstatus, sizeof(int32_t)));
  KGW_CUDA(cudaMalloc(&F->divlist, sizeof(int32_t) * std::max<size_t>(ivlist.size(), 1)));
  KGW_CUDA(cudaMemcpy(F->divlist, ivlist.data(), sizeof(int32_t) * ivlist.size(), cudaMemcpyHostToDevice));
  for (FamClass &C : F->cls) {
    KGW_CUDA(cudaMalloc(&C.dwork, sizeof(int32_t)));
    KGW_CUDA(cudaMalloc(&C.dlist, sizeof(int32_t) * std::max<size_t>(C.list.size(), 1)));
    KGW_CUDA(cudaMemcpy(C.dlist, C.list.data(), sizeof(int32_t) * C.list.size(), cudaMemcpyHostToDevice));
  }
  KGW_CUDA(cudaMalloc(&F->dranges, sizeof(KgwFamRange) * std::max<size_t>(ranges.size(), 1)));
  KGW_CUDA(cudaMemcpy(F->dranges, ranges.data(), sizeof(KgwFamRange) * ranges.size(), cudaMemcpyHostToDevice));
  KGW_CUDA(cudaMemcpy(F->dmembers, F->h_members.data(), sizeof(int32_t) * F->members_total, cudaMemcpyHostToDevice));
  KGW_CUDA(cudaMemcpy(F->divstart, iv_start.data(), sizeof(int32_t) * iv_start.size(), cudaMemcpyHostToDevice));
  KGW_CUDA(cudaMemcpy(F->divs, ivs.data(), sizeof(KgwFamInterval) * ivs.size(), cudaMemcpyHostToDevice));
  KGW_CUDA(cudaMemcpy(F->dloc, loc.data(), sizeof(KgwFamLocus) * (p + 16), cudaMemcpyHostToDevice));
  KGW_CUDA(cudaMemcpy(F->demtab, emtab.data(), sizeof(double) * emtab.size(), cudaMemcpyHostToDevice));
  KGW_CUDA(cudaMemcpy(F->dref_global, prob->ref_global, sizeof(int32_t) * (size_t)N * K, cudaMemcpyHostToDevice));
  KGW_CUDA(cudaMemset(F->dstatus, 0, sizeof(int32_t)));

  // one warp per resident family, as many CTAs as the kernel's registers / shared memory keep resident;
  // scratch slots bounded by 80 % of the memory still free (the sampler took at most a quarter before).  The big
  // class is sized first (few families, large slots: at most 40 % of the free memory), the small class takes the rest.
  F->SF = (K + 31) / 32 <= 1 ? 1 : ((K + 31) / 32 <= 2 ? 2 : ((K + 31) / 32 <= 4 ? 4 : 8));
  const size_t KP = (size_t)F->SF * 32;
  // the sampler's scratch moves into the largest message buffer (kgw_plan::scratch_aliased): its slab goes back first,
  // so that the families are sized with that memory
  bool alias = pl->n_haps > 0 && pl->slab != nullptr;
  { const char *e = getenv("KGW_ALIAS_SCRATCH"); if (e && e[0] == '0') alias = false; }            // development
  { const char *e = getenv("KGW_FAMILY_OVERLAP"); if (e && e[0] == '1') alias = false; }           // (both kernels at once)
  if (alias) {
    KGW_CUDA(cudaStreamSynchronize(pl->stream));   // (the memsets of the latent-path bytes)
    pl->release(pl->slab);
  }
  size_t free_b = 0, total_b = 0;
  KGW_CUDA(cudaMemGetInfo(&free_b, &total_b));
  free_b += g_pool.parked(pl->device);
  static const int class_order[3] = {1, 0, 2};
  for (int ci = 0; ci < 3; ci++) {
    const int c = class_order[ci];
    FamClass &C = F->cls[c];
    if (C.count == 0) continue;
    const bool big = (c == 1), pairs = (c == 2);
    size_t per_slot = (size_t)C.n_max * p * ((size_t)K * 9 + 2) + (size_t)C.n_max * C.rtot * K * sizeof(double);
    if (big) per_slot += (size_t)2 * C.n_max * KP * sizeof(double) + (size_t)5 * C.n_max * sizeof(int32_t);
    // (the sampler sized its scratch with the families in mind); the small class leaves the pairs their share
    double budget = (double)free_b * (big ? 0.4 : 0.95);
    if (c == 0 && F->cls[2].count > 0) budget *= (double)C.members / (double)(C.members + F->cls[2].members);
    // CTAs of four warps; a big class whose slots are too large for four of them runs single-warp CTAs
    C.threads = (big && (double)per_slot * 4 > budget) ? 32 : 128;
    const int wpc = C.threads / 32;
    C.smem = (size_t)8 * 256 * sizeof(double) + (big ? 0 : (size_t)wpc * 2 * C.n_max * KP * sizeof(double));
    int ctas_per_sm = 4;
    {
      fam_kernel_fn fnq = fam_kernel_for(F->SF, uniform, big);
      KGW_CUDA(cudaFuncSetAttribute(fnq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C.smem));
      KGW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, fnq, C.threads, C.smem));
      ctas_per_sm = std::max(1, ctas_per_sm);
    }
    const long long slots_mem = (long long)(budget / (double)per_slot);
    if (slots_mem < wpc) return fail(KGW_ERR_NOMEM, "not enough device memory for the message tables of one IBD family");
    size_t nslots;
    if (pairs) {
      // a slot per pair of the batch in flight (the phases are separate launches): as many as fit, whole CTAs
      nslots = (size_t)std::min<long long>(C.count, slots_mem);
      if (const char *e = getenv("KGW_PAIR_SLOTS")) {   // development / tests: force several batches
        const long long v = atoll(e);
        if (v > 0) nslots = (size_t)std::min<long long>((long long)nslots, v);
      }
      C.grid = (int)((nslots + wpc - 1) / wpc);
    } else {
      const long long want = std::min<long long>(C.count, (long long)dp.multiProcessorCount * ctas_per_sm * wpc);
      long long ctas = (want + wpc - 1) / wpc;
      if (ctas * wpc > slots_mem) ctas = slots_mem / wpc;
      C.grid = (int)std::max<long long>(1, ctas);
      nslots = (size_t)C.grid * wpc;
    }
    C.nslots = nslots;
    KGW_CUDA(pl->alloc(&C.dmsg, sizeof(double) * (C.n_max * (size_t)p * K * nslots + 20 * (size_t)K + 64)));   // (rows of padding: kgw_pairs.cuh reads blocks of rows ahead)   // pooled with the plan's buffers
    KGW_CUDA(pl->alloc(&C.dmrc, sizeof(double) * C.n_max * (size_t)C.rtot * K * nslots));
    KGW_CUDA(pl->alloc(&C.dpat, (size_t)C.n_max * p * K * nslots + 20 * (size_t)K + 64));
    KGW_CUDA(pl->alloc(&C.dbytes, (size_t)2 * C.n_max * p * nslots + 64));
    if (big) {
      KGW_CUDA(pl->alloc(&C.dpmg, sizeof(double) * 2 * C.n_max * KP * nslots));
      KGW_CUDA(pl->alloc(&C.dflg, sizeof(int32_t) * 5 * C.n_max * nslots));
    }
    free_b -= std::min(free_b, per_slot * nslots);
  }
  if (alias) {
    FamClass *big = nullptr;
    size_t big_bytes = 0;
    for (FamClass &C : F->cls) {
      const size_t bts = C.dmsg ? sizeof(double) * C.n_max * (size_t)p * K * C.nslots : 0;
      if (bts > big_bytes) { big_bytes = bts; big = &C; }
    }
    if (big && big_bytes >= pl->slab_bytes) {
      pl->point_scratch(reinterpret_cast<unsigned char *>(big->dmsg));
      pl->scratch_aliased = true;
    } else {   // few families: the scratch keeps a buffer of its own after all
      KGW_CUDA(pl->alloc(&pl->slab, pl->slab_bytes));
      pl->point_scratch(pl->slab);
      KGW_CUDA(cudaMemsetAsync(pl->dz, 0, pl->z_bytes, pl->stream));
      KGW_CUDA(cudaMemsetAsync(pl->dzk, 0, pl->z_bytes, pl->stream));
    }
  }
  {
    const char *ov = getenv("KGW_FAMILY_OVERLAP");   // development: "0" / "1"
    F->overlap = ov && ov[0] == '1';   // measured on B200: no gain, the persistent sampler CTAs queue behind the family CTAs
    KGW_CUDA(cudaStreamCreateWithFlags(&F->side, cudaStreamNonBlocking));
    KGW_CUDA(cudaEventCreateWithFlags(&F->ev_fork, cudaEventDisableTiming));
    KGW_CUDA(cudaEventCreateWithFlags(&F->ev_join, cudaEventDisableTiming));
  }

  KgwFamParams &P = F->P;
  P.N = N; P.p = p; P.K = K; P.n_families = 0; P.stride_w = pl->stride_w; P.n_max = 0;   // (class fields: fam_launch)
  P.uniform_mu = uniform ? 1 : 0;
  P.Hw = pl->dHw; P.ref_global = F->dref_global; P.fams = F->dfams; P.members = F->dmembers; P.iv_start = F->divstart;
  P.ivs = F->divs; P.ivlist = F->divlist; P.loc = F->dloc; P.emtab = F->demtab; P.Hk = pl->dHk; P.W = pl->W;
  P.phase_cycles = nullptr;
#ifdef KGW_PHASES
  KGW_CUDA(cudaMalloc(&F->dphase, sizeof(long long) * 8));
  KGW_CUDA(cudaMemset(F->dphase, 0, sizeof(long long) * 8));
  P.phase_cycles = F->dphase;
#endif
  P.ranges = F->dranges; P.seed = pl->P.seed; P.injected_u = pl->dinj; P.dbg_z = P.dbg_zk = nullptr; P.status = F->dstatus;
  return KGW_OK;
}

int fam_launch(kgw_plan *pl, cudaStream_t st) {
  FamilyState *F = pl->fam;
  if (!F || F->n_families == 0) return KGW_OK;
  for (int c = 0; c < 3; c++) {   // one launch per class that has families (pairs: one per phase and batch)
    FamClass &C = F->cls[c];
    if (C.count == 0) continue;
    fam_kernel_fn fn = fam_kernel_for(F->SF, F->P.uniform_mu != 0, c == 1);
    KGW_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C.smem));
    KgwFamParams Pc = F->P;
    Pc.n_families = C.count; Pc.fam_list = C.dlist; Pc.n_max = C.n_max; Pc.rtot = C.rtot;
    Pc.msg = C.dmsg; Pc.mrc = C.dmrc; Pc.pat = C.dpat; Pc.bytes = C.dbytes; Pc.pmg = C.dpmg; Pc.flg = C.dflg; Pc.work = C.dwork;
    Pc.phase_mask = KGW_FPH_ALL; Pc.slot_by_family = 0; Pc.f_base = 0;
    if (c != 2) {
      KGW_CUDA(cudaMemsetAsync(C.dwork, 0, sizeof(int32_t), st));
      fn<<<C.grid, C.threads, C.smem, st>>>(Pc);
      KGW_CUDA(cudaGetLastError());
      continue;
    }
    // pairs: batches of as many pairs as there are slots; belief propagation one warp per pair, then the two
    // chain-sequential phases with several chains per warp (kgw_pairs.cuh), then checks and alleles
    Pc.slot_by_family = 1;
    const int wpc = C.threads / 32;
    for (int f0 = 0; f0 < C.count; f0 += (int)C.nslots) {
      const int f1 = std::min<int>(C.count, f0 + (int)C.nslots);
      const int grid = (f1 - f0 + wpc - 1) / wpc;
      Pc.f_base = f0; Pc.n_families = f1;
      const bool fast_ko = F->pair_kernels && F->pair_ops_ok;
      auto family_phases = [&](int mask) -> int {
        Pc.phase_mask = mask;
        KGW_CUDA(cudaMemsetAsync(C.dwork, 0, sizeof(int32_t), st));
        fn<<<grid, C.threads, C.smem, st>>>(Pc);
        KGW_CUDA(cudaGetLastError());
        return KGW_OK;
      };
      int rc = KGW_OK;
      if ((rc = family_phases(KGW_FPH_INIT | KGW_FPH_BP)) != KGW_OK) return rc;
      if (F->pair_kernels) {
        if (kgw_pair_draw_launch(Pc, pl->sm_count, st) != KGW_OK) return fail(KGW_ERR_CUDA, "launch of the pair draw kernel failed");
      } else if ((rc = family_phases(KGW_FPH_DRAW)) != KGW_OK) return rc;
      if (fast_ko) {
        const int b = f0 / (int)C.nslots;
        const int o0 = F->pair_op_start[b], o1 = F->pair_op_start[b + 1];
        KGW_CUDA(cudaMemsetAsync(C.dwork, 0, sizeof(int32_t), st));
        if (kgw_pair_ko_launch(Pc, F->dpair_ops + o0, o1 - o0, pl->sm_count, st) != KGW_OK) return fail(KGW_ERR_CUDA, "launch of the pair knockoff kernel failed");
      } else if ((rc = family_phases(KGW_FPH_KO)) != KGW_OK) return rc;
      if ((rc = family_phases(KGW_FPH_OUT | (fast_ko ? KGW_FPH_KOCOPY : 0))) != KGW_OK) return rc;
    }
  }
  return KGW_OK;
}

int fam_check_status(kgw_plan *pl) {
  FamilyState *F = pl->fam;
  if (!F || F->n_families == 0) return KGW_OK;
  int32_t st = 0;
  KGW_CUDA(cudaMemcpy(&st, F->dstatus, sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (st == -10) return fail(KGW_ERR_INVALID, "Missing knockoff variants (family_knockoffs.cpp:272-277)");
  if (st == -11) return fail(KGW_ERR_INVALID, "Inconsistency found on Zk (family_knockoffs.cpp:280-291)");
  return KGW_OK;
}

int upload_partition(kgw_plan *pl, const int32_t *groups, bool hmm_too) {
  Tables T;
  int rc = build_tables(pl->p, pl->K, pl->W, pl->b.data(), pl->mu.data(), groups, pl->win_start.data(),
                        pl->win_end.data(), T, hmm_too);
  if (rc != KGW_OK) return rc;
  if (hmm_too) {
    // the kernel stages whole 32-locus words of the per-locus tables with bulk copies: padded to whole words
    const size_t p32 = ((size_t)pl->p + 31) / 32 * 32;
    KGW_CUDA(pl->alloc(&pl->dhmm, sizeof(KgwLocusHMM) * p32));
    KGW_CUDA(pl->alloc(&pl->demit, sizeof(KgwEmit) * p32));
    KGW_CUDA(cudaMemsetAsync(pl->dhmm, 0, sizeof(KgwLocusHMM) * p32, pl->stream));
    KGW_CUDA(cudaMemsetAsync(pl->demit, 0, sizeof(KgwEmit) * p32, pl->stream));
    KGW_CUDA(cudaMemcpyAsync(pl->dhmm, T.hmm.data(), sizeof(KgwLocusHMM) * pl->p, cudaMemcpyHostToDevice, pl->stream));
    KGW_CUDA(cudaMemcpyAsync(pl->demit, T.emit.data(), sizeof(KgwEmit) * pl->p, cudaMemcpyHostToDevice, pl->stream));
    KGW_CUDA(pl->alloc(&pl->dpass, sizeof(KgwPass) * pl->W));
  }
  if (T.ko.size() > pl->ko_capacity) {
    // 32 rows of padding on either side: the kernel copies whole words of a pass's rows, and a pass may start or end
    // inside a word (the padding rows are never used)
    KGW_CUDA(cudaStreamSynchronize(pl->stream));
    pl->release(pl->dko_base);
    KGW_CUDA(pl->alloc(&pl->dko_base, sizeof(KgwLocusKO) * (T.ko.size() + 64)));
    KGW_CUDA(cudaMemsetAsync(pl->dko_base, 0, sizeof(KgwLocusKO) * (T.ko.size() + 64), pl->stream));
    pl->dko = pl->dko_base + 32;
    pl->ko_capacity = T.ko.size();
  }
  KGW_CUDA(cudaMemcpyAsync(pl->dko, T.ko.data(), sizeof(KgwLocusKO) * T.ko.size(), cudaMemcpyHostToDevice, pl->stream));
  KGW_CUDA(cudaMemcpyAsync(pl->dpass, T.passes.data(), sizeof(KgwPass) * T.passes.size(), cudaMemcpyHostToDevice, pl->stream));
  KGW_CUDA(cudaStreamSynchronize(pl->stream));  // T goes out of scope
  pl->P.hmm = pl->dhmm;
  pl->P.emit = pl->demit;
  pl->P.ko = pl->dko;
  pl->P.passes = pl->dpass;
  pl->P.n_passes = (int)T.passes.size();
  if (pl->fam && pl->fam->n_families > 0) return fam_upload_partition(pl, groups);
  return KGW_OK;
}

// The two branches write disjoint rows of Hk and share only read-only inputs, so they may run side by side
// (KGW_FAMILY_OVERLAP=1: family kernel first on a side stream, the sampler fills the remaining registers / shared
// memory).  Default: one after the other -- measured on a C5-shaped shard (512 pairs + 12 000 unrelated haplotypes
// x 60 000 SNPs, K=100): 671 ms in sequence, 681 ms side by side.
int launch(kgw_plan *pl) {
  FamilyState *F = pl->fam;
  const bool both = pl->n_haps > 0 && F && F->n_families > 0;
  if (both && F->overlap) {
    KGW_CUDA(cudaEventRecord(F->ev_fork, pl->stream));
    KGW_CUDA(cudaStreamWaitEvent(F->side, F->ev_fork, 0));
    int rc = fam_launch(pl, F->side);
    if (rc != KGW_OK) return rc;
    KGW_CUDA(cudaEventRecord(F->ev_join, F->side));
  }
  if (pl->n_haps > 0) {
    if (pl->scratch_aliased) {   // the family kernels of the previous run wrote over the latent-path bytes
      KGW_CUDA(cudaMemsetAsync(pl->dz, 0, pl->z_bytes, pl->stream));
      KGW_CUDA(cudaMemsetAsync(pl->dzk, 0, pl->z_bytes, pl->stream));
    }
    pl->fn<<<pl->grid, pl->threads, pl->smem, pl->stream>>>(pl->P);
    KGW_CUDA(cudaGetLastError());
  }
  if (both && F->overlap) {
    KGW_CUDA(cudaStreamWaitEvent(pl->stream, F->ev_join, 0));
    return KGW_OK;
  }
  return fam_launch(pl, pl->stream);
}

}  // namespace

namespace {

// ------------------------------------------------------------------------------------------------
// Several GPUs in one process (kgw_options.n_devices > 1).  The reference fans Knockoffs::generate out over
// worker threads inside the call (generate_unrelated, knockoffs.cpp:573-634: contiguous slices of the sorted
// unrelated list; generate_related, :329-431: families dealt out in descending size); here the same split goes
// over devices.  Haplotypes are independent given H, the references and the tables (:654-692) and families are
// independent of each other (:440-445), so there is no exchange between devices: every device holds H and the
// references, runs an ordinary plan on its share and copies its rows of Hk straight into the caller's buffer
// (disjoint rows).  A family costs KGW_FAMILY_COST unrelated haplotypes per member (measured).
constexpr double KGW_FAMILY_COST = 5.0;

struct SubProblem {
  kgw_problem prob;
  kgw_families fam;
  std::vector<int32_t> fam_off, fam_members, seg_off, seg_jmin, seg_jmax, seg_ids_off, seg_ids;
};

// the split itself (also exported as kgw_partition_work): cut[nd+1] into the sorted unrelated list, fam_dev[nf]
void partition_work(int nu, int nf, const int32_t *fam_off, int nd, std::vector<int> &cut, std::vector<int> &fam_dev) {
  std::vector<double> load(nd, 0.0);
  fam_dev.assign(nf, 0);
  std::vector<int> order(nf);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return fam_off[a + 1] - fam_off[a] > fam_off[b + 1] - fam_off[b]; });
  for (int f : order) {   // largest family first, to the least loaded device
    const int d = (int)(std::min_element(load.begin(), load.end()) - load.begin());
    fam_dev[f] = d;
    load[d] += KGW_FAMILY_COST * (fam_off[f + 1] - fam_off[f]);
  }
  const double total = std::accumulate(load.begin(), load.end(), 0.0) + nu;
  std::vector<double> want(nd);
  double wsum = 0.0;
  for (int d = 0; d < nd; d++) { want[d] = std::max(0.0, total / nd - load[d]); wsum += want[d]; }
  cut.assign(nd + 1, 0);
  double acc = 0.0;
  for (int d = 0; d < nd; d++) {
    acc += wsum > 0.0 ? want[d] * (nu / wsum) : (double)nu / nd;
    cut[d + 1] = std::max(cut[d], std::min(nu, (int)std::llround(acc)));
  }
  cut[nd] = nu;
}

void split_work(const kgw_problem *prob, int nd, std::vector<SubProblem> &out) {
  out.assign(nd, SubProblem());
  const kgw_families *fi = (prob->families && prob->families->n_families > 0) ? prob->families : nullptr;
  const int nf = fi ? fi->n_families : 0;
  std::vector<int> cut, fam_dev;
  partition_work(prob->n_unrelated, nf, fi ? fi->fam_off : nullptr, nd, cut, fam_dev);
  std::vector<std::vector<int>> fam_of(nd);
  for (int f = 0; f < nf; f++) fam_of[fam_dev[f]].push_back(f);   // the caller's family order inside a device
  for (int d = 0; d < nd; d++) {
    SubProblem &S = out[d];
    S.prob = *prob;
    S.prob.unrelated = prob->unrelated ? prob->unrelated + cut[d] : nullptr;
    S.prob.n_unrelated = cut[d + 1] - cut[d];
    S.prob.families = nullptr;
    if (fam_of[d].empty()) continue;
    S.fam_off.push_back(0); S.seg_off.push_back(0); S.seg_ids_off.push_back(0);
    for (int f : fam_of[d]) {
      S.fam_members.insert(S.fam_members.end(), fi->fam_members + fi->fam_off[f], fi->fam_members + fi->fam_off[f + 1]);
      S.fam_off.push_back((int32_t)S.fam_members.size());
      for (int sg = fi->seg_off[f]; sg < fi->seg_off[f + 1]; sg++) {
        S.seg_jmin.push_back(fi->seg_jmin[sg]);
        S.seg_jmax.push_back(fi->seg_jmax[sg]);
        S.seg_ids.insert(S.seg_ids.end(), fi->seg_ids + fi->seg_ids_off[sg], fi->seg_ids + fi->seg_ids_off[sg + 1]);
        S.seg_ids_off.push_back((int32_t)S.seg_ids.size());
      }
      S.seg_off.push_back((int32_t)S.seg_jmin.size());
    }
    S.fam = kgw_families{};
    S.fam.n_families = (int32_t)fam_of[d].size();
    S.fam.fam_off = S.fam_off.data(); S.fam.fam_members = S.fam_members.data(); S.fam.seg_off = S.seg_off.data();
    S.fam.seg_jmin = S.seg_jmin.data(); S.fam.seg_jmax = S.seg_jmax.data();
    S.fam.seg_ids_off = S.seg_ids_off.data(); S.fam.seg_ids = S.seg_ids.data();
  }
  for (int d = 0; d < nd; d++) if (!out[d].fam_off.empty()) out[d].prob.families = &out[d].fam;   // (after the vector stopped moving)
}

// fn(d) on one host thread per device; the first failure's code and message are returned in the caller's thread
template <typename F>
int for_each_device(int nd, F fn) {
  std::vector<int> rc(nd, KGW_OK);
  std::vector<std::string> msg(nd);
  std::vector<std::thread> th;
  for (int d = 0; d < nd; d++)
    th.emplace_back([&, d]() {
      rc[d] = fn(d);
      if (rc[d] != KGW_OK) msg[d] = g_err;   // thread-local text of this worker
    });
  for (auto &t : th) t.join();
  for (int d = 0; d < nd; d++)
    if (rc[d] != KGW_OK) return fail(rc[d], "device shard " + std::to_string(d) + ": " + msg[d]);
  return KGW_OK;
}

int create_multi(const kgw_problem *prob, const kgw_options &o, kgw_plan **plan_out) {
  const int nd = o.n_devices;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(KGW_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  if (o.stream) return fail(KGW_ERR_INVALID, "a caller stream belongs to one device: leave `stream` NULL with n_devices > 1");
  if (prob->injected_u == nullptr && prob->n_unrelated > 0 && !prob->unrelated) return fail(KGW_ERR_INVALID, "bad unrelated list");
  std::vector<int> devs(nd);
  for (int d = 0; d < nd; d++) {
    devs[d] = o.devices ? o.devices[d] : d;
    if (devs[d] < 0 || devs[d] >= ndev) return fail(KGW_ERR_INVALID, "bad device ordinal in `devices`");
  }
  std::vector<SubProblem> sp;
  split_work(prob, nd, sp);
  kgw_plan *pl = new kgw_plan();
  pl->subs.assign(nd, nullptr);
  pl->N = prob->N; pl->p = prob->p; pl->K = prob->K; pl->W = prob->W; pl->row_bytes = prob->row_bytes;
  pl->n_haps = prob->n_unrelated;
  int rc = for_each_device(nd, [&](int d) {
    kgw_options od = o;
    od.device = devs[d];
    od.n_devices = 0;
    od.devices = nullptr;
    return (int)kgw_plan_create(&sp[d].prob, &od, &pl->subs[d]);
  });
  if (rc != KGW_OK) { delete pl; return rc; }
  pl->device = devs[0];
  *plan_out = pl;
  return KGW_OK;
}

// Hk rows of the other devices -> the first device's matrix (what a .bed writer or an NCCL sender on that device
// needs): rows packed on the source device, one peer copy over NVLink by the copy engines, scattered on arrival
__global__ void rows_scatter_pitched(const uint8_t *__restrict__ src, size_t row_len, const int32_t *__restrict__ rows,
                                     uint8_t *__restrict__ dst, size_t dst_pitch, size_t n_rows) {
  const size_t total = n_rows * row_len;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / row_len, c = i - r * row_len;
    dst[(size_t)rows[r] * dst_pitch + c] = src[i];
  }
}

int gather_to_primary(kgw_plan *pl) {
  if (pl->primary_has_all) return KGW_OK;
  kgw_plan *p0 = pl->subs[0];
  const size_t rb = (size_t)(pl->p + 7) / 8;
  for (size_t d = 1; d < pl->subs.size(); d++) {
    kgw_plan *ps = pl->subs[d];
    const size_t n = ps->out_rows.size();
    if (n == 0) continue;
    KGW_CUDA(cudaSetDevice(ps->device));
    if (n * rb > ps->stage_bytes) {
      ps->release(ps->dstage);
      KGW_CUDA(ps->alloc(&ps->dstage, n * rb));
      ps->stage_bytes = n * rb;
    }
    if (!ps->dout_rows) {
      KGW_CUDA(ps->alloc(&ps->dout_rows, sizeof(int32_t) * n));
      KGW_CUDA(cudaMemcpyAsync(ps->dout_rows, ps->out_rows.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, ps->stream));
    }
    rows_from_pitched<<<ps->sm_count * 8, 256, 0, ps->stream>>>((const uint8_t *)ps->dHk, (size_t)ps->stride_w * 4, ps->dout_rows, 0,
                                                                 ps->dstage, rb, n);
    KGW_CUDA(cudaGetLastError());
    KGW_CUDA(cudaStreamSynchronize(ps->stream));
    KGW_CUDA(cudaSetDevice(p0->device));
    uint8_t *tmp = nullptr;
    int32_t *trows = nullptr;
    KGW_CUDA(kgw_malloc(&tmp, n * rb));
    KGW_CUDA(kgw_malloc(&trows, sizeof(int32_t) * n));
    KGW_CUDA(cudaMemcpyPeerAsync(tmp, p0->device, ps->dstage, ps->device, n * rb, p0->stream));
    KGW_CUDA(cudaMemcpyAsync(trows, ps->out_rows.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, p0->stream));
    rows_scatter_pitched<<<p0->sm_count * 8, 256, 0, p0->stream>>>(tmp, rb, trows, (uint8_t *)p0->dHk, (size_t)p0->stride_w * 4, n);
    KGW_CUDA(cudaGetLastError());
    KGW_CUDA(cudaStreamSynchronize(p0->stream));
    cudaFree(tmp);
    cudaFree(trows);
  }
  pl->primary_has_all = true;
  return KGW_OK;
}

}  // namespace

// shared with the other translation units (kgw_common.h)
cudaError_t kgw_dev_malloc_(void **out, size_t bytes) {
  cudaError_t e = cudaMalloc(out, bytes);
  if (e == cudaErrorMemoryAllocation) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (g_pool.clear(dev) > 0) {
      cudaGetLastError();
      e = cudaMalloc(out, bytes);
    }
  }
  return e;
}
int kgw_fail_(int code, const std::string &msg) { return fail(code, msg); }
void *kgw_plan_device_h_(kgw_plan *pl, int *N, int *p, int *stride_w, int *device, void **stream) {
  if (!pl->subs.empty()) pl = pl->subs[0];   // every device holds the whole H
  *N = pl->N; *p = pl->p; *stride_w = pl->stride_w; *device = pl->device; *stream = (void *)pl->stream;
  return (void *)pl->dHw;
}

extern "C" {

int32_t kgw_abi_version(void) { return KGW_ABI_VERSION; }
int32_t kgw_partition_work(int32_t n_unrelated, int32_t n_families, const int32_t *fam_off, int32_t n_devices,
                           int32_t *unrelated_cut, int32_t *family_device) {
  if (n_unrelated < 0 || n_families < 0 || n_devices < 1 || !unrelated_cut || (n_families > 0 && (!fam_off || !family_device)))
    return fail(KGW_ERR_INVALID, "bad argument");
  std::vector<int> cut, fam_dev;
  partition_work(n_unrelated, n_families, fam_off, n_devices, cut, fam_dev);
  for (int d = 0; d <= n_devices; d++) unrelated_cut[d] = cut[d];
  for (int f = 0; f < n_families; f++) family_device[f] = fam_dev[f];
  return KGW_OK;
}
int32_t kgw_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
const char *kgw_last_error(void) { return g_err.c_str(); }

double kgw_uniform(uint64_t seed, uint32_t hap, uint32_t stream, uint32_t locus) {
  const kgw_u32x4 r = kgw_philox_block(seed, hap, stream, locus >> 2);
  const uint32_t w4[4] = {r.x, r.y, r.z, r.w};
  return kgw_word_to_uniform(w4[locus & 3u]);
}

int32_t kgw_plan_create(const kgw_problem *prob, const kgw_options *opt, kgw_plan **plan_out) {
  if (!prob || !plan_out) return fail(KGW_ERR_INVALID, "null argument");
  StageTimer tm(nullptr);
  *plan_out = nullptr;
  if (prob->abi_version != KGW_ABI_VERSION) return fail(KGW_ERR_INVALID, "abi_version mismatch");
  const int N = prob->N, p = prob->p, K = prob->K, W = prob->W;
  if (N <= 0 || p <= 0 || K <= 0 || W <= 0) return fail(KGW_ERR_INVALID, "N, p, K, W must be positive");
  if (prob->row_bytes < (p + 7) / 8) return fail(KGW_ERR_INVALID, "row_bytes < ceil(p/8)");
  if (!prob->H || !prob->ref_local || !prob->b || !prob->mu || !prob->groups || !prob->win_start || !prob->win_end)
    return fail(KGW_ERR_INVALID, "null input buffer");
  if (prob->n_unrelated < 0 || (prob->n_unrelated > 0 && !prob->unrelated))
    return fail(KGW_ERR_INVALID, "bad unrelated list");
  if (K > 256) return fail(KGW_ERR_UNSUPPORTED, "K > 256 is not built (latent states are stored as bytes)");
  kgw_options o{};
  if (opt) o = *opt;
  if (o.n_devices < 0 || o.n_devices > 64) return fail(KGW_ERR_INVALID, "bad n_devices");
  if (o.n_devices > 1) return create_multi(prob, o, plan_out);
  if (o.n_devices == 1 && o.devices) o.device = o.devices[0];
  if (o.lanes_per_hap != 0 && o.lanes_per_hap != 1 && o.lanes_per_hap != 2 && o.lanes_per_hap != 4 && o.lanes_per_hap != 8)
    return fail(KGW_ERR_INVALID, "lanes_per_hap must be 0 (auto), 1, 2, 4 or 8");
  if (o.block_loci != 0 && o.block_loci != 2 && o.block_loci != 4 && o.block_loci != 8)
    return fail(KGW_ERR_INVALID, "block_loci must be 0 (auto), 2, 4 or 8");
  if (o.segment_loci < 0 || (o.segment_loci % 32) != 0)
    return fail(KGW_ERR_INVALID, "segment_loci must be 0 (auto) or a positive multiple of 32");
  const int B = o.block_loci ? o.block_loci : 4;
  if (!pick_variant(K, o.lanes_per_hap, B))
    return fail(KGW_ERR_UNSUPPORTED, "no kernel variant for this K / lanes_per_hap / block_loci");
  for (int i = 0; i < prob->n_unrelated; i++) {
    if (prob->unrelated[i] < 0 || prob->unrelated[i] >= N) return fail(KGW_ERR_INVALID, "unrelated index out of range");
  }
  {   // (minimum and maximum instead of a test per entry: the loop vectorises; N * W * K entries per call)
    int32_t mn = 0, mx = 0;
    const int32_t *r = prob->ref_local;
    const size_t nref = (size_t)N * W * K;
    for (size_t t = 0; t < nref; t++) { mn = std::min(mn, r[t]); mx = std::max(mx, r[t]); }
    if (mn < 0 || mx >= N) return fail(KGW_ERR_INVALID, "reference index out of range");
  }

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(KGW_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  if (o.device < 0 || o.device >= ndev) return fail(KGW_ERR_INVALID, "bad device ordinal");
  KGW_CUDA(cudaSetDevice(o.device));
  DevInfo dp;   // (cudaGetDeviceProperties costs tens of milliseconds per call; three attributes suffice)
  KGW_CUDA(cudaDeviceGetAttribute(&dp.major, cudaDevAttrComputeCapabilityMajor, o.device));
  KGW_CUDA(cudaDeviceGetAttribute(&dp.minor, cudaDevAttrComputeCapabilityMinor, o.device));
  KGW_CUDA(cudaDeviceGetAttribute(&dp.multiProcessorCount, cudaDevAttrMultiProcessorCount, o.device));
  if (dp.major != 10) return fail(KGW_ERR_CUDA, "kernels are built for sm_100a only (found sm_" + std::to_string(dp.major) + std::to_string(dp.minor) + ")");

  kgw_plan *pl = new kgw_plan();
  pl->device = o.device;
  pl->sm_count = dp.multiProcessorCount;
  pl->N = N; pl->p = p; pl->K = K; pl->W = W;
  pl->n_haps = prob->n_unrelated;
  pl->row_bytes = prob->row_bytes;
  pl->stride_w = (int)(((size_t)prob->row_bytes + 15) / 16 * 4);  // rows padded to 16 B
  pl->b.assign(prob->b, prob->b + p);
  pl->mu.assign(prob->mu, prob->mu + p);
  pl->win_start.assign(prob->win_start, prob->win_start + W);
  pl->win_end.assign(prob->win_end, prob->win_end + W);
  pl->haps.assign(prob->unrelated, prob->unrelated + prob->n_unrelated);

  auto guard_fail = [&](int rc) { delete pl; return rc; };
#define KGW_CUDA_PL(call)                                                                                  \
  do {                                                                                                     \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess) return guard_fail(fail(KGW_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_))); \
  } while (0)

  if (o.stream) {
    pl->stream = (cudaStream_t)o.stream;
  } else {
    KGW_CUDA_PL(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
    pl->own_stream = true;
  }
  KGW_CUDA_PL(cudaEventCreate(&pl->ev0));
  KGW_CUDA_PL(cudaEventCreate(&pl->ev1));
  tm.st = pl->stream;
  tm.lap("validate + stream");

  // variant, occupancy-driven persistent grid (one scratch slot per resident warp); the chromosome is cut
  // into segments when the per-slot scratch of a whole pass would not fit in 60 % of the free memory
  size_t free_b = 0, total_b = 0;   // free memory incl. the parked buffers (taken back or released by pool_alloc)
  {
    const bool fam = prob->families && prob->families->n_families > 0;
    const double rough = fam ? 1e30 : (double)prob->n_unrelated * (double)p * (4.0 * K + 16.0) + 2.0 * (double)N * (double)prob->row_bytes;
    KGW_CUDA_PL(mem_available(pl->device, rough, &free_b, &total_b));
  }
  tm.lap("free memory");
  const size_t hw_bytes_est = 2 * (size_t)N * (((size_t)prob->row_bytes + 15) / 16 * 16) + sizeof(int32_t) * (size_t)N * W * K;
  // Variant (lanes per haplotype, states per lane), compiled form and CTA shape: the candidate with the lowest
  // modelled time.  Each form is compiled twice -- `fn` for the densest residency (12 or 16 warps per SM, 168 / 128
  // registers) and `fn_wide` for 8 warps per SM with the full register budget -- and K in (32, 50] can also run with
  // ONE lane per haplotype (50 states per lane, 32 haplotypes per warp, no lane of a haplotype repeating the draw
  // logic of another: 62 instead of 76 warp instructions per haplotype-SNP, but 40 KB of shared memory per warp);
  // likewise K in (64, 100] on two lanes and K in (128, 200] on four.
  // Model: rounds x time of a warp, where rounds = ceil(batches / resident warps) and the time of a warp grows with
  // the number of warps that share its scheduler (measured on B200 at K = 50, ms per 10 000 loci):
  //   two lanes, wide   20.0 alone, 28.2 with a second warp      two lanes, dense   31.6 (two), 38.4 (three)
  //   one lane          26.7 alone, 35.5 with a second warp
  //   one lane, slim    about 8 % more than the line above (the message row of a block is fetched when the block starts)
  // Times of a warp count per batch, and a one-lane batch holds twice the haplotypes of a two-lane batch.
  struct Choice { const KgwVariant *var; kgw_kernel_fn fn; size_t smem_warp; int wpc, occ; double cost; };
  auto evaluate = [&](const KgwVariant *v, bool one_round_only, Choice *best) -> cudaError_t {
    const int Gv = 32 / v->L;
    const int n_tasks = (pl->n_haps + Gv - 1) / Gv;
    const char *force = getenv("KGW_FORCE_WIDE");   // development: "0" / "1"
    for (int form = 0; form < 2; form++) {
      kgw_kernel_fn fn = form ? v->fn_wide : v->fn;
      const size_t smem_warp = form ? v->smem_warp_wide : v->smem_warp;
      if (form == 1 && v->fn_wide == v->fn) continue;
      if (force && v->fn_wide != v->fn && (force[0] == '1') != (form == 1)) continue;
      cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_warp * v->max_wpc));
      if (e != cudaSuccess) return e;
      // single-warp CTAs spread a partial wave most evenly over the SMs; CTAs of max_wpc warps fit more warps per SM
      // (the 1 KB per-CTA shared-memory reservation is paid once) and are used when the single-warp grid cannot hold
      // every haplotype batch at once
      int wpc = 1, occ = 0;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 32, smem_warp);
      if (e != cudaSuccess) return e;
      if (occ * dp.multiProcessorCount < n_tasks && v->max_wpc > 1) {
        int occ2 = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, fn, 32 * v->max_wpc, smem_warp * v->max_wpc);
        if (e != cudaSuccess) return e;
        if (occ2 * v->max_wpc > occ) { wpc = v->max_wpc; occ = occ2; }
      }
      if (occ <= 0) continue;
      if (o.ctas_per_sm > 0) occ = std::min(occ, o.ctas_per_sm);
      const double slots = (double)occ * wpc * dp.multiProcessorCount;
      const double rounds = std::ceil(n_tasks / slots);
      const bool one_lane = (v->S > 32);   // "fat" lanes: 50 states each (one lane per haplotype at K <= 50, two at K <= 100, four at K <= 200)
      // one lane per haplotype: the double-buffered form keeps 5 warps per SM and is only taken when every batch is
      // resident at once; the slim form (one message-row buffer, 8 warps per SM) runs long lists of batches
      if (one_round_only && one_lane && form == 1 && rounds > 1.0) continue;
      const double per_sm = std::min<double>(n_tasks, slots) / dp.multiProcessorCount;
      const double w = std::max(1.0, std::ceil(per_sm / 4.0));          // warps sharing the busiest scheduler
      const bool wide_regs = (form == 1) || (v->fn_wide == v->fn);
      const double t1 = one_lane ? (form == 0 ? 26.7 * 1.08 : 26.7) : (wide_regs ? 20.0 : 24.8);
      const double al = one_lane ? 0.33 : (wide_regs ? 0.41 : 0.274);
      const double cost = rounds * t1 * (1.0 + al * (w - 1.0));
      if (!best->var || cost < best->cost) *best = Choice{v, fn, smem_warp, wpc, occ, cost};
    }
    return cudaSuccess;
  };
  Choice ch{nullptr, nullptr, 0, 1, 0, 0.0};
  KGW_CUDA_PL(evaluate(pick_variant(K, o.lanes_per_hap, B), false, &ch));
  if (o.lanes_per_hap == 0) {
    // the same K on half the lanes, 50 states per lane, where that fits (K in (32, 50], (64, 100], (128, 200])
    const KgwVariant *v0 = pick_variant(K, 0, B);
    if (v0 && v0->L > 1 && (K + v0->L / 2 - 1) / (v0->L / 2) <= 50) {
      const KgwVariant *v1 = pick_variant(K, v0->L / 2, B);
      if (v1) KGW_CUDA_PL(evaluate(v1, true, &ch));
    }
  }
  if (!ch.var) return guard_fail(fail(KGW_ERR_CUDA, "kernel does not fit on an SM"));
  const KgwVariant *var = ch.var;
  const int G = 32 / var->L;
  const size_t p32 = ((size_t)p + 31) / 32 * 32;
  const size_t SPR = (size_t)(var->S + 1) / 2 * 2, SPW = (size_t)(var->S + 3) / 4 * 4;
  const size_t ROW = (size_t)var->L * SPR * G, TILE = (size_t)var->L * SPW * G;
  const size_t CSW = (var->S <= 32 ? 4 : 8) * 32;   // partial sums per locus and warp (kgw::Cfg::CSW)
  size_t slots = 0;
  {
    const int n_tasks = (pl->n_haps + G - 1) / G;
    const int wpc = ch.wpc;
    const int grid = std::max(1, std::min(ch.occ * dp.multiProcessorCount, (n_tasks + wpc - 1) / wpc));
    slots = (size_t)grid * wpc;
    pl->fn = ch.fn;
    pl->threads = 32 * wpc;
    pl->smem = ch.smem_warp * wpc;
    pl->var = var; pl->B = B; pl->G = G; pl->n_tasks = n_tasks; pl->grid = grid;
    pl->nblk = (int)((p + B - 1) / B);
  }
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  // bytes per slot: fixed part (match words, latent paths) and the part proportional to the segment length
  const size_t fixed_slot = up(sizeof(uint32_t) * (size_t)pl->stride_w * TILE) + 2 * up(p32 * G);
  auto seg_slot = [&](size_t ps) {   // rows every B loci, partial sums, totals, segment-top rows
    const size_t nseg = (p32 + ps - 1) / ps;
    return up(sizeof(double) * (ps / B) * ROW) + up(sizeof(double) * ps * CSW) + up(sizeof(double) * ps * G) +
           up(sizeof(double) * (nseg > 1 ? nseg : 1) * ROW);
  };
  size_t PS = p32;
  if (o.segment_loci > 0) {
    PS = std::min<size_t>(p32, (size_t)o.segment_loci);
  } else {
    // With IBD families in the same call most of the memory goes to their message slots (108 MB per pair at
    // 60 000 SNPs, K=100): residency is what bounds that kernel -- a family that is not resident waits for a whole
    // family-time -- while the sampler loses only a few per cent when its passes are cut into short segments
    // (487 ms instead of 467 ms for 3 744 instead of 30 016 loci per segment, C5-shaped shard).  So the families'
    // wish comes first: one slot per family up to the kernel's residency; the sampler takes what is left, between
    // 6 % and 25 % of the free memory.
    const bool with_families = prob->families && prob->families->n_families > 0;
    double budget = 0.6 * (double)free_b - (double)hw_bytes_est;
    if (with_families) {
      const kgw_families *fi = prob->families;
      double want = 0.0;
      long long resident = (long long)dp.multiProcessorCount * 16;   // 4 CTAs of 4 warps per SM
      for (int f = 0; f < fi->n_families && resident > 0; f++, resident--)
        want += (double)(fi->fam_off[f + 1] - fi->fam_off[f]) * (double)p * ((double)K * 9 + 2) * 1.02;
      budget = std::max(0.06 * (double)free_b, std::min(0.25 * (double)free_b, 0.9 * (double)free_b - want)) - (double)hw_bytes_est;
    }
    while (PS > 1024 && (double)slots * (double)(fixed_slot + seg_slot(PS)) > budget) PS = ((PS / 2) + 31) / 32 * 32;
  }
  pl->seg_loci = (int)PS;
  pl->nseg = (int)((p32 + PS - 1) / PS);
  pl->scratch_bytes = slots * (fixed_slot + seg_slot(PS));

  tm.lap("variant + occupancy");
  // inputs
  const size_t hw_bytes = (size_t)N * pl->stride_w * 4;
  KGW_CUDA_PL(pl->alloc(&pl->dHw, hw_bytes));
  KGW_CUDA_PL(pl->alloc(&pl->dHk, hw_bytes));
  KGW_CUDA_PL(cudaMemsetAsync(pl->dHk, 0, hw_bytes, pl->stream));
  pl->stage_bytes = (size_t)N * (size_t)prob->row_bytes;
  KGW_CUDA_PL(pl->alloc(&pl->dstage, pl->stage_bytes));
  KGW_CUDA_PL(cudaMemcpyAsync(pl->dstage, prob->H, pl->stage_bytes, cudaMemcpyHostToDevice, pl->stream));
  rows_to_pitched<<<dp.multiProcessorCount * 8, 256, 0, pl->stream>>>(pl->dstage, (size_t)prob->row_bytes, (uint8_t *)pl->dHw,
                                                                    (size_t)pl->stride_w * 4, (size_t)(p + 7) / 8, (size_t)N);
  KGW_CUDA_PL(cudaGetLastError());
  tm.lap("H upload + re-pitch");
  KGW_CUDA_PL(pl->alloc(&pl->dref, sizeof(int32_t) * (size_t)N * W * K));
  KGW_CUDA_PL(cudaMemcpyAsync(pl->dref, prob->ref_local, sizeof(int32_t) * (size_t)N * W * K, cudaMemcpyHostToDevice, pl->stream));
  KGW_CUDA_PL(pl->alloc(&pl->dhaps, sizeof(int32_t) * std::max(1, pl->n_haps)));
  if (pl->n_haps)
    KGW_CUDA_PL(cudaMemcpyAsync(pl->dhaps, pl->haps.data(), sizeof(int32_t) * pl->n_haps, cudaMemcpyHostToDevice, pl->stream));
  if (prob->injected_u) {
    const size_t ub = sizeof(double) * (size_t)N * W * 3 * p;
    KGW_CUDA_PL(pl->alloc(&pl->dinj, ub));
    KGW_CUDA_PL(cudaMemcpyAsync(pl->dinj, prob->injected_u, ub, cudaMemcpyHostToDevice, pl->stream));
  }
  tm.lap("refs / haps upload");
  // scratch, one slot per resident warp
  const size_t mw_stride = up(sizeof(uint32_t) * (size_t)pl->stride_w * TILE) / sizeof(uint32_t);
  const size_t ckpt_stride = up(sizeof(double) * (PS / B) * ROW) / sizeof(double);
  const size_t cs_stride = up(sizeof(double) * PS * CSW) / sizeof(double);
  const size_t se_stride = up(sizeof(double) * PS * G) / sizeof(double);
  const size_t segck_stride = up(sizeof(double) * (pl->nseg > 1 ? pl->nseg : 1) * ROW) / sizeof(double);
  const size_t z_stride = up(p32 * G);
  {
    const size_t b_mw = mw_stride * sizeof(uint32_t) * slots, b_ck = ckpt_stride * sizeof(double) * slots;
    const size_t b_cs = cs_stride * sizeof(double) * slots, b_se = se_stride * sizeof(double) * slots;
    const size_t b_sg = segck_stride * sizeof(double) * slots, b_z = z_stride * slots;
    const size_t need = b_mw + b_ck + b_cs + b_se + b_sg + 2 * b_z;
    KGW_CUDA_PL(pl->alloc(&pl->slab, need));
    pl->slab_bytes = need;
    pl->slab_off[0] = b_mw; pl->slab_off[1] = b_mw + b_ck; pl->slab_off[2] = b_mw + b_ck + b_cs;
    pl->slab_off[3] = b_mw + b_ck + b_cs + b_se; pl->slab_off[4] = pl->slab_off[3] + b_sg; pl->slab_off[5] = pl->slab_off[4] + b_z;
    pl->z_bytes = b_z;
    unsigned char *base = (unsigned char *)pl->slab;
    pl->dmw = (uint32_t *)base; base += b_mw;
    pl->dckpt = (double *)base; base += b_ck;
    pl->dcs = (double *)base; base += b_cs;
    pl->dse = (double *)base; base += b_se;
    pl->dsegck = (double *)base; base += b_sg;
    pl->dz = base; base += b_z;
    pl->dzk = base;
  }
  KGW_CUDA_PL(cudaMemsetAsync(pl->dz, 0, z_stride * slots, pl->stream));
  KGW_CUDA_PL(cudaMemsetAsync(pl->dzk, 0, z_stride * slots, pl->stream));

  tm.lap("scratch slab");
  KgwParams &P = pl->P;
  P.N = N; P.p = p; P.K = K; P.W = W;
  P.n_haps = pl->n_haps;
  P.stride_w = pl->stride_w;
  P.n_tasks = pl->n_tasks;
  P.Hw = pl->dHw;
  P.ref_local = pl->dref;
  P.haps = pl->dhaps;
  P.Hk = pl->dHk;
  P.seg_loci = pl->seg_loci;
  P.alpha = 1.0 / K;   // knockoffs.cpp:75
  P.mwbuf = pl->dmw;
  P.mw_stride = mw_stride;
  P.ckpt = pl->dckpt;
  P.ckpt_stride = ckpt_stride;
  P.segck = pl->dsegck;
  P.segck_stride = segck_stride;
  P.csbuf = pl->dcs;
  P.cs_stride = cs_stride;
  P.sebuf = pl->dse;
  P.se_stride = se_stride;
  P.zbuf = pl->dz;
  P.zkbuf = pl->dzk;
  P.z_stride = z_stride;
  P.seed = prob->seed;
  P.injected_u = pl->dinj;
  P.dbg_z = P.dbg_zk = nullptr;
  P.dbg_beta = nullptr;
  P.dbg_nblk = 0;
  P.skip_sweeps = o.reserved;
  P.phase_cycles = nullptr;
#ifdef KGW_PHASES
  KGW_CUDA_PL(cudaMalloc(&P.phase_cycles, sizeof(long long) * 16));
  KGW_CUDA_PL(cudaMemset(P.phase_cycles, 0, sizeof(long long) * 16));
#endif  // development timing experiments only (0 in production)

  pl->out_rows = pl->haps;
  if (prob->families && prob->families->n_families > 0) {
    int rcf = fam_create(pl, prob, dp);
    if (rcf != KGW_OK) return guard_fail(rcf);
    pl->out_rows.insert(pl->out_rows.end(), pl->fam->h_members.begin(), pl->fam->h_members.end());
    std::sort(pl->out_rows.begin(), pl->out_rows.end());
    pl->out_rows.erase(std::unique(pl->out_rows.begin(), pl->out_rows.end()), pl->out_rows.end());
  }
  int rc = upload_partition(pl, prob->groups, true);
  if (rc != KGW_OK) return guard_fail(rc);
  KGW_CUDA_PL(cudaStreamSynchronize(pl->stream));
  tm.lap("tables");
  *plan_out = pl;
  return KGW_OK;
}

int32_t kgw_plan_set_groups(kgw_plan *pl, const int32_t *groups) {
  if (!pl || !groups) return fail(KGW_ERR_INVALID, "null argument");
  if (!pl->subs.empty()) {
    pl->primary_has_all = false;
    return for_each_device((int)pl->subs.size(), [&](int d) { return (int)kgw_plan_set_groups(pl->subs[d], groups); });
  }
  KGW_CUDA(cudaSetDevice(pl->device));
  return upload_partition(pl, groups, false);
}

int32_t kgw_plan_run(kgw_plan *pl) {
  if (!pl) return fail(KGW_ERR_INVALID, "null plan");
  if (!pl->subs.empty()) {   // launches are asynchronous: one after the other from this thread
    pl->primary_has_all = false;
    for (kgw_plan *s : pl->subs) {
      int rc = kgw_plan_run(s);
      if (rc != KGW_OK) return rc;
    }
    return KGW_OK;
  }
  KGW_CUDA(cudaSetDevice(pl->device));
  return launch(pl);
}

int32_t kgw_plan_run_timed(kgw_plan *pl, int32_t iters, float *ms_out) {
  if (!pl || iters <= 0 || !ms_out) return fail(KGW_ERR_INVALID, "bad argument");
  if (!pl->subs.empty()) {   // every device times its own stream; the job takes as long as the slowest device
    pl->primary_has_all = false;
    for (kgw_plan *s : pl->subs) {
      KGW_CUDA(cudaSetDevice(s->device));
      KGW_CUDA(cudaStreamSynchronize(s->stream));
    }
    for (kgw_plan *s : pl->subs) {
      KGW_CUDA(cudaSetDevice(s->device));
      KGW_CUDA(cudaEventRecord(s->ev0, s->stream));
      for (int i = 0; i < iters; i++) {
        int rc = launch(s);
        if (rc != KGW_OK) return rc;
      }
      KGW_CUDA(cudaEventRecord(s->ev1, s->stream));
    }
    float worst = 0.f;
    for (kgw_plan *s : pl->subs) {
      KGW_CUDA(cudaSetDevice(s->device));
      KGW_CUDA(cudaEventSynchronize(s->ev1));
      float ms = 0.f;
      KGW_CUDA(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
      worst = std::max(worst, ms);
    }
    *ms_out = worst / iters;
    return KGW_OK;
  }
  KGW_CUDA(cudaSetDevice(pl->device));
  KGW_CUDA(cudaStreamSynchronize(pl->stream));
  KGW_CUDA(cudaEventRecord(pl->ev0, pl->stream));
  for (int i = 0; i < iters; i++) {
    int rc = launch(pl);
    if (rc != KGW_OK) return rc;
  }
  KGW_CUDA(cudaEventRecord(pl->ev1, pl->stream));
  KGW_CUDA(cudaEventSynchronize(pl->ev1));
  float ms = 0.f;
  KGW_CUDA(cudaEventElapsedTime(&ms, pl->ev0, pl->ev1));
  *ms_out = ms / iters;
  return KGW_OK;
}

int32_t kgw_plan_sync(kgw_plan *pl) {
  if (!pl) return fail(KGW_ERR_INVALID, "null plan");
  if (!pl->subs.empty()) {
    for (kgw_plan *s : pl->subs) {
      int rc = kgw_plan_sync(s);
      if (rc != KGW_OK) return rc;
    }
    return KGW_OK;
  }
  KGW_CUDA(cudaSetDevice(pl->device));
  KGW_CUDA(cudaStreamSynchronize(pl->stream));
  return KGW_OK;
}

void *kgw_plan_device_hk(kgw_plan *pl) {
  if (!pl) return nullptr;
  if (!pl->subs.empty()) {   // rows of the other devices are brought to the first device's matrix first
    for (kgw_plan *s : pl->subs) if (kgw_plan_sync(s) != KGW_OK) return nullptr;
    if (gather_to_primary(pl) != KGW_OK) return nullptr;
    return (void *)pl->subs[0]->dHk;
  }
  return (void *)pl->dHk;
}

int32_t kgw_plan_info(kgw_plan *pl, kgw_plan_info_t *out) {
  if (!pl || !out) return fail(KGW_ERR_INVALID, "null argument");
  if (!pl->subs.empty()) {   // shape of the first device's plan; launches and scratch summed over the devices
    int rc = kgw_plan_info(pl->subs[0], out);
    if (rc != KGW_OK) return rc;
    for (size_t d = 1; d < pl->subs.size(); d++) {
      kgw_plan_info_t t;
      rc = kgw_plan_info(pl->subs[d], &t);
      if (rc != KGW_OK) return rc;
      out->launches_per_run += t.launches_per_run;
      out->scratch_bytes += t.scratch_bytes;
    }
    out->reserved = (int32_t)pl->subs.size();
    return KGW_OK;
  }
  out->launches_per_run = (pl->n_haps > 0 ? 1 : 0) + (pl->fam ? (pl->fam->cls[0].count > 0) + (pl->fam->cls[1].count > 0) + 4 * (int)((pl->fam->cls[2].count + std::max<size_t>(pl->fam->cls[2].nslots, 1) - 1) / std::max<size_t>(pl->fam->cls[2].nslots, 1)) : 0);
  out->grid = pl->grid;
  out->block = pl->threads;
  out->lanes_per_hap = pl->var->L;
  out->states_per_lane = pl->var->S;
  out->block_loci = pl->B;
  out->segment_loci = pl->seg_loci;
  out->reserved = 0;
  out->scratch_bytes = (int64_t)pl->scratch_bytes;
  out->row_stride_bytes = (int64_t)pl->stride_w * 4;
  out->smem_bytes = (int64_t)pl->smem;
  return KGW_OK;
}

// debug outputs are produced by one extra run with the debug pointers set
static int run_debug(kgw_plan *pl, kgw_debug *dbg) {
  const size_t np = (size_t)pl->n_haps * pl->p;
  if (dbg->z) KGW_CUDA(cudaMalloc(&pl->ddbg_z, sizeof(int32_t) * std::max<size_t>(np, 1)));
  if (dbg->zk) KGW_CUDA(cudaMalloc(&pl->ddbg_zk, sizeof(int32_t) * std::max<size_t>(np, 1)));
  int nck = 0;
  if (dbg->beta_ckpt) {
    nck = std::min(dbg->n_ckpt, pl->nblk);
    KGW_CUDA(cudaMalloc(&pl->ddbg_beta, sizeof(double) * std::max<size_t>((size_t)pl->n_haps * nck * pl->K, 1)));
    KGW_CUDA(cudaMemsetAsync(pl->ddbg_beta, 0, sizeof(double) * (size_t)pl->n_haps * nck * pl->K, pl->stream));
  }
  pl->P.dbg_z = pl->ddbg_z;
  pl->P.dbg_zk = pl->ddbg_zk;
  pl->P.dbg_beta = pl->ddbg_beta;
  pl->P.dbg_nblk = nck;
  FamilyState *F = pl->fam;
  const size_t fnp = F ? (size_t)F->members_total * pl->p : 0;
  if (F && dbg->fam_z) KGW_CUDA(cudaMalloc(&F->ddbg_z, sizeof(int32_t) * std::max<size_t>(fnp, 1)));
  if (F && dbg->fam_zk) KGW_CUDA(cudaMalloc(&F->ddbg_zk, sizeof(int32_t) * std::max<size_t>(fnp, 1)));
  if (F) { F->P.dbg_z = F->ddbg_z; F->P.dbg_zk = F->ddbg_zk; }
  int rc = launch(pl);
  if (F) { F->P.dbg_z = F->P.dbg_zk = nullptr; }
  pl->P.dbg_z = pl->P.dbg_zk = nullptr;
  pl->P.dbg_beta = nullptr;
  pl->P.dbg_nblk = 0;
  if (rc != KGW_OK) return rc;
  KGW_CUDA(cudaStreamSynchronize(pl->stream));
  if (dbg->z) KGW_CUDA(cudaMemcpy(dbg->z, pl->ddbg_z, sizeof(int32_t) * np, cudaMemcpyDeviceToHost));
  if (dbg->zk) KGW_CUDA(cudaMemcpy(dbg->zk, pl->ddbg_zk, sizeof(int32_t) * np, cudaMemcpyDeviceToHost));
  if (dbg->beta_ckpt) {
    KGW_CUDA(cudaMemcpy(dbg->beta_ckpt, pl->ddbg_beta, sizeof(double) * (size_t)pl->n_haps * nck * pl->K, cudaMemcpyDeviceToHost));
    // checkpoint n of the last pass holds beta at locus min((n+1)*B, je) - 1 (js..je = last pass's padded range)
    const int W = pl->W;
    const int js = (W > 1) ? pl->win_start[W - 2] : pl->win_start[W - 1];
    const int je = pl->win_end[W - 1];
    if (dbg->ckpt_locus) {
      for (int n = 0; n < nck; n++) {
        const int top = std::min((n + 1) * pl->B, je) - 1;
        dbg->ckpt_locus[n] = (n >= js / pl->B && top >= js) ? top : -1;
      }
    }
  }
  dbg->n_ckpt = nck;
  if (F && dbg->fam_z) KGW_CUDA(cudaMemcpy(dbg->fam_z, F->ddbg_z, sizeof(int32_t) * fnp, cudaMemcpyDeviceToHost));
  if (F && dbg->fam_zk) KGW_CUDA(cudaMemcpy(dbg->fam_zk, F->ddbg_zk, sizeof(int32_t) * fnp, cudaMemcpyDeviceToHost));
  if (F) { cudaFree(F->ddbg_z); F->ddbg_z = nullptr; cudaFree(F->ddbg_zk); F->ddbg_zk = nullptr; }
  cudaFree(pl->ddbg_z); pl->ddbg_z = nullptr;
  cudaFree(pl->ddbg_zk); pl->ddbg_zk = nullptr;
  cudaFree(pl->ddbg_beta); pl->ddbg_beta = nullptr;
  return KGW_OK;
}

int32_t kgw_plan_download(kgw_plan *pl, uint8_t *hk_out, kgw_debug *dbg) {
  if (!pl) return fail(KGW_ERR_INVALID, "null plan");
  if (!pl->subs.empty()) {
    if (dbg && (dbg->z || dbg->zk || dbg->beta_ckpt || dbg->fam_z || dbg->fam_zk))
      return fail(KGW_ERR_UNSUPPORTED, "debug outputs are per device: ask for them with n_devices <= 1");
    // every device copies its own rows into the caller's matrix (disjoint rows), one host thread per device
    return for_each_device((int)pl->subs.size(), [&](int d) { return (int)kgw_plan_download(pl->subs[d], hk_out, nullptr); });
  }
  KGW_CUDA(cudaSetDevice(pl->device));
  if (dbg && (dbg->z || dbg->zk || dbg->beta_ckpt || dbg->fam_z || dbg->fam_zk)) {
    int rc = run_debug(pl, dbg);
    if (rc != KGW_OK) return rc;
  }
  KGW_CUDA(cudaStreamSynchronize(pl->stream));
  int rcs = fam_check_status(pl);
  if (rcs != KGW_OK) return rcs;
  const std::vector<int32_t> &rows = pl->out_rows;
  if (hk_out && !rows.empty()) {
    // only the rows of the processed haplotypes are written back (generate() touches only those):
    // packed on the device into [n][row bytes], one contiguous D2H copy, scattered on the host if the
    // rows are not consecutive
    StageTimer tm(pl->stream);
    const size_t rb = (size_t)(pl->p + 7) / 8;
    const size_t n = rows.size();
    bool contiguous = true;
    for (size_t i = 1; i < n; i++)
      if (rows[i] != rows[i - 1] + 1) { contiguous = false; break; }
    if (n * rb > pl->stage_bytes) {
      pl->release(pl->dstage);
      KGW_CUDA(pl->alloc(&pl->dstage, n * rb));
      pl->stage_bytes = n * rb;
    }
    if (!contiguous && !pl->dout_rows) {
      KGW_CUDA(pl->alloc(&pl->dout_rows, sizeof(int32_t) * n));
      KGW_CUDA(cudaMemcpyAsync(pl->dout_rows, rows.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, pl->stream));
    }
    rows_from_pitched<<<pl->sm_count * 8, 256, 0, pl->stream>>>((const uint8_t *)pl->dHk, (size_t)pl->stride_w * 4,
                                                       contiguous ? nullptr : pl->dout_rows, rows[0], pl->dstage, rb, n);
    KGW_CUDA(cudaGetLastError());
    if (contiguous && (size_t)pl->row_bytes == rb) {
      KGW_CUDA(cudaMemcpyAsync(hk_out + (size_t)rows[0] * pl->row_bytes, pl->dstage, n * rb, cudaMemcpyDeviceToHost, pl->stream));
      KGW_CUDA(cudaStreamSynchronize(pl->stream));
    } else {
      std::vector<uint8_t> tmp(n * rb);
      KGW_CUDA(cudaMemcpyAsync(tmp.data(), pl->dstage, n * rb, cudaMemcpyDeviceToHost, pl->stream));
      KGW_CUDA(cudaStreamSynchronize(pl->stream));
      for (size_t i = 0; i < n; i++) memcpy(hk_out + (size_t)rows[i] * pl->row_bytes, tmp.data() + i * rb, rb);
    }
    tm.lap("Hk pack + download");
  }
  return KGW_OK;
}

void kgw_plan_destroy(kgw_plan *pl) { delete pl; }

void kgw_release_workspace(void) { g_pool.clear(); }

int32_t kgw_generate(const kgw_problem *prob, const kgw_options *opt, uint8_t *hk_out, kgw_debug *dbg) {
  kgw_plan *pl = nullptr;
  int rc = kgw_plan_create(prob, opt, &pl);
  if (rc != KGW_OK) return rc;
  StageTimer tm(pl->stream);
  rc = kgw_plan_run(pl);
  tm.lap("kernels");
  if (rc == KGW_OK) rc = kgw_plan_download(pl, hk_out, dbg);
  tm.t0 = std::chrono::steady_clock::now();
  kgw_plan_destroy(pl);
  tm.st = nullptr;
  tm.lap("destroy");
  return rc;
}

}  // extern "C"
