// knockoffgwas_b200/csrc/kgw_bed.cu -- device-side body of the PLINK .bed written after Knockoffs::generate().
//
// Replaces the data loops of writeBED / writeBED_variant (snpknock2/src/plink_interface.cpp:10-35, 67-103):
// the reference walks every (sample, SNP) pair with two single-bit reads of the haplotype-major chaplotype
// rows, on one thread, per output file.  Here the packed haplotype-major matrices that are already in HBM
// (H and the freshly generated Hk) are transposed to SNP-major 2-bit genotype columns on the device:
//   sample s, SNP j:  b0 = H[2s][j], b1 = H[2s+1][j];  code = b0 & b1 | (b0 | b1) << 1   (:20-28)
//   column = ceil(N/2 / 4) bytes, sample s in bits 2(s%4) .. 2(s%4)+1 of byte s/4       (:17-33)
//   SNP j -> columns 2j, 2j+1 = (H, Hk), or (Hk, H) when swap[j]                        (:92-99)
// One CTA transposes a tile of 256 loci x 256 haplotypes.  A thread owns 8 haplotypes x 64 loci: eight 8-byte
// loads, an 8x8 byte transpose (PRMT) and eight 8x8 bit transposes (three mask-shift steps on a 64-bit
// word) give, per locus, the byte of its 8 haplotypes' bits; even/odd bit pairing turns that byte into the
// four 2-bit codes of 4 samples = exactly one .bed byte.  The CTA assembles 32 contiguous output bytes per
// locus in shared memory and writes them as 16-byte (or 4-byte / single-byte) stores, one full 32-byte
// sector per (locus, tile).
#include "kgw_common.h"
#include "../../include/kgw.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

int kgw_fail_(int code, const std::string &msg);                   // kgw_api.cu
void *kgw_plan_device_h_(kgw_plan *pl, int *N, int *p, int *stride_w, int *device, void **stream);   // kgw_api.cu

namespace {

constexpr int BED_LOCI = 512;    // loci per tile: 64 contiguous bytes of every input row
constexpr int BED_HAPS = 1024;   // haplotypes per tile: 128 groups of 8 = 128 contiguous bytes of every output column
constexpr int BED_THREADS = 256;
constexpr int BED_PITCH = BED_HAPS / 8;        // bytes per locus row of the shared tile

// 8x8 bit-matrix transpose of a 64-bit word whose byte r holds row r, bit c = column c (LSB first):
// afterwards byte c holds column c, bit r = row r  (Hacker's Delight 7-3).
__device__ __forceinline__ unsigned long long transpose8x8(unsigned long long x) {
  x = (x & 0xAA55AA55AA55AA55ull) | ((x & 0x00AA00AA00AA00AAull) << 7) | ((x >> 7) & 0x00AA00AA00AA00AAull);
  x = (x & 0xCCCC3333CCCC3333ull) | ((x & 0x0000CCCC0000CCCCull) << 14) | ((x >> 14) & 0x0000CCCC0000CCCCull);
  x = (x & 0xF0F0F0F00F0F0F0Full) | ((x & 0x00000000F0F0F0F0ull) << 28) | ((x >> 28) & 0x00000000F0F0F0F0ull);
  return x;
}

// Hw / Hkw: [N][stride_w] packed rows (LSB-first, stride_w a multiple of 4 words); either may be null.
// out: [ncols][nb] with nb = ceil(N/2/4); col_of(j, which) = 2j + (which ^ swap[j]) when both matrices are
// written, j when only one is.
// Persistent grid over (matrix, haplotype tile, locus tile).  A tile is 1024 haplotypes x 512 loci: every input row
// contributes 64 contiguous bytes (two full 32-byte sectors) and every output column receives 128 contiguous bytes (a
// full line) -- partial-sector writes into columns that lie N/8 bytes apart were what held the first version of this
// kernel to 0.43 TB/s on a 87 500 x 60 000 matrix.
__global__ void __launch_bounds__(BED_THREADS) bed_encode_kernel(const uint32_t *__restrict__ Hw, const uint32_t *__restrict__ Hkw,
                                                                 const uint8_t *__restrict__ swap, uint8_t *__restrict__ out, int N,
                                                                 int p, int stride_w, long long nb) {
  extern __shared__ __align__(16) uint8_t tile[];   // [BED_LOCI][BED_PITCH]: [locus][haplotype group] = one .bed byte each
  const int chunks_per_row = stride_w / 2;           // 64-bit chunks in a packed row
  const bool both = (Hw != nullptr) && (Hkw != nullptr);
  const long long tiles_l = ((long long)p + BED_LOCI - 1) / BED_LOCI, tiles_h = ((long long)N + BED_HAPS - 1) / BED_HAPS;
  const long long n_tiles = tiles_l * tiles_h * 2;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int which = (int)(t / (tiles_l * tiles_h));
    const uint32_t *M = which == 0 ? Hw : Hkw;
    if (M == nullptr) continue;
    const long long tl = (t % (tiles_l * tiles_h)) % tiles_l, th = (t % (tiles_l * tiles_h)) / tiles_l;
    const long long h0 = th * BED_HAPS;
    __syncthreads();   // the previous tile has been written out
    // a task = 8 haplotypes x 64 loci: (BED_HAPS / 8) x (BED_LOCI / 64) of them per tile; consecutive threads take
    // consecutive 64-locus chunks of the same rows (one warp reads 4 haplotype groups x 64 contiguous bytes per row).
    // The eight chunks of a warp write tile rows that lie 64 rows apart (the same bank): the 32-bit words of a tile row
    // are stored at word index w ^ (chunk << 2), undone when the row is read out.
    constexpr int CH = BED_LOCI / 64;
    for (int task = threadIdx.x; task < (BED_HAPS / 8) * CH; task += BED_THREADS) {
      const int cq = task % CH, gq = task / CH;
      const long long chunk = tl * CH + cq;
      uint32_t lo[8], hi[8];
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const long long hap = h0 + 8 * gq + r;
        uint2 v = make_uint2(0u, 0u);
        if (hap < N && chunk < chunks_per_row) v = __ldg(reinterpret_cast<const uint2 *>(M + (size_t)hap * stride_w) + chunk);
        lo[r] = v.x;
        hi[r] = v.y;
      }
#pragma unroll
      for (int b = 0; b < 8; b++) {
        // byte b of every row -> one 64-bit word (byte r = haplotype r, bit c = locus 8b + c)
        const uint32_t *w = (b < 4) ? lo : hi;
        const unsigned sel = (unsigned)(b & 3) | ((4u + (unsigned)(b & 3)) << 4);
        const uint32_t t01 = __byte_perm(w[0], w[1], sel), t23 = __byte_perm(w[2], w[3], sel);
        const uint32_t t45 = __byte_perm(w[4], w[5], sel), t67 = __byte_perm(w[6], w[7], sel);
        const unsigned long long x = (unsigned long long)__byte_perm(t01, t23, 0x5410) |
                                     ((unsigned long long)__byte_perm(t45, t67, 0x5410) << 32);
        const unsigned long long y = transpose8x8(x);   // byte c = the 8 haplotypes' bits at locus 8b + c
        const unsigned long long even = y & 0x5555555555555555ull, odd = (y >> 1) & 0x5555555555555555ull;
        const unsigned long long code = (even & odd) | ((even | odd) << 1);   // four 2-bit genotypes per byte (:20-28)
#pragma unroll
        for (int c = 0; c < 8; c++)
          tile[(size_t)(64 * cq + 8 * b + c) * BED_PITCH + 4 * ((gq >> 2) ^ (cq << 2)) + (gq & 3)] = (uint8_t)(code >> (8 * c));
      }
    }
    __syncthreads();
    // BED_LOCI columns x 128 bytes: a warp writes one column piece per instruction (32 lanes x 4 bytes = one line)
    const long long byte0 = h0 / 8;   // first output byte of this tile within the column
    const int n = (int)((nb - byte0) < (BED_HAPS / 8) ? (nb - byte0) : (BED_HAPS / 8));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int l = wid; l < BED_LOCI; l += BED_THREADS / 32) {
      const long long j = tl * BED_LOCI + l;
      if (j >= p || n <= 0) continue;
      const long long col = both ? 2 * j + (which ^ ((swap != nullptr && swap[j]) ? 1 : 0)) : j;
      uint8_t *dst = out + (size_t)col * (size_t)nb + (size_t)byte0;
      const uint8_t *src = tile + (size_t)l * BED_PITCH;
      const int sw = ((l >> 6) & 7) << 2;   // word swizzle of this tile row
      if (n == BED_HAPS / 8 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
        reinterpret_cast<uint32_t *>(dst)[lane] = reinterpret_cast<const uint32_t *>(src)[lane ^ sw];
      } else {
        for (int i = lane; i < n; i += 32) dst[i] = src[4 * ((i >> 2) ^ sw) + (i & 3)];
      }
    }
  }
}

}  // namespace

namespace {

#define BED_CUDA(call)                                                                                 \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) return kgw_fail_(KGW_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

int encode_on_device(const uint32_t *dH, const uint32_t *dHk, int N, int p, int stride_w, const uint8_t *swap_host,
                     uint8_t *out_host, uint8_t *out_dev_in, cudaStream_t st, float *kernel_ms) {
  if ((N & 1) != 0) return kgw_fail_(KGW_ERR_INVALID, "the number of haplotypes must be even (plink_interface.cpp:259)");
  const long long nb = ((long long)N / 2 + 3) / 4;
  const long long ncols = (dH && dHk) ? 2LL * p : p;
  const size_t out_bytes = (size_t)ncols * (size_t)nb;
  uint8_t *dswap = nullptr, *dout = out_dev_in;
  if (swap_host && dH && dHk) {
    BED_CUDA(kgw_malloc(&dswap, (size_t)p));
    BED_CUDA(cudaMemcpyAsync(dswap, swap_host, (size_t)p, cudaMemcpyHostToDevice, st));
  }
  if (!dout) BED_CUDA(kgw_malloc(&dout, out_bytes));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (kernel_ms) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
  static bool attr_set = false;
  const size_t smem = (size_t)BED_LOCI * BED_PITCH;
  if (!attr_set) {
    BED_CUDA(cudaFuncSetAttribute(bed_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long n_tiles = (((long long)p + BED_LOCI - 1) / BED_LOCI) * (((long long)N + BED_HAPS - 1) / BED_HAPS) * 2;
  const int grid = (int)std::min<long long>(n_tiles, (long long)sms * 3);   // persistent: 3 CTAs of 72 KB per SM
  bed_encode_kernel<<<grid, BED_THREADS, smem, st>>>(dH, dHk, dswap, dout, N, p, stride_w, nb);
  BED_CUDA(cudaGetLastError());
  if (kernel_ms) {
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(kernel_ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  if (out_host) BED_CUDA(cudaMemcpyAsync(out_host, dout, out_bytes, cudaMemcpyDeviceToHost, st));
  BED_CUDA(cudaStreamSynchronize(st));
  cudaFree(dswap);
  if (!out_dev_in) cudaFree(dout);
  return KGW_OK;
}

}  // namespace

extern "C" {

int64_t kgw_bed_column_bytes(int32_t N) { return ((int64_t)N / 2 + 3) / 4; }

int32_t kgw_plan_encode_bed(kgw_plan *plan, const uint8_t *swap, int32_t include_genotypes, uint8_t *bed_out, float *kernel_ms) {
  if (!plan || !bed_out) return kgw_fail_(KGW_ERR_INVALID, "null argument");
  int N = 0, p = 0, stride_w = 0, device = 0;
  void *stream = nullptr;
  void *dHk = kgw_plan_device_hk(plan);
  void *dH = kgw_plan_device_h_(plan, &N, &p, &stride_w, &device, &stream);
  BED_CUDA(cudaSetDevice(device));
  return encode_on_device(include_genotypes ? (const uint32_t *)dH : nullptr, (const uint32_t *)dHk, N, p, stride_w,
                          include_genotypes ? swap : nullptr, bed_out, nullptr, (cudaStream_t)stream, kernel_ms);
}

int32_t kgw_bed_encode(int32_t N, int32_t p, int64_t row_bytes, const uint8_t *H, const uint8_t *Hk, const uint8_t *swap,
                       int32_t device, uint8_t *bed_out) {
  if (N <= 0 || p <= 0 || (!H && !Hk) || !bed_out) return kgw_fail_(KGW_ERR_INVALID, "bad argument");
  if (row_bytes < (p + 7) / 8) return kgw_fail_(KGW_ERR_INVALID, "row_bytes < ceil(p/8)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return kgw_fail_(KGW_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  BED_CUDA(cudaSetDevice(device));
  const int stride_w = (int)(((size_t)row_bytes + 15) / 16 * 4);
  const size_t bytes = (size_t)N * stride_w * 4;
  uint32_t *d[2] = {nullptr, nullptr};
  const uint8_t *src[2] = {H, Hk};
  for (int m = 0; m < 2; m++) {
    if (!src[m]) continue;
    BED_CUDA(kgw_malloc(&d[m], bytes));
    BED_CUDA(cudaMemset(d[m], 0, bytes));
    BED_CUDA(cudaMemcpy2D(d[m], (size_t)stride_w * 4, src[m], (size_t)row_bytes, (size_t)(p + 7) / 8, N, cudaMemcpyHostToDevice));
  }
  const int rc = encode_on_device(d[0], d[1], N, p, stride_w, swap, bed_out, nullptr, nullptr, nullptr);
  cudaFree(d[0]);
  cudaFree(d[1]);
  return rc;
}

}  // extern "C"
