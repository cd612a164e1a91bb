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
constexpr int BED_HAPS = 1024;   // haplotypes a tile owns: 128 groups of 8 = 128 bytes of every output column
constexpr int BED_OVER = 128;    // ... and reads beyond them (the 16 bytes that complete its last aligned 16-byte chunk)
constexpr int BED_THREADS = 256;
constexpr int BED_PITCH = (BED_HAPS + BED_OVER) / 8;   // bytes per locus row of the shared tile (144)
constexpr int BED_PW = BED_PITCH / 4;                  // ... in 32-bit words (36)

// 8x8 bit-matrix transpose of a 64-bit word (xh:xl) whose byte r holds row r, bit c = column c (LSB first): afterwards
// byte c holds column c, bit r = row r (Hacker's Delight 7-3, the exchange form).  The first two steps never move a bit
// across the 32-bit halves (the masks clear what a shift would carry over), the third exchanges the high nibbles of the
// low word with the low nibbles of the high word: 22 32-bit instructions.
__device__ __forceinline__ void transpose8x8(uint32_t &xl, uint32_t &xh) {
  uint32_t t;
  t = (xl ^ (xl >> 7)) & 0x00AA00AAu;  xl = xl ^ t ^ (t << 7);
  t = (xh ^ (xh >> 7)) & 0x00AA00AAu;  xh = xh ^ t ^ (t << 7);
  t = (xl ^ (xl >> 14)) & 0x0000CCCCu; xl = xl ^ t ^ (t << 14);
  t = (xh ^ (xh >> 14)) & 0x0000CCCCu; xh = xh ^ t ^ (t << 14);
  t = (xl ^ __funnelshift_r(xl, xh, 28)) & 0xF0F0F0F0u;
  xl = xl ^ t ^ (t << 28);
  xh = xh ^ (t >> 4);
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
  __shared__ uint8_t swp[BED_LOCI];
  const int chunks_per_row = stride_w / 2;           // 64-bit chunks in a packed row
  const bool both = (Hw != nullptr) && (Hkw != nullptr);
  const long long tiles_l = ((long long)p + BED_LOCI - 1) / BED_LOCI, tiles_h = ((long long)N + BED_HAPS - 1) / BED_HAPS;
  const long long n_tiles = tiles_l * tiles_h * 2;
  // word w of the tile row of a locus in 64-locus chunk cq sits at (w + 4 cq) mod BED_PW: the eight chunks of a warp write
  // rows that lie 64 rows apart (the same bank without the rotation)
  auto rot = [](int w, int cq) { const int x = w + 4 * cq; return x >= BED_PW ? x - BED_PW : x; };
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int which = (int)(t / (tiles_l * tiles_h));
    const uint32_t *M = which == 0 ? Hw : Hkw;
    if (M == nullptr) continue;
    const long long tl = (t % (tiles_l * tiles_h)) % tiles_l, th = (t % (tiles_l * tiles_h)) / tiles_l;
    const long long h0 = th * BED_HAPS;
    __syncthreads();   // the previous tile has been written out
    for (int l = threadIdx.x; l < BED_LOCI; l += BED_THREADS) {   // column order of the tile's loci (:92-99)
      const long long j = tl * BED_LOCI + l;
      swp[l] = (swap != nullptr && j < p) ? swap[j] : (uint8_t)0;
    }
    // a task = 8 haplotypes x 64 loci: (BED_PITCH) x (BED_LOCI / 64) of them per tile; consecutive threads take
    // consecutive 64-locus chunks of the same rows (one warp reads 4 haplotype groups x 64 contiguous bytes per row)
    constexpr int CH = BED_LOCI / 64;
    for (int task = threadIdx.x; task < BED_PITCH * CH; task += BED_THREADS) {
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
      // the genotype of sample s at a locus is b0 & b1 | (b0 | b1) << 1 (:20-28): formed on whole words first, so that the
      // eight "rows" that go into the transpose are (and, or) of the four haplotype pairs and a transposed byte IS the
      // .bed byte of 4 samples (bits 2s, 2s+1)
#pragma unroll
      for (int sidx = 0; sidx < 4; sidx++) {
        const uint32_t a0 = lo[2 * sidx], a1 = lo[2 * sidx + 1], c0 = hi[2 * sidx], c1 = hi[2 * sidx + 1];
        lo[2 * sidx] = a0 & a1; lo[2 * sidx + 1] = a0 | a1;
        hi[2 * sidx] = c0 & c1; hi[2 * sidx + 1] = c0 | c1;
      }
      const int wcol = 4 * rot(gq >> 2, cq) + (gq & 3);
#pragma unroll
      for (int b = 0; b < 8; b++) {
        // byte b of every row -> one 64-bit word (byte r = row r, bit c = locus 8b + c)
        const uint32_t *w = (b < 4) ? lo : hi;
        const unsigned sel = (unsigned)(b & 3) | ((4u + (unsigned)(b & 3)) << 4);
        const uint32_t t01 = __byte_perm(w[0], w[1], sel), t23 = __byte_perm(w[2], w[3], sel);
        const uint32_t t45 = __byte_perm(w[4], w[5], sel), t67 = __byte_perm(w[6], w[7], sel);
        uint32_t yl = __byte_perm(t01, t23, 0x5410), yh = __byte_perm(t45, t67, 0x5410);
        transpose8x8(yl, yh);   // byte c = the .bed byte of the 4 samples at locus 8b + c
        uint8_t *trow = tile + (size_t)(64 * cq + 8 * b) * BED_PITCH + wcol;
#pragma unroll
        for (int c = 0; c < 4; c++) {
          trow[(size_t)c * BED_PITCH] = (uint8_t)(yl >> (8 * c));
          trow[(size_t)(c + 4) * BED_PITCH] = (uint8_t)(yh >> (8 * c));
        }
      }
    }
    __syncthreads();
    // Output.  A column is ceil(N/8) bytes long, so a tile's 128 bytes of a column start at any alignment.  The tile writes
    // the aligned 16-byte chunks that START inside its 128 bytes -- the last one runs up to 15 bytes into the next tile's
    // range, which is why the tile was built 128 haplotypes wider -- so interior tiles store nothing but whole chunks:
    // eight lanes per column, one chunk per lane (four funnel shifts over five words of the tile row), four columns per
    // warp pass.  Only the first tile of a column stores the bytes before the first boundary and only the last one those
    // after the last whole chunk, singly.
    const long long byte0 = h0 / 8;            // first output byte of this tile within the column
    const int n_own = (int)((nb - byte0) < (BED_HAPS / 8) ? (nb - byte0) : (BED_HAPS / 8));   // bytes of the column this tile owns
    const int n_have = (int)((nb - byte0) < BED_PITCH ? (nb - byte0) : BED_PITCH);            // ... and holds
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = lane & 7;
    if (n_own > 0) {
      for (int l = wid * 4 + (lane >> 3); l < BED_LOCI; l += (BED_THREADS / 32) * 4) {
        const long long j = tl * BED_LOCI + l;
        if (j >= p) continue;
        const long long col = both ? 2 * j + (which ^ (int)swp[l]) : j;
        uint8_t *dst = out + (size_t)col * (size_t)nb + (size_t)byte0;
        const uint32_t *srcw = reinterpret_cast<const uint32_t *>(tile + (size_t)l * BED_PITCH);
        const uint8_t *src = tile + (size_t)l * BED_PITCH;
        const int cql = (l >> 6) & 7;
        const int head = (int)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);   // bytes before the first boundary
        // whole chunks that start inside the owned bytes and end inside the column
        const int c1 = n_own > head ? (n_own - head + 15) >> 4 : 0, c2 = n_have >= head ? (n_have - head) >> 4 : 0;
        const int kk = min(8, min(c1, c2));
        const int o = head + 16 * g;            // first tile byte of this lane's chunk
        if (g < kk) {
          const int q = o >> 2, sh = 8 * (o & 3);
          uint32_t wv[5];
#pragma unroll
          for (int tt = 0; tt < 5; tt++) wv[tt] = srcw[rot(min(q + tt, BED_PW - 1), cql)];
          uint4 v;
          v.x = __funnelshift_r(wv[0], wv[1], sh);
          v.y = __funnelshift_r(wv[1], wv[2], sh);
          v.z = __funnelshift_r(wv[2], wv[3], sh);
          v.w = __funnelshift_r(wv[3], wv[4], sh);
          *reinterpret_cast<uint4 *>(dst + o) = v;
        }
        if (th == 0) {   // the column starts here: nothing before it writes the bytes up to the first boundary
          for (int i = g; i < min(head, n_have); i += 8) dst[i] = src[4 * rot(i >> 2, cql) + (i & 3)];
        }
        const int next = head + 16 * kk;
        if (next < n_own) {   // the column ends before the chunk that would cover these bytes does: the rest of it, singly
          for (int i = next + g; i < n_have; i += 8) dst[i] = src[4 * rot(i >> 2, cql) + (i & 3)];
        }
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
