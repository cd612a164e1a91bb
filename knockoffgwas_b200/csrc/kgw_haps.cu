// knockoffgwas_b200/csrc/kgw_haps.cu -- `.haps` text -> packed haplotype matrix on the device
// (the text half of the input path).
//
// Replaces the data loop of HapsReader::read (snpknock2/src/haps_reader.cpp:93-118): for every file row whose
// legend id was requested the reference std::find()s the id among the requested ones, tokenizes the row
// (sutils::tokenize, utils.cpp:234-246: tokens are maximal runs of non-blank characters, blank = ' ' or '\t'; a
// trailing '\r' is stripped from the LAST token only) into std::strings and sets
//     haps[i][j] = (tokens[hap_idx[i]] == "1"),   hap_idx[2s] = 2 * file sample, hap_idx[2s+1] = 2 * file sample + 1
// in a vector<vector<bool>> that is transposed into chaplotype rows afterwards (:120-124).
// Here the text goes to HBM as it is.  Pass 1 (one CTA per requested row) finds the token starts, numbers them
// with a block-wide prefix sum and leaves one bit per token (token == "1"); pass 2 gathers the requested token
// columns and transposes them into the [2 n_samples][stride] bit matrix (lane = haplotype, 128 loci per warp, one
// 16-byte store per lane).  HBM-bound byte work: ~2 bytes read per (file haplotype, SNP), 1 bit written per cell.
#include "kgw_common.h"
#include "../../include/kgw.h"

#include <cub/block/block_scan.cuh>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

int kgw_fail_(int code, const std::string &msg);   // kgw_api.cu

namespace {

constexpr int TOK_THREADS = 256;
constexpr int TOK_ITEMS = 16;   // bytes per thread and tile

__device__ __forceinline__ bool is_blank(unsigned c) { return c == ' ' || c == '\t'; }

// Pass 1.  rowbits[j][t >> 5] bit (t & 31) = (token t of the row of column j is "1"); ntok[j] = tokens in the row.
__global__ void __launch_bounds__(TOK_THREADS) haps_tokenize_kernel(const uint8_t *__restrict__ text, const long long *__restrict__ row_off,
                                                                    const int *__restrict__ row_len, int tok_words,
                                                                    uint32_t *__restrict__ rowbits, int *__restrict__ ntok) {
  typedef cub::BlockScan<int, TOK_THREADS> Scan;
  __shared__ typename Scan::TempStorage tmp;
  const int j = blockIdx.x;
  const long long abs0 = row_off[j];
  const uint8_t *line = text + abs0;
  const int len = row_len[j];
  uint32_t *bits = rowbits + (size_t)j * tok_words;
  int running = 0;
  // tiles are aligned to 16 bytes of the text buffer (cudaMalloc'ed, so 16-byte aligned itself): one 16-byte load per
  // thread and tile plus the single bytes either side; positions below are relative to the start of the row
  const int lead = (int)(abs0 & 15);   // bytes of the first chunk that belong to the previous row
  for (int base = -lead; base < len; base += TOK_THREADS * TOK_ITEMS) {
    const int b0 = base + threadIdx.x * TOK_ITEMS;
    unsigned c[TOK_ITEMS + 2];   // c[0] = byte before the thread's first, c[TOK_ITEMS + 1] = byte after its last
    if (b0 < len && b0 + TOK_ITEMS > 0) {
      const uint4 v = *reinterpret_cast<const uint4 *>(line + b0);
      const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < TOK_ITEMS; q++) {
        const int b = b0 + q;
        c[q + 1] = (b >= 0 && b < len) ? ((w[q >> 2] >> (8 * (q & 3))) & 0xffu) : ' ';
      }
      c[0] = (b0 - 1 >= 0 && b0 - 1 < len) ? line[b0 - 1] : ' ';
      c[TOK_ITEMS + 1] = (b0 + TOK_ITEMS < len) ? line[b0 + TOK_ITEMS] : ' ';
    } else {
#pragma unroll
      for (int q = 0; q < TOK_ITEMS + 2; q++) c[q] = ' ';
    }
    int starts = 0;
#pragma unroll
    for (int q = 1; q <= TOK_ITEMS; q++) starts += (!is_blank(c[q]) && is_blank(c[q - 1])) ? 1 : 0;
    int before = 0, total = 0;
    Scan(tmp).ExclusiveSum(starts, before, total);
    __syncthreads();
    int t = running + before;
#pragma unroll
    for (int q = 1; q <= TOK_ITEMS; q++) {
      const int b = b0 + q - 1;
      if (!is_blank(c[q]) && is_blank(c[q - 1])) {
        bool one = false;
        if (c[q] == '1') {
          // the token ends after this character, or it is "1\r" and nothing but blanks follows (last token of the row)
          const unsigned nx = c[q + 1];
          if (b + 1 >= len || is_blank(nx)) {
            one = true;
          } else if (nx == '\r') {
            one = true;
            for (int x = b + 2; x < len; x++) if (!is_blank(line[x])) { one = false; break; }
          }
        }
        if (one) atomicOr(bits + (t >> 5), 1u << (t & 31));
        t++;
      }
    }
    running += total;
  }
  if (threadIdx.x == 0) ntok[j] = running;
}

// Pass 2.  One warp: 32 output haplotypes x 128 output loci.
__global__ void __launch_bounds__(256) haps_gather_kernel(const uint32_t *__restrict__ rowbits, const int *__restrict__ ntok, int tok_words,
                                                          const int32_t *__restrict__ sample_idx, int n_haps, int p, int stride_w,
                                                          uint32_t *__restrict__ Hw, int *__restrict__ status) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int hap = (blockIdx.y * 8 + warp) * 32 + lane;
  const int w0 = blockIdx.x * 4;
  if (hap >= n_haps) return;
  const int t = 2 * sample_idx[hap >> 1] + (hap & 1);   // token column of this haplotype (haps_reader.cpp:79-81)
  uint32_t words[4] = {0u, 0u, 0u, 0u};
  bool bad = false;
#pragma unroll
  for (int ww = 0; ww < 4; ww++) {
    uint32_t acc = 0;
#pragma unroll 8
    for (int jj = 0; jj < 32; jj++) {
      const int j = ((w0 + ww) << 5) + jj;
      if (j >= p) break;
      if (t >= ntok[j]) { bad = true; continue; }
      acc |= ((rowbits[(size_t)j * tok_words + (t >> 5)] >> (t & 31)) & 1u) << jj;
    }
    words[ww] = acc;
  }
  uint32_t *dst = Hw + (size_t)hap * stride_w + w0;
  if (w0 + 4 <= stride_w) {
    *reinterpret_cast<uint4 *>(dst) = make_uint4(words[0], words[1], words[2], words[3]);
  } else {
    for (int ww = 0; ww < 4 && w0 + ww < stride_w; ww++) dst[ww] = words[ww];
  }
  if (bad) atomicOr(status, 1);
}

}  // namespace

extern "C" {

int32_t kgw_haps_read(const uint8_t *text, int64_t text_bytes, int32_t n_samples, const int32_t *sample_idx, int32_t p,
                      const int32_t *row_of_col, int32_t device, int64_t row_bytes, uint8_t *H_out, float *kernel_ms) {
  if (!text || !sample_idx || !row_of_col || !H_out) return kgw_fail_(KGW_ERR_INVALID, "null argument");
  if (n_samples < 1 || p < 1 || text_bytes < 1) return kgw_fail_(KGW_ERR_INVALID, "n_samples, p and the text length must be positive");
  if (row_bytes < (p + 7) / 8) return kgw_fail_(KGW_ERR_INVALID, "row_bytes < ceil(p/8)");
  for (int i = 0; i < n_samples; i++)
    if (sample_idx[i] < 0 || sample_idx[i] > (1 << 29)) return kgw_fail_(KGW_ERR_INVALID, "sample index out of range");
  // rows as std::getline(fd, buffer, '\n') yields them (haps_reader.cpp:95)
  std::vector<long long> line_off;
  std::vector<int> line_len;
  for (long long pos = 0; pos < text_bytes;) {
    const void *nl = memchr(text + pos, '\n', (size_t)(text_bytes - pos));
    const long long end = nl ? (const uint8_t *)nl - text : text_bytes;
    if (end - pos > (1ll << 30)) return kgw_fail_(KGW_ERR_UNSUPPORTED, "a .haps row longer than 1 GiB");
    line_off.push_back(pos);
    line_len.push_back((int)(end - pos));
    pos = end + 1;
  }
  std::vector<long long> col_off(p);
  std::vector<int> col_len(p);
  int max_len = 0;
  for (int j = 0; j < p; j++) {
    const int r = row_of_col[j];
    if (r < 0 || (size_t)r >= line_off.size()) return kgw_fail_(KGW_ERR_INVALID, "row index outside the .haps file");
    col_off[j] = line_off[r];
    col_len[j] = line_len[r];
    max_len = std::max(max_len, line_len[r]);
  }
  const int tok_words = std::max(1, ((max_len + 1) / 2 + 31) / 32);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return kgw_fail_(KGW_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  uint8_t *dtext = nullptr;
  long long *doff = nullptr;
  int *dlen = nullptr, *dntok = nullptr, *dstatus = nullptr;
  int32_t *dsidx = nullptr;
  uint32_t *dbits = nullptr, *dH = nullptr;
#define KGW_CUDA_H(call)                                                                                          \
  do {                                                                                                            \
    cudaError_t e_ = (call);                                                                                      \
    if (e_ != cudaSuccess) {                                                                                      \
      cudaFree(dtext); cudaFree(doff); cudaFree(dlen); cudaFree(dntok); cudaFree(dstatus); cudaFree(dsidx); cudaFree(dbits); cudaFree(dH); \
      return kgw_fail_(KGW_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                         \
    }                                                                                                             \
  } while (0)
  const int n_haps = 2 * n_samples;
  const int stride_w = (((p + 31) / 32) + 3) & ~3;
  KGW_CUDA_H(cudaSetDevice(device));
  KGW_CUDA_H(kgw_malloc(&dtext, (size_t)text_bytes + 32));   // the tokenizer reads whole 16-byte chunks
  KGW_CUDA_H(kgw_malloc(&doff, sizeof(long long) * p));
  KGW_CUDA_H(kgw_malloc(&dlen, sizeof(int) * p));
  KGW_CUDA_H(kgw_malloc(&dntok, sizeof(int) * p));
  KGW_CUDA_H(kgw_malloc(&dstatus, sizeof(int)));
  KGW_CUDA_H(kgw_malloc(&dsidx, sizeof(int32_t) * n_samples));
  KGW_CUDA_H(kgw_malloc(&dbits, sizeof(uint32_t) * (size_t)p * tok_words));
  KGW_CUDA_H(kgw_malloc(&dH, sizeof(uint32_t) * (size_t)n_haps * stride_w));
  KGW_CUDA_H(cudaMemcpy(dtext, text, (size_t)text_bytes, cudaMemcpyHostToDevice));
  KGW_CUDA_H(cudaMemcpy(doff, col_off.data(), sizeof(long long) * p, cudaMemcpyHostToDevice));
  KGW_CUDA_H(cudaMemcpy(dlen, col_len.data(), sizeof(int) * p, cudaMemcpyHostToDevice));
  KGW_CUDA_H(cudaMemcpy(dsidx, sample_idx, sizeof(int32_t) * n_samples, cudaMemcpyHostToDevice));
  KGW_CUDA_H(cudaMemset(dbits, 0, sizeof(uint32_t) * (size_t)p * tok_words));
  KGW_CUDA_H(cudaMemset(dstatus, 0, sizeof(int)));
  cudaEvent_t e0, e1;
  KGW_CUDA_H(cudaEventCreate(&e0));
  KGW_CUDA_H(cudaEventCreate(&e1));
  KGW_CUDA_H(cudaEventRecord(e0));
  haps_tokenize_kernel<<<p, TOK_THREADS>>>(dtext, doff, dlen, tok_words, dbits, dntok);
  const dim3 grid((unsigned)((stride_w + 3) / 4), (unsigned)((n_haps + 255) / 256));
  haps_gather_kernel<<<grid, 256>>>(dbits, dntok, tok_words, dsidx, n_haps, p, stride_w, dH, dstatus);
  KGW_CUDA_H(cudaGetLastError());
  KGW_CUDA_H(cudaEventRecord(e1));
  KGW_CUDA_H(cudaEventSynchronize(e1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (kernel_ms) *kernel_ms = ms;
  int st = 0;
  KGW_CUDA_H(cudaMemcpy(&st, dstatus, sizeof(int), cudaMemcpyDeviceToHost));
  if (st == 0)
    KGW_CUDA_H(cudaMemcpy2D(H_out, (size_t)row_bytes, dH, sizeof(uint32_t) * (size_t)stride_w, (size_t)((p + 7) / 8), (size_t)n_haps,
                            cudaMemcpyDeviceToHost));
  cudaFree(dtext); cudaFree(doff); cudaFree(dlen); cudaFree(dntok); cudaFree(dstatus); cudaFree(dsidx); cudaFree(dbits); cudaFree(dH);
#undef KGW_CUDA_H
  if (st) return kgw_fail_(KGW_ERR_INVALID, "a requested .haps row has fewer tokens than the requested haplotype columns (haps_reader.cpp:109 would index past the end of `tokens`)");
  return KGW_OK;
}

}  // extern "C"
