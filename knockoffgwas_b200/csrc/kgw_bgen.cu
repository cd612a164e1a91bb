// knockoffgwas_b200/csrc/kgw_bgen.cu -- phased BGEN -> packed haplotype matrix on the device
// (the step before the path, producer of Haplotypes::H).
//
// Replaces the data loop of BgenReader::read_worker (snpknock2/src/bgen_reader.cpp:139-261): the reference
// walks the file variant by variant, std::find()s the rsid among the requested ones (O(p^2) string compares),
// expands the probability block into vector<vector<double>> (one double per haplotype and allele), converts
// with prob_to_hap (:48-63) and sets single bits of the haplotype-major chaplotype rows (:207-210).
// Here the host only indexes the variant blocks and inflates the selected ones (zstd / zlib, host threads);
// the uncompressed BGEN v1.2 "layout 2" probability blocks go to HBM as they are and one kernel decodes
// and transposes them into the [2 n_samples][stride] bit matrix the sampler reads:
//   block = N u32 | alleles u16 | min ploidy u8 | max ploidy u8 | N x (missing << 7 | ploidy) u8 |
//           phased u8 | B u8 | 2N values of B bits (P(first allele) of haplotype 2s, 2s+1), little-endian bit stream
//   prob_to_hap: P(first allele) == 1 -> bit 0;  == 0 -> bit 1;  anything else (or a missing sample) is
//   "Incorrect input" in the reference (a throw that nothing catches) and an error status here.
// HBM-bound byte work: per variant ~ (2 B/8 + 1) N bytes read, 2N/8 bytes written.
#include "kgw_common.h"
#include "../../include/kgw.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <climits>
#include <string>
#include <thread>
#include <vector>

int kgw_fail_(int code, const std::string &msg);   // kgw_api.cu

namespace {

enum { BGEN_BAD_HEADER = 1, BGEN_BAD_PLOIDY = 2, BGEN_BAD_PROB = 4, BGEN_NOT_8BIT = 8 /* not an error: selects the general kernel */ };

// Pass 1, one thread per output column: checks the block header and leaves (offset of the probability stream |
// B << 56) for pass 2.
__global__ void __launch_bounds__(256) bgen_header_kernel(const uint8_t *__restrict__ blocks, const long long *__restrict__ col_off,
                                                          const long long *__restrict__ col_len, int n_samples_bgen, int p,
                                                          unsigned long long *__restrict__ colinfo, int *__restrict__ status) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p) return;
  const uint8_t *blk = blocks + col_off[j];
  const uint32_t nfile = (uint32_t)blk[0] | ((uint32_t)blk[1] << 8) | ((uint32_t)blk[2] << 16) | ((uint32_t)blk[3] << 24);
  const uint32_t nall = (uint32_t)blk[4] | ((uint32_t)blk[5] << 8);
  const uint32_t pmin = blk[6], pmax = blk[7];
  uint32_t phased = 0, B = 0;
  if (nfile == (uint32_t)n_samples_bgen) { phased = blk[8 + n_samples_bgen]; B = blk[9 + n_samples_bgen]; }
  // the block must hold its header, the ploidy bytes and 2 N values of B bits
  const long long need = 10ll + n_samples_bgen + ((2ll * n_samples_bgen * (long long)B + 7) >> 3);
  if (nfile != (uint32_t)n_samples_bgen || nall != 2u || pmin != 2u || pmax != 2u || phased != 1u || B < 1u || B > 32u ||
      col_len[j] < need) {
    atomicOr(status, BGEN_BAD_HEADER);
    B = 0;
  }
  if (B != 8u) atomicOr(status, BGEN_NOT_8BIT);
  colinfo[j] = (unsigned long long)(col_off[j] + 10 + n_samples_bgen) | ((unsigned long long)B << 56);
}

// Pass 2.  One warp: 32 output haplotypes (lane = haplotype) x 128 output loci (4 words, one 16-byte store per
// lane); the 128 column descriptors of the CTA sit in shared memory.  Consecutive lanes read consecutive
// haplotype values of the same variant block (coalesced when the kept samples are consecutive in the file, the
// usual case); per (haplotype, variant) the kernel loads the sample's ploidy byte and its B-bit value.
__global__ void __launch_bounds__(256) bgen_unpack_kernel(const uint8_t *__restrict__ blocks, const unsigned long long *__restrict__ colinfo,
                                                          const int32_t *__restrict__ sample_idx, int n_samples_bgen, int n_haps,
                                                          int p, int stride_w, uint32_t *__restrict__ Hw, int *__restrict__ status) {
  __shared__ unsigned long long info[128];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int hap = (blockIdx.y * 8 + warp) * 32 + lane;
  const int w0 = blockIdx.x * 4;   // first 32-locus word of this CTA's 128 loci
  if (threadIdx.x < 128) {
    const int j = (w0 << 5) + threadIdx.x;
    info[threadIdx.x] = (j < p) ? colinfo[j] : 0ull;
  }
  __syncthreads();
  const bool live = hap < n_haps;
  const int s_file = live ? sample_idx[hap >> 1] : 0;
  const long long q = 2ll * s_file + (hap & 1);   // index of the haplotype's value in a block
  int bad = 0;
  uint32_t words[4] = {0u, 0u, 0u, 0u};
  if (live) {
#pragma unroll
    for (int ww = 0; ww < 4; ww++) {
      uint32_t acc = 0;
#pragma unroll 8
      for (int jj = 0; jj < 32; jj++) {
        const unsigned long long ci = info[ww * 32 + jj];
        const uint32_t B = (uint32_t)(ci >> 56);
        if (B == 0u) continue;   // beyond p, or a header pass 1 rejected
        const uint8_t *pr = blocks + (ci & 0x00ffffffffffffffull);
        if (pr[s_file - n_samples_bgen - 2] != 2u) bad |= BGEN_BAD_PLOIDY;   // ploidy bytes end 2 bytes before the stream
        uint32_t v;
        if (B == 8u) {
          v = pr[q];
        } else {
          const long long bit0 = q * (long long)B;
          const uint8_t *b0 = pr + (bit0 >> 3);
          const int sh = (int)(bit0 & 7);
          const int nbytes = (sh + (int)B + 7) >> 3;   // <= 5
          unsigned long long acc64 = 0;
          for (int t = 0; t < nbytes; t++) acc64 |= (unsigned long long)b0[t] << (8 * t);
          v = (uint32_t)((acc64 >> sh) & ((B == 32u) ? 0xffffffffull : ((1ull << B) - 1ull)));
        }
        const uint32_t ones = (B == 32u) ? 0xffffffffu : ((1u << B) - 1u);
        if (v == 0u) acc |= 1u << jj;            // P(first allele) = 0 -> second allele -> bit 1
        else if (v != ones) bad |= BGEN_BAD_PROB;
      }
      words[ww] = acc;
    }
    uint32_t *dst = Hw + (size_t)hap * stride_w + w0;
    if (w0 + 4 <= stride_w) {
      *reinterpret_cast<uint4 *>(dst) = make_uint4(words[0], words[1], words[2], words[3]);
    } else {
      for (int ww = 0; ww < 4 && w0 + ww < stride_w; ww++) dst[ww] = words[ww];
    }
  }
  if (bad) atomicOr(status, bad);
}

// Pass 2, fast path: 8-bit probabilities (what UK Biobank and qctool write) and kept samples that are consecutive in the
// file (sample_idx[i] = s0 + i).  A CTA owns 256 consecutive haplotypes x 128 loci; for 64 variants at a time it
// copies the 256 value bytes and the 128 ploidy bytes of its haplotypes to shared memory with 16-byte loads
// (aligned down; the slack bytes are skipped when reading), then every thread packs the bits of its haplotype.
constexpr int BG_TILE = 64;          // variants staged at a time
constexpr int BG_VAL_CHUNKS = 17;    // 16-byte chunks covering 256 bytes at any alignment
constexpr int BG_PL_CHUNKS = 9;      // ... 128 bytes
__global__ void __launch_bounds__(256) bgen_unpack8_kernel(const uint8_t *__restrict__ blocks, const unsigned long long *__restrict__ colinfo,
                                                           int s0, int n_samples_bgen, int n_haps, int p, int stride_w,
                                                           uint32_t *__restrict__ Hw, int *__restrict__ status) {
  __shared__ uint4 sval[BG_TILE][BG_VAL_CHUNKS];
  __shared__ uint4 spl[BG_TILE][BG_PL_CHUNKS];
  __shared__ unsigned long long info[BG_TILE];
  const int tid = threadIdx.x;
  const int hap0 = blockIdx.y * 256;           // first output haplotype of the CTA (even)
  const int hap = hap0 + tid;
  const int w0 = blockIdx.x * 4;
  const long long q0 = 2ll * s0 + hap0;        // its value index in a block; its sample is s0 + hap0 / 2
  const long long pl0 = (long long)s0 + hap0 / 2 - n_samples_bgen - 2;   // offset of that sample's ploidy byte from the stream
  int bad = 0;
  uint32_t words[4] = {0u, 0u, 0u, 0u};
  for (int half = 0; half < 2; half++) {
    const int jbase = (w0 + 2 * half) << 5;
    __syncthreads();
    if (tid < BG_TILE) info[tid] = (jbase + tid < p) ? colinfo[jbase + tid] : 0ull;
    __syncthreads();
    for (int x = tid; x < BG_TILE * BG_VAL_CHUNKS; x += 256) {
      const int v = x / BG_VAL_CHUNKS, c = x - v * BG_VAL_CHUNKS;
      const unsigned long long ci = info[v];
      if ((ci >> 56) != 0ull) {
        const unsigned long long a = ((ci & 0x00ffffffffffffffull) + (unsigned long long)q0) & ~15ull;
        sval[v][c] = *reinterpret_cast<const uint4 *>(blocks + a + 16ull * c);
      }
    }
    for (int x = tid; x < BG_TILE * BG_PL_CHUNKS; x += 256) {
      const int v = x / BG_PL_CHUNKS, c = x - v * BG_PL_CHUNKS;
      const unsigned long long ci = info[v];
      if ((ci >> 56) != 0ull) {
        const unsigned long long a = (unsigned long long)((long long)(ci & 0x00ffffffffffffffull) + pl0) & ~15ull;
        spl[v][c] = *reinterpret_cast<const uint4 *>(blocks + a + 16ull * c);
      }
    }
    __syncthreads();
    if (hap < n_haps) {
#pragma unroll
      for (int ww = 0; ww < 2; ww++) {
        uint32_t acc = 0;
#pragma unroll 8
        for (int jj = 0; jj < 32; jj++) {
          const int v = ww * 32 + jj;
          const unsigned long long ci = info[v];
          if ((ci >> 56) == 0ull) continue;
          const unsigned long long st = ci & 0x00ffffffffffffffull;
          const unsigned val = reinterpret_cast<const uint8_t *>(&sval[v][0])[((st + (unsigned long long)q0) & 15ull) + tid];
          const unsigned pl = reinterpret_cast<const uint8_t *>(&spl[v][0])[((unsigned long long)((long long)st + pl0) & 15ull) + (tid >> 1)];
          if (pl != 2u) bad |= BGEN_BAD_PLOIDY;
          if (val == 0u) acc |= 1u << jj;
          else if (val != 255u) bad |= BGEN_BAD_PROB;
        }
        words[2 * half + ww] = acc;
      }
    }
  }
  if (hap < n_haps) {
    uint32_t *dst = Hw + (size_t)hap * stride_w + w0;
    if (w0 + 4 <= stride_w) {
      *reinterpret_cast<uint4 *>(dst) = make_uint4(words[0], words[1], words[2], words[3]);
    } else {
      for (int ww = 0; ww < 4 && w0 + ww < stride_w; ww++) dst[ww] = words[ww];
    }
  }
  if (bad) atomicOr(status, bad);
}

// ---- host side: container walk and block decompression -------------------------------------------------
typedef size_t (*zstd_decompress_fn)(void *, size_t, const void *, size_t);
typedef unsigned (*zstd_iserror_fn)(size_t);
typedef int (*zlib_uncompress_fn)(unsigned char *, unsigned long *, const unsigned char *, unsigned long);

// ---- zlib-compressed probability blocks on the device.  Every variant's block is its own zlib stream (RFC 1950 around
// RFC 1951 deflate; what the UK Biobank's BGEN v1.2 files use): the streams are independent, so one thread inflates one
// block -- canonical Huffman decoding one bit at a time from the code-length counts (no look-up tables to build: the
// per-thread state is 1.4 KB of local memory), stored / fixed / dynamic blocks, LZ77 copies within the thread's own
// output.  Slow per stream, but a chromosome has tens of thousands of them and the compressed file is what crosses
// PCIe.  The Adler-32 of the output is checked as zlib's uncompress() does.
struct InflateState {
  const uint8_t *in;
  long long inlen, incnt;
  uint8_t *out;
  long long outlen, outcnt;
  unsigned long long bitbuf;   // whole bytes of the input, not yet consumed bits from bit 0 up
  int bitcnt;
  int err;
};
struct Huff { short count[16]; short *symbol; };
struct HuffCounts { int c[16]; };   // the counts in registers while symbols are decoded (every index is static)

__device__ __forceinline__ void inf_refill(InflateState &s) {
  while (s.bitcnt <= 56 && s.incnt < s.inlen) {
    s.bitbuf |= (unsigned long long)s.in[s.incnt++] << s.bitcnt;
    s.bitcnt += 8;
  }
}
// gives the whole bytes still in the bit buffer back to the input (stored blocks and the trailer start on a byte boundary)
__device__ __forceinline__ void inf_byte_align(InflateState &s) {
  s.incnt -= s.bitcnt >> 3;
  s.bitbuf = 0;
  s.bitcnt = 0;
}
__device__ __forceinline__ int inf_bits(InflateState &s, int need) {   // need <= 13
  if (s.bitcnt < need) {
    inf_refill(s);
    if (s.bitcnt < need) { s.err = 1; return 0; }
  }
  const int v = (int)(s.bitbuf & ((1ull << need) - 1ull));
  s.bitbuf >>= need;
  s.bitcnt -= need;
  return v;
}
__device__ __forceinline__ void inf_counts(HuffCounts &hc, const Huff &h) {
#pragma unroll
  for (int len = 0; len < 16; len++) hc.c[len] = h.count[len];
}
// canonical Huffman decoding, one code length after the other (count[len] codes of each length, in symbol order)
__device__ __forceinline__ int inf_decode(InflateState &s, const HuffCounts &hc, const short *symbol) {
  if (s.bitcnt < 15) inf_refill(s);
  unsigned bb = (unsigned)s.bitbuf;
  int code = 0, first = 0, index = 0;
#pragma unroll
  for (int len = 1; len <= 15; len++) {
    code |= (int)(bb & 1u);
    bb >>= 1;
    const int count = hc.c[len];
    if (code - count < first) {
      if (len > s.bitcnt) { s.err = 1; return -1; }   // ran into the end of the input
      s.bitbuf >>= len;
      s.bitcnt -= len;
      return symbol[index + (code - first)];
    }
    index += count;
    first += count;
    first <<= 1;
    code <<= 1;
  }
  s.err = 2;
  return -1;
}
// First decoding level: the next INF_LBITS (INF_DBITS) bits of the stream index a table of (symbol << 4 | code length)
// in shared memory; codes longer than that, invalid codes and the end of the input take the loop above.
constexpr int INF_LBITS = 9, INF_DBITS = 6;
__device__ __forceinline__ int inf_decode_fast(InflateState &s, const unsigned short *tab, int bits, const HuffCounts &hc, const short *symbol) {
  if (s.bitcnt < 15) inf_refill(s);
  const unsigned e = tab[(unsigned)s.bitbuf & ((1u << bits) - 1u)];
  const int l = (int)(e & 15u);
  if (l != 0 && l <= s.bitcnt) {
    s.bitbuf >>= l;
    s.bitcnt -= l;
    return (int)(e >> 4);
  }
  return inf_decode(s, hc, symbol);
}
// fills the (zeroed) table from the code lengths: canonical codes in symbol order within a length (RFC 1951 3.2.2),
// stored bit-reversed because the stream carries Huffman codes most significant bit first
__device__ void inf_fast_table(unsigned short *tab, int bits, const Huff &h, const short *length, int n) {
  int next[16];
  int code = 0;
  next[0] = 0;
  for (int len = 1; len <= 15; len++) {
    code = (code + (len > 1 ? h.count[len - 1] : 0)) << 1;
    next[len] = code;
  }
  for (int sym = 0; sym < n; sym++) {
    const int l = length[sym];
    if (l == 0) continue;
    const unsigned c = (unsigned)next[l]++;
    if (l > bits) continue;
    const unsigned r = __brev(c) >> (32 - l);
    for (unsigned idx = r; idx < (1u << bits); idx += 1u << l) tab[idx] = (unsigned short)((sym << 4) | l);
  }
}
// canonical code from the code lengths; returns 0 for a complete code, > 0 for an incomplete one, < 0 for over-subscription
__device__ int inf_construct(Huff &h, const short *length, int n) {
  for (int len = 0; len <= 15; len++) h.count[len] = 0;
  for (int sym = 0; sym < n; sym++) h.count[length[sym]]++;
  if (h.count[0] == n) return 0;
  int left = 1;
  for (int len = 1; len <= 15; len++) {
    left <<= 1;
    left -= h.count[len];
    if (left < 0) return left;
  }
  short offs[16];
  offs[1] = 0;
  for (int len = 1; len < 15; len++) offs[len + 1] = offs[len] + h.count[len];
  for (int sym = 0; sym < n; sym++)
    if (length[sym] != 0) h.symbol[offs[length[sym]]++] = (short)sym;
  return left;
}
// One warp per block of one variant: lane 0 walks the bit stream (Huffman decoding is serial), every lane takes part
// in the copies (matches, stored blocks) and in the Adler-32 of the result.  s is lane 0's; the others only use s.out.
__device__ __forceinline__ void inf_codes(InflateState &s, const Huff &lencode, const Huff &distcode, const unsigned short *ltab,
                                          const unsigned short *dtab, int lane) {
  HuffCounts lc, dc;
  inf_counts(lc, lencode);
  inf_counts(dc, distcode);
  const short lens[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  const short lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  const short dists[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  const short dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  for (;;) {
    int len = 0, dist = 0, stop = 0;
    if (lane == 0) {
      for (;;) {   // literals until the next match, the end of the block or an error
        int symbol = inf_decode_fast(s, ltab, INF_LBITS, lc, lencode.symbol);
        if (s.err) break;
        if (symbol < 256) {
          if (s.outcnt == s.outlen) { s.err = 3; break; }
          s.out[s.outcnt++] = (uint8_t)symbol;
          continue;
        }
        if (symbol == 256) { stop = 1; break; }
        symbol -= 257;
        if (symbol >= 29) { s.err = 4; break; }
        len = lens[symbol] + inf_bits(s, lext[symbol]);
        symbol = inf_decode_fast(s, dtab, INF_DBITS, dc, distcode.symbol);
        if (s.err) break;
        if (symbol >= 30) { s.err = 4; break; }
        dist = dists[symbol] + inf_bits(s, dext[symbol]);
        if (s.err) break;
        if ((long long)dist > s.outcnt) { s.err = 5; break; }
        if (s.outcnt + len > s.outlen) { s.err = 3; break; }
        break;
      }
      if (s.err) stop = 1;
    }
    stop = __shfl_sync(0xffffffffu, stop, 0);
    if (stop) return;
    len = __shfl_sync(0xffffffffu, len, 0);
    dist = __shfl_sync(0xffffffffu, dist, 0);
    const long long o = __shfl_sync(0xffffffffu, s.outcnt, 0);
    __syncwarp();   // lane 0's literals are visible to the lanes that copy from them
    uint8_t *dst = s.out + o;
    const uint8_t *src = dst - dist;
    if (dist >= len) for (int i = lane; i < len; i += 32) dst[i] = src[i];
    else for (int i = lane; i < len; i += 32) dst[i] = src[i % dist];   // an overlapping match repeats its last dist bytes
    __syncwarp();
    if (lane == 0) s.outcnt = o + len;
  }
}
// code lengths of a dynamic block (RFC 1951 3.2.7) into the two decoding tables; s.err on a malformed header
__device__ void inf_dynamic_tables(InflateState &s, short *lengths, Huff &lencode, Huff &distcode, int &nlen_out, int &ndist_out) {
  const short order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  const int nlen = inf_bits(s, 5) + 257, ndist = inf_bits(s, 5) + 1, ncode = inf_bits(s, 4) + 4;
  if (s.err) return;
  if (nlen > 286 || ndist > 30) { s.err = 8; return; }
  nlen_out = nlen;
  ndist_out = ndist;
  int index = 0;
  for (; index < ncode; index++) lengths[order[index]] = (short)inf_bits(s, 3);
  for (; index < 19; index++) lengths[order[index]] = 0;
  if (s.err) return;
  if (inf_construct(lencode, lengths, 19) != 0) { s.err = 8; return; }
  HuffCounts cc;
  inf_counts(cc, lencode);
  index = 0;
  while (index < nlen + ndist) {
    int symbol = inf_decode(s, cc, lencode.symbol);
    if (s.err) return;
    if (symbol < 16) lengths[index++] = (short)symbol;
    else {
      int len = 0, rep;
      if (symbol == 16) {
        if (index == 0) { s.err = 8; return; }
        len = lengths[index - 1];
        rep = 3 + inf_bits(s, 2);
      } else if (symbol == 17) rep = 3 + inf_bits(s, 3);
      else rep = 11 + inf_bits(s, 7);
      if (s.err) return;
      if (index + rep > nlen + ndist) { s.err = 8; return; }
      while (rep--) lengths[index++] = (short)len;
    }
  }
  if (lengths[256] == 0) { s.err = 8; return; }
  const int e1 = inf_construct(lencode, lengths, nlen);
  if (e1 != 0 && (e1 < 0 || nlen != lencode.count[0] + lencode.count[1])) { s.err = 8; return; }   // incomplete: one 1-bit code only
  const int e2 = inf_construct(distcode, lengths + nlen, ndist);
  if (e2 != 0 && (e2 < 0 || ndist != distcode.count[0] + distcode.count[1])) { s.err = 8; return; }
}
constexpr int INF_WARPS = 4;
__global__ void __launch_bounds__(32 * INF_WARPS, 8) bgen_inflate_kernel(const uint8_t *__restrict__ comp, const long long *__restrict__ comp_off,
                                                                      const long long *__restrict__ comp_len, uint8_t *blocks,
                                                                      const long long *__restrict__ col_off, const long long *__restrict__ col_len, int p,
                                                                      int *__restrict__ status) {
  const int j = blockIdx.x * INF_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= p) return;
  InflateState s;
  s.in = comp + comp_off[j]; s.inlen = comp_len[j]; s.incnt = 0;
  s.out = blocks + col_off[j]; s.outlen = col_len[j]; s.outcnt = 0;
  s.bitbuf = 0; s.bitcnt = 0; s.err = 0;
  __shared__ unsigned short fast_tab[INF_WARPS][(1 << INF_LBITS) + (1 << INF_DBITS)];
  unsigned short *ltab = fast_tab[threadIdx.x >> 5], *dtab = ltab + (1 << INF_LBITS);
  short lengths[320], lensym[288], distsym[30];
  Huff lencode, distcode;
  lencode.symbol = lensym;
  distcode.symbol = distsym;
  // zlib header (RFC 1950): deflate, window <= 32 KB, no preset dictionary
  if (s.inlen < 6) s.err = 1;
  else {
    const unsigned cmf = s.in[0], flg = s.in[1];
    if ((cmf & 0x0f) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31 != 0 || (flg & 0x20)) s.err = 6;
    s.incnt = 2;
  }
  for (;;) {
    int last = 0, type = -1;
    if (lane == 0 && !s.err) {
      last = inf_bits(s, 1);
      type = inf_bits(s, 2);
      if (s.err) type = -1;
    }
    type = __shfl_sync(0xffffffffu, type, 0);
    last = __shfl_sync(0xffffffffu, last, 0);
    if (type < 0) break;
    if (type == 0) {   // stored
      unsigned len = 0xffffffffu;
      long long ip = 0, op = 0;
      if (lane == 0) {
        inf_byte_align(s);
        if (s.incnt + 4 > s.inlen) s.err = 1;
        else {
          const unsigned l = s.in[s.incnt] | (s.in[s.incnt + 1] << 8), nl = s.in[s.incnt + 2] | (s.in[s.incnt + 3] << 8);
          s.incnt += 4;
          if (l != (~nl & 0xffffu)) s.err = 7;
          else if (s.incnt + l > s.inlen) s.err = 1;
          else if (s.outcnt + l > s.outlen) s.err = 3;
          else { len = l; ip = s.incnt; op = s.outcnt; s.incnt += l; s.outcnt += l; }
        }
      }
      len = __shfl_sync(0xffffffffu, len, 0);
      if (len == 0xffffffffu) break;
      ip = __shfl_sync(0xffffffffu, ip, 0);
      op = __shfl_sync(0xffffffffu, op, 0);
      for (unsigned i = lane; i < len; i += 32) s.out[op + i] = s.in[ip + i];
      __syncwarp();
    } else if (type == 3) {
      if (lane == 0) s.err = 9;
      break;
    } else {
      for (int e = lane; e < (1 << INF_LBITS) + (1 << INF_DBITS); e += 32) ltab[e] = 0;
      __syncwarp();
      if (lane == 0) {
        if (type == 1) {   // fixed codes
          int sym = 0;
          for (; sym < 144; sym++) lengths[sym] = 8;
          for (; sym < 256; sym++) lengths[sym] = 9;
          for (; sym < 280; sym++) lengths[sym] = 7;
          for (; sym < 288; sym++) lengths[sym] = 8;
          inf_construct(lencode, lengths, 288);
          inf_fast_table(ltab, INF_LBITS, lencode, lengths, 288);
          for (sym = 0; sym < 30; sym++) lengths[sym] = 5;
          inf_construct(distcode, lengths, 30);
          inf_fast_table(dtab, INF_DBITS, distcode, lengths, 30);
        } else {
          int nlen = 0, ndist = 0;
          inf_dynamic_tables(s, lengths, lencode, distcode, nlen, ndist);
          if (!s.err) {
            inf_fast_table(ltab, INF_LBITS, lencode, lengths, nlen);
            inf_fast_table(dtab, INF_DBITS, distcode, lengths + nlen, ndist);
          }
        }
      }
      if (__shfl_sync(0xffffffffu, s.err, 0)) break;
      inf_codes(s, lencode, distcode, ltab, dtab, lane);
      if (__shfl_sync(0xffffffffu, s.err, 0)) break;
    }
    if (last) break;
  }
  if (lane == 0 && !s.err && s.outcnt != s.outlen) s.err = 10;
  if (__shfl_sync(0xffffffffu, s.err, 0) == 0) {
    // Adler-32 of the output against the stream's trailer (big endian): every lane sums one stretch of the block,
    // lane 0 chains the stretches (after a stretch of n bytes, a += sum d_i, b += n a + sum (n - i) d_i)
    __syncwarp();
    const long long L = s.outlen, chunk = (((L + 31) / 32) + 3) & ~3ll;
    const long long lo = min(L, lane * chunk), hi = min(L, lo + chunk);
    unsigned long long s1 = 0, s2 = 0;
    long long i = lo;
    for (; i + 4 <= hi; i += 4) {   // lo and the blocks are 4-byte aligned
      const unsigned w = *reinterpret_cast<const unsigned *>(s.out + i);
      const unsigned long long rem = (unsigned long long)(hi - i);
      const unsigned d0 = w & 255u, d1 = (w >> 8) & 255u, d2 = (w >> 16) & 255u, d3 = w >> 24;
      s1 += d0 + d1 + d2 + d3;
      s2 += rem * (d0 + d1 + d2 + d3) - (d1 + 2 * d2 + 3 * d3);
    }
    for (; i < hi; i++) { const unsigned d = s.out[i]; s1 += d; s2 += (unsigned long long)(hi - i) * d; }
    unsigned long long a = 1, b = 0;
    for (int l = 0; l < 32; l++) {
      const unsigned long long t1 = __shfl_sync(0xffffffffu, s1, l), t2 = __shfl_sync(0xffffffffu, s2, l);
      const unsigned long long n = (unsigned long long)__shfl_sync(0xffffffffu, hi - lo, l);
      b = (b + (n % 65521ull) * a + t2) % 65521ull;
      a = (a + t1) % 65521ull;
    }
    if (lane == 0) {   // the trailer follows the deflate data at the next byte boundary
      inf_byte_align(s);
      if (s.incnt + 4 > s.inlen) s.err = 1;
      else {
        const unsigned want = ((unsigned)s.in[s.incnt] << 24) | ((unsigned)s.in[s.incnt + 1] << 16) | ((unsigned)s.in[s.incnt + 2] << 8) | s.in[s.incnt + 3];
        if (want != (((unsigned)b << 16) | (unsigned)a)) s.err = 11;
      }
    }
  }
  if (lane == 0 && s.err) atomicOr(status, 1);
}

struct Codecs {
  zstd_decompress_fn zstd = nullptr;
  zstd_iserror_fn zstd_err = nullptr;
  zlib_uncompress_fn zlib = nullptr;
};

Codecs &codecs() {
  static Codecs c;
  static std::atomic<bool> done{false};
  if (!done.load()) {
    if (void *h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_GLOBAL)) {
      c.zstd = (zstd_decompress_fn)dlsym(h, "ZSTD_decompress");
      c.zstd_err = (zstd_iserror_fn)dlsym(h, "ZSTD_isError");
    }
    if (void *h = dlopen("libz.so.1", RTLD_NOW | RTLD_GLOBAL)) c.zlib = (zlib_uncompress_fn)dlsym(h, "uncompress");
    done.store(true);
  }
  return c;
}

inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

struct VariantBlock { int64_t data_off; int64_t comp_len; int64_t raw_len; };

// header + variant identifying data (BGEN v1.2 spec); fills the probability-block index in file order
int bgen_index(const uint8_t *f, int64_t nbytes, uint32_t *n_samples, uint32_t *compression, std::vector<VariantBlock> &out) {
  if (nbytes < 24) return kgw_fail_(KGW_ERR_INVALID, "BGEN image shorter than its header");
  const uint32_t offset = rd32(f), hlen = rd32(f + 4), M = rd32(f + 8), N = rd32(f + 12);
  if (hlen < 20 || 4ll + hlen > nbytes || 4ll + offset > nbytes) return kgw_fail_(KGW_ERR_INVALID, "BGEN header lengths outside the file");
  if (memcmp(f + 16, "bgen", 4) != 0 && rd32(f + 16) != 0u) return kgw_fail_(KGW_ERR_INVALID, "not a BGEN file (magic)");
  const uint32_t flags = rd32(f + 4 + hlen - 4);
  *compression = flags & 3u;
  const uint32_t layout = (flags >> 2) & 15u;
  if (layout != 2u) return kgw_fail_(KGW_ERR_UNSUPPORTED, "only BGEN layout 2 (v1.2) is built; the reference data and UK Biobank use it");
  if (*compression > 2u) return kgw_fail_(KGW_ERR_INVALID, "unknown BGEN compression flag");
  *n_samples = N;
  out.resize(M);
  int64_t pos = 4ll + offset;
  for (uint32_t v = 0; v < M; v++) {
    for (int t = 0; t < 3; t++) {   // variant id, rsid, chromosome
      if (pos + 2 > nbytes) return kgw_fail_(KGW_ERR_INVALID, "truncated BGEN variant block");
      pos += 2 + rd16(f + pos);
    }
    if (pos + 6 > nbytes) return kgw_fail_(KGW_ERR_INVALID, "truncated BGEN variant block");
    const uint16_t nall = rd16(f + pos + 4);
    pos += 6;
    for (uint16_t a = 0; a < nall; a++) {
      if (pos + 4 > nbytes) return kgw_fail_(KGW_ERR_INVALID, "truncated BGEN variant block");
      pos += 4 + (int64_t)rd32(f + pos);
    }
    if (pos + 4 > nbytes) return kgw_fail_(KGW_ERR_INVALID, "truncated BGEN variant block");
    const int64_t C = rd32(f + pos);
    pos += 4;
    if (pos + C > nbytes) return kgw_fail_(KGW_ERR_INVALID, "BGEN probability block runs past the end of the file");
    if (*compression == 0u) out[v] = {pos, C, C};
    else {
      if (C < 4) return kgw_fail_(KGW_ERR_INVALID, "BGEN compressed block shorter than its length field");
      out[v] = {pos + 4, C - 4, (int64_t)rd32(f + pos)};
    }
    pos += C;
  }
  return KGW_OK;
}

}  // namespace

extern "C" {

int32_t kgw_bgen_variant_count(const uint8_t *file, int64_t file_bytes, int32_t *n_variants, int32_t *n_samples) {
  if (!file || file_bytes < 24) return kgw_fail_(KGW_ERR_INVALID, "BGEN image shorter than its header");
  if (n_variants) *n_variants = (int32_t)rd32(file + 8);
  if (n_samples) *n_samples = (int32_t)rd32(file + 12);
  return KGW_OK;
}

int32_t kgw_bgen_rsids(const uint8_t *file, int64_t file_bytes, int32_t n_variants, int64_t *rsid_off, int32_t *rsid_len) {
  if (!file || !rsid_off || !rsid_len || file_bytes < 24) return kgw_fail_(KGW_ERR_INVALID, "null argument or BGEN image shorter than its header");
  const uint32_t offset = rd32(file), M = rd32(file + 8);
  if ((uint32_t)n_variants != M) return kgw_fail_(KGW_ERR_INVALID, "n_variants differs from the BGEN header");
  const uint32_t hlen = rd32(file + 4);
  if (hlen < 20 || 4ll + hlen > file_bytes || 4ll + offset > file_bytes) return kgw_fail_(KGW_ERR_INVALID, "BGEN header lengths outside the file");
  const uint32_t flags = rd32(file + 4 + hlen - 4);
  if (((flags >> 2) & 15u) != 2u) return kgw_fail_(KGW_ERR_UNSUPPORTED, "only BGEN layout 2 (v1.2) is built");
  int64_t pos = 4ll + offset;
  for (uint32_t v = 0; v < M; v++) {
    if (pos + 2 > file_bytes) return kgw_fail_(KGW_ERR_INVALID, "truncated BGEN variant block");
    pos += 2 + rd16(file + pos);                               // variant id
    if (pos + 2 > file_bytes) return kgw_fail_(KGW_ERR_INVALID, "truncated BGEN variant block");
    rsid_len[v] = rd16(file + pos);
    rsid_off[v] = pos + 2;
    pos += 2 + rsid_len[v];
    if (pos + 2 > file_bytes) return kgw_fail_(KGW_ERR_INVALID, "truncated BGEN variant block");
    pos += 2 + rd16(file + pos);                               // chromosome
    if (pos + 6 > file_bytes) return kgw_fail_(KGW_ERR_INVALID, "truncated BGEN variant block");
    const uint16_t nall = rd16(file + pos + 4);
    pos += 6;
    for (uint16_t a = 0; a < nall; a++) {
      if (pos + 4 > file_bytes) return kgw_fail_(KGW_ERR_INVALID, "truncated BGEN variant block");
      pos += 4 + (int64_t)rd32(file + pos);
    }
    if (pos + 4 > file_bytes) return kgw_fail_(KGW_ERR_INVALID, "truncated BGEN variant block");
    pos += 4 + (int64_t)rd32(file + pos);
    if (pos > file_bytes) return kgw_fail_(KGW_ERR_INVALID, "BGEN probability block runs past the end of the file");
  }
  return KGW_OK;
}

int32_t kgw_bgen_read(const uint8_t *file, int64_t file_bytes, int32_t n_samples, const int32_t *sample_idx, int32_t p,
                      const int32_t *variant_of_col, int32_t host_threads, int32_t device, int64_t row_bytes, uint8_t *H_out,
                      float *kernel_ms) {
  if (!file || !sample_idx || !variant_of_col || !H_out) return kgw_fail_(KGW_ERR_INVALID, "null argument");
  if (n_samples < 1 || p < 1) return kgw_fail_(KGW_ERR_INVALID, "n_samples and p must be positive");
  if (row_bytes < (p + 7) / 8) return kgw_fail_(KGW_ERR_INVALID, "row_bytes < ceil(p/8)");
  uint32_t nfile = 0, compression = 0;
  std::vector<VariantBlock> idx;
  int rc = bgen_index(file, file_bytes, &nfile, &compression, idx);
  if (rc != KGW_OK) return rc;
  for (int i = 0; i < n_samples; i++)
    if (sample_idx[i] < 0 || (uint32_t)sample_idx[i] >= nfile) return kgw_fail_(KGW_ERR_INVALID, "sample index outside the BGEN sample block");
  // Uncompressed files go to the device as they are (one copy of the span that holds the requested blocks);
  // compressed blocks are inflated into a staging buffer first, 16-byte aligned one after the other.
  std::vector<long long> col_off(p), col_len(p);
  long long total = 0, span_lo = file_bytes, span_hi = 0;
  for (int j = 0; j < p; j++) {
    const int v = variant_of_col[j];
    if (v < 0 || (size_t)v >= idx.size()) return kgw_fail_(KGW_ERR_INVALID, "variant index outside the BGEN file");
    if (idx[v].raw_len < 10 + (int64_t)nfile) return kgw_fail_(KGW_ERR_INVALID, "BGEN probability block shorter than its header");
    col_len[j] = idx[v].raw_len;
    if (compression == 0u) {
      span_lo = std::min<long long>(span_lo, idx[v].data_off);
      span_hi = std::max<long long>(span_hi, idx[v].data_off + idx[v].raw_len);
    } else {
      col_off[j] = total;
      total += (idx[v].raw_len + 15) & ~15ll;
    }
  }
  if (compression == 0u) {
    for (int j = 0; j < p; j++) col_off[j] = idx[variant_of_col[j]].data_off - span_lo;
    total = span_hi - span_lo;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return kgw_fail_(KGW_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
#define KGW_CUDA_B(call)                                                                                     \
  do {                                                                                                       \
    cudaError_t e_ = (call);                                                                                 \
    if (e_ != cudaSuccess) {                                                                                 \
      free(stage); cudaFree(dblocks); cudaFree(dcomp); cudaFree(dcoff); cudaFree(dclen); cudaFree(doff); cudaFree(dlen); cudaFree(dinfo); cudaFree(dsidx); cudaFree(dH); cudaFree(dstatus); \
      return kgw_fail_(KGW_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                    \
    }                                                                                                        \
  } while (0)
  uint8_t *stage = nullptr, *dblocks = nullptr;
  uint8_t *dcomp = nullptr;
  long long *dcoff = nullptr, *dclen = nullptr;
  long long *doff = nullptr, *dlen = nullptr;
  unsigned long long *dinfo = nullptr;
  int32_t *dsidx = nullptr;
  uint32_t *dH = nullptr;
  int *dstatus = nullptr;
  KGW_CUDA_B(cudaSetDevice(device));
  // zlib blocks are inflated on the device (bgen_inflate_kernel); KGW_BGEN_HOST_INFLATE=1 keeps them on the host threads
  bool device_inflate = (compression == 1u);
  { const char *e = getenv("KGW_BGEN_HOST_INFLATE"); if (e && e[0] == '1') device_inflate = false; }
  std::vector<long long> comp_off, comp_len;
  long long total_comp = 0;
  const uint8_t *comp_src = nullptr;   // what is uploaded for the device inflate: the file's own span, or a packed copy
  if (device_inflate) {
    comp_off.resize(p); comp_len.resize(p);
    long long lo = LLONG_MAX, hi = 0, packed = 0;
    for (int j = 0; j < p; j++) {
      const VariantBlock &vb = idx[variant_of_col[j]];
      comp_len[j] = vb.comp_len;
      lo = std::min<long long>(lo, vb.data_off);
      hi = std::max<long long>(hi, vb.data_off + vb.comp_len);
      packed += (vb.comp_len + 15) & ~15ll;
    }
    if (hi - lo <= packed + packed / 2) {   // most of the span is wanted: no staging copy
      for (int j = 0; j < p; j++) comp_off[j] = idx[variant_of_col[j]].data_off - lo;
      comp_src = file + lo;
      total_comp = hi - lo;
    } else {
      stage = (uint8_t *)malloc((size_t)packed + 16);   // the compressed blocks, one after the other
      if (!stage) return kgw_fail_(KGW_ERR_NOMEM, "staging buffer for the compressed BGEN blocks");
      for (int j = 0; j < p; j++) {
        const VariantBlock &vb = idx[variant_of_col[j]];
        comp_off[j] = total_comp;
        memcpy(stage + total_comp, file + vb.data_off, (size_t)vb.comp_len);
        total_comp += (vb.comp_len + 15) & ~15ll;
      }
      comp_src = stage;
    }
  } else if (compression != 0u) {
    stage = (uint8_t *)malloc((size_t)total + 16);
    if (!stage) return kgw_fail_(KGW_ERR_NOMEM, "staging buffer for the inflated BGEN blocks");
  }

  // inflate the selected blocks (host threads; the reference does this inside genfile::bgen, one variant at a time)
  const Codecs &cd = codecs();
  if (compression == 2u && (!cd.zstd || !cd.zstd_err)) { free(stage); return kgw_fail_(KGW_ERR_UNSUPPORTED, "zstd-compressed BGEN but libzstd.so.1 could not be loaded"); }
  if (compression == 1u && !device_inflate && !cd.zlib) { free(stage); return kgw_fail_(KGW_ERR_UNSUPPORTED, "zlib-compressed BGEN but libz.so.1 could not be loaded"); }
  int nt = host_threads > 0 ? host_threads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min(nt, std::min(p, 64)));
  std::atomic<int> next{0}, failed{0};
  auto worker = [&]() {
    for (;;) {
      const int j0 = next.fetch_add(64);
      if (j0 >= p) break;
      for (int j = j0; j < std::min(p, j0 + 64); j++) {
        const VariantBlock &vb = idx[variant_of_col[j]];
        uint8_t *dst = stage + col_off[j];
        if (compression == 2u) {
          const size_t got = cd.zstd(dst, (size_t)vb.raw_len, file + vb.data_off, (size_t)vb.comp_len);
          if (cd.zstd_err(got) || got != (size_t)vb.raw_len) failed.store(1);
        } else {
          unsigned long dl = (unsigned long)vb.raw_len;
          if (cd.zlib(dst, &dl, file + vb.data_off, (unsigned long)vb.comp_len) != 0 || dl != (unsigned long)vb.raw_len) failed.store(1);
        }
      }
    }
  };
  if (compression != 0u && !device_inflate) {
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(worker);
    worker();
    for (auto &t : th) t.join();
  }
  if (failed.load()) { free(stage); return kgw_fail_(KGW_ERR_INVALID, "a BGEN probability block did not decompress to its stated length"); }

  const int n_haps = 2 * n_samples;
  const int stride_w = (((p + 31) / 32) + 3) & ~3;   // rows padded to 16 bytes
  KGW_CUDA_B(kgw_malloc(&dblocks, (size_t)total + 1024));   // the staged 16-byte loads of the fast path may run past the last block
  KGW_CUDA_B(kgw_malloc(&doff, sizeof(long long) * p));
  KGW_CUDA_B(kgw_malloc(&dinfo, sizeof(unsigned long long) * p));
  KGW_CUDA_B(kgw_malloc(&dlen, sizeof(long long) * p));
  KGW_CUDA_B(cudaMemcpy(dlen, col_len.data(), sizeof(long long) * p, cudaMemcpyHostToDevice));
  KGW_CUDA_B(kgw_malloc(&dsidx, sizeof(int32_t) * n_samples));
  KGW_CUDA_B(kgw_malloc(&dH, sizeof(uint32_t) * (size_t)n_haps * stride_w));
  KGW_CUDA_B(kgw_malloc(&dstatus, sizeof(int)));
  KGW_CUDA_B(cudaMemset(dstatus, 0, sizeof(int)));
  cudaEvent_t e0, e1;
  KGW_CUDA_B(cudaEventCreate(&e0));
  KGW_CUDA_B(cudaEventCreate(&e1));
  KGW_CUDA_B(cudaMemcpy(doff, col_off.data(), sizeof(long long) * p, cudaMemcpyHostToDevice));
  if (device_inflate) {
    KGW_CUDA_B(kgw_malloc(&dcomp, (size_t)total_comp + 16));
    KGW_CUDA_B(kgw_malloc(&dcoff, sizeof(long long) * p));
    KGW_CUDA_B(kgw_malloc(&dclen, sizeof(long long) * p));
    KGW_CUDA_B(cudaMemcpy(dcomp, comp_src, (size_t)total_comp, cudaMemcpyHostToDevice));
    KGW_CUDA_B(cudaMemcpy(dcoff, comp_off.data(), sizeof(long long) * p, cudaMemcpyHostToDevice));
    KGW_CUDA_B(cudaMemcpy(dclen, comp_len.data(), sizeof(long long) * p, cudaMemcpyHostToDevice));
    KGW_CUDA_B(cudaEventRecord(e0));   // the reported device time covers the inflate kernel
    bgen_inflate_kernel<<<(p + INF_WARPS - 1) / INF_WARPS, 32 * INF_WARPS>>>(dcomp, dcoff, dclen, dblocks, doff, dlen, p, dstatus);
    KGW_CUDA_B(cudaGetLastError());
    int st_inf = 0;
    KGW_CUDA_B(cudaMemcpy(&st_inf, dstatus, sizeof(int), cudaMemcpyDeviceToHost));
    if (st_inf) {
      free(stage); cudaFree(dblocks); cudaFree(dcomp); cudaFree(dcoff); cudaFree(dclen); cudaFree(doff); cudaFree(dlen); cudaFree(dinfo); cudaFree(dsidx); cudaFree(dH); cudaFree(dstatus);
      return kgw_fail_(KGW_ERR_INVALID, "a BGEN probability block did not decompress to its stated length");
    }
  } else {
    KGW_CUDA_B(cudaMemcpy(dblocks, compression == 0u ? file + span_lo : stage, (size_t)total, cudaMemcpyHostToDevice));
  }
  KGW_CUDA_B(cudaMemcpy(dsidx, sample_idx, sizeof(int32_t) * n_samples, cudaMemcpyHostToDevice));
  KGW_CUDA_B(cudaMemset(dstatus, 0, sizeof(int)));
  const dim3 grid((unsigned)((stride_w + 3) / 4), (unsigned)((n_haps + 255) / 256));
  if (!device_inflate) KGW_CUDA_B(cudaEventRecord(e0));
  bgen_header_kernel<<<(p + 255) / 256, 256>>>(dblocks, doff, dlen, (int)nfile, p, dinfo, dstatus);
  int st_hdr = 0;
  KGW_CUDA_B(cudaMemcpy(&st_hdr, dstatus, sizeof(int), cudaMemcpyDeviceToHost));
  bool consecutive = true;
  for (int i = 1; i < n_samples; i++) consecutive &= (sample_idx[i] == sample_idx[0] + i);
  if (!(st_hdr & BGEN_NOT_8BIT) && consecutive)
    bgen_unpack8_kernel<<<grid, 256>>>(dblocks, dinfo, sample_idx[0], (int)nfile, n_haps, p, stride_w, dH, dstatus);
  else
    bgen_unpack_kernel<<<grid, 256>>>(dblocks, dinfo, dsidx, (int)nfile, n_haps, p, stride_w, dH, dstatus);
  KGW_CUDA_B(cudaGetLastError());
  KGW_CUDA_B(cudaEventRecord(e1));
  KGW_CUDA_B(cudaEventSynchronize(e1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (kernel_ms) *kernel_ms = ms;
  int st = 0;
  KGW_CUDA_B(cudaMemcpy(&st, dstatus, sizeof(int), cudaMemcpyDeviceToHost));
  st &= ~BGEN_NOT_8BIT;
  if (st == 0)
    KGW_CUDA_B(cudaMemcpy2D(H_out, (size_t)row_bytes, dH, sizeof(uint32_t) * (size_t)stride_w, (size_t)((p + 7) / 8), (size_t)n_haps,
                            cudaMemcpyDeviceToHost));
  free(stage); cudaFree(dblocks); cudaFree(dcomp); cudaFree(dcoff); cudaFree(dclen); cudaFree(doff); cudaFree(dlen); cudaFree(dinfo); cudaFree(dsidx); cudaFree(dH); cudaFree(dstatus);
#undef KGW_CUDA_B
  if (st & BGEN_BAD_HEADER)
    return kgw_fail_(KGW_ERR_INVALID, "BGEN probability block is not phased diploid biallelic layout-2 data of the stated sample count");
  if (st & (BGEN_BAD_PLOIDY | BGEN_BAD_PROB))
    return kgw_fail_(KGW_ERR_INVALID, "Incorrect input: a haplotype probability is not exactly 0 or 1, or a sample is missing (bgen_reader.cpp:48-63)");
  // bits of the last partial byte beyond p stay 0, as in chaplotype (chaplotype.cpp:25-29)
  return KGW_OK;
}

}  // extern "C"
