// knockoffgwas_b200/csrc/kgw_variants.h -- registry of the compiled sampler-kernel instantiations.
#pragma once
#include <stddef.h>
struct KgwParams;
typedef void (*kgw_kernel_fn)(const KgwParams);
struct KgwVariant {
  int L, S, B;              // lanes per haplotype, states per lane, checkpoint spacing
  bool pad;                 // false: compiled for K == L*S exactly (no padded states)
  int max_wpc;              // the kernel may be launched with CTAs of 1 .. max_wpc warps
  kgw_kernel_fn fn;         // compiled for the densest residency the shared memory allows (12 or 16 warps per SM)
  kgw_kernel_fn fn_wide;    // compiled for 8 warps per SM (255 registers per thread); == fn when that is the densest
  size_t smem_warp;         // dynamic shared memory per warp of `fn`
  size_t smem_warp_wide;    // ... and of `fn_wide`
};
const KgwVariant *kgw_variants_L1(int *n);
const KgwVariant *kgw_variants_L2(int *n);
const KgwVariant *kgw_variants_L4(int *n);
const KgwVariant *kgw_variants_L8(int *n);
