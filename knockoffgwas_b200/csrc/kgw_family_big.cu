// knockoffgwas_b200/csrc/kgw_family_big.cu -- the BIG class of the related-branch kernel (kgw_family.cuh): families of more
// than KGW_FAM_MAXN haplotypes or with nodes of more than KGW_FAM_MAXNB IBD neighbours.  Its own translation unit so
// that the build stays as long as its slowest unit.
#include "kgw_family.cuh"

typedef void (*fam_kernel_fn)(const KgwFamParams);
fam_kernel_fn kgw_family_kernel_big_(int SF, bool uniform_mu) {
  if (uniform_mu)
    return SF == 1 ? kgw::family_kernel<1, true, true> : (SF == 2 ? kgw::family_kernel<2, true, true> : (SF == 4 ? kgw::family_kernel<4, true, true> : kgw::family_kernel<8, true, true>));
  return SF == 1 ? kgw::family_kernel<1, false, true> : (SF == 2 ? kgw::family_kernel<2, false, true> : (SF == 4 ? kgw::family_kernel<4, false, true> : kgw::family_kernel<8, false, true>));
}
