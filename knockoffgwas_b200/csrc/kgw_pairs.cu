// knockoffgwas_b200/csrc/kgw_pairs.cu -- launchers of the pair class's chain-sequential kernels (kgw_pairs.cuh); its own
// translation unit so that the build stays as long as its slowest unit.
#include "kgw_pairs.cuh"
#include "../../include/kgw.h"
#include <algorithm>

namespace {

// lanes per chain and states per lane for K states: K <= L * S
template <typename F>
int pair_dispatch(int K, F &&fn) {
  if (K <= 32) return fn(std::integral_constant<int, 8>{}, std::integral_constant<int, 4>{});
  if (K <= 56) return fn(std::integral_constant<int, 8>{}, std::integral_constant<int, 7>{});
  if (K <= 104) return fn(std::integral_constant<int, 8>{}, std::integral_constant<int, 13>{});
  if (K <= 128) return fn(std::integral_constant<int, 8>{}, std::integral_constant<int, 16>{});
  return fn(std::integral_constant<int, 16>{}, std::integral_constant<int, 16>{});
}

}  // namespace

// final draws of the pairs P.f_base .. P.n_families-1 (slots by family)
int kgw_pair_draw_launch(const KgwFamParams &P, cudaStream_t st) {
  const int nb = P.n_families - P.f_base;
  if (nb <= 0) return KGW_OK;
  return pair_dispatch(P.K, [&](auto Lc, auto Sc) -> int {
    constexpr int L = decltype(Lc)::value, S = decltype(Sc)::value;
    constexpr int PW = 32 / (2 * L);
    const int grid = (nb + 4 * PW - 1) / (4 * PW);
    if (P.uniform_mu) kgw::pair_draw_kernel<L, S, true><<<grid, 128, 0, st>>>(P);
    else kgw::pair_draw_kernel<L, S, false><<<grid, 128, 0, st>>>(P);
    return cudaGetLastError() == cudaSuccess ? KGW_OK : KGW_ERR_CUDA;
  });
}

// the DRAW operations ops[0 .. n_ops) of the same batch; P.work is zero
int kgw_pair_ko_launch(const KgwFamParams &P, const KgwPairOp *ops, int n_ops, int sm_count, cudaStream_t st) {
  if (n_ops <= 0) return KGW_OK;
  return pair_dispatch(P.K, [&](auto Lc, auto Sc) -> int {
    constexpr int L = decltype(Lc)::value, S = decltype(Sc)::value;
    constexpr int GW = 32 / L;   // lane groups per warp
    int per_sm = 4;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kgw::pair_ko_kernel<L, S>, 128, 0) != cudaSuccess) return KGW_ERR_CUDA;
    const int grid = std::max(1, std::min((n_ops + 4 * GW - 1) / (4 * GW), sm_count * std::max(1, per_sm)));
    kgw::pair_ko_kernel<L, S><<<grid, 128, 0, st>>>(P, ops, n_ops);
    return cudaGetLastError() == cudaSuccess ? KGW_OK : KGW_ERR_CUDA;
  });
}
