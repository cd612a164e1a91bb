// knockoffgwas_b200/csrc/kgw_pairs.cu -- launchers of the pair class's chain-sequential kernels (kgw_pairs.cuh); its own
// translation unit so that the build stays as long as its slowest unit.
#include "kgw_pairs.cuh"
#include "../../include/kgw.h"
#include <algorithm>
#include <cstdlib>

namespace {

// lanes per chain and states per lane for K states: K <= L * S.  `wide`: sixteen lanes per chain (one pair per warp in
// the draw kernel) -- twice the warps for the same pairs, which is what a short list of pairs needs: these kernels are
// bound by the dependent-issue latency of a warp, and two warps per scheduler leave its issue slots 40 % busy.
template <typename F>
int pair_dispatch(int K, bool wide, F &&fn) {
  if (wide && K <= 64) return fn(std::integral_constant<int, 16>{}, std::integral_constant<int, 4>{});
  if (wide && K <= 112) return fn(std::integral_constant<int, 16>{}, std::integral_constant<int, 7>{});
  if (K <= 32) return fn(std::integral_constant<int, 8>{}, std::integral_constant<int, 4>{});
  if (K <= 56) return fn(std::integral_constant<int, 8>{}, std::integral_constant<int, 7>{});
  if (K <= 104) return fn(std::integral_constant<int, 8>{}, std::integral_constant<int, 13>{});
  if (K <= 128) return fn(std::integral_constant<int, 8>{}, std::integral_constant<int, 16>{});
  return fn(std::integral_constant<int, 16>{}, std::integral_constant<int, 16>{});
}

}  // namespace

// final draws of the pairs P.f_base .. P.n_families-1 (slots by family)
int kgw_pair_draw_launch(const KgwFamParams &P, int sm_count, cudaStream_t st) {
  const int nb = P.n_families - P.f_base;
  if (nb <= 0) return KGW_OK;
  bool wide = nb / 2 < 3 * 4 * sm_count;   // fewer than three warps per scheduler with two pairs per warp
  if (const char *e = getenv("KGW_PAIR_WIDE")) wide = (e[0] == '1');   // development / tests: force either form
  return pair_dispatch(P.K, wide, [&](auto Lc, auto Sc) -> int {
    constexpr int L = decltype(Lc)::value, S = decltype(Sc)::value;
    constexpr int PW = 32 / (2 * L), G = 2 * PW;
    const int grid = (nb + 4 * PW - 1) / (4 * PW);
    // loci per staged block: what keeps a warp's two stages near 14 KB (8 warps per SM)
    int RB = 8;
    auto warp_bytes = [&](int rb) { return 2 * (size_t)(G * (kgw::PairStage::buf((unsigned)rb * P.K * 8) + kgw::PairStage::buf((unsigned)rb * P.K)) + rb * sizeof(KgwFamLocus)) + 16; };
    while (RB > 1 && warp_bytes(RB) > (wide ? 8 : 15) * 1024) RB /= 2;   // (wide: 16+ warps per SM resident)
    const size_t smem = 4 * warp_bytes(RB);
    auto fn = P.uniform_mu ? kgw::pair_draw_kernel<L, S, true> : kgw::pair_draw_kernel<L, S, false>;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return KGW_ERR_CUDA;
    fn<<<grid, 128, smem, st>>>(P, RB);
    return cudaGetLastError() == cudaSuccess ? KGW_OK : KGW_ERR_CUDA;
  });
}

// the DRAW operations ops[0 .. n_ops) of the same batch; P.work is zero
int kgw_pair_ko_launch(const KgwFamParams &P, const KgwPairOp *ops, int n_ops, int sm_count, cudaStream_t st) {
  if (n_ops <= 0) return KGW_OK;
  // sixteen lanes per operation beyond 56 states: thirteen or more states per lane cost 230+ registers (two warps per
  // scheduler) and a long chain of dependent adds per draw, and the kernel then waits instead of issuing
  bool wide = P.K > 56;
  if (const char *e = getenv("KGW_PAIR_KO_WIDE")) wide = (e[0] == '1');   // development
  return pair_dispatch(P.K, wide, [&](auto Lc, auto Sc) -> int {
    constexpr int L = decltype(Lc)::value, S = decltype(Sc)::value;
    constexpr int GW = 32 / L;   // lane groups per warp
    int per_sm = 4;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kgw::pair_ko_kernel<L, S>, 128, 0) != cudaSuccess) return KGW_ERR_CUDA;
    const int grid = std::max(1, std::min((n_ops + 4 * GW - 1) / (4 * GW), sm_count * std::max(1, per_sm)));
    kgw::pair_ko_kernel<L, S><<<grid, 128, 0, st>>>(P, ops, n_ops);
    return cudaGetLastError() == cudaSuccess ? KGW_OK : KGW_ERR_CUDA;
  });
}
