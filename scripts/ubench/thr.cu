// development micro-benchmark: FP64 throughput of one warp vs number of independent chains, looped
#include <cstdio>
#include <cuda_runtime.h>
template <int NCH>
__device__ __forceinline__ long long run(double &acc, double x, double y, int iters) {
  double r[NCH];
#pragma unroll
  for (int i = 0; i < NCH; i++) r[i] = x + i;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int q = 0; q < NCH; q++) r[q] = fma(r[q], y, x);
    }
  }
  long long t1 = clock64();
#pragma unroll
  for (int i = 0; i < NCH; i++) acc += r[i];
  return t1 - t0;
}
__global__ void k(double *out, long long *cyc, double x, double y, int iters) {
  double acc = 0;
  long long c1 = run<1>(acc, x, y, iters), c2 = run<2>(acc, x, y, iters), c4 = run<4>(acc, x, y, iters), c8 = run<8>(acc, x, y, iters),
            c16 = run<16>(acc, x, y, iters), c25 = run<25>(acc, x, y, iters), c32 = run<32>(acc, x, y, iters);
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = c1; cyc[1] = c2; cyc[2] = c4; cyc[3] = c8; cyc[4] = c16; cyc[5] = c25; cyc[6] = c32; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 8 * 1024 * 1024); cudaMallocManaged(&cyc, 8 * 16);
  const int iters = 256;
  int nch[] = {1, 2, 4, 8, 16, 25, 32};
  for (int warps : {1, 4, 8, 16}) {
    for (int rep = 0; rep < 2; rep++) { k<<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999, iters); cudaDeviceSynchronize(); }
    printf("warps per SM = %d (cycles per DFMA warp-instruction, per warp):", warps);
    for (int i = 0; i < 7; i++) printf("  %dch %.2f", nch[i], (double)cyc[i] / (iters * 4.0 * nch[i]));
    printf("\n");
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
