// development micro-benchmark: dependent-chain latencies on sm_100a (one warp, clock64)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double x, double y) {
  __shared__ double sm[1024];
  sm[threadIdx.x] = x + threadIdx.x;
  __syncwarp();
  double a = x;
  long long t0, t1;
  // DFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) a = fma(a, y, x);
  t1 = clock64(); cyc[0] = t1 - t0;
  // DADD chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) a = a + y;
  t1 = clock64(); cyc[1] = t1 - t0;
  // DMUL chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) a = a * y;
  t1 = clock64(); cyc[2] = t1 - t0;
  // 4 independent DFMA chains (throughput)
  double b = x + 1, c = x + 2, d = x + 3;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) { a = fma(a, y, x); b = fma(b, y, x); c = fma(c, y, x); d = fma(d, y, x); }
  t1 = clock64(); cyc[3] = t1 - t0;
  a += b + c + d;
  // DADD + DSETP + predicated IADD (the search loop)
  int cnt = 0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) { a = a + y; cnt += (a <= x) ? 1 : 0; }
  t1 = clock64(); cyc[4] = t1 - t0;
  // shuffle chain (64-bit)
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) a = __shfl_xor_sync(0xffffffffu, a, 16) + y;
  t1 = clock64(); cyc[5] = t1 - t0;
  // LDS dependent chain
  int idx = threadIdx.x;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) { double v = sm[idx]; idx = ((int)v) & 1023; }
  t1 = clock64(); cyc[6] = t1 - t0;
  // rcp.approx + 2 NR
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); double e = fma(-a, r, 1.0); r = fma(r, e, r); e = fma(-a, r, 1.0); a = fma(r, e, r) + 1.5; }
  t1 = clock64(); cyc[7] = t1 - t0;
  // 1.0/x division
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) a = 1.0 / a + 1.5;
  t1 = clock64(); cyc[8] = t1 - t0;
  // 16 independent DFMA chains, 8 warps? (single warp throughput)
  double r[16];
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = x + i;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int q = 0; q < 16; q++) r[q] = fma(r[q], y, x);
  }
  t1 = clock64(); cyc[9] = t1 - t0;
#pragma unroll
  for (int i = 0; i < 16; i++) a += r[i];
  out[threadIdx.x] = a + cnt + idx;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 8 * 1024); cudaMallocManaged(&cyc, 8 * 16);
  for (int rep = 0; rep < 2; rep++) { k<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999); cudaDeviceSynchronize(); }
  const char *names[] = {"DFMA dep chain /op", "DADD dep chain /op", "DMUL dep chain /op", "4 indep DFMA /op", "DADD+DSETP+IADD /iter", "SHFL64+DADD /iter", "LDS64 dep /iter", "rcp+2NR(+add) /iter", "1.0/x (+add) /iter", "16 indep DFMA /op"};
  int div[] = {256, 256, 256, 1024, 256, 64, 64, 64, 64, 1024};
  for (int i = 0; i < 10; i++) printf("%-28s %.2f cycles\n", names[i], (double)cyc[i] / div[i]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
