// microbenchmark: FFMA issue rate per SM vs warps per scheduler, register-tiled 8x8 pattern (a[i]*b[j] + acc[i][j])
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters, float x) {
  float acc[8][8], a[8], b[8];
  for (int i = 0; i < 8; ++i) { a[i] = x + i + threadIdx.x; b[i] = x * i - threadIdx.x; for (int j = 0; j < 8; ++j) acc[i][j] = 0.f; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] += 1e-9f; b[i] -= 1e-9f; }
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j];
  if (s == 12345.f) out[0] = s;
}
int main() {
  float* d; cudaMalloc(&d, 4);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int nw : {4, 8, 12, 16}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    k<<<148, 32 * nw>>>(d, 100, 1.f); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<148, 32 * nw>>>(d, iters, 1.f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ffma = (double)iters * 4 * 64 * 32 * nw;           // thread-FFMAs per SM
    double cyc = ms * 1e-3 * 1.965e9;
    printf("warps/SM %2d (%d per scheduler): %.1f FFMA/clk/SM (peak 128) %.3f ms\n", nw, nw / 4, ffma / cyc, ms);
  }
  return 0;
}
