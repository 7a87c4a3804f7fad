#include <cstdio>
#include <cuda_runtime.h>
template <int RM, int CN>
__global__ void k(float* out, long long* cyc, int iters, float x) {
  float acc[RM][CN], a[RM], b[CN];
  for (int i = 0; i < RM; ++i) a[i] = x + i + threadIdx.x;
  for (int j = 0; j < CN; ++j) b[j] = x * j - threadIdx.x;
  for (int i = 0; i < RM; ++i) for (int j = 0; j < CN; ++j) acc[i][j] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < CN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      a[u % RM] += 1e-9f; b[u % CN] -= 1e-9f;
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < RM; ++i) for (int j = 0; j < CN; ++j) s += acc[i][j];
  if (s == 12345.f) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int RM, int CN> void run(float* d, long long* dc) {
  for (int nw : {4, 8, 12, 16}) {
    int iters = 4000;
    k<RM, CN><<<148, 32 * nw>>>(d, dc, 100, 1.f); cudaDeviceSynchronize();
    k<RM, CN><<<148, 32 * nw>>>(d, dc, iters, 1.f); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    double ffma = (double)iters * 8 * RM * CN * 32 * nw;
    printf("tile %dx%d warps/SM %2d: %.1f FFMA/clk/SM (cycles %lld)\n", RM, CN, nw, ffma / (double)c, c);
  }
}
int main() {
  float* d; cudaMalloc(&d, 4); long long* dc; cudaMalloc(&dc, 8);
  run<8, 8>(d, dc); run<4, 8>(d, dc); run<8, 4>(d, dc); run<12, 8>(d, dc); run<4, 4>(d, dc);
  return 0;
}
