// how many CTAs of 576 threads / 227 KB shared memory can be co-resident for cluster sizes 1, 2, 4, 8 (148 SMs, GPC layout)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(576, 1) k(float* p) { extern __shared__ float s[]; if (p) p[0] = s[0]; }
int main() {
  const int smem = 232064;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148 / cs * cs); cfg.blockDim = dim3(576); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster %2d: max active clusters %d -> %d CTAs (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
