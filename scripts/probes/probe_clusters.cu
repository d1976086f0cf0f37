// Probe: how many clusters of size 1/2/4/8 (one CTA per SM, ~200 KB dynamic smem) can be
// co-resident on this GPU, plus the L2 / clock facts the retrieval kernel's schedule depends on.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dummy(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("sms %d l2 %d MB persistL2max %zu clock_khz %d mem_khz %d\n", prop.multiProcessorCount, prop.l2CacheSize >> 20,
         (size_t)prop.persistingL2CacheMaxSize, prop.clockRate, prop.memoryClockRate);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148 * 4 / cs * cs); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster %d: max active clusters %d (%d SMs) err=%s\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
