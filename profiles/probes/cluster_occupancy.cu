// Probe: how many thread-block clusters of size 8 / 16 can be co-resident on this GPU for a
// 1-CTA-per-SM kernel (large dynamic shared memory)?  Build: nvcc -arch=sm_100a -o probe probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* p) { extern __shared__ float s[]; if (p) p[0] = s[0]; }
int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("%s SMs=%d smemPerSM=%zu\n", prop.name, prop.multiProcessorCount, prop.sharedMemPerMultiprocessor);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int smem_kb : {32, 100, 166, 214}) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
    for (int cs : {1, 2, 4, 8, 16}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem_kb * 1024;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
      printf("smem %3d KB cluster %2d -> max active clusters %d (%d CTAs) %s\n", smem_kb, cs, n, n * cs, e ? cudaGetErrorString(e) : "");
    }
  }
  return 0;
}
