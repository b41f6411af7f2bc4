// Probe: kernel-to-kernel gap when a kernel with a 200 KB dynamic shared-memory footprint is followed by a small
// kernel, as a function of the small kernel's preferred shared-memory carve-out (does the SM reconfigure its
// L1 / shared split between the two?).  Gap = first globaltimer of B minus last globaltimer of A.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 profiles/probes/carveout_gap.cu -o profiles/probes/bin/carveout_gap
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void big(unsigned long long* last, int spin) {
  extern __shared__ float sm[];
  sm[threadIdx.x] = threadIdx.x;
  __syncthreads();
  float a = sm[(threadIdx.x + 1) % blockDim.x];
  for (int i = 0; i < spin; ++i) a = a * 1.0001f + 0.5f;
  if (a == 123.f) sm[0] = a;
  __syncthreads();
  if (threadIdx.x == 0) atomicMax(last, gt());
}
__global__ void small_k(unsigned long long* first, float* x) {
  if (threadIdx.x == 0) atomicMin(first, gt());
  x[blockIdx.x * blockDim.x + threadIdx.x] += 1.f;
}
int main() {
  unsigned long long *d; cudaMalloc(&d, 16);
  float* x; cudaMalloc(&x, 1184 * 256 * 4); cudaMemset(x, 0, 1184 * 256 * 4);
  cudaFuncSetAttribute(big, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int carve : {-1, 100}) {
    cudaFuncSetAttribute(small_k, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    for (int grid_b : {64, 1184}) {
      double tot = 0; int n = 0;
      for (int it = 0; it < 30; ++it) {
        unsigned long long init[2] = {0ull, ~0ull};
        cudaMemcpy(d, init, 16, cudaMemcpyHostToDevice);
        big<<<148, 512, 200 * 1024>>>(d, 20000);
        small_k<<<grid_b, 256>>>(d + 1, x);
        cudaDeviceSynchronize();
        unsigned long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (it >= 5) { tot += (double)(h[1] - h[0]); ++n; }
      }
      printf("carve-out %4d, small grid %4d: gap after the 200 KB-smem kernel %.2f us\n", carve, grid_b, tot / n / 1e3);
    }
  }
  // reference: small after small
  {
    double tot = 0; int n = 0;
    for (int it = 0; it < 30; ++it) {
      unsigned long long init[2] = {0ull, ~0ull};
      cudaMemcpy(d, init, 16, cudaMemcpyHostToDevice);
      big<<<148, 512, 4096>>>(d, 20000);
      small_k<<<1184, 256>>>(d + 1, x);
      cudaDeviceSynchronize();
      unsigned long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      if (it >= 5) { tot += (double)(h[1] - h[0]); ++n; }
    }
    printf("reference (predecessor with 4 KB smem): gap %.2f us\n", tot / n / 1e3);
  }
  {
    cudaFuncSetAttribute(small_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    double tot = 0; int n = 0;
    for (int it = 0; it < 30; ++it) {
      unsigned long long init[2] = {0ull, ~0ull};
      cudaMemcpy(d, init, 16, cudaMemcpyHostToDevice);
      big<<<148, 512, 200 * 1024>>>(d, 20000);
      small_k<<<148, 256, 200 * 1024>>>(d + 1, x);
      cudaDeviceSynchronize();
      unsigned long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      if (it >= 5) { tot += (double)(h[1] - h[0]); ++n; }
    }
    printf("follower also launched with 200 KB of dynamic smem: gap %.2f us\n", tot / n / 1e3);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
  return 0;
}
