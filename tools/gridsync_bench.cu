// micro-benchmark: cost of cooperative grid.sync() and of the vsum pattern on this GPU
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
__global__ void k_sync(int n, double *out) {
  cg::grid_group g = cg::this_grid();
  for (int i = 0; i < n; ++i) g.sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = 1;
}
__global__ void k_vsum(int n, double *part, double *out) {
  cg::grid_group g = cg::this_grid();
  __shared__ double tot[8];
  double acc = 0;
  for (int i = 0; i < n; ++i) {
    if (threadIdx.x < 4) part[blockIdx.x * 320 + threadIdx.x] = i + blockIdx.x;
    g.sync();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (w < 4) { double t = 0; for (int b = lane; b < gridDim.x; b += 32) t += part[b * 320 + w];
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o); if (lane == 0) tot[w] = t; }
    __syncthreads();
    acc += tot[0];
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = acc;
}
int main() {
  int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  double *out, *part; cudaMalloc(&out, 8); cudaMalloc(&part, 148 * 320 * 8 * 2);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads : {256, 1024}) for (int blocks : {p.multiProcessorCount, p.multiProcessorCount / 2, 32}) {
    int n = 2000; void *a1[] = {&n, &out};
    cudaLaunchCooperativeKernel((void *)k_sync, dim3(blocks), dim3(threads), a1, 0, 0); cudaDeviceSynchronize();
    cudaEventRecord(e0); cudaLaunchCooperativeKernel((void *)k_sync, dim3(blocks), dim3(threads), a1, 0, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    void *a2[] = {&n, &part, &out};
    cudaEventRecord(e0); cudaLaunchCooperativeKernel((void *)k_vsum, dim3(blocks), dim3(threads), a2, 0, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms2; cudaEventElapsedTime(&ms2, e0, e1);
    printf("blocks %d threads %d: grid.sync %.2f us, vsum round %.2f us  (%s)\n", blocks, threads, ms * 1e3 / n, ms2 * 1e3 / n, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
