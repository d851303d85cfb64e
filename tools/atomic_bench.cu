// Micro-benchmark: FP64 atomicAdd throughput into 36-double blocks (the Schur accumulation pattern).
//   A: one lane per (pair -> block), 36 sequential REDs per lane (32 different sectors per instruction)
//   B: one warp per pair, lanes 0..35 cover the block (9-10 sectors per instruction)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/atomic_bench.cu -o tools/atomic_bench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
__global__ void varA(const int *__restrict__ blk, int npairs, double *__restrict__ S) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x; if (p >= npairs) return;
  double *b = S + 36 * (size_t)blk[p];
  #pragma unroll
  for (int e = 0; e < 36; ++e) atomicAdd(b + e, 1.0 + e);
}
__global__ void varB(const int *__restrict__ blk, int npairs, double *__restrict__ S) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
  for (int p0 = warp * 32; p0 < npairs; p0 += nw * 32) {
    const int mine = p0 + lane < npairs ? blk[p0 + lane] : -1;
    for (int q = 0; q < 32; ++q) {
      const int bi = __shfl_sync(0xffffffffu, mine, q); if (bi < 0) break;
      double *b = S + 36 * (size_t)bi;
      atomicAdd(b + lane, 1.0 + lane);
      if (lane < 4) atomicAdd(b + 32 + lane, 33.0 + lane);
    }
  }
}
int main() {
  const int npairs = 5500000, nblocks = 28000;
  std::vector<int> h(npairs); srand(1);
  // pairs of consecutive "points" hit nearby blocks (as in the real scene), otherwise random
  for (int i = 0; i < npairs; ++i) h[i] = (int)(((long long)rand() * 7919 + i / 55) % nblocks);
  int *d; double *S; cudaMalloc(&d, npairs * 4); cudaMalloc(&S, (size_t)nblocks * 36 * 8);
  cudaMemcpy(d, h.data(), npairs * 4, cudaMemcpyHostToDevice); cudaMemset(S, 0, (size_t)nblocks * 36 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    float ms;
    cudaEventRecord(e0); varA<<<(npairs + 127) / 128, 128>>>(d, npairs, S); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("A lane-per-block : %.3f ms  %.1f G atomics/s\n", ms, npairs * 36.0 / ms / 1e6);
    cudaEventRecord(e0); varB<<<148 * 8, 256>>>(d, npairs, S); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("B warp-per-block : %.3f ms  %.1f G atomics/s\n", ms, npairs * 36.0 / ms / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
